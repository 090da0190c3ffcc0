/* =====================================================================================
 * TEST INFRASTRUCTURE — CPU ORACLE.  NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's per-step env hot path
 * (ChuaCheowHuan/gym-continuousDoubleAuction @ /root/reference, all paths below relative to
 * gym_continuousDoubleAuction/envs/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (csrc/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this file against the
 * UNMODIFIED Python reference (imported through oracle/ref_stub.py) on identical seeds and
 * actions in this container, and tests/test_oracle_golden.py checks it against fixtures in
 * tests/golden/ that were generated from the reference by oracle/gen_golden.py.
 *
 * It deliberately keeps the reference's DATA STRUCTURES (sorted price levels, a FIFO
 * doubly-linked list per level, an insertion-ordered order map per side) so that it is an
 * independent check on the CUDA path, which uses a flat order pool + warp reductions.
 *
 * Decimal -> int64: with tick_size integral every price, size and trade value is an integer
 * and |pos|*VWAP obeys an exact integer recurrence (cost basis C, see orc_process_acc), so the
 * ledger is held in int64.  The reference's Decimal(prec 28) carries <=1e-20 residues from the
 * VWAP division; they are invisible at the 1e-6 obs/reward tolerance.
 * ===================================================================================== */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "np_rng.h"
#ifdef ORC_DEC128
#include "dec128.h"   /* fixed-width (unsigned __int128) form of the same arithmetic: the shape of the future device ledger */
#else
#include "dec28.h"
static int dec_range_errors = 0;
#endif

#define ORC_K 10          /* k_rows  (config/tunable_constants.json: observation_layout) */
#define ORC_SNAP 42       /* book_rows*k_rows + extra_dim */
#define ORC_MAX_AGENTS 32
#define ORC_MAX_HIST 16
#define ORC_MAX_FILLS 256

/* status bits (sticky per market) */
#define ORC_ST_FILL_OVERFLOW 1u
#define ORC_ST_BAD_SIZE 2u
#define ORC_ST_BAD_ACTION 4u
#define ORC_ST_LEDGER_MISMATCH 8u   /* decimal_ledger mode: a Decimal(28) value is not within 1e-9 of its exact-integer twin */

typedef struct {
    int32_t num_agents, n_hist, max_step, tick;
    int64_t init_cash;
    int32_t min_size, mkt_max_size, limit_size_multiple;
    int32_t price_lo, price_hi; /* initial_price_min / initial_price_max (inclusive) */
    double order_penalty, trade_penalty, drawdown_penalty, passive_bonus, loss_multiplier;
    int32_t decimal_ledger; /* 1: ALSO keep the money fields as Decimal(prec 28) exactly like the reference (dec28.h) and take the
                               cash-gate / bankruptcy / high-water-mark decisions on them; 0 (default): exact int64 ledger only */
} OrcConfig;

/* ---- orderbook/order.py:4-36, orderlist.py, ordertree.py ---------------------------------- */
typedef struct {
    int64_t qty, price;
    int32_t trader;
    int64_t order_id, ts;
    int next, prev;         /* FIFO list inside the price level (orderlist.py:45-78) */
    int map_next, map_prev; /* insertion order of OrderTree.order_map (ordertree.py:15,54,77) */
    int level_price_valid;
} OrcOrder;

typedef struct {
    int64_t price;
    int head, tail, length;
    int64_t volume;
} OrcLevel;

typedef struct {
    OrcOrder *orders; int cap, free_head;
    OrcLevel *levels; int n_levels, cap_levels; /* ascending price == SortedDict order */
    int map_head, map_tail, n_orders;
    int64_t volume;
} OrcTree;

typedef struct {
    int32_t time, price, qty, maker, maker_oid, maker_left, taker, taker_side;
} OrcFill;

/* ---- account/account.py:12-82 ------------------------------------------------------------- */
typedef struct {
    int64_t cash, hold, pv, C, nav, prev_nav, max_nav;
    int64_t pos;
    int32_t num_trades, num_trades_step, num_passive_fills_step, order_step_placed, num_rejected_step;
    int32_t is_pass;
    double reward, terms[5], drawdown;
    struct OrcDec *d;   /* decimal_ledger mode: the reference's Decimal fields of this account (kept OUT of this struct so that the
                           default build of the CPU baseline steps exactly the memory it always did); NULL otherwise */
} OrcAcct;
typedef struct OrcDec { dec cash, hold, pv, vwap, nav, prev_nav, max_nav; } OrcDec;

typedef struct {
    OrcTree bids, asks;
    int64_t time, next_order_id;
    int tape_nonempty; int64_t tape_last;
    int64_t last_price;
    OrcAcct acc[ORC_MAX_AGENTS];
    int32_t last_act[ORC_MAX_AGENTS][4];   /* the reference's LOB_actions of the last step: type, side, size, price per agent; side -1 = pass / absent */
    int32_t t_step; uint32_t done_mask;
    orc_rng rng;
    float raw[4 * ORC_K];
    float hist[ORC_MAX_HIST][ORC_SNAP];
    OrcFill fills[ORC_MAX_FILLS]; int n_fills;
    uint32_t status;
    int64_t best_bid, best_ask; /* set_market_snapshot, 0 = None */
} OrcMarket;

typedef struct {
    OrcConfig cfg;
    int M;
    OrcMarket *mk;
    OrcDec *dec;   /* [M][ORC_MAX_AGENTS] Decimal twins (decimal_ledger mode), else NULL */
} OrcEnv;

/* ---------------- tree primitives ---------------------------------------------------------- */
static void tree_init(OrcTree *t) {
    memset(t, 0, sizeof(*t));
    t->cap = 64;
    t->orders = (OrcOrder *)malloc(sizeof(OrcOrder) * t->cap);
    for (int i = 0; i < t->cap; ++i) t->orders[i].next = i + 1;
    t->orders[t->cap - 1].next = -1;
    t->free_head = 0;
    t->cap_levels = 32;
    t->levels = (OrcLevel *)malloc(sizeof(OrcLevel) * t->cap_levels);
    t->map_head = t->map_tail = -1;
}
static void tree_free(OrcTree *t) { free(t->orders); free(t->levels); }
static void tree_clear(OrcTree *t) { tree_free(t); tree_init(t); }

static int tree_alloc(OrcTree *t) {
    if (t->free_head < 0) {
        int old = t->cap;
        t->cap *= 2;
        t->orders = (OrcOrder *)realloc(t->orders, sizeof(OrcOrder) * t->cap);
        for (int i = old; i < t->cap; ++i) t->orders[i].next = i + 1;
        t->orders[t->cap - 1].next = -1;
        t->free_head = old;
    }
    int i = t->free_head;
    t->free_head = t->orders[i].next;
    return i;
}
static int tree_find_level(const OrcTree *t, int64_t price) {
    for (int i = 0; i < t->n_levels; ++i) if (t->levels[i].price == price) return i;
    return -1;
}
/* ordertree.py:29-32 create_price */
static int tree_create_level(OrcTree *t, int64_t price) {
    if (t->n_levels == t->cap_levels) {
        t->cap_levels *= 2;
        t->levels = (OrcLevel *)realloc(t->levels, sizeof(OrcLevel) * t->cap_levels);
    }
    int pos = 0;
    while (pos < t->n_levels && t->levels[pos].price < price) ++pos;
    memmove(&t->levels[pos + 1], &t->levels[pos], sizeof(OrcLevel) * (t->n_levels - pos));
    t->levels[pos].price = price;
    t->levels[pos].head = t->levels[pos].tail = -1;
    t->levels[pos].length = 0;
    t->levels[pos].volume = 0;
    t->n_levels++;
    return pos;
}
/* ordertree.py:44-55 insert_order  (+ orderlist.py:45-57 append_order) */
static int tree_insert(OrcTree *t, int64_t price, int64_t qty, int trader, int64_t order_id, int64_t ts) {
    int li = tree_find_level(t, price);
    if (li < 0) li = tree_create_level(t, price);
    int oi = tree_alloc(t);
    OrcOrder *o = &t->orders[oi];
    o->qty = qty; o->price = price; o->trader = trader; o->order_id = order_id; o->ts = ts;
    OrcLevel *L = &t->levels[li];
    o->next = -1; o->prev = L->tail;
    if (L->tail >= 0) t->orders[L->tail].next = oi; else L->head = oi;
    L->tail = oi;
    L->length++; L->volume += qty;
    o->map_next = -1; o->map_prev = t->map_tail;
    if (t->map_tail >= 0) t->orders[t->map_tail].map_next = oi; else t->map_head = oi;
    t->map_tail = oi;
    t->n_orders++; t->volume += qty;
    return oi;
}
/* ordertree.py:70-77 remove_order_by_id (+ orderlist.py:59-78 remove_order, ordertree.py:34-36) */
static void tree_remove(OrcTree *t, int oi) {
    OrcOrder *o = &t->orders[oi];
    int li = tree_find_level(t, o->price);
    OrcLevel *L = &t->levels[li];
    t->n_orders--; t->volume -= o->qty;
    L->volume -= o->qty; L->length--;
    if (o->prev >= 0) t->orders[o->prev].next = o->next; else L->head = o->next;
    if (o->next >= 0) t->orders[o->next].prev = o->prev; else L->tail = o->prev;
    if (L->length == 0) {
        memmove(&t->levels[li], &t->levels[li + 1], sizeof(OrcLevel) * (t->n_levels - li - 1));
        t->n_levels--;
    }
    if (o->map_prev >= 0) t->orders[o->map_prev].map_next = o->map_next; else t->map_head = o->map_next;
    if (o->map_next >= 0) t->orders[o->map_next].map_prev = o->map_prev; else t->map_tail = o->map_prev;
    o->next = t->free_head;
    t->free_head = oi;
}

/* ---------------- ledger: account.py:124-231, cash_processor.py, calculate.py --------------- */
/* party: 0 = init_party, 1 = counter_party.  side: 0 bid, 1 ask (that party's side). */
#define DI(x) dec_from_i64((int64_t)(x))
static void cash_increase(OrcAcct *a, int party, int64_t v) { /* cash_processor.py:31-36 */
    if (party == 0) a->cash -= v; else a->hold -= v;
    if (a->d) { if (party == 0) a->d->cash = dec_sub(a->d->cash, DI(v)); else a->d->hold = dec_sub(a->d->hold, DI(v)); }
}
static void cash_decrease(OrcAcct *a, int party, int64_t v) { /* cash_processor.py:38-45 */
    if (party == 0) a->cash += v; else { a->cash += v; a->hold -= v; a->cash += v; }
    if (a->d) {
        if (party == 0) a->d->cash = dec_add(a->d->cash, DI(v));
        else { a->d->cash = dec_add(a->d->cash, DI(v)); a->d->hold = dec_sub(a->d->hold, DI(v)); a->d->cash = dec_add(a->d->cash, DI(v)); }
    }
}
/* escrow moves with exact integer values: cash_processor.py:15-29 (dir = +1: cash -> hold), :55-62 / :85-97 (dir = -1: hold -> cash;
 * the reference subtracts from cash_on_hold first, then adds to cash) */
static void escrow_move(OrcAcct *a, int64_t v, int dir) {
    if (dir > 0) { a->cash -= v; a->hold += v; } else { a->hold -= v; a->cash += v; }
    if (a->d) {
        if (dir > 0) { a->d->cash = dec_sub(a->d->cash, DI(v)); a->d->hold = dec_add(a->d->hold, DI(v)); }
        else { a->d->hold = dec_sub(a->d->hold, DI(v)); a->d->cash = dec_add(a->d->cash, DI(v)); }
    }
}
/* calculate.py:24-33 cal_profit + "position_val = raw_val + profit" on the Decimal fields */
static void dl_set_pv(OrcAcct *a, int is_long, dec raw, dec mkt) {
    dec profit = is_long ? dec_sub(mkt, raw) : dec_sub(raw, mkt);
    a->d->pv = dec_add(raw, profit);
}
/* account.py:135-149 _covered: position_val = raw + profit; cash += position_val - mkt_val */
static int64_t acct_covered(OrcAcct *a, int is_long, int64_t price) {
    int64_t ap = a->pos < 0 ? -a->pos : a->pos;
    int64_t raw = a->C, mkt = ap * price;
    int64_t profit = is_long ? mkt - raw : raw - mkt; /* calculate.py:24-33 */
    a->pv = raw + profit;
    a->cash += a->pv - mkt; /* cash_processor.py:47-53 size_zero_cash_transfer */
    a->pv = 0; a->C = 0;
    if (a->d) {
        dec draw = dec_mul(DI(ap), a->d->vwap), dmkt = DI(mkt);
        dl_set_pv(a, is_long, draw, dmkt);
        a->d->cash = dec_add(a->d->cash, dec_sub(a->d->pv, dmkt));
        a->d->pv = dec_zero(); a->d->vwap = dec_zero();
    }
    return mkt;
}
static void acct_size_increase(OrcAcct *a, int is_long, int party, int64_t q, int64_t price, int64_t tv) {
    /* account.py:124-133 : VWAP' = (|pos|*VWAP + tv)/total  =>  C' = C + tv */
    int64_t ap = a->pos < 0 ? -a->pos : a->pos;
    int64_t total = ap + q;
    a->C += tv;
    int64_t raw = a->C, mkt = total * price;
    a->pv = raw + (is_long ? mkt - raw : raw - mkt);
    if (a->d) {
        a->d->vwap = dec_div(dec_add(dec_mul(DI(ap), a->d->vwap), DI(tv)), DI(total));
        dl_set_pv(a, is_long, dec_mul(DI(total), a->d->vwap), DI(mkt));
    }
    cash_increase(a, party, tv);
}
static void acct_size_decrease(OrcAcct *a, int is_long, int party, int64_t q, int64_t price, int64_t tv) {
    /* account.py:151-161 */
    int64_t ap = a->pos < 0 ? -a->pos : a->pos;
    int64_t left = ap - q;
    if (left > 0) {
        a->C -= tv; /* VWAP' = (|pos|*VWAP - tv)/left */
        int64_t raw = a->C, mkt = left * price;
        a->pv = raw + (is_long ? mkt - raw : raw - mkt);
        if (a->d) {
            a->d->vwap = dec_div(dec_sub(dec_mul(DI(ap), a->d->vwap), DI(tv)), DI(left));
            dl_set_pv(a, is_long, dec_mul(DI(left), a->d->vwap), DI(mkt));
        }
    } else {
        acct_covered(a, is_long, price);
    }
    cash_decrease(a, party, tv);
}
static void acct_covered_side_chg(OrcAcct *a, int is_long, int party, int64_t q, int64_t price) {
    /* account.py:163-171 */
    int64_t ap = a->pos < 0 ? -a->pos : a->pos;
    int64_t mkt = acct_covered(a, is_long, price);
    cash_decrease(a, party, mkt);
    int64_t new_size = q - ap;
    a->pv = new_size * price;
    a->C = new_size * price; /* VWAP = price */
    if (a->d) { a->d->pv = DI(new_size * price); a->d->vwap = DI(price); }
    cash_increase(a, party, a->pv);
}
/* account.py:215-231 process_acc */
static void orc_process_acc(OrcAcct *a, int party, int side, int64_t q, int64_t price) {
    a->num_trades++; a->num_trades_step++;
    if (party == 1) a->num_passive_fills_step++;
    int64_t tv = q * price;
    if (a->pos > 0) { /* account.py:178-185 _net_long */
        if (side == 0) acct_size_increase(a, 1, party, q, price, tv);
        else if (a->pos >= q) acct_size_decrease(a, 1, party, q, price, tv);
        else acct_covered_side_chg(a, 1, party, q, price);
    } else if (a->pos < 0) { /* account.py:187-194 _net_short */
        if (side == 1) acct_size_increase(a, 0, party, q, price, tv);
        else if (-a->pos >= q) acct_size_decrease(a, 0, party, q, price, tv);
        else acct_covered_side_chg(a, 0, party, q, price);
    } else { /* account.py:173-176 _neutral */
        a->pv += tv; a->C = tv;
        if (a->d) { a->d->pv = dec_add(a->d->pv, DI(tv)); a->d->vwap = DI(price); }
        cash_increase(a, party, tv);
    }
    /* account.py:196-213 _update_net_position */
    if (side == 0) a->pos += q; else a->pos -= q;
}
/* calculate.py:35-55 mark_to_mkt */
static void orc_mtm(OrcAcct *a, int64_t p) {
    int64_t ap = a->pos < 0 ? -a->pos : a->pos;
    int64_t profit = a->pos >= 0 ? ap * p - a->C : a->C - ap * p;
    a->pv = a->C + profit;
    a->prev_nav = a->nav;
    a->nav = a->cash + a->hold + a->pv;
    if (!a->d) { if (a->nav > a->max_nav) a->max_nav = a->nav; return; }
    dec diff = a->pos >= 0 ? dec_sub(DI(p), a->d->vwap) : dec_sub(a->d->vwap, DI(p));   /* calculate.py:44-45 */
    dec dprofit = dec_mul(DI(ap), diff);
    a->d->pv = dec_add(dec_mul(DI(ap), a->d->vwap), dprofit);
    a->d->prev_nav = a->d->nav;
    a->d->nav = dec_add(dec_add(a->d->cash, a->d->hold), a->d->pv);                        /* calculate.py:12 */
    if (dec_cmp(a->d->nav, a->d->max_nav) > 0) { a->d->max_nav = a->d->nav; a->max_nav = a->nav; }
}

/* ---------------- matching: orderbook.py:61-194 -------------------------------------------- */
typedef struct { int trader; int side; /* of the incoming quote */ } OrcQuoteCtx;

static void record_fill(OrcMarket *mk, int64_t price, int64_t qty, int maker, int64_t maker_oid,
                        int64_t maker_left, int taker, int taker_side) {
    mk->tape_nonempty = 1; mk->tape_last = price; /* orderbook.py:140 tape.append */
    if (mk->n_fills < ORC_MAX_FILLS) {
        OrcFill *f = &mk->fills[mk->n_fills];
        f->time = (int32_t)mk->time; f->price = (int32_t)price; f->qty = (int32_t)qty;
        f->maker = maker; f->maker_oid = (int32_t)maker_oid; f->maker_left = (int32_t)maker_left;
        f->taker = taker; f->taker_side = taker_side;
    } else mk->status |= ORC_ST_FILL_OVERFLOW;
    mk->n_fills++;
}

/* orderbook.py:61-142 process_order_list over ONE price level (index li of tree `book`).
 * Returns quantity still to trade.  Appends fills to mk->fills (== trades list). */
static int64_t process_order_list(OrcMarket *mk, OrcTree *book, int64_t level_price, int64_t qty,
                                  int taker, int taker_side) {
    for (;;) {
        int li = tree_find_level(book, level_price);
        if (li < 0 || qty <= 0) break; /* len(order_list) > 0 and quantity_to_trade > 0 */
        int hi = book->levels[li].head;
        OrcOrder *h = &book->orders[hi];
        int64_t traded, left = -1;
        int maker = h->trader; int64_t oid = h->order_id, price = h->price;
        if (qty < h->qty) { /* :73-85 partial: resting qty reduced in place, timestamp kept */
            traded = qty;
            h->qty -= qty; left = h->qty;
            book->levels[li].volume -= qty; book->volume -= qty;
            qty = 0;
        } else if (qty == h->qty) { /* :86-92 */
            traded = qty;
            tree_remove(book, hi);
            qty = 0;
        } else { /* :93-100 */
            traded = h->qty;
            tree_remove(book, hi);
            qty -= traded;
        }
        record_fill(mk, price, traded, maker, oid, left, taker, taker_side);
    }
    return qty;
}

/* orderbook.py:144-160 process_market_order */
static void process_market(OrcMarket *mk, int side, int64_t qty, int trader) {
    OrcTree *opp = side == 0 ? &mk->asks : &mk->bids;
    while (qty > 0 && opp->n_orders > 0) {
        int64_t best = side == 0 ? opp->levels[0].price : opp->levels[opp->n_levels - 1].price;
        qty = process_order_list(mk, opp, best, qty, trader, side);
    }
}
/* orderbook.py:162-194 process_limit_order.  Returns residue qty (0 => nothing rests). */
static int64_t process_limit(OrcMarket *mk, int side, int64_t qty, int64_t price, int trader,
                             int64_t order_id, int64_t ts) {
    OrcTree *opp = side == 0 ? &mk->asks : &mk->bids;
    OrcTree *own = side == 0 ? &mk->bids : &mk->asks;
    while (opp->n_orders > 0 && qty > 0) {
        int64_t best = side == 0 ? opp->levels[0].price : opp->levels[opp->n_levels - 1].price;
        if (side == 0 ? !(price >= best) : !(price <= best)) break;
        qty = process_order_list(mk, opp, best, qty, trader, side);
    }
    if (qty > 0) tree_insert(own, price, qty, trader, order_id, ts);
    return qty;
}

/* ---------------- trader.py ---------------------------------------------------------------- */
/* trader.py:254-287 _get_order_ID.  type: 1 limit, 2 modify, 3 cancel. Returns order index or -1 */
static int get_order_id(OrcTree *t, int trader, int type, int64_t price) {
    int best = -1;
    if (type == 2) { /* oldest timestamp, first in map order on ties (python min) */
        for (int i = t->map_head; i >= 0; i = t->orders[i].map_next)
            if (t->orders[i].trader == trader && (best < 0 || t->orders[i].ts < t->orders[best].ts)) best = i;
        return best;
    }
    for (int i = t->map_head; i >= 0; i = t->orders[i].map_next)
        if (t->orders[i].trader == trader && t->orders[i].price == price) return i;
    return -1;
}

/* trader.py:108-151 _order_approved */
static int order_approved(OrcMarket *mk, OrcAcct *a, int side, int64_t size, int is_market, int64_t price) {
    if (a->d ? dec_sign(a->d->nav) <= 0 : a->nav <= 0) return 0;
    int64_t opening;
    if ((side == 0 && a->pos >= 0) || (side == 1 && a->pos <= 0)) opening = size;
    else { int64_t ap = a->pos < 0 ? -a->pos : a->pos; opening = size - ap; if (opening < 0) opening = 0; }
    if (opening <= 0) return 1;
    int64_t est;
    if (is_market) {
        OrcTree *opp = side == 0 ? &mk->asks : &mk->bids;
        if (opp->n_levels > 0) est = side == 0 ? opp->levels[0].price : opp->levels[opp->n_levels - 1].price;
        else est = mk->tape_nonempty ? mk->tape_last : 1;
    } else est = price;
    if (a->d) return dec_cmp(a->d->cash, DI(opening * est)) >= 0;   /* on the Decimal cash, residues included */
    return a->cash >= opening * est;
}

/* trader.py:303-328 _process_trades over fills [f0, n_fills) */
static void process_trades(OrcMarket *mk, int f0, int self_id) {
    int n = mk->n_fills < ORC_MAX_FILLS ? mk->n_fills : ORC_MAX_FILLS;
    for (int i = f0; i < n; ++i) {
        OrcFill *f = &mk->fills[i];
        int64_t tv = (int64_t)f->qty * f->price;
        if (f->maker != f->taker) {
            orc_process_acc(&mk->acc[f->maker], 1, 1 - f->taker_side, f->qty, f->price);
            orc_process_acc(&mk->acc[self_id], 0, f->taker_side, f->qty, f->price);
        } else { /* cash_processor.py:55-62 init_is_counter_cash_transfer */
            escrow_move(&mk->acc[self_id], tv, -1);
        }
    }
}
/* orderbook.py:210-266 modify_order (after trader.py:219-235 released the old escrow).
 * Returns residue (price, qty) through rp and rq; rq=0 => none. */
static void modify_order(OrcMarket *mk, int side, int oi, int64_t new_price, int64_t new_qty,
                         int64_t *rp, int64_t *rq) {
    OrcTree *t = side == 0 ? &mk->bids : &mk->asks;
    mk->time++;
    OrcOrder *o = &t->orders[oi];
    if (new_price == o->price && new_qty <= o->qty) { /* :245-248 in place, priority kept, ts refreshed */
        int li = tree_find_level(t, o->price);
        t->levels[li].volume -= (o->qty - new_qty);
        t->volume += new_qty - o->qty;
        o->qty = new_qty; o->ts = mk->time;
        *rp = new_price; *rq = new_qty;
        return;
    }
    int trader = o->trader; int64_t oid = o->order_id;
    tree_remove(t, oi);
    *rq = process_limit(mk, side, new_qty, new_price, trader, oid, mk->time);
    *rp = new_price;
}

/* trader.py:49-106 place_order.  type: 0 market 1 limit 2 modify 3 cancel */
static void place_order(OrcEnv *e, OrcMarket *mk, int id, int type, int side, int64_t size, int64_t price) {
    OrcAcct *a = &mk->acc[id];
    if (!order_approved(mk, a, side, size, type == 0, price)) { a->num_rejected_step++; return; }
    if (type == 0 || type == 1) a->order_step_placed = 1;
    OrcTree *own = side == 0 ? &mk->bids : &mk->asks;
    int f0 = mk->n_fills;
    int64_t rp = 0, rq = 0;
    if (size <= 0 && (type == 0 || type == 1)) { mk->status |= ORC_ST_BAD_SIZE; return; } /* sys.exit in ref */
    if (type == 0) { /* orderbook.py:33-46 */
        mk->time++; mk->next_order_id++;
        process_market(mk, side, size, id);
    } else if (type == 1) { /* trader.py:189-203 */
        int oi = get_order_id(own, id, 1, price);
        if (oi < 0) {
            mk->time++; mk->next_order_id++;
            rq = process_limit(mk, side, size, price, id, mk->next_order_id, mk->time); rp = price;
        } else {
            OrcOrder *o = &own->orders[oi];
            int64_t ov = o->price * o->qty; escrow_move(a, ov, -1); /* cancel_cash_transfer */
            modify_order(mk, side, oi, price, size, &rp, &rq);
        }
    } else if (type == 2) { /* trader.py:205-217 */
        int oi = get_order_id(own, id, 2, price);
        if (oi >= 0) {
            OrcOrder *o = &own->orders[oi];
            int64_t ov = o->price * o->qty; escrow_move(a, ov, -1);
            modify_order(mk, side, oi, price, size, &rp, &rq);
        }
    } else { /* trader.py:237-252 */
        int oi = get_order_id(own, id, 3, price);
        if (oi >= 0) {
            OrcOrder *o = &own->orders[oi];
            int64_t ov = o->price * o->qty;
            mk->time++; /* orderbook.py:196-208 */
            tree_remove(own, oi);
            escrow_move(a, ov, -1);
        }
    }
    if (mk->n_fills > f0) process_trades(mk, f0, id);
    if (rq > 0) escrow_move(a, rp * rq, +1); /* cash_processor.py:15-29 */
    (void)e;
}

/* ---------------- state_helper.py:113-214 set_agg_LOB -------------------------------------- */
static void set_agg_lob(const OrcEnv *e, OrcMarket *mk, float *snap /*42*/) {
    double bp[ORC_K] = {0}, bs[ORC_K] = {0}, ap[ORC_K] = {0}, as[ORC_K] = {0};
    for (int k = 0; k < ORC_K && k < mk->bids.n_levels; ++k) {
        const OrcLevel *L = &mk->bids.levels[mk->bids.n_levels - 1 - k];
        bp[k] = (double)L->price; bs[k] = (double)L->volume;
    }
    for (int k = 0; k < ORC_K && k < mk->asks.n_levels; ++k) {
        const OrcLevel *L = &mk->asks.levels[k];
        ap[k] = -(double)L->price; as[k] = -(double)L->volume;
    }
    for (int k = 0; k < ORC_K; ++k) {
        mk->raw[k] = (float)bp[k]; mk->raw[ORC_K + k] = (float)bs[k];
        mk->raw[2 * ORC_K + k] = (float)ap[k]; mk->raw[3 * ORC_K + k] = (float)as[k];
    }
    double l1_bid = bp[0] > 0 ? bp[0] : 0.0;
    double l1_ask = ap[0] != 0 ? fabs(ap[0]) : 0.0;
    double M;
    if (l1_bid > 0 && l1_ask > 0) M = (l1_bid + l1_ask) / 2.0;
    else if (l1_bid > 0) M = l1_bid;
    else if (l1_ask > 0) M = l1_ask;
    else { M = (double)mk->last_price; if (M <= 0) M = 100.0; }
    for (int k = 0; k < ORC_K; ++k) {
        snap[k] = (float)(bp[k] > 0 ? (M - bp[k]) / M : 0.0);
        snap[ORC_K + k] = (float)(bs[k] > 0 ? sqrt(bs[k]) : 0.0);
        snap[2 * ORC_K + k] = (float)(ap[k] != 0 ? -((fabs(ap[k]) - M) / M) : 0.0);
        snap[3 * ORC_K + k] = (float)(as[k] != 0 ? -sqrt(fabs(as[k])) : 0.0);
    }
    snap[40] = (float)log(M);
    if (l1_bid > 0 && l1_ask > 0) {
        double st = (l1_ask - l1_bid) / (double)e->cfg.tick;
        snap[41] = (float)log1p(st > 0.0 ? st : 0.0);
    } else snap[41] = 0.0f;
}

/* ---------------- env: continuousDoubleAuction_env.py:175-231 reset ------------------------ */
static void market_reset(OrcEnv *e, OrcMarket *mk, int reseed, uint64_t seed) {
    const OrcConfig *c = &e->cfg;
    if (reseed) orc_rng_seed(&mk->rng, seed);
    tree_clear(&mk->bids); tree_clear(&mk->asks);
    mk->time = 0; mk->next_order_id = 0; mk->tape_nonempty = 0; mk->tape_last = 0;
    mk->t_step = 0; mk->done_mask = 0; mk->n_fills = 0; mk->status = 0;
    mk->best_bid = mk->best_ask = 0;
    mk->last_price = orc_integers(&mk->rng, c->price_lo, (int64_t)c->price_hi + 1);
    for (int i = 0; i < c->num_agents; ++i) { /* account.py:55-82 */
        OrcAcct *a = &mk->acc[i];
        memset(a, 0, sizeof(*a));
        a->cash = a->nav = a->prev_nav = a->max_nav = c->init_cash;
        if (e->dec) {
            a->d = &e->dec[(size_t)(mk - e->mk) * ORC_MAX_AGENTS + i];
            memset(a->d, 0, sizeof(*a->d));
            a->d->cash = a->d->nav = a->d->prev_nav = a->d->max_nav = DI(c->init_cash);
        }
    }
    float snap[ORC_SNAP];
    set_agg_lob(e, mk, snap); /* state_helper.py:66-78 */
    for (int h = 0; h < c->n_hist; ++h) memcpy(mk->hist[h], snap, sizeof(snap));
}

/* action decode: action_helper.py:241-283, :311-339, :341-397 */
typedef struct { int id, side, type; int64_t size, price; } OrcAct;

static void market_step(OrcEnv *e, OrcMarket *mk, const int32_t *cat, const float *mean,
                        const float *sigma, const int32_t *pcode, const int32_t *poff,
                        float *obs, double *reward, uint8_t *term, uint8_t *trunc) {
    const OrcConfig *c = &e->cfg;
    const int A = c->num_agents;
    float snap[ORC_SNAP];
    mk->n_fills = 0;
    for (int i = 0; i < A; ++i) { /* exchg_helper.py:116-120 zeroes these at the end of the previous step */
        OrcAcct *a = &mk->acc[i];
        a->num_trades_step = a->num_passive_fills_step = a->order_step_placed = a->num_rejected_step = 0;
    }
    set_agg_lob(e, mk, snap); /* continuousDoubleAuction_env.py:274 (refreshes mk->raw) */

    const double mkt_mul = (c->mkt_max_size - c->min_size) / 2.0;                       /* action_helper.py:46 */
    const double lim_mul = ((double)c->mkt_max_size * c->limit_size_multiple - c->min_size) / 2.0; /* :47 */
    OrcAct acts[ORC_MAX_AGENTS]; int n = 0;
    for (int i = 0; i < A; ++i) { /* set_actions :145-172, dict order == agent order */
        mk->acc[i].is_pass = 0;
        mk->last_act[i][0] = -1; mk->last_act[i][1] = -1; mk->last_act[i][2] = 0; mk->last_act[i][3] = -1;
        int cg = cat[i];
        if (cg < 0) continue;           /* agent absent from the action dict: no RNG draw */
        if (cg > 8) { mk->status |= ORC_ST_BAD_ACTION; cg = 0; }
        int side = cg == 0 ? -1 : (cg <= 4 ? 0 : 1);
        int type = cg == 0 ? 0 : (cg - 1) & 3;
        /* _set_size :311-339: loc = mul * mean computed in float32 (python float x f32 array) */
        float loc = (float)(type == 0 ? mkt_mul : lim_mul) * mean[i];
        double z = orc_standard_normal(&mk->rng);
        volatile double sz = (double)sigma[i] * z; /* numpy: loc + scale*z, two roundings */
        double x = (double)loc + sz;
        int64_t size = (int64_t)rint(fabs(x)) + c->min_size; /* :339, :276 */
        if (side < 0) { mk->acc[i].is_pass = 1; continue; }
        int64_t price = -1;
        if (type != 0) { /* _set_price :341-397 reads the frozen pre-step raw top-K */
            int lvl = pcode[i]; int off = poff[i] - 1;
            if (lvl < 0 || lvl >= ORC_K || poff[i] < 0 || poff[i] > 2) { mk->status |= ORC_ST_BAD_ACTION; lvl = 0; off = 0; }
            float p = side == 0 ? mk->raw[lvl] : fabsf(mk->raw[2 * ORC_K + lvl]);
            int64_t base;
            if (side == 0) { base = p == 0 ? mk->last_price - (int64_t)(lvl + 1) * c->tick : (int64_t)fabsf(p); price = base + (int64_t)off * c->tick; }
            else           { base = p == 0 ? mk->last_price + (int64_t)(lvl + 1) * c->tick : (int64_t)p;        price = base - (int64_t)off * c->tick; }
            if (price < c->tick) price = c->tick;
        }
        acts[n].id = i; acts[n].side = side; acts[n].type = type; acts[n].size = size; acts[n].price = price;
        mk->last_act[i][0] = type; mk->last_act[i][1] = side; mk->last_act[i][2] = (int32_t)size; mk->last_act[i][3] = (int32_t)price;
        ++n;
    }
    int perm[ORC_MAX_AGENTS];
    orc_permutation(&mk->rng, n, perm); /* rand_exec_seq :174-199 */
    for (int k = 0; k < n; ++k) { /* do_actions :201-239 */
        const OrcAct *a = &acts[perm[k]];
        place_order(e, mk, a->id, a->type, a->side, a->size, a->price);
    }
    if (mk->tape_nonempty) { /* exchg_helper.py:56-66 */
        mk->last_price = mk->tape_last;
        for (int i = 0; i < A; ++i) orc_mtm(&mk->acc[i], mk->tape_last);
    }
    /* state_helper.py:80-92 prep_next_state */
    set_agg_lob(e, mk, snap);
    for (int h = 0; h + 1 < c->n_hist; ++h) memcpy(mk->hist[h], mk->hist[h + 1], sizeof(snap));
    memcpy(mk->hist[c->n_hist - 1], snap, sizeof(snap));
    for (int h = 0; h < c->n_hist; ++h) memcpy(obs + h * ORC_SNAP, mk->hist[h], sizeof(snap));
    /* exchg_helper.py:68-91 */
    mk->best_bid = mk->bids.n_levels ? mk->bids.levels[mk->bids.n_levels - 1].price : 0;
    mk->best_ask = mk->asks.n_levels ? mk->asks.levels[0].price : 0;
    for (int i = 0; i < A; ++i) { /* reward_helper.py:35-103 ; done_helper.py:3-18 */
        OrcAcct *a = &mk->acc[i];
        double nav_change = (double)(a->nav - a->prev_nav);
        int64_t ddi = a->max_nav - a->nav; if (ddi < 0) ddi = 0;
        double dd = (double)ddi;
        if (a->d) {   /* float(nav - prev_nav), float(max(0, max_nav - nav)) on the Decimal fields */
            nav_change = dec_to_double(dec_sub(a->d->nav, a->d->prev_nav));
            dec ddd = dec_sub(a->d->max_nav, a->d->nav);
            dd = dec_sign(ddd) > 0 ? dec_to_double(ddd) : 0.0;
        }
        double nav_term = nav_change * (nav_change < 0 ? c->loss_multiplier : 1.0);
        volatile double t0 = nav_term;
        volatile double t1 = -(c->order_penalty * (double)a->order_step_placed);
        volatile double t2 = -(c->trade_penalty * (double)a->num_trades_step);
        volatile double t3 = -(c->drawdown_penalty * dd);
        volatile double t4 = c->passive_bonus * (double)a->num_passive_fills_step;
        volatile double r = 0.0;
        r = r + t0; r = r + t1; r = r + t2; r = r + t3; r = r + t4;
        a->reward = r; a->drawdown = dd;
        a->terms[0] = t0; a->terms[1] = t1; a->terms[2] = t2; a->terms[3] = t3; a->terms[4] = t4;
        reward[i] = r;
        if (a->d ? dec_sign(a->d->nav) <= 0 : a->nav <= 0) mk->done_mask |= (1u << i);
    }
    uint32_t all = A >= 32 ? 0xffffffffu : ((1u << A) - 1u);
    *term = (mk->done_mask & all) == all;            /* done_helper.py:20-55 */
    *trunc = (mk->t_step + 1 >= c->max_step);
    mk->t_step++;
}

/* ======================================= C API ============================================ */
void *orc_create(const OrcConfig *cfg, int M) {
    if (cfg->num_agents < 1 || cfg->num_agents > ORC_MAX_AGENTS || cfg->n_hist < 1 || cfg->n_hist > ORC_MAX_HIST || cfg->tick < 1) return NULL;
    OrcEnv *e = (OrcEnv *)calloc(1, sizeof(OrcEnv));
    e->cfg = *cfg; e->M = M;
    e->mk = (OrcMarket *)calloc((size_t)M, sizeof(OrcMarket));
    if (cfg->decimal_ledger) e->dec = (OrcDec *)calloc((size_t)M * ORC_MAX_AGENTS, sizeof(OrcDec));
    for (int m = 0; m < M; ++m) { tree_init(&e->mk[m].bids); tree_init(&e->mk[m].asks); }
    return e;
}
void orc_destroy(void *h) {
    OrcEnv *e = (OrcEnv *)h;
    for (int m = 0; m < e->M; ++m) { tree_free(&e->mk[m].bids); tree_free(&e->mk[m].asks); }
    free(e->mk); free(e->dec); free(e);
}
/* seeds: NULL => keep every stream (reset(seed=None)); mask: NULL => all markets. */
void orc_reset(void *h, const uint64_t *seeds, const uint8_t *mask, float *obs) {
    OrcEnv *e = (OrcEnv *)h;
    const int W = e->cfg.n_hist * ORC_SNAP;
    for (int m = 0; m < e->M; ++m) {
        if (mask && !mask[m]) continue;
        market_reset(e, &e->mk[m], seeds != NULL, seeds ? seeds[m] : 0);
        if (obs) for (int k = 0; k < e->cfg.n_hist; ++k) memcpy(obs + (size_t)m * W + k * ORC_SNAP, e->mk[m].hist[k], sizeof(float) * ORC_SNAP);
    }
}
/* Arrays are [M][A] row-major; obs [M][n_hist*42]; reward [M][A]; flags [M].
 * orc_rollout runs T consecutive steps (actions [T][M][A]); markets are independent, so each
 * thread owns a contiguous slice of markets and runs all T steps on it without barriers — the
 * reference's best case (one env per process, N processes).  Outputs hold the LAST step. */
/* Uniform random policy of the fused device rollout (cda_rollout_random): the shape of the reference's RandomRLModule
 * (train/model/model_handler.py:38-78: category U{0..8}, price U{0..9}, price_offset U{0..2}, size_mean U(-1,1), size_sigma U(0,1)) drawn
 * from a counter-based generator keyed by (policy_seed, market, t_step, agent) — NOT from the env stream, which stays numpy-exact.
 * Restated here so that the device rollout has a CPU twin: same actions, then the ordinary market_step. */
static uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static void orc_policy_actions(uint64_t seed, int m, uint32_t t_step, int A, int32_t *cat, float *mean, float *sigma, int32_t *pcode, int32_t *poff) {
    for (int a = 0; a < A; ++a) {
        const uint64_t h = orc_splitmix64(seed ^ orc_splitmix64(((uint64_t)m << 32) ^ ((uint64_t)t_step * 64ULL + (uint64_t)a)));
        cat[a] = (int32_t)(((h & 0xffffu) * 9u) >> 16);
        pcode[a] = (int32_t)((((h >> 16) & 0xffffu) * 10u) >> 16);
        poff[a] = (int32_t)((((h >> 32) & 0xffffu) * 3u) >> 16);
        const uint64_t h2 = orc_splitmix64(h);
        mean[a] = (float)((double)(h2 & 0xffffffu) * (2.0 / 16777216.0) - 1.0);
        sigma[a] = (float)((double)((h2 >> 24) & 0xffffffu) * (1.0 / 16777216.0));
    }
}
typedef struct {
    OrcEnv *e; int m0, m1, T;
    int policy; uint64_t policy_seed;   /* policy != 0: actions come from orc_policy_actions instead of the arrays */
    const int32_t *cat; const float *mean; const float *sigma; const int32_t *pcode; const int32_t *poff;
    float *obs; double *reward; uint8_t *term; uint8_t *trunc;
} OrcJob;
static void *orc_job_run(void *arg) {
    OrcJob *j = (OrcJob *)arg; OrcEnv *e = j->e;
    const int A = e->cfg.num_agents, W = e->cfg.n_hist * ORC_SNAP;
    const size_t MA = (size_t)e->M * A;
    /* market-outer, time-inner: a market's state stays in the core's cache for all T steps (markets are independent, so the order
     * of (t, m) pairs does not matter; this is the CPU path's best case and keeps it from slowing down as M grows) */
    for (int m = j->m0; m < j->m1; ++m)
        for (int t = 0; t < j->T; ++t) {
            size_t o = (size_t)t * MA + (size_t)m * A, r = (size_t)m * A;
            if (j->policy) {
                int32_t pc[ORC_MAX_AGENTS], pp[ORC_MAX_AGENTS], po[ORC_MAX_AGENTS]; float pm[ORC_MAX_AGENTS], ps[ORC_MAX_AGENTS];
                orc_policy_actions(j->policy_seed, m, (uint32_t)e->mk[m].t_step, A, pc, pm, ps, pp, po);
                market_step(e, &e->mk[m], pc, pm, ps, pp, po, j->obs + (size_t)m * W, j->reward + r, j->term + m, j->trunc + m);
                continue;
            }
            market_step(e, &e->mk[m], j->cat + o, j->mean + o, j->sigma + o, j->pcode + o, j->poff + o,
                        j->obs + (size_t)m * W, j->reward + r, j->term + m, j->trunc + m);
        }
    return NULL;
}
static void orc_rollout_impl(void *h, int T, int policy, uint64_t policy_seed, const int32_t *cat, const float *mean, const float *sigma, const int32_t *pcode,
                 const int32_t *poff, float *obs, double *reward, uint8_t *term, uint8_t *trunc, int nthreads) {
    OrcEnv *e = (OrcEnv *)h;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > e->M) nthreads = e->M;
    if (nthreads > 256) nthreads = 256;
    OrcJob jobs[256]; pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        OrcJob *j = &jobs[k];
        j->e = e; j->T = T; j->policy = policy; j->policy_seed = policy_seed;
        j->m0 = (int)((int64_t)e->M * k / nthreads); j->m1 = (int)((int64_t)e->M * (k + 1) / nthreads);
        j->cat = cat; j->mean = mean; j->sigma = sigma; j->pcode = pcode; j->poff = poff;
        j->obs = obs; j->reward = reward; j->term = term; j->trunc = trunc;
    }
    if (nthreads == 1) { orc_job_run(&jobs[0]); return; }
    for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, orc_job_run, &jobs[k]);
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
}
void orc_rollout(void *h, int T, const int32_t *cat, const float *mean, const float *sigma, const int32_t *pcode,
                 const int32_t *poff, float *obs, double *reward, uint8_t *term, uint8_t *trunc, int nthreads) {
    orc_rollout_impl(h, T, 0, 0, cat, mean, sigma, pcode, poff, obs, reward, term, trunc, nthreads);
}
/* T steps of every market under the uniform random policy (CPU twin of cda_rollout_random); outputs hold the LAST step */
void orc_rollout_random(void *h, int T, uint64_t policy_seed, float *obs, double *reward, uint8_t *term, uint8_t *trunc, int nthreads) {
    orc_rollout_impl(h, T, 1, policy_seed, NULL, NULL, NULL, NULL, NULL, obs, reward, term, trunc, nthreads);
}
void orc_step(void *h, const int32_t *cat, const float *mean, const float *sigma, const int32_t *pcode,
              const int32_t *poff, float *obs, double *reward, uint8_t *term, uint8_t *trunc, int nthreads) {
    orc_rollout(h, 1, cat, mean, sigma, pcode, poff, obs, reward, term, trunc, nthreads);
}

/* ---- state dumps for bit-exact comparison ---- */
/* Book side in PRIORITY order (bids: price desc, asks: price asc; FIFO inside a level).
 * out rows: price, qty, trader, order_id, ts.  Returns number of orders. */
int orc_dump_book(void *h, int m, int side, int64_t *out, int max_rows) {
    OrcEnv *e = (OrcEnv *)h; OrcTree *t = side == 0 ? &e->mk[m].bids : &e->mk[m].asks;
    int n = 0;
    for (int l = 0; l < t->n_levels; ++l) {
        const OrcLevel *L = &t->levels[side == 0 ? t->n_levels - 1 - l : l];
        for (int i = L->head; i >= 0; i = t->orders[i].next) {
            if (n < max_rows) {
                const OrcOrder *o = &t->orders[i];
                out[n * 5 + 0] = o->price; out[n * 5 + 1] = o->qty; out[n * 5 + 2] = o->trader;
                out[n * 5 + 3] = o->order_id; out[n * 5 + 4] = o->ts;
            }
            ++n;
        }
    }
    return n;
}
/* decoded actions of the last step (continuousDoubleAuction_env.py:285 LOB_actions): out[A][4] = type, side, size, price; side -1 = pass / absent */
void orc_dump_actions(void *h, int m, int32_t *out) {
    OrcEnv *e = (OrcEnv *)h;
    memcpy(out, e->mk[m].last_act, sizeof(int32_t) * 4 * (size_t)e->cfg.num_agents);
}
/* order ids in order_map iteration order */
int orc_dump_map(void *h, int m, int side, int64_t *out, int max_rows) {
    OrcEnv *e = (OrcEnv *)h; OrcTree *t = side == 0 ? &e->mk[m].bids : &e->mk[m].asks;
    int n = 0;
    for (int i = t->map_head; i >= 0; i = t->orders[i].map_next) { if (n < max_rows) out[n] = t->orders[i].order_id; ++n; }
    return n;
}
/* scalars: time, next_order_id, last_price, tape_nonempty, t_step, done_mask, status, n_fills, best_bid, best_ask */
void orc_dump_scalars(void *h, int m, int64_t *out) {
    OrcEnv *e = (OrcEnv *)h; OrcMarket *k = &e->mk[m];
    out[0] = k->time; out[1] = k->next_order_id; out[2] = k->last_price; out[3] = k->tape_nonempty;
    out[4] = k->t_step; out[5] = k->done_mask; out[6] = k->status; out[7] = k->n_fills;
    out[8] = k->best_bid; out[9] = k->best_ask;
}
/* accounts rows: cash, hold, pv, C, nav, prev_nav, max_nav, pos, num_trades, trades_step, passive_step, placed, rejected, is_pass */
void orc_dump_accounts(void *h, int m, int64_t *out) {
    OrcEnv *e = (OrcEnv *)h; OrcMarket *k = &e->mk[m];
    for (int i = 0; i < e->cfg.num_agents; ++i) {
        const OrcAcct *a = &k->acc[i]; int64_t *o = out + i * 14;
        o[0] = a->cash; o[1] = a->hold; o[2] = a->pv; o[3] = a->C; o[4] = a->nav; o[5] = a->prev_nav; o[6] = a->max_nav;
        if (a->d) {   /* the Decimal twins, to the nearest integer (what ref_runner does with the reference's fields); they must agree */
            int64_t ap = a->pos < 0 ? -a->pos : a->pos;
            int64_t dv[7] = {dec_to_i64_nearest(a->d->cash), dec_to_i64_nearest(a->d->hold), dec_to_i64_nearest(a->d->pv),
                             dec_to_i64_nearest(dec_mul(DI(ap), a->d->vwap)), dec_to_i64_nearest(a->d->nav),
                             dec_to_i64_nearest(a->d->prev_nav), dec_to_i64_nearest(a->d->max_nav)};
            for (int j = 0; j < 7; ++j) { if (dv[j] != o[j]) k->status |= ORC_ST_LEDGER_MISMATCH; o[j] = dv[j]; }
            if (dec_range_errors) k->status |= ORC_ST_LEDGER_MISMATCH;   /* dec128.h: an operation did not fit 128 bits */
        }
        o[7] = a->pos; o[8] = a->num_trades; o[9] = a->num_trades_step; o[10] = a->num_passive_fills_step;
        o[11] = a->order_step_placed; o[12] = a->num_rejected_step; o[13] = a->is_pass;
    }
}
void orc_dump_reward_terms(void *h, int m, double *out /*[A][6]: 5 terms + drawdown*/) {
    OrcEnv *e = (OrcEnv *)h; OrcMarket *k = &e->mk[m];
    for (int i = 0; i < e->cfg.num_agents; ++i) { for (int j = 0; j < 5; ++j) out[i * 6 + j] = k->acc[i].terms[j]; out[i * 6 + 5] = k->acc[i].drawdown; }
}
/* fills of the last step; rows of 8 int32 (see OrcFill). Returns count. */
int orc_dump_fills(void *h, int m, int32_t *out, int max_rows) {
    OrcEnv *e = (OrcEnv *)h; OrcMarket *k = &e->mk[m];
    int n = k->n_fills < ORC_MAX_FILLS ? k->n_fills : ORC_MAX_FILLS;
    for (int i = 0; i < n && i < max_rows; ++i) memcpy(out + i * 8, &k->fills[i], sizeof(OrcFill));
    return k->n_fills;
}
/* rng state: state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger */
void orc_dump_rng(void *h, int m, uint64_t *out) {
    OrcEnv *e = (OrcEnv *)h; orc_rng *r = &e->mk[m].rng;
    out[0] = (uint64_t)(r->state >> 64); out[1] = (uint64_t)r->state; out[2] = (uint64_t)(r->inc >> 64); out[3] = (uint64_t)r->inc;
    out[4] = r->has_uint32; out[5] = r->uinteger;
}
/* RNG self-test hooks (tests/test_np_rng.py) */
void orc_test_seed(uint64_t seed, uint64_t *out) {
    orc_rng r; orc_rng_seed(&r, seed);
    out[0] = (uint64_t)(r.state >> 64); out[1] = (uint64_t)r.state; out[2] = (uint64_t)(r.inc >> 64); out[3] = (uint64_t)r.inc;
}
/* ops[i]: 0 normal, -1 integers(lo,hi), k>=2 permutation(k) (written to outp + 32*i) */
void orc_test_stream(uint64_t seed, int n, const int32_t *ops, double *outn, int32_t *outp, int lo, int hi) {
    orc_rng r; orc_rng_seed(&r, seed);
    for (int i = 0; i < n; ++i) {
        if (ops[i] == 0) outn[i] = orc_standard_normal(&r);
        else if (ops[i] == -1) outn[i] = (double)orc_integers(&r, lo, hi);
        else { int p[ORC_MAX_AGENTS]; orc_permutation(&r, ops[i], p); for (int k = 0; k < ops[i]; ++k) outp[32 * i + k] = p[k]; }
    }
}

/* dec28.h test entry: op is one of  + - * / (divide) c (compare: result "-1", "0" or "1"); operands and result are decimal strings */
void orc_dec_op(char op, const char *a, const char *b, char *out, int cap) {
    dec x = dec_from_str(a), y = dec_from_str(b), r;
    if (op == 'c') { snprintf(out, (size_t)cap, "%d", dec_cmp(x, y)); return; }
    r = op == '+' ? dec_add(x, y) : op == '-' ? dec_sub(x, y) : op == '*' ? dec_mul(x, y) : dec_div(x, y);
    dec_to_str(r, out, (size_t)cap);
}
double orc_dec_to_double(const char *a) { return dec_to_double(dec_from_str(a)); }
int orc_dec_range_errors(void) { return dec_range_errors; }
void orc_dec_reset_range_errors(void) { dec_range_errors = 0; }
/* decimal_ledger mode: the seven Decimal fields of every agent of market m as strings, 48 bytes each:
 * cash, cash_on_hold, position_val, VWAP, nav, prev_nav, max_nav (account.py:12-53) */
void orc_dump_accounts_dec(void *h, int m, char *out /*[A][7][48]*/) {
    OrcEnv *e = (OrcEnv *)h; OrcMarket *k = &e->mk[m];
    for (int i = 0; i < e->cfg.num_agents; ++i) {
        const OrcAcct *a = &k->acc[i];
        const dec f[7] = {a->d->cash, a->d->hold, a->d->pv, a->d->vwap, a->d->nav, a->d->prev_nav, a->d->max_nav};
        for (int j = 0; j < 7; ++j) dec_to_str(f[j], out + ((size_t)i * 7 + j) * 48, 48);
    }
}
