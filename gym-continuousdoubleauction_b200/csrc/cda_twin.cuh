// cda_twin.cuh — the reference's Decimal(28) ledger residues, kept EXACTLY but off the hot path ("deferred Decimal twin").
//
// Why: the reference holds money as decimal.Decimal (prec 28).  VWAP = cost / |position| (account.py:124-161) rarely terminates, so
// |position| * VWAP differs from the exact cost basis by ~1e-24, the difference leaks into `cash` when a position is covered
// (account.py:135-149, cash_processor.py:47-53), and the gate `cash >= opening * est_price` (trader.py:108-151) — or the bankruptcy
// test `nav <= 0` — at EXACT integer equality is then decided by the sign of that residue.  The step kernel runs the exact int64
// ledger (cost basis recurrence, cda_kernels.cuh acct_fill); the integers decide everything except those ties.
//
// What decides a tie is path dependent (every Decimal operation rounds at 28 significant digits), so the residues must be carried
// forward operation by operation.  Doing that inside the step would put ~700 instructions of 128-bit arithmetic per fill on the
// critical path of a warp in which two lanes are active.  Instead:
//   * per agent, the twin state is { Decimal VWAP, Decimal cash (only while it carries a residue: "tracked"), position } as of the
//     last flush, plus a JOURNAL of the ledger events since then (8 bytes each: FILL {party, side, qty, price}, and — only while cash is
//     tracked — ESCROW {signed value} / CASHSYNC {integer cash});
//   * the step kernel only APPENDS events (a few instructions on the two lanes of a fill; no call inside the matching loop);
//   * cda_twin_flush_kernel replays the journals with one THREAD per (market, agent) — 32 active lanes per warp instead of two — every
//     CDA_TWIN_FLUSH_STEPS steps, on the same stream, between steps;
//   * when a tie does occur (about one agent-step in 10^5..10^6 on low-cash configurations, never seen at the default cash), the
//     step leaves its action loop, the trader's lane replays its own journal (cold, out of line) and the action is retried with the
//     answer.
// position_val and nav are pure functions of (cash, hold, |position|, VWAP, last price) — the reference recomputes position_val from
// VWAP at every fill and every mark-to-market (calculate.py:24-55) and never accumulates it — so they are derived on demand.
// The replay mirrors oracle/cda_oracle.c's decimal_ledger mode operation for operation (same order of additions: rounding depends
// on it); that mode reproduces the reference's Decimal fields exactly (tests/test_oracle_vs_reference.py).
#pragma once
#include "cda_dec128.cuh"

#define CDA_JRN_E 48                 /* journal entries per agent (8 bytes each) */
#define CDA_JRN_RESERVE 16           /* entries kept free for ONE step (an agent in more fills than this in one step is flagged) */
#define CDA_TWIN_BYTES 64            /* twin state per agent */
#define CDA_TWIN_FLUSH_STEPS 10      /* host-side replay cadence (steps): ~7 events per agent at the BASELINE fill rates, 32 fit */
// twin-flags word per agent (u32 in the accounts block, persistent): journal length, cash tracked, sign of a zero-integer NAV
#define CDA_TWF_JN_MASK 0xffu
#define CDA_TWF_TRACKED 0x100u
#define CDA_TWF_NAVSIGN_SHIFT 9      /* 2 bits: 0 unknown / integer NAV != 0, 1 nav > 0, 2 nav == 0 exactly, 3 nav < 0 */
#define CDA_TWF_RANGE 0x1000u        /* an operation left the 128-bit domain (reported as CDA_ST_DEC_RANGE) */

#define CDA_EV_FILL 0ULL
#define CDA_EV_ESCROW 1ULL
#define CDA_EV_CASHSYNC 2ULL
CDA_HD unsigned long long cda_ev_fill(int party, int side, unsigned q, unsigned p) {
    return (CDA_EV_FILL << 62) | ((unsigned long long)(p & 0xffffffu) << 34) | ((unsigned long long)q << 2) | ((unsigned long long)(side & 1) << 1) | (unsigned long long)(party & 1);
}
CDA_HD unsigned long long cda_ev_value(unsigned long long tag, long long v) { return (tag << 62) | ((unsigned long long)v & 0x3fffffffffffffffULL); }
CDA_HD long long cda_ev_signed(unsigned long long ev) { return (long long)(ev << 2) >> 2; }

struct CdaTwinStored {               // 64 bytes per agent in the market block
    unsigned long long vwap_lo, vwap_hi; int vwap_exp, vwap_sign;
    unsigned long long cash_lo, cash_hi; int cash_exp, cash_sign;
    int pos; unsigned flags;         // flags bit 0: cash tracked
    unsigned long long pad;
};
struct CdaTwin { CdaDec vwap, cash; long long pos; int tracked; int err; };

CDA_HD CdaDec cda_dec_make(unsigned long long lo, unsigned long long hi, int exp, int sign) { CdaDec d; d.c = ((cda_u128)hi << 64) | lo; d.exp = exp; d.sign = sign; return d; }
CDA_HD void cda_twin_load(CdaTwin &t, const CdaTwinStored *s) {
    t.vwap = cda_dec_make(s->vwap_lo, s->vwap_hi, s->vwap_exp, s->vwap_sign);
    t.cash = cda_dec_make(s->cash_lo, s->cash_hi, s->cash_exp, s->cash_sign);
    t.pos = s->pos; t.tracked = (int)(s->flags & 1u); t.err = 0;
}
CDA_HD void cda_twin_store(const CdaTwin &t, CdaTwinStored *s) {
    s->vwap_lo = (unsigned long long)t.vwap.c; s->vwap_hi = (unsigned long long)(t.vwap.c >> 64); s->vwap_exp = t.vwap.exp; s->vwap_sign = t.vwap.sign;
    s->cash_lo = (unsigned long long)t.cash.c; s->cash_hi = (unsigned long long)(t.cash.c >> 64); s->cash_exp = t.cash.exp; s->cash_sign = t.cash.sign;
    s->pos = (int)t.pos; s->flags = (unsigned)t.tracked;
}
#define CDA_DI(x) cda_dec_from_i64((long long)(x))
CDA_HD void cda_twin_cash_add(CdaTwin &t, long long v) { if (t.tracked) t.cash = cda_dec_add(t.cash, CDA_DI(v)); }
// account.py:135-149 _covered: position_val = raw + profit; cash += position_val - mkt_val (cash_processor.py:47-53); VWAP = 0
CDA_HDN long long cda_twin_covered(CdaTwin &t, int is_long, long long ap, long long price) {
    const long long mkt = ap * price;
    if (t.tracked) {                                           // (always, in the step kernel's protocol: CASHSYNC precedes a covering FILL)
        const CdaDec draw = cda_dec_mul(CDA_DI(ap), t.vwap, &t.err), dmkt = CDA_DI(mkt);
        const CdaDec profit = is_long ? cda_dec_sub(dmkt, draw) : cda_dec_sub(draw, dmkt);     // calculate.py:24-33
        const CdaDec pv = cda_dec_add(draw, profit);
        t.cash = cda_dec_add(t.cash, cda_dec_sub(pv, dmkt));
    }
    t.vwap = cda_dec_zero();
    return mkt;
}
// account.py:215-231 process_acc on the Decimal fields; party 0 = init_party, 1 = counter_party; side 0 bid, 1 ask.
// Written so that the expensive case — a VWAP update, increase or partial decrease alike — is ONE instruction stream (the replay kernel
// runs 32 agents per warp: lanes in different branches serialise).
CDA_HDN void cda_twin_fill(CdaTwin &t, int party, int side, long long q, long long price) {
    const long long tv = q * price, pos = t.pos, ap = pos < 0 ? -pos : pos;
    const int is_long = pos > 0, same = (side == 0) == is_long;
    t.pos = pos + (side == 0 ? q : -q);
    if (pos != 0 && (same || ap > q)) {
        // :124-133 _size_increase  VWAP' = (|pos| * VWAP + tv) / (|pos| + q);  :151-161 _size_decrease (left > 0)  VWAP' = (|pos| * VWAP - tv) / (|pos| - q)
        const CdaDec raw = cda_dec_mul(CDA_DI(ap), t.vwap, &t.err);
        t.vwap = cda_dec_div(cda_dec_add(raw, CDA_DI(same ? tv : -tv)), CDA_DI(same ? ap + q : ap - q), &t.err);
        if (t.tracked) {                                       // cash_processor.py:31-45
            if (same) { if (party == 0) cda_twin_cash_add(t, -tv); }
            else { cda_twin_cash_add(t, tv); if (party == 1) cda_twin_cash_add(t, tv); }   // passive: cash += v, hold -= v, cash += v
        }
        return;
    }
    if (pos == 0) {                                            // :173-176 _neutral
        t.vwap = CDA_DI(price);
        if (party == 0) cda_twin_cash_add(t, -tv);             // (the passive side pays out of cash_on_hold)
        return;
    }
    // the position is covered: flat (:151-161 with left == 0) or flipped (:163-171 _covered_side_chg)
    const long long mkt = cda_twin_covered(t, is_long, ap, price);
    if (ap == q) {
        cda_twin_cash_add(t, tv);
        if (party == 1) cda_twin_cash_add(t, tv);
    } else {
        cda_twin_cash_add(t, mkt);
        if (party == 1) cda_twin_cash_add(t, mkt);
        t.vwap = CDA_DI(price);
        if (party == 0) cda_twin_cash_add(t, -(q - ap) * price);
    }
}
CDA_HD void cda_twin_apply(CdaTwin &t, unsigned long long ev) {
    const unsigned long long tag = ev >> 62;
    if (tag == CDA_EV_FILL) cda_twin_fill(t, (int)(ev & 1ULL), (int)((ev >> 1) & 1ULL), (long long)((ev >> 2) & 0xffffffffULL), (long long)((ev >> 34) & 0xffffffULL));
    else if (tag == CDA_EV_ESCROW) cda_twin_cash_add(t, cda_ev_signed(ev));
    else { t.tracked = 1; t.cash = CDA_DI(cda_ev_signed(ev)); }
}
// after a replay: cash that has become an integer again needs no tracking (integer additions are exact and commute)
CDA_HD void cda_twin_settle(CdaTwin &t) {
    if (!t.tracked) return;
    if (t.cash.c == 0 || t.cash.exp >= 0) { t.tracked = 0; return; }
    const CdaDec s = cda_dec_strip(t.cash);
    if (s.exp >= 0) t.tracked = 0;
}
// calculate.py:35-55 mark_to_mkt on the Decimal fields: the sign (-1, 0, +1) of nav = (cash + cash_on_hold) + position_val
CDA_HDN int cda_twin_nav_sign(CdaTwin &t, long long cash_int, long long hold, long long price) {
    const long long pos = t.pos, ap = pos < 0 ? -pos : pos;
    const CdaDec cash = t.tracked ? t.cash : CDA_DI(cash_int);
    const CdaDec diff = pos >= 0 ? cda_dec_sub(CDA_DI(price), t.vwap) : cda_dec_sub(t.vwap, CDA_DI(price));
    const CdaDec pv = cda_dec_add(cda_dec_mul(CDA_DI(ap), t.vwap, &t.err), cda_dec_mul(CDA_DI(ap), diff, &t.err));
    const CdaDec nav = cda_dec_add(cda_dec_add(cash, CDA_DI(hold)), pv);
    return nav.c == 0 ? 0 : (nav.sign ? -1 : 1);
}

#if defined(__CUDACC__)
// Replay `jn` journal entries of one agent onto its stored twin (global memory), store it back.  Returns the twin's tracked flag in
// bit 0 and a range error in bit 1.  query: 0 none; 1 -> bits 8..9 = sign code of (cash - qarg) (1 greater, 2 equal, 3 less);
// 2 -> bits 8..9 = sign code of nav at price qarg2 with integer cash qarg and cash_on_hold qhold.
__device__ __noinline__ unsigned cda_twin_replay(CdaTwinStored *st, const unsigned long long *jr, int jn, int query, long long qarg, long long qhold, long long qarg2) {
    CdaTwin t;
    cda_twin_load(t, st);
    for (int i = 0; i < jn; ++i) cda_twin_apply(t, jr[i]);
    cda_twin_settle(t);
    unsigned out = 0;
    if (query == 1) {
        const int c = t.tracked ? cda_dec_cmp(t.cash, CDA_DI(qarg)) : 0;       // untracked: the Decimal cash IS the integer cash (== qarg at a tie)
        out = (c > 0 ? 1u : c == 0 ? 2u : 3u) << 8;
    } else if (query == 2) {
        const int c = cda_twin_nav_sign(t, qarg, qhold, qarg2);
        out = (c > 0 ? 1u : c == 0 ? 2u : 3u) << 8;
    }
    cda_twin_store(t, st);
    return out | (unsigned)t.tracked | (t.err ? 2u : 0u);
}
#endif
