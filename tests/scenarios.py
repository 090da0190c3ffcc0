"""Hand-scripted single-market scenarios (RNG-independent: sigma = 0, one non-pass action per step),
modelled on the reference's known-answer unit tests: test_accounting.py, test_modify_order.py,
test_cash_check.py, test_orderbook_crossed_book.py, test_new_action_space.py.
A backend exposes reset_one(seed) / step_one(cat, mean, sigma, price, off) / dump_one()."""
import numpy as np

A = 4
BID_MKT, BID_LMT, BID_MOD, BID_CAN, ASK_MKT, ASK_LMT, ASK_MOD, ASK_CAN = 1, 2, 3, 4, 5, 6, 7, 8
PASSIVE, JOIN, AGGR = 0, 1, 2


def cfg(**kw):
    c = dict(num_of_agents=A, init_cash=1000, max_step=1000, n_hist=4, initial_price_min=100, initial_price_max=100)
    c.update(kw)
    return c


class Script:
    def __init__(self, be, c):
        self.be = be
        self.cfg = c
        be.reset_one(0)

    def act(self, agent, cat, size, level=0, off=JOIN):
        """One agent acts, the rest pass.  size -> size_mean so that rint(|mul*mean|) + min_size == size."""
        mul = 49.5 if cat in (BID_MKT, ASK_MKT) else 499.5
        cats = np.zeros(A, np.int32); mean = np.zeros(A, np.float32); sig = np.zeros(A, np.float32)
        pr = np.zeros(A, np.int32); of = np.ones(A, np.int32)
        cats[agent], pr[agent], of[agent] = cat, level, off
        mean[agent] = np.float32((size - 1) / mul)
        out = self.be.step_one(cats, mean, sig, pr, of)
        self.d = self.be.dump_one()
        return out

    def acc(self, i):
        a = self.d["accounts"][i]
        return dict(cash=int(a[0]), hold=int(a[1]), pv=int(a[2]), nav=int(a[4]), pos=int(a[7]), trades=int(a[8]),
                    placed=int(a[11]), rejected=int(a[12]))

    def book(self, side):
        return [tuple(int(x) for x in r) for r in self.d[side]]   # (price, qty, trader, order_id, ts)


def limit_hold_and_cancel(be):
    """test_accounting.py:23-78 — escrow on a resting limit, released by cancel."""
    s = Script(be, cfg())
    s.act(0, BID_LMT, 1, level=0, off=AGGR)            # ghost bid level 0 = 99, aggressive +1 -> 100
    assert s.book("bids") == [(100, 1, 0, 1, 1)]
    assert s.acc(0) == dict(cash=900, hold=100, pv=0, nav=1000, pos=0, trades=0, placed=1, rejected=0)
    s.act(1, ASK_LMT, 1, level=0, off=PASSIVE)         # ghost ask level 0 = 101, passive +1 -> 102
    assert s.book("asks") == [(102, 1, 1, 2, 2)] and s.acc(1)["cash"] == 898 and s.acc(1)["hold"] == 102
    # cancel needs the price: best bid is now a real level (100), offset join
    s.act(0, BID_CAN, 1, level=0, off=JOIN)
    assert s.book("bids") == [] and s.acc(0)["cash"] == 1000 and s.acc(0)["hold"] == 0
    assert s.d["time"] == 3 and s.d["next_order_id"] == 2           # cancel advances time, not the id counter


def market_short_and_partial_fill(be):
    """test_accounting.py:80-140 — market sell into a resting bid; partial fill keeps the remainder escrowed."""
    s = Script(be, cfg())
    s.act(0, BID_LMT, 2, level=0, off=AGGR)            # bid 2 @ 100
    s.act(1, ASK_MKT, 1)                               # market sell 1
    assert s.d["fills"].tolist() == [[2, 100, 1, 0, 1, 1, 1, 1]]
    assert s.book("bids") == [(100, 1, 0, 1, 1)]                      # shrinks in place, keeps timestamp 1
    a0, a1 = s.acc(0), s.acc(1)
    assert (a0["cash"], a0["hold"], a0["pv"], a0["pos"], a0["nav"]) == (800, 100, 100, 1, 1000)
    assert (a1["cash"], a1["pv"], a1["pos"], a1["nav"]) == (900, 100, -1, 1000)
    assert s.d["last_price"] == 100


def position_flip_uses_cash_gate_for_opening_leg_only(be):
    """test_cash_check.py — closing needs no cash; a flip is gated on the opening leg only."""
    s = Script(be, cfg(init_cash=1000))
    s.act(0, BID_LMT, 5, level=0, off=AGGR)            # bid 5 @ 100 (hold 500)
    s.act(1, ASK_MKT, 5)                               # a1 short 5 @ 100, cash 500
    assert s.acc(1)["pos"] == -5 and s.acc(1)["cash"] == 500
    s.act(2, ASK_LMT, 9, level=0, off=JOIN)            # ask 9 @ 101 (ghost level from last_price 100)
    # a1 buys 9 at market: closes 5 (no cash needed), opens 4 -> needs 4*101 = 404 <= 500: approved
    s.act(1, BID_MKT, 9)
    assert s.acc(1)["pos"] == 4 and s.acc(1)["rejected"] == 0
    # now a3 (cash 1000) tries to buy 10 @ market with best ask none -> tape fallback price 101 -> 1010 > 1000: rejected
    s.act(3, BID_MKT, 10)
    assert s.acc(3)["rejected"] == 1 and s.acc(3)["placed"] == 0 and s.acc(3)["pos"] == 0


def modify_scenarios(be):
    """test_modify_order.py:19-86 + test_orderbook_crossed_book.py — in-place decrease keeps priority and
    refreshes the timestamp; a price change re-queues; a crossing modify trades."""
    s = Script(be, cfg(init_cash=100000))               # (the cash gate runs BEFORE the old escrow is released)
    s.act(0, BID_LMT, 10, level=0, off=AGGR)           # oid1 bid 10 @ 100 ts1
    s.act(1, BID_LMT, 4, level=0, off=JOIN)            # oid2 bid 4 @ 100 ts2 (level 0 is real now: join 100)
    # limit at a price the trader already rests at = upsert-modify; smaller qty: in place
    s.act(0, BID_LMT, 6, level=0, off=JOIN)
    assert s.book("bids") == [(100, 6, 0, 1, 3), (100, 4, 1, 2, 2)]    # still first in the queue, ts refreshed
    assert s.d["time"] == 3 and s.d["next_order_id"] == 2 and s.acc(0)["hold"] == 600
    # larger qty at the same price: removed and re-queued behind oid2, same order id
    s.act(0, BID_LMT, 8, level=0, off=JOIN)
    assert s.book("bids") == [(100, 4, 1, 2, 2), (100, 8, 0, 1, 4)] and s.acc(0)["hold"] == 800
    # 'modify' picks the trader's oldest-timestamp order whatever its price, moves it to 99
    s.act(0, BID_MOD, 3, level=0, off=PASSIVE)
    assert s.book("bids") == [(100, 4, 1, 2, 2), (99, 3, 0, 1, 5)] and s.acc(0)["hold"] == 297
    # a resting ask, then a modify that crosses it must trade (never leave best_bid >= best_ask)
    s.act(2, ASK_LMT, 2, level=0, off=JOIN)            # ask 2 @ 101
    s.act(0, BID_MOD, 5, level=0, off=AGGR)            # bid level 0 = 100 -> 101: crosses, fills 2, rests 3 @ 101
    assert s.d["fills"].tolist() == [[7, 101, 2, 2, 3, -1, 0, 0]]
    assert s.book("bids")[0] == (101, 3, 0, 1, 7) and s.book("asks") == []
    assert s.d["best_bid"] == 101 and s.d["best_ask"] == 0


def self_trade_moves_escrow_only(be):
    """trader.py:321-322 / cash_processor.py:55-62 — initiator == counter party: escrow back to cash, no
    position, no trade count (but the tape still moves last_price)."""
    s = Script(be, cfg())
    s.act(0, ASK_LMT, 3, level=0, off=JOIN)            # ask 3 @ 101
    s.act(0, BID_MKT, 2)                               # buys from itself
    a0 = s.acc(0)
    assert a0["pos"] == 0 and a0["trades"] == 0 and a0["hold"] == 101 and a0["cash"] == 899 and a0["nav"] == 1000
    assert s.d["last_price"] == 101 and s.book("asks") == [(101, 1, 0, 1, 1)]


def empty_book_market_order_is_a_noop_but_counts(be):
    """test_orderbook_new.py empty-book market; orderbook.py:39-44 time/id still advance."""
    s = Script(be, cfg())
    s.act(2, ASK_MKT, 7)
    assert s.d["time"] == 1 and s.d["next_order_id"] == 1 and s.d["fills"].shape[0] == 0
    assert s.acc(2)["placed"] == 1 and s.acc(2)["pos"] == 0


def tick_two_prices_stay_on_grid(be):
    """test_new_action_space.py:172-180 — ghost levels and offsets step by tick_size."""
    s = Script(be, cfg(tick_size=2))
    s.act(0, BID_LMT, 1, level=2, off=JOIN)            # ghost: 100 - 3*2 = 94
    s.act(1, ASK_LMT, 1, level=1, off=PASSIVE)         # ghost: 100 + 2*2 = 104, passive +2 -> 106
    assert s.book("bids") == [(94, 1, 0, 1, 1)] and s.book("asks") == [(106, 1, 1, 2, 2)]


ALL = [limit_hold_and_cancel, market_short_and_partial_fill, position_flip_uses_cash_gate_for_opening_leg_only,
       modify_scenarios, self_trade_moves_escrow_only, empty_book_market_order_is_a_noop_but_counts,
       tick_two_prices_stay_on_grid]
