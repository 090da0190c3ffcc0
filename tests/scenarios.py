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
        # backends answer with [1, ...] batches (oracle, CUDA: numpy / torch) or plain rows (live reference)
        arr = [np.asarray(x.cpu() if hasattr(x, "cpu") else x) for x in out]
        self.obs = arr[0].reshape(-1).astype(np.float32)
        self.rew = arr[1].reshape(-1).astype(np.float64)
        self.term, self.trunc = bool(arr[2].reshape(-1)[0]), bool(arr[3].reshape(-1)[0])
        return out

    def snap(self, j=-1):
        """j-th of the n_hist stacked 42-float snapshots (oldest first; -1 = newest)."""
        return self.obs.reshape(-1, 42)[j]

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


def reward_terms_through_a_round_trip(be):
    """reward_helper.py:35-103 with calculate.py:35-55 — order/trade penalties, passive bonus, mark-to-market of a long and
    a short, the 1.5x loss multiplier and the drawdown penalty, all by hand (the reference's unit vectors 39.9 / -170.0,
    test_reward_logic.py:58-114, poke the helper with mocked accounts; this drives the same terms through step())."""
    s = Script(be, cfg(init_cash=100000))
    s.act(0, BID_LMT, 10, level=0, off=AGGR)           # bid 10 @ 100: one order placed, no tape yet -> no mark-to-market
    assert np.allclose(s.rew, [-0.1, 0, 0, 0], atol=1e-12)
    s.act(1, ASK_MKT, 10)                              # a1 sells 10 @ 100 to a0
    # a1: placed 1, one trade: -0.1 - 0.05 ; a0 (passive): one trade, one passive fill: -0.05 + 0.1 ; NAVs unchanged at 100
    assert np.allclose(s.rew, [0.05, -0.15, 0, 0], atol=1e-12)
    assert s.acc(0)["nav"] == 100000 and s.acc(1)["nav"] == 100000 and s.acc(1)["pv"] == 1000
    s.act(2, ASK_LMT, 5, level=0, off=JOIN)            # ghost ask level 0 = last_price + 1 = 101
    assert s.book("asks") == [(101, 5, 2, 3, 3)] and np.allclose(s.rew, [0, 0, -0.1, 0], atol=1e-12)
    s.act(3, BID_MKT, 5)                               # a3 buys 5 @ 101: last_price 101
    # a0 long 10 from 100: nav +10 -> +10 ; a1 short 10 from 100: nav -10 -> -15 (x1.5) and drawdown 10 -> -2 ; a2 passive
    # short 5 @ 101: -0.05 + 0.1 ; a3: -0.1 - 0.05
    assert np.allclose(s.rew, [10.0, -17.0, 0.05, -0.15], atol=1e-12)
    assert [s.acc(i)["nav"] for i in range(4)] == [100010, 99990, 100000, 100000]
    assert s.acc(1)["pv"] == 990 and s.acc(0)["pv"] == 1010 and s.acc(2)["pv"] == 505 and s.acc(3)["pv"] == 505
    s.act(1, BID_MKT, 1)                               # empty ask side: nothing trades; a1 keeps its drawdown: -0.1 - 0.2*10
    assert np.allclose(s.rew, [0, -2.1, 0, 0], atol=1e-12)


def snapshot_formulas_and_history_window(be):
    """state_helper.py:113-214 + :66-111 — level aggregation, midpoint normalisation, sqrt volumes, sign conventions, the two
    market features, the empty-book fallbacks and the oldest-to-newest stacking (test_obs_normalization.py:142-250,
    test_obs_market_features.py:76-178, test_observation_history.py:9-43)."""
    s = Script(be, cfg(init_cash=100000))
    first = s.be.reset_one(0)
    first = np.asarray(first.cpu() if hasattr(first, "cpu") else first, np.float32).reshape(-1, 42)
    empty = np.zeros(42, np.float32); empty[40] = np.float32(np.log(100.0))      # empty book: zeros, log(anchor), 0
    assert all(np.array_equal(f, empty) for f in first) and first.shape == (4, 42)   # reset pads n_hist identical frames
    s.act(0, BID_LMT, 3, level=0, off=JOIN)            # bid 3 @ 99
    one_sided = s.snap().copy()
    # only bids: M = best bid (99) -> normalised price 0, size sqrt(3); log_mid = log 99; spread feature 0
    exp = np.zeros(42, np.float32); exp[10] = np.float32(np.sqrt(3.0)); exp[40] = np.float32(np.log(99.0))
    assert np.array_equal(one_sided, exp)
    s.act(1, BID_LMT, 4, level=1, off=JOIN)            # level 1 is empty: ghost 100 - 2 = 98
    s.act(2, BID_LMT, 5, level=1, off=JOIN)            # level 1 is real now (98): joins it
    s.act(3, ASK_LMT, 16, level=1, off=JOIN)           # ghost ask 100 + 2 = 102
    M = 100.5
    exp = np.zeros(42, np.float64)
    exp[0], exp[1] = (M - 99) / M, (M - 98) / M        # bid prices: (M - p) / M, best first
    exp[10], exp[11] = np.sqrt(3.0), 3.0               # bid sizes: sqrt(level volume); 4 + 5 aggregate into one level
    exp[20] = -((102 - M) / M)                         # ask prices: -((p - M) / M)
    exp[30] = -4.0                                     # ask sizes: -sqrt(16)
    exp[40], exp[41] = np.log(M), np.log1p(3.0)        # log mid, log1p(spread in ticks)
    assert np.array_equal(s.snap(), exp.astype(np.float32))
    # the stack slides: 4 steps after the reset the oldest frame is the one-sided book of step 1
    assert np.array_equal(s.snap(0), one_sided)
    assert s.d["best_bid"] == 99 and s.d["best_ask"] == 102 and not s.term and not s.trunc


def position_flips_aggressor_and_passive(be):
    """test_accounting.py:216-320 — one trade flips the aggressor short->long and the passive side long->short: the covered part
    settles at the old cost basis (short side: cash += 2*cost - 2*mkt), the remainder opens at the trade price."""
    s = Script(be, cfg(init_cash=100000))
    s.act(0, BID_LMT, 5, level=0, off=AGGR)            # bid 5 @ 100
    s.act(1, ASK_MKT, 5)                               # a0 long 5 @ 100, a1 short 5 @ 100
    s.act(0, ASK_LMT, 8, level=0, off=JOIN)            # ghost ask 101: a0 offers 8 (3 more than it holds)
    assert (s.acc(0)["cash"], s.acc(0)["hold"]) == (98692, 808)
    s.act(1, BID_MKT, 8)                               # a1 covers 5 and opens 3 long; a0 sells its 5 and opens 3 short
    assert s.d["fills"].tolist() == [[4, 101, 8, 0, 3, -1, 1, 0]]     # maker order id 3: the market order of step 2 took id 2
    a0, a1 = s.acc(0), s.acc(1)
    assert (a1["pos"], a1["cash"], a1["hold"], a1["pv"], a1["nav"]) == (3, 99692, 0, 303, 99995)
    assert (a0["pos"], a0["cash"], a0["hold"], a0["pv"], a0["nav"]) == (-3, 99702, 0, 303, 100005)
    assert np.allclose(s.rew, [5.0 - 0.05 + 0.1, -5 * 1.5 - 0.1 - 0.05 - 0.2 * 5, 0, 0], atol=1e-12)


def truncation_lands_on_max_step(be):
    """done_helper.py:3-55 / test_env_lifecycle.py:94-122 — truncated flips exactly when t_step reaches max_step; nobody is done."""
    s = Script(be, cfg(max_step=3))
    flags = []
    for _ in range(3):
        s.act(0, 0, 1)                                 # everybody passes
        flags.append((s.term, s.trunc))
    assert flags == [(False, False), (False, False), (False, True)]
    assert np.array_equal(s.rew, np.zeros(4))


ALL = [position_flips_aggressor_and_passive, reward_terms_through_a_round_trip, snapshot_formulas_and_history_window, truncation_lands_on_max_step, limit_hold_and_cancel, market_short_and_partial_fill, position_flip_uses_cash_gate_for_opening_leg_only,
       modify_scenarios, self_trade_moves_escrow_only, empty_book_market_order_is_a_noop_but_counts,
       tick_two_prices_stay_on_grid]
