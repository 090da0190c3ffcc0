"""Shared helpers for parity tests (oracle <-> reference <-> GPU)."""
import numpy as np

BOOK_KEYS = ("bids", "asks", "bids_map", "asks_map")
SCALAR_KEYS = ("time", "next_order_id", "last_price", "t_step", "done_mask", "best_bid", "best_ask")
# account columns shared by every dump: cash, hold, pv, C, nav, prev_nav, max_nav, pos, num_trades,
# trades_step, passive_step, placed, rejected, is_pass
ACC_COLS = 14


def assert_dump_equal(a, b, ctx="", fills=True, rng=True):
    """Bit-exact comparison of two canonical dumps (integer book state, ledger, fills, RNG)."""
    for k in BOOK_KEYS:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"{ctx}: {k} differs\n{a[k]}\n{b[k]}"
    for k in SCALAR_KEYS:
        assert int(a[k]) == int(b[k]), f"{ctx}: {k} differs: {a[k]} vs {b[k]}"
    assert np.array_equal(a["accounts"][:, :ACC_COLS], b["accounts"][:, :ACC_COLS]), \
        f"{ctx}: accounts differ\n{a['accounts']}\n{b['accounts']}"
    if fills and "fills" in a and "fills" in b:
        assert int(a["n_fills"]) == int(b["n_fills"]), f"{ctx}: n_fills {a['n_fills']} vs {b['n_fills']}"
        assert np.array_equal(a["fills"], b["fills"]), f"{ctx}: fills differ\n{a['fills']}\n{b['fills']}"
    if rng:
        assert np.array_equal(np.asarray(a["rng"], np.uint64), np.asarray(b["rng"], np.uint64)), f"{ctx}: rng state differs"
