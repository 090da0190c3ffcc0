"""TEST INFRASTRUCTURE — drives the UNMODIFIED Python reference (through ref_stub) and dumps its
state in the canonical schema shared with oracle/cda_oracle.py::OracleEnv.dump and the GPU env.

Only usable where /root/reference exists (this container); used by
tests/test_oracle_vs_reference.py and oracle/gen_golden.py.
"""
from decimal import Decimal

import numpy as np

from . import ref_stub


def actions_to_dict(cat, mean, sigma, price, off):
    """[A] arrays -> the reference's action dict, in agent order (absent agent: cat < 0)."""
    d = {}
    for i in range(len(cat)):
        if int(cat[i]) < 0:
            continue
        d[f"agent_{i}"] = {
            "category": int(cat[i]),
            "size_mean": np.array([mean[i]], dtype=np.float32),
            "size_sigma": np.array([sigma[i]], dtype=np.float32),
            "price": int(price[i]),
            "price_offset": int(off[i]),
        }
    return d


def _int(x, what):
    d = Decimal(x) if not isinstance(x, Decimal) else x
    r = d.to_integral_value()
    if abs(d - r) > Decimal("1e-9"):
        raise AssertionError(f"{what} = {x!r} is not integral (integer-ledger assumption broken)")
    return int(r)


def dump_reference(env, infos=None):
    lob = env.LOB
    out = {}
    for name, tree, rev in (("bids", lob.bids, True), ("asks", lob.asks, False)):
        rows = []
        items = list(tree.price_map.items())
        if rev:
            items = items[::-1]
        for price, olist in items:
            o = olist.head_order
            vol = 0
            n = 0
            while o is not None:
                rows.append((_int(o.price, "price"), _int(o.quantity, "qty"), int(o.trade_id),
                             int(o.order_id), int(o.timestamp)))
                vol += _int(o.quantity, "qty")
                n += 1
                o = o.next_order
            assert n == len(olist) and vol == _int(olist.volume, "level volume")
        out[name] = np.array(rows, np.int64).reshape(-1, 5)
        out[name + "_map"] = np.array(list(tree.order_map.keys()), np.int64)
    out["time"] = int(lob.time)
    out["next_order_id"] = int(lob.next_order_id)
    out["last_price"] = _int(env.last_price, "last_price")
    out["tape_nonempty"] = int(len(lob.tape) > 0)
    out["t_step"] = int(env.t_step)
    mask = 0
    for a in env.done_set:
        mask |= 1 << int(a.split("_")[1])
    out["done_mask"] = mask
    bb, ba = lob.get_best_bid(), lob.get_best_ask()
    out["best_bid"] = 0 if bb is None else _int(bb, "best_bid")
    out["best_ask"] = 0 if ba is None else _int(ba, "best_ask")
    acc = np.zeros((len(env.traders), 14), np.int64)
    rt = np.zeros((len(env.traders), 6), np.float64)
    for i, tr in enumerate(env.traders):
        a = tr.acc
        acc[i, 0] = _int(a.cash, "cash")
        acc[i, 1] = _int(a.cash_on_hold, "hold")
        acc[i, 2] = _int(a.position_val, "position_val")
        acc[i, 3] = _int(abs(a.net_position) * a.VWAP, "|pos|*VWAP")
        acc[i, 4] = _int(a.nav, "nav")
        acc[i, 5] = _int(a.prev_nav, "prev_nav")
        acc[i, 6] = _int(a.max_nav, "max_nav")
        acc[i, 7] = int(a.net_position)
        acc[i, 8] = int(a.num_trades)
        if infos is not None and f"agent_{i}" in infos and infos[f"agent_{i}"]:
            inf = infos[f"agent_{i}"]
            acc[i, 9] = inf["num_trades_step"]
            acc[i, 10] = inf["num_passive_fills_step"]
            acc[i, 11] = inf["order_step_placed"]
            acc[i, 12] = inf["num_rejected_step"]
            acc[i, 13] = int(inf["is_pass_action"])
            t = inf["reward_terms"]
            rt[i, :5] = [t["nav_term"], t["order_penalty"], t["trade_penalty"],
                         t["drawdown_penalty"], t["passive_bonus"]]
            rt[i, 5] = inf["drawdown"]
    out["accounts"] = acc
    out["reward_terms"] = rt
    fills = []
    for trades in (env.seq_trades or []):
        for tr in trades:
            cp, ip = tr["counter_party"], tr["init_party"]
            left = cp["new_book_quantity"]
            fills.append((int(tr["time"]), _int(tr["price"], "fill price"), _int(tr["quantity"], "fill qty"),
                          int(cp["ID"]), int(cp["order_id"]), -1 if left is None else _int(left, "left"),
                          int(ip["ID"]), 0 if ip["side"] == "bid" else 1))
    out["fills"] = np.array(fills, np.int32).reshape(-1, 8)
    out["n_fills"] = len(fills)
    st = env.np_random.bit_generator.state
    s, inc = st["state"]["state"], st["state"]["inc"]
    m64 = (1 << 64) - 1
    out["rng"] = np.array([s >> 64, s & m64, inc >> 64, inc & m64, st["has_uint32"], st["uinteger"]],
                          np.uint64)
    return out


class ReferenceMarket:
    """One reference env == one market; tensor-style step for easy comparison."""

    def __init__(self, config):
        self.env = ref_stub.make_reference_env(config)
        self.A = self.env.num_of_agents
        self.last_infos = None

    def reset(self, seed=None):
        obs, _ = self.env.reset(seed=seed)
        self.last_infos = None
        return obs["agent_0"].copy()

    def step(self, cat, mean, sigma, price, off):
        o, r, te, tr, inf = self.env.step(actions_to_dict(cat, mean, sigma, price, off))
        self.last_infos = inf
        obs = o["agent_0"].copy()
        rew = np.array([r[f"agent_{i}"] for i in range(self.A)], np.float64)
        return obs, rew, bool(te["__all__"]), bool(tr["__all__"])

    def dump(self):
        return dump_reference(self.env, self.last_infos)
