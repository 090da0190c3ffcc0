"""TEST INFRASTRUCTURE — an oracle-backed stand-in for VecCDAEnv (the subset of its surface the dict adapter uses).

Why: the reference's own surface-level tests live under /root/reference (this container: no GPU), the GPU box has no /root/reference.  To run
those tests against the ADAPTER (`continuousDoubleAuctionEnv`: action-dict packing, the info dict, spaces, done/truncation flags, lazy
state attributes) here, its engine is swapped for the CPU oracle through this shim.  The CUDA engine itself is pinned to the same oracle,
bit for bit, by the `-m gpu` parity tests — so "adapter over oracle here" + "CUDA == oracle there" covers the product path end to end.
The product never imports this module."""
import numpy as np

from oracle.cda_oracle import OracleEnv


class _T:
    """The few tensor methods the adapter calls on what VecCDAEnv returns."""

    def __init__(self, a):
        self.a = np.asarray(a)

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def item(self):
        return self.a.item()

    def __getitem__(self, i):
        r = self.a[i]
        return _T(r) if isinstance(r, np.ndarray) else _T(np.asarray(r))

    def __int__(self):
        return int(self.a)


_FIELDS = ("cash", "cash_on_hold", "position_val", "cost_basis", "nav", "prev_nav", "max_nav", "net_position", "num_trades", "num_trades_step",
           "num_passive_fills_step", "order_step_placed", "num_rejected_step", "is_pass_action")   # column order of OracleEnv.dump()["accounts"]


class OracleVec:
    def __init__(self, config=None, num_markets=1, device=0, order_capacity=0, fill_capacity=0, status_policy="raise", decimal_ledger=False):
        self.M, self.fill_capacity, self.decimal_ledger = int(num_markets), int(fill_capacity), bool(decimal_ledger)
        cfg = {k: v for k, v in (config or {}).items() if not k.startswith("_")}
        self._o = OracleEnv(cfg, self.M, decimal_ledger=decimal_ledger, dec128=decimal_ledger)
        self.A, self.W, self.n_hist = self._o.A, self._o.W, self._o.n_hist
        self._seeded = False

    def reset(self, seed=None, mask=None):
        seeds = None
        if seed is not None:
            seeds = np.arange(self.M, dtype=np.uint64) + np.uint64(seed) if isinstance(seed, (int, np.integer)) else np.asarray(seed, np.uint64)
        return _T(self._o.reset(seeds=seeds, mask=mask).copy())

    def step_host(self, category, size_mean, size_sigma, price, price_offset, sync=True):
        o, r, te, tr = self._o.step(category, size_mean, size_sigma, price, price_offset)
        return o.copy(), r.copy(), te.copy(), tr.copy()

    def _dumps(self):
        return [self._o.dump(m) for m in range(self.M)]

    def info_all(self):
        d = self._dumps()
        out = {name: _T(np.stack([x["accounts"][:, j] for x in d])) for j, name in enumerate(_FIELDS)}
        out["market"] = _T(np.array([[x["last_price"], x["best_bid"], x["best_ask"], x["time"], x["next_order_id"], x["t_step"], x["done_mask"], x["status"]]
                                     for x in d], np.int64))
        return out

    def info(self, field):
        return self.info_all()[field]

    def status(self):
        return _T(self.info_all()["market"].a[:, 7])

    def check_status(self, include_fill_log=False):
        bits = int(np.bitwise_or.reduce(self.status().a)) & (~0 if include_fill_log else 29)
        if bits:
            raise RuntimeError(f"oracle market status {bits}")
        return 0

    def enable_action_log(self, on=True):
        self._log = bool(on)

    def last_actions(self):
        return _T(np.stack([self._o.last_actions(m) for m in range(self.M)]))

    def dump(self, m=0):
        return self._o.dump(m)

    def fills(self):
        d = self._dumps()
        cap = max(self.fill_capacity, 1)
        f = np.zeros((self.M, cap, 8), np.int32)
        n = np.zeros(self.M, np.int32)
        for m, x in enumerate(d):
            n[m] = x["n_fills"]; k = min(int(n[m]), cap, len(x["fills"])); f[m, :k] = x["fills"][:k]
        return _T(f), _T(n)

    def decimal_fields(self, markets=None):
        names = {"cash": "cash", "VWAP": "VWAP", "cash_on_hold": "cash_on_hold", "position_val": "position_val", "nav": "nav"}
        return {m: {k: self._o.dump_decimal(m)[v] for k, v in names.items()} for m in (range(self.M) if markets is None else markets)}

    def close(self):
        pass
