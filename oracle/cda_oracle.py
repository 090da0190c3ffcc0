"""TEST INFRASTRUCTURE — ctypes wrapper over the CPU oracle (oracle/cda_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcda_oracle.so")
_SO128 = os.path.join(_HERE, "_build", "libcda_oracle_dec128.so")   # decimal ledger on unsigned __int128 (dec128.h)

DEFAULTS = dict(  # config/env_defaults.json of the reference (standalone defaults)
    num_of_agents=4, init_cash=1_000_000, tick_size=1, max_step=64, n_hist=4,
    initial_price_min=10, initial_price_max=100, min_size=1, mkt_max_size=100,
    limit_size_multiple=10, order_penalty=0.1, trade_penalty=0.05, drawdown_penalty=0.2,
    passive_bonus=0.1, loss_multiplier=1.5,
)

SNAP = 42


class OrcConfig(ctypes.Structure):
    _fields_ = [
        ("num_agents", ctypes.c_int32), ("n_hist", ctypes.c_int32), ("max_step", ctypes.c_int32),
        ("tick", ctypes.c_int32), ("init_cash", ctypes.c_int64), ("min_size", ctypes.c_int32),
        ("mkt_max_size", ctypes.c_int32), ("limit_size_multiple", ctypes.c_int32),
        ("price_lo", ctypes.c_int32), ("price_hi", ctypes.c_int32),
        ("order_penalty", ctypes.c_double), ("trade_penalty", ctypes.c_double),
        ("drawdown_penalty", ctypes.c_double), ("passive_bonus", ctypes.c_double),
        ("loss_multiplier", ctypes.c_double), ("decimal_ledger", ctypes.c_int32),
    ]


def build(force=False):
    """Compile the oracle with gcc (no-op when up to date)."""
    srcs = [os.path.join(_HERE, f) for f in ("cda_oracle.c", "np_rng.h", "zig_tables.h", "dec28.h", "dec128.h")]
    if (not force and os.path.exists(_SO) and os.path.exists(_SO128)
            and all(min(os.path.getmtime(_SO), os.path.getmtime(_SO128)) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_libs = {}


def lib(dec128=False):
    """dec128=True: the build whose decimal ledger runs on unsigned __int128 (oracle/dec128.h)."""
    if dec128 not in _libs:
        build()
        L = ctypes.CDLL(_SO128 if dec128 else _SO)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(OrcConfig), ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        _libs[dec128] = L
    return _libs[dec128]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class OracleEnv:
    """M independent markets stepped on the CPU by the C oracle (tensor-style API)."""

    def __init__(self, config=None, num_markets=1, decimal_ledger=False, dec128=False):
        """decimal_ledger=True: the oracle ALSO keeps the reference's Decimal(prec 28) money fields (oracle/dec28.h) and decides the
        cash gate / bankruptcy / high-water mark on them — reproduces the reference even where its ~1e-24 VWAP residues
        flip a `cash >= order value` test at exact equality (the exact-integer ledger, default, is what the CUDA env runs)."""
        cfg = dict(DEFAULTS)
        cfg.update(config or {})
        self.cfg = cfg
        self.M = int(num_markets)
        self.A = int(cfg["num_of_agents"])
        self.n_hist = int(cfg["n_hist"])
        self.W = self.n_hist * SNAP
        if float(cfg["tick_size"]) != int(cfg["tick_size"]):
            raise ValueError("oracle supports integral tick_size only")
        c = OrcConfig(self.A, self.n_hist, int(cfg["max_step"]), int(cfg["tick_size"]),
                      int(cfg["init_cash"]), int(cfg["min_size"]), int(cfg["mkt_max_size"]),
                      int(cfg["limit_size_multiple"]), int(cfg["initial_price_min"]),
                      int(cfg["initial_price_max"]), float(cfg["order_penalty"]),
                      float(cfg["trade_penalty"]), float(cfg["drawdown_penalty"]),
                      float(cfg["passive_bonus"]), float(cfg["loss_multiplier"]), 1 if decimal_ledger else 0)
        self._L = lib(dec128)
        if decimal_ledger:
            self._L.orc_dec_reset_range_errors()   # (a process-wide counter of dec128.h; an env reports it through its status word)
        self._h = self._L.orc_create(ctypes.byref(c), self.M)
        if not self._h:
            raise ValueError("orc_create rejected the config")
        self.obs = np.zeros((self.M, self.W), np.float32)
        self.reward = np.zeros((self.M, self.A), np.float64)
        self.terminated = np.zeros(self.M, np.uint8)
        self.truncated = np.zeros(self.M, np.uint8)

    def dump_decimal(self, m=0):
        """decimal_ledger mode: {field: [Decimal per agent]} of market m — the reference's own Decimal values, residues included."""
        from decimal import Decimal
        buf = ctypes.create_string_buffer(self.A * 7 * 48)
        self._L.orc_dump_accounts_dec(ctypes.c_void_p(self._h), int(m), buf)
        names = ("cash", "cash_on_hold", "position_val", "VWAP", "nav", "prev_nav", "max_nav")
        raw = buf.raw
        get = lambda i, j: Decimal(raw[(i * 7 + j) * 48:(i * 7 + j + 1) * 48].split(b"\0", 1)[0].decode())
        return {n: [get(i, j) for i in range(self.A)] for j, n in enumerate(names)}

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(ctypes.c_void_p(self._h))
            self._h = None

    def reset(self, seeds=None, mask=None):
        s = None if seeds is None else np.ascontiguousarray(seeds, np.uint64)
        k = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._L.orc_reset(ctypes.c_void_p(self._h), _p(s), _p(k), _p(self.obs))
        return self.obs

    @staticmethod
    def _acts(category, size_mean, size_sigma, price, price_offset):
        return (np.ascontiguousarray(category, np.int32), np.ascontiguousarray(size_mean, np.float32),
                np.ascontiguousarray(size_sigma, np.float32), np.ascontiguousarray(price, np.int32),
                np.ascontiguousarray(price_offset, np.int32))

    def step(self, category, size_mean, size_sigma, price, price_offset, nthreads=1):
        a = self._acts(category, size_mean, size_sigma, price, price_offset)
        assert a[0].shape == (self.M, self.A)
        self._L.orc_step(ctypes.c_void_p(self._h), *[_p(x) for x in a], _p(self.obs), _p(self.reward),
                         _p(self.terminated), _p(self.truncated), int(nthreads))
        return self.obs, self.reward, self.terminated, self.truncated

    def rollout(self, category, size_mean, size_sigma, price, price_offset, nthreads=1):
        """actions shaped [T, M, A]; outputs hold the last step."""
        a = self._acts(category, size_mean, size_sigma, price, price_offset)
        T = a[0].shape[0]
        assert a[0].shape == (T, self.M, self.A)
        self._L.orc_rollout(ctypes.c_void_p(self._h), T, *[_p(x) for x in a], _p(self.obs),
                            _p(self.reward), _p(self.terminated), _p(self.truncated), int(nthreads))
        return self.obs, self.reward, self.terminated, self.truncated

    def rollout_random(self, num_steps, policy_seed=0, nthreads=1):
        """CPU twin of VecCDAEnv.rollout_random: the same counter-based uniform policy, then the ordinary step."""
        self._L.orc_rollout_random(ctypes.c_void_p(self._h), int(num_steps), ctypes.c_uint64(policy_seed), _p(self.obs), _p(self.reward),
                                   _p(self.terminated), _p(self.truncated), int(nthreads))
        return self.obs, self.reward, self.terminated, self.truncated

    def last_actions(self, m=0):
        """Decoded actions of the last step (the reference's LOB_actions): int32 [A, 4] = type, side, size, price; side -1 = pass / absent."""
        out = np.zeros((self.A, 4), np.int32)
        self._L.orc_dump_actions(ctypes.c_void_p(self._h), int(m), _p(out))
        return out

    # ---- canonical state dump (same schema as ref_runner.dump_reference and the GPU env) ----
    def dump(self, m=0):
        h = ctypes.c_void_p(self._h)
        out = {}
        for side, name in ((0, "bids"), (1, "asks")):
            buf = np.zeros((4096, 5), np.int64)
            n = self._L.orc_dump_book(h, m, side, _p(buf), 4096)
            out[name] = buf[:n].copy()
            mp = np.zeros(4096, np.int64)
            n2 = self._L.orc_dump_map(h, m, side, _p(mp), 4096)
            out[name + "_map"] = mp[:n2].copy()
        sc = np.zeros(10, np.int64)
        self._L.orc_dump_scalars(h, m, _p(sc))
        keys = ("time", "next_order_id", "last_price", "tape_nonempty", "t_step", "done_mask",
                "status", "n_fills", "best_bid", "best_ask")
        out.update({k: int(v) for k, v in zip(keys, sc)})
        acc = np.zeros((self.A, 14), np.int64)
        self._L.orc_dump_accounts(h, m, _p(acc))
        out["accounts"] = acc
        rt = np.zeros((self.A, 6), np.float64)
        self._L.orc_dump_reward_terms(h, m, _p(rt))
        out["reward_terms"] = rt
        fl = np.zeros((256, 8), np.int32)
        n = self._L.orc_dump_fills(h, m, _p(fl), 256)
        out["fills"] = fl[:min(n, 256)].copy()
        rg = np.zeros(6, np.uint64)
        self._L.orc_dump_rng(h, m, _p(rg))
        out["rng"] = rg
        return out


ACC_COLS = ("cash", "hold", "pv", "C", "nav", "prev_nav", "max_nav", "pos", "num_trades",
            "num_trades_step", "num_passive_fills_step", "order_step_placed", "num_rejected_step",
            "is_pass")
FILL_COLS = ("time", "price", "qty", "maker", "maker_oid", "maker_left", "taker", "taker_side")
