"""The CPU oracle against golden trajectories recorded from the UNMODIFIED Python reference
(tests/golden/*.npz, produced by oracle/gen_golden.py in the container that has /root/reference).
Integer state bit-exact; obs/reward within 1e-6 (normally bit-equal)."""
import glob
import os

import numpy as np
import pytest

from oracle.cda_oracle import OracleEnv

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "traj_*.npz")))


def load_cfg(g):
    cfg = {str(k): float(v) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    return {k: (int(v) if float(v).is_integer() else v) for k, v in cfg.items()}


def check_backend_against_golden(make_env, path, exact_obs=False):
    g = np.load(path)
    cfg = load_cfg(g)
    env = make_env(cfg)
    obs0 = env.reset_one(int(g["seed"]))
    assert np.array_equal(obs0, g["obs0"]), "reset obs"
    T = g["cat"].shape[0]
    for t in range(T):
        o, r, te, tr = env.step_one(g["cat"][t], g["mean"][t], g["sigma"][t], g["price"][t], g["off"][t])
        assert np.abs(o.astype(np.float64) - g["obs"][t]).max() <= 1e-6, f"obs t={t}"
        assert np.abs(r - g["reward"][t]).max() <= 1e-6, f"reward t={t}"
        assert int(te) == int(g["terminated"][t]) and int(tr) == int(g["truncated"][t]), f"flags t={t}"
        d = env.dump_one()
        assert (d["time"], d["next_order_id"], d["last_price"]) == tuple(int(x) for x in g["scalars"][t]), f"scalars t={t}"
        assert np.array_equal(d["accounts"][:, :9], g["accounts"][t][:, :9]), f"ledger t={t}"
        assert np.array_equal(d["accounts"][:, 9:14], g["accounts"][t][:, 9:14]), f"step counters t={t}"
        f0, f1 = int(g["fill_ptr"][t]), int(g["fill_ptr"][t + 1])
        assert np.array_equal(d["fills"], g["fills"][f0:f1]), f"fills t={t}"
    d = env.dump_one()
    for k in ("bids", "asks", "bids_map", "asks_map"):
        assert np.array_equal(d[k], g["final_" + k]), k
    assert np.array_equal(np.asarray(d["rng"], np.uint64), g["final_rng"]), "numpy RNG stream state"


class OracleOne:
    def __init__(self, cfg):
        self.e = OracleEnv(cfg, 1)

    def reset_one(self, seed):
        return self.e.reset(seeds=[seed])[0].copy()

    def step_one(self, *a):
        o, r, te, tr = self.e.step(*[np.asarray(x)[None] for x in a])
        return o[0].copy(), r[0].copy(), te[0], tr[0]

    def dump_one(self):
        return self.e.dump(0)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    check_backend_against_golden(OracleOne, path)


def test_golden_fixtures_present():
    assert len(GOLD) >= 8


def test_appendix_d_hand_checked_values():
    """RNG-independent scenario, numbers checkable by hand."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_appendix_d.npz"))
    assert np.allclose(g["reward"][3], [6.0, 0.05, -10.2, -0.15000000000000002], rtol=0, atol=1e-12)
    assert list(g["scalars"][4]) == [5, 4, 50]           # modify: time advances, next_order_id does not
    # fills (time, price, qty, maker, maker_oid, maker_left, taker, taker_side)
    assert g["fills"].tolist() == [[3, 51, 6, 0, 1, 5, 2, 0], [4, 50, 6, 1, 2, -1, 3, 1]]
    # accounts after t=3: a0 (cash 999439, hold 255, pv 312, pos -6, nav 1000006)
    a0 = g["accounts"][3][0]
    assert (a0[0], a0[1], a0[2], a0[7], a0[4]) == (999439, 255, 312, -6, 1000006)
