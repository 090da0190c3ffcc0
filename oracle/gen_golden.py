#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate golden trajectories from the UNMODIFIED Python reference.

Runs /root/reference's continuousDoubleAuctionEnv (through oracle/ref_stub.py) on seeded
configs/actions and stores, per case, everything a parity test needs:
    config, seed, the action arrays, per-step obs f32[T,168] / reward f64[T,A] / flags,
    per-step fill records, per-step (time, next_order_id, last_price), and the final canonical
    dump (book in priority order, order_map order, ledger, RNG state).
The reference tree does not exist on the GPU box, so these fixtures are what travels:
tests/test_oracle_golden.py checks the C oracle against them on CPU, tests/test_gpu_golden.py
checks the CUDA env against them on the B200.

Also records the hand-checkable RNG-independent case of SURVEY.md Appendix D and the
known-answer vectors of the reference's own unit tests that fit the env API.

Usage (this container only):  python -m oracle.gen_golden
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_runner import ReferenceMarket  # noqa: E402

MIXES = {
    "uniform": np.full(9, 1.0 / 9.0),
    "limit_market": np.array([.10, .15, .30, 0, 0, .15, .30, 0, 0]),
    "modify_heavy": np.array([.05, .05, .15, .30, .05, .05, .15, .15, .05]),
}

CASES = [
    # name, config overrides, seed, T, mix, absent probability
    ("uniform_a4", dict(num_of_agents=4), 1000, 96, "uniform", 0.0),
    ("limit_market_a4", dict(num_of_agents=4), 2001, 96, "limit_market", 0.0),
    ("modify_heavy_a8", dict(num_of_agents=8), 3002, 96, "modify_heavy", 0.0),
    ("low_cash_a4", dict(num_of_agents=4, init_cash=3000), 4003, 120, "uniform", 0.0),
    ("nhist2_a5_absent", dict(num_of_agents=5, n_hist=2), 5004, 80, "uniform", 0.25),
    ("nhist6_a3_fixed_anchor", dict(num_of_agents=3, n_hist=6, initial_price_min=50, initial_price_max=50), 6005, 64, "limit_market", 0.0),
    ("trunc_a4_maxstep10", dict(num_of_agents=4, max_step=10), 7006, 12, "uniform", 0.0),
    # configuration corners (added after the first set; existing fixtures are never rewritten unless --force)
    ("single_agent_a1", dict(num_of_agents=1), 8007, 64, "uniform", 0.0),                      # self-trades only, no permutation draw
    ("two_agents_a2", dict(num_of_agents=2), 8108, 80, "limit_market", 0.0),
    ("sixteen_agents_a16", dict(num_of_agents=16), 8209, 48, "uniform", 0.0),                  # capacity-256 kernel
    ("tick5_a4", dict(num_of_agents=4, tick_size=5), 8310, 80, "uniform", 0.0),                # integral tick != 1
    ("sizes_a4", dict(num_of_agents=4, min_size=3, mkt_max_size=20, limit_size_multiple=4), 8411, 80, "uniform", 0.0),
    ("reward_coeffs_a4", dict(num_of_agents=4, order_penalty=0.3, trade_penalty=0.01, drawdown_penalty=0.5, passive_bonus=0.25,
                              loss_multiplier=2.0), 8512, 80, "limit_market", 0.0),
    ("low_anchor_a4", dict(num_of_agents=4, initial_price_min=1, initial_price_max=3), 8613, 80, "uniform", 0.0),   # prices clamp at one tick
    ("high_anchor_a4", dict(num_of_agents=4, initial_price_min=4000, initial_price_max=5000, init_cash=50_000_000), 8714, 80, "modify_heavy", 0.0),
]


def gen_actions(rng, T, A, mix, absent_p):
    p = MIXES[mix]
    cat = rng.choice(9, size=(T, A), p=p / p.sum()).astype(np.int32)
    mean = rng.uniform(-1, 1, (T, A)).astype(np.float32)
    sigma = rng.uniform(0, 1, (T, A)).astype(np.float32)
    price = rng.integers(0, 10, (T, A)).astype(np.int32)
    off = rng.integers(0, 3, (T, A)).astype(np.int32)
    if absent_p > 0:
        cat[rng.random((T, A)) < absent_p] = -1
    return cat, mean, sigma, price, off


def run_case(cfg_over, seed, acts):
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=100_000, n_hist=4)
    cfg.update(cfg_over)
    ref = ReferenceMarket(cfg)
    obs0 = ref.reset(seed=seed)
    T, A = acts[0].shape
    obs = np.zeros((T, obs0.shape[0]), np.float32)
    rew = np.zeros((T, A), np.float64)
    term = np.zeros(T, np.uint8)
    trunc = np.zeros(T, np.uint8)
    scal = np.zeros((T, 3), np.int64)
    fills, fill_ptr = [], [0]
    acc = np.zeros((T, A, 14), np.int64)
    for t in range(T):
        o, r, te, tr = ref.step(*[a[t] for a in acts])
        obs[t], rew[t], term[t], trunc[t] = o, r, te, tr
        d = ref.dump()
        scal[t] = (d["time"], d["next_order_id"], d["last_price"])
        acc[t] = d["accounts"]
        fills.append(d["fills"])
        fill_ptr.append(fill_ptr[-1] + d["fills"].shape[0])
    final = ref.dump()
    out = dict(cfg_keys=np.array(sorted(cfg)), cfg_vals=np.array([float(cfg[k]) for k in sorted(cfg)]),
               seed=np.int64(seed), obs0=obs0, obs=obs, reward=rew, terminated=term, truncated=trunc,
               scalars=scal, accounts=acc, fills=np.concatenate(fills, 0) if fills else np.zeros((0, 8), np.int32),
               fill_ptr=np.array(fill_ptr, np.int64),
               cat=acts[0], mean=acts[1], sigma=acts[2], price=acts[3], off=acts[4],
               final_bids=final["bids"], final_asks=final["asks"], final_bids_map=final["bids_map"],
               final_asks_map=final["asks_map"], final_rng=final["rng"])
    return out


def appendix_d_case():
    """RNG-independent, hand-checkable smoke golden."""
    A, T = 4, 5
    cat = np.zeros((T, A), np.int32); mean = np.zeros((T, A), np.float32); sigma = np.zeros((T, A), np.float32)
    price = np.zeros((T, A), np.int32); off = np.ones((T, A), np.int32)
    cat[0, 0], price[0, 0], off[0, 0], mean[0, 0] = 6, 0, 1, 0.02
    cat[1, 1], price[1, 1], off[1, 1], mean[1, 1] = 2, 0, 2, 0.01
    cat[2, 2], mean[2, 2] = 1, 0.1
    cat[3, 3], mean[3, 3] = 5, 0.2
    cat[4, 0], price[4, 0], off[4, 0], mean[4, 0] = 7, 3, 0, 0.004
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=8, initial_price_min=50, initial_price_max=50)
    return cfg, 0, (cat, mean, sigma, price, off)


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    force = "--force" in sys.argv
    for name, cfg_over, seed, T, mix, absent in CASES:
        if os.path.exists(os.path.join(outdir, f"traj_{name}.npz")) and not force:
            continue
        A = cfg_over.get("num_of_agents", 4)
        acts = gen_actions(np.random.default_rng(seed + 7), T, A, mix, absent)
        out = run_case(cfg_over, seed, acts)
        np.savez_compressed(os.path.join(outdir, f"traj_{name}.npz"), **out)
        print(name, "T", T, "fills", out["fills"].shape[0], "time", out["scalars"][-1, 0])
    if force or not os.path.exists(os.path.join(outdir, "traj_appendix_d.npz")):
        cfg, seed, acts = appendix_d_case()
        out = run_case(cfg, seed, acts)
        np.savez_compressed(os.path.join(outdir, "traj_appendix_d.npz"), **out)
        print("appendix_d rewards:\n", out["reward"])


if __name__ == "__main__":
    main()
