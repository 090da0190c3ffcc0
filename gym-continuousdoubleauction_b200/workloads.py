"""Synthetic action workloads of BASELINE.json / SURVEY.md §8d (seeded, numpy on the host).

category codes follow the reference's _CATEGORY_MAP (action_helper.py:12-22):
0 pass, 1-4 bid {market, limit, modify, cancel}, 5-8 ask {market, limit, modify, cancel}.
"""
import numpy as np

MIXES = {
    # config 2: uniform random actions (CDA_rand / RandomRLModule shaped)
    "uniform": np.full(9, 1.0 / 9.0),
    # config 3: scripted limit + market mix
    "limit_market": np.array([.10, .15, .30, 0, 0, .15, .30, 0, 0]),
    # config 4: modify-order heavy mix
    "modify_heavy": np.array([.05, .05, .15, .30, .05, .05, .15, .15, .05]),
}


def make_actions(seed, steps, num_markets, num_agents, mix="uniform"):
    """Returns 5 arrays shaped [steps, M, A]: category i32, size_mean f32, size_sigma f32,
    price i32, price_offset i32.  (Vectorised float32 draws: generating the actions must not dominate
    the CPU baseline it feeds.)"""
    rng = np.random.default_rng(seed)
    shape = (steps, num_markets, num_agents)
    p = MIXES[mix]
    cum = np.cumsum(p / p.sum()).astype(np.float32)
    cum[-1] = 1.0
    cat = np.minimum(np.searchsorted(cum, rng.random(shape, dtype=np.float32), side="right"), 8).astype(np.int32)
    mean = rng.random(shape, dtype=np.float32) * np.float32(2.0) - np.float32(1.0)
    sigma = rng.random(shape, dtype=np.float32)
    price = rng.integers(0, 10, shape, dtype=np.int32)
    off = rng.integers(0, 3, shape, dtype=np.int32)
    return cat, mean, sigma, price, off
