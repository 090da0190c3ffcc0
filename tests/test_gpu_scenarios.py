"""Known-answer scenarios + sticky status bits on the CUDA env."""
import numpy as np
import pytest
import torch

import scenarios
from test_scenarios_oracle import run
import gym_continuousdoubleauction_b200 as cda

pytestmark = pytest.mark.gpu


class GpuBackend:
    def __init__(self, cfg):
        self.e = cda.VecCDAEnv(cfg, num_markets=1, fill_capacity=32)

    def reset_one(self, seed):
        return self.e.reset(seed=[seed]).cpu().numpy()[0]

    def step_one(self, cat, mean, sigma, price, off):
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)[None]).cuda()
        return self.e.step(t(cat, np.int32), t(mean, np.float32), t(sigma, np.float32), t(price, np.int32), t(off, np.int32))

    def dump_one(self):
        return self.e.dump(0)


@pytest.mark.parametrize("scn", scenarios.ALL, ids=[f.__name__ for f in scenarios.ALL])
def test_scenario_on_gpu(scn):
    run(scn, GpuBackend)


def test_pool_overflow_sets_sticky_status_and_raises():
    """64-order capacity, 8 agents stacking limit bids at distinct ghost levels: the reference would
    keep growing; we flag the market (sticky) instead of corrupting it."""
    env = cda.VecCDAEnv(dict(num_of_agents=8, max_step=100000, initial_price_min=5000, initial_price_max=5000),
                        num_markets=2, order_capacity=64, status_policy="ignore")
    env.reset(seed=1)
    M, A = 2, 8
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(120):
        cat = torch.full((M, A), 2, dtype=torch.int32, device="cuda")          # bid limits only: nothing ever trades
        cat[1] = 0                                                              # market 1 only passes
        mean = torch.zeros((M, A), dtype=torch.float32, device="cuda")
        sig = torch.zeros_like(mean)
        price = torch.randint(0, 10, (M, A), device="cuda", generator=g, dtype=torch.int32)
        off = torch.randint(0, 3, (M, A), device="cuda", generator=g, dtype=torch.int32)
        env.step(cat, mean, sig, price, off)
    st = env.status().cpu().numpy()
    assert st[0] & 1 and st[1] == 0
    with pytest.raises(RuntimeError, match="order pool overflow"):
        env.check_status()
    env.close()


def test_bad_action_sets_status_not_crash():
    env = cda.VecCDAEnv(dict(num_of_agents=4), num_markets=1, status_policy="ignore")
    env.reset(seed=1)
    z = lambda v, dt: torch.full((1, 4), v, dtype=dt, device="cuda")
    env.step(z(11, torch.int32), z(0.0, torch.float32), z(0.0, torch.float32), z(3, torch.int32), z(1, torch.int32))
    assert int(env.status()[0].item()) & 4
    env.close()


def test_status_flag_makes_step_raise_without_being_asked():
    """a pool overflow must not pass silently.  Default policy: the first step() after the offending step has
    completed raises (the kernel sets a pinned flag word; polling it costs nothing while every market is clean)."""
    env = cda.VecCDAEnv(dict(num_of_agents=8, max_step=100000, initial_price_min=5000, initial_price_max=5000),
                        num_markets=2, order_capacity=64)
    env.reset(seed=1)
    M, A = 2, 8
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    with pytest.raises(RuntimeError, match="order pool overflow"):
        for t in range(200):
            cat = torch.full((M, A), 2, dtype=torch.int32, device="cuda")
            z = torch.zeros((M, A), dtype=torch.float32, device="cuda")
            env.step(cat, z, z, torch.randint(0, 10, (M, A), device="cuda", generator=g, dtype=torch.int32),
                     torch.randint(0, 3, (M, A), device="cuda", generator=g, dtype=torch.int32))
            torch.cuda.synchronize()
    env.close()


def test_fill_log_overflow_is_not_fatal_in_the_dict_adapter():
    """one aggressive order sweeping more resting orders than the fill log holds truncates the LOG only; the dict
    adapter (fill_capacity 2 here) keeps stepping and its ledger stays exact (NAV conserved)."""
    env = cda.continuousDoubleAuctionEnv(dict(num_of_agents=4, max_step=1000, initial_price_min=50, initial_price_max=50, fill_capacity=2))
    env.reset(seed=3)
    one = lambda c, m=0.0: {"category": c, "size_mean": np.array([m], np.float32), "size_sigma": np.array([0.0], np.float32), "price": 0, "price_offset": 1}
    for p in range(3):      # three makers rest one small ask each at the same ghost level, in three steps
        env.step({f"agent_{p}": one(6, 0.0)})                     # size_mean 0 -> one unit each, all at the same price
    _, _, _, _, infos = env.step({"agent_3": one(1, 1.0)})       # a market buy of 50.5 -> 51 units sweeps all three
    assert len(env.fills()) == 2 and env._vec_status & 2          # log truncated, flagged ...
    env.step({"agent_0": one(0)})                                 # ... and the env keeps going
    assert sum(int(infos[a]["NAV"]) for a in env.agents) == 4 * 1_000_000
    env.close()


def test_fill_tape_keeps_the_fills_of_every_step_of_a_fused_rollout():
    """only the last step's fills survived a launch.  With fill_tape the log is a ring across steps: after a fused 40-step
    rollout (ONE launch) plus 10 single steps it holds the market's most recent fills, equal to the oracle's per-step fills concatenated."""
    from oracle.cda_oracle import OracleEnv
    cfg = dict(num_of_agents=4, max_step=100000)
    M, cap = 32, 16
    env = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=cap, fill_tape=True)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(17)
    env.reset(seed=seeds); orc.reset(seeds=seeds)
    hist = [[] for _ in range(M)]
    def orc_steps(n):
        for _ in range(n):
            orc.rollout_random(1, policy_seed=3)
            for m in range(M):
                d = orc.dump(m)
                hist[m] += [tuple(r) for r in d["fills"][:d["n_fills"]]]
    env.rollout_random(40, policy_seed=3); orc_steps(40)
    for _ in range(10):
        env.rollout_random(1, policy_seed=3)
    orc_steps(10)
    some = 0
    for m in range(M):
        rows, total = env.tape(m)
        assert total == len(hist[m]) and [tuple(r) for r in rows.tolist()] == [tuple(int(x) for x in r) for r in hist[m][-cap:]], m
        some += total > cap
    assert some > 0                                      # the ring did wrap for some markets
    env.reset(seed=seeds)
    assert env.tape(0)[1] == 0
    env.close()


def test_rollout_random_conserves_nav_and_matches_stepwise_rng_state():
    """The fused T-step rollout is the same state machine: NAV is conserved exactly and the env RNG stream
    advances exactly as a step-by-step run with the same (device-generated) policy would."""
    cfg = dict(num_of_agents=4, max_step=100000)
    a = cda.VecCDAEnv(cfg, num_markets=64); b = cda.VecCDAEnv(cfg, num_markets=64)
    a.reset(seed=5); b.reset(seed=5)
    a.rollout_random(48, policy_seed=9)
    for _ in range(48):
        b.rollout_random(1, policy_seed=9)
    assert torch.equal(a.obs, b.obs) and torch.equal(a.reward, b.reward)
    for m in (0, 31, 63):
        da, db = a.dump(m), b.dump(m)
        assert np.array_equal(da["bids"], db["bids"]) and np.array_equal(da["asks"], db["asks"])
        assert np.array_equal(da["rng"], db["rng"]) and np.array_equal(da["accounts"], db["accounts"])
    nav = a.info("nav").sum(1)
    assert bool((nav == 4 * 1_000_000).all())
    a.close(); b.close()
