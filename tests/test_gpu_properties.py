"""Size-independent properties at BASELINE.json's full sizes (where a step-by-step oracle comparison of
every market would take too long): exact NAV conservation, never a crossed book, observation sign
conventions, determinism, and independence of the result from how markets are sharded."""
import numpy as np
import pytest
import torch

import gym_continuousdoubleauction_b200 as cda

pytestmark = pytest.mark.gpu

MIX = {"limit_market": [.10, .15, .30, 0, 0, .15, .30, 0, 0], "modify_heavy": [.05, .05, .15, .30, .05, .05, .15, .15, .05],
       "uniform": [1 / 9] * 9}


def device_actions(M, A, mix, gen):
    p = torch.tensor(MIX[mix], device="cuda")
    cat = torch.multinomial(p, M * A, replacement=True, generator=gen).to(torch.int32).view(M, A)
    return (cat, torch.rand((M, A), device="cuda", generator=gen) * 2 - 1, torch.rand((M, A), device="cuda", generator=gen),
            torch.randint(0, 10, (M, A), device="cuda", generator=gen, dtype=torch.int32),
            torch.randint(0, 3, (M, A), device="cuda", generator=gen, dtype=torch.int32))


@pytest.mark.parametrize("A,M,mix,T", [(4, 4096, "limit_market", 96), (8, 8192, "modify_heavy", 64), (4, 32768, "uniform", 24)])
def test_invariants_at_baseline_sizes(A, M, mix, T):
    init = 1_000_000
    env = cda.VecCDAEnv(dict(num_of_agents=A, init_cash=init, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    traded = torch.zeros(M, dtype=torch.bool, device="cuda")
    for t in range(T):
        obs, rew, term, trunc = env.step(*device_actions(M, A, mix, g))
        if t % 8 == 7 or t == T - 1:
            info = env.info_all()
            mk = info["market"]
            bb, ba = mk[:, 1], mk[:, 2]
            assert bool(((bb == 0) | (ba == 0) | (bb < ba)).all()), "crossed book"          # test_orderbook_crossed_book.py
            traded |= info["num_trades"].sum(1) > 0
            nav = info["nav"].sum(1)
            # NAV is marked only once a market has traded; from then on it is conserved EXACTLY
            assert bool((nav == A * init).all()), "NAV not conserved"                           # league callback :679-704
            assert bool((info["cash"] + info["cash_on_hold"] + info["position_val"] == info["nav"])[traded].all())
            assert bool((info["net_position"].sum(1) == 0).all()), "positions must net to zero"
            o = obs.view(M, env.n_hist, 42)[:, -1]
            assert bool((o[:, 0:10] >= 0).all() and (o[:, 10:20] >= 0).all() and (o[:, 20:30] <= 0).all() and (o[:, 30:40] <= 0).all())
            assert bool(torch.isfinite(obs).all() and torch.isfinite(rew).all())
            both = (bb > 0) & (ba > 0)
            assert bool((o[both, 41] >= 0.693).all()) and bool((o[~both, 41] == 0).all())      # log1p(spread>=1 tick) / sentinel
    assert int(env.status().max().item()) == 0
    env.close()


def test_same_seed_same_actions_bitwise_deterministic():
    outs = []
    for rep in range(2):
        env = cda.VecCDAEnv(dict(num_of_agents=4, max_step=1 << 30), num_markets=4096)
        env.reset(seed=77)
        g = torch.Generator(device="cuda"); g.manual_seed(9)
        for t in range(48):
            obs, rew, _, _ = env.step(*device_actions(4096, 4, "uniform", g))
        outs.append((obs.clone(), rew.clone(), env.info("nav").clone()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_result_independent_of_sharding():
    """One env of 8192 markets == two envs of 4096 with global-id seeds and sliced actions."""
    M = 8192
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [device_actions(M, 4, "limit_market", g) for _ in range(40)]
    whole = cda.VecCDAEnv(dict(num_of_agents=4, max_step=1 << 30), num_markets=M)
    whole.reset(seed=np.arange(M, dtype=np.uint64) + 1000)
    halves = [cda.VecCDAEnv(dict(num_of_agents=4, max_step=1 << 30), num_markets=M // 2) for _ in range(2)]
    for h, env in enumerate(halves):
        env.reset(seed=np.arange(h * M // 2, (h + 1) * M // 2, dtype=np.uint64) + 1000)
    for a in acts:
        ow, rw, _, _ = whole.step(*a)
        for h, env in enumerate(halves):
            sl = slice(h * M // 2, (h + 1) * M // 2)
            oh, rh, _, _ = env.step(*[x[sl].contiguous() for x in a])
            assert torch.equal(oh, ow[sl]) and torch.equal(rh, rw[sl])
    whole.close()
    for env in halves:
        env.close()
