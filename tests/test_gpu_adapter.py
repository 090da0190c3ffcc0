"""The reference-compatible dict adapter (continuousDoubleAuctionEnv surface) on the GPU path.
Modelled on the reference's own tests: test_info_dict.py (keys, terms sum exactly to reward, NAV
string), test_seeding.py (same seed + same actions => identical), test_env_lifecycle.py
(truncation lands on max_step), test_observation_history.py (all agents share one obs)."""
import json
import os
from decimal import Decimal

import numpy as np
import pytest

import gym_continuousdoubleauction_b200 as cda

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

INFO_KEYS = {"reward", "NAV", "num_trades", "net_position", "VWAP", "cash", "cash_on_hold", "position_val",
             "drawdown", "max_nav", "num_trades_step", "num_passive_fills_step", "order_step_placed",
             "num_rejected_step", "is_pass_action", "reward_terms", "last_price", "best_bid", "best_ask", "spread",
             "model_action"}


def acts_at(g, t, A):
    d = {}
    for i in range(A):
        if g["cat"][t, i] < 0:
            continue
        d[f"agent_{i}"] = {"category": int(g["cat"][t, i]), "size_mean": np.array([g["mean"][t, i]], np.float32),
                           "size_sigma": np.array([g["sigma"][t, i]], np.float32), "price": int(g["price"][t, i]),
                           "price_offset": int(g["off"][t, i])}
    return d


@pytest.mark.parametrize("name", ["uniform_a4", "modify_heavy_a8", "nhist2_a5_absent", "appendix_d"])
def test_dict_api_reproduces_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
    cfg = {str(k): (int(v) if float(v).is_integer() else float(v)) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    env = cda.continuousDoubleAuctionEnv(cfg)
    A = env.num_of_agents
    obs, infos = env.reset(seed=int(g["seed"]))
    assert set(obs) == set(env.agents) and all(v == {} for v in infos.values())
    assert np.array_equal(obs["agent_0"], g["obs0"])
    for t in range(g["cat"].shape[0]):
        o, r, te, tr, info = env.step(acts_at(g, t, A))
        assert o["agent_0"] is o[f"agent_{A - 1}"]                         # one shared array
        assert o["agent_0"].dtype == np.float32 and o["agent_0"].shape == (env.n_hist * 42,)
        assert np.abs(o["agent_0"].astype(np.float64) - g["obs"][t]).max() <= 1e-6
        assert np.abs(np.array([r[a] for a in env.agents]) - g["reward"][t]).max() <= 1e-6
        assert te["__all__"] == bool(g["terminated"][t]) and tr["__all__"] == bool(g["truncated"][t])
        assert all(te[a] is False and tr[a] is False for a in env.agents)
        for i, a in enumerate(env.agents):
            inf = info[a]
            assert set(inf) - {"model_action"} == INFO_KEYS - {"model_action"}
            acc = g["accounts"][t][i]
            assert Decimal(inf["NAV"]) == int(acc[4]) and inf["net_position"] == int(acc[7])
            assert inf["cash"] == float(acc[0]) and inf["cash_on_hold"] == float(acc[1]) and inf["position_val"] == float(acc[2])
            assert inf["num_trades"] == int(acc[8]) and inf["num_trades_step"] == int(acc[9])
            assert inf["is_pass_action"] == bool(acc[13])
            s = 0.0
            for v in inf["reward_terms"].values():                          # naive left-to-right sum, exactly
                s += v
            assert s == inf["reward"] == r[a]
            json.dumps(inf)                                                 # JSON-safe
        assert info["agent_0"]["last_price"] == float(g["scalars"][t][2])
    fills = env.fills()
    f0, f1 = int(g["fill_ptr"][-2]), int(g["fill_ptr"][-1])
    assert np.array_equal(fills, g["fills"][f0:f1])
    env.close()


def test_nav_conservation_and_random_driver():
    from gym_continuousdoubleauction_b200.cda_rand import run_random
    out = run_random(num_agents=4, max_step=200, init_cash=1_000_000, seed=3)
    assert out["steps"] == 200
    assert out["total_nav"] == 4 * 1_000_000      # SelfPlayCallback's |sum NAV - init_cash*A| <= 1e-6 check, exactly


def test_same_seed_same_actions_identical_and_seed_none_differs():
    def run(seed_second):
        env = cda.continuousDoubleAuctionEnv({"num_of_agents": 4, "max_step": 50})
        rng = np.random.default_rng(0)
        env.reset(seed=123)
        outs = []
        for ep in range(2):
            if ep == 1:
                env.reset(seed=seed_second)
            for t in range(30):
                a = {f"agent_{i}": {"category": int(rng.integers(0, 9)), "size_mean": rng.uniform(-1, 1, 1).astype(np.float32),
                                    "size_sigma": rng.uniform(0, 1, 1).astype(np.float32), "price": int(rng.integers(0, 10)),
                                    "price_offset": int(rng.integers(0, 3))} for i in range(4)}
                o, r, *_ = env.step(a)
                outs.append((o["agent_0"].copy(), dict(r)))
        env.close()
        return outs
    a, b, c = run(123), run(123), run(None)
    assert all(np.array_equal(x[0], y[0]) and x[1] == y[1] for x, y in zip(a, b))
    assert not all(np.array_equal(x[0], y[0]) for x, y in zip(a[30:], c[30:]))   # seed=None keeps the stream


@pytest.mark.parametrize("max_step", [1, 2, 5, 10, 64])
def test_truncation_lands_exactly_on_max_step(max_step):
    env = cda.continuousDoubleAuctionEnv({"num_of_agents": 2, "max_step": max_step})
    env.reset(seed=1)
    n = 0
    while True:
        _, _, te, tr, _ = env.step({"agent_0": {"category": 0, "size_mean": np.zeros(1, np.float32), "size_sigma": np.zeros(1, np.float32)}})
        n += 1
        if tr["__all__"]:
            break
    assert n == max_step
    env.close()


def test_spaces_and_attributes():
    env = cda.continuousDoubleAuctionEnv({"num_of_agents": 3, "n_hist": 6})
    assert env.agents == env.possible_agents == ["agent_0", "agent_1", "agent_2"]
    assert env.get_observation_space("agent_1").shape == (6 * 42,)
    s = env.get_action_space("agent_0").sample()
    assert set(s) == {"category", "size_mean", "size_sigma", "price", "price_offset"}
    assert env.num_of_agents == 3 and env.init_cash == 1000000 and env.n_hist == 6
    with pytest.raises(ValueError):
        cda.continuousDoubleAuctionEnv({"tick_size": 0.5})
    env.close()


def test_vector_dict_env_matches_one_object_per_market():
    """VectorCDAEnv (num_envs markets, one launch per step, lazy infos) == num_envs continuousDoubleAuctionEnv objects:
    observations, rewards, flags and every info key, through partial action dicts, a per-sub-env reset and truncation."""
    cfg = dict(num_of_agents=4, init_cash=20_000, max_step=25, n_hist=4)
    M, T = 3, 45
    vec = cda.VectorCDAEnv(cfg, num_envs=M)
    singles = [cda.continuousDoubleAuctionEnv(cfg) for _ in range(M)]
    ov, iv = vec.reset(seed=100)
    for m in range(M):
        o1, i1 = singles[m].reset(seed=100 + m)
        assert np.array_equal(ov[m]["agent_0"], o1["agent_0"]) and iv[m] == i1
        assert ov[m]["agent_0"] is ov[m]["agent_3"]                      # one array shared by the agents of a market
    rng = np.random.default_rng(3)
    space = singles[0].action_spaces["agent_0"]; space.seed(5)
    stale = None
    for t in range(T):
        acts = []
        for m in range(M):
            d = {}
            for i in range(4):
                if rng.random() < 0.15:
                    continue                                             # partial dict: this agent is absent
                d[f"agent_{i}"] = space.sample()
            acts.append(d)
        if t == 20:                                                      # episode boundary of sub-env 1 only
            o_a, _ = vec.reset_at(1)
            o_b, _ = singles[1].reset(seed=None)
            assert np.array_equal(o_a["agent_0"], o_b["agent_0"])
        o, r, te, tr, inf = vec.step(acts)
        for m in range(M):
            o1, r1, te1, tr1, inf1 = singles[m].step(acts[m])
            assert np.array_equal(o[m]["agent_0"], o1["agent_0"]), (t, m)
            assert r[m] == r1 and te[m] == te1 and tr[m] == tr1, (t, m)
            if t % 6 == 0 or tr1["__all__"]:
                assert set(inf[m]) == set(inf1)
                for a in inf1:
                    assert dict(inf[m])[a] == inf1[a], (t, m, a)
        if t == 3:
            stale = inf[0]
    with pytest.raises(RuntimeError):
        stale["agent_0"]                                                 # lazy info of an old step: refused, not silently wrong
    vec.close()
    for s in singles:
        s.close()


def test_lob_actions_traders_view_and_np_random_snapshot_match_the_oracle():
    """The attributes the reference's tests / callbacks read off the env object: LOB_actions (decoded actions of the last step, in dict
    order, pass actions filtered out), traders[i].acc.<field>, min_tick, and np_random — a snapshot generator positioned exactly where the
    device stream is (its next draws are what the oracle's stream draws next)."""
    from oracle.cda_oracle import OracleEnv
    from gym_continuousdoubleauction_b200.workloads import make_actions
    cfg = dict(num_of_agents=5, init_cash=50_000, max_step=500, tick_size=2, initial_price_min=40, initial_price_max=90, is_render=False)
    env = cda.continuousDoubleAuctionEnv(cfg)
    orc = OracleEnv({k: v for k, v in cfg.items() if k != "is_render"}, 1)
    env.reset(seed=77); orc.reset(seeds=[77])
    assert env.min_tick == 2 and env.LOB_actions is None
    acts = make_actions(5, 60, 1, 5, "uniform")
    types, sides = ("market", "limit", "modify", "cancel"), ("bid", "ask")
    for t in range(60):
        d = {f"agent_{i}": {"category": int(acts[0][t, 0, i]), "size_mean": np.array([acts[1][t, 0, i]], np.float32),
                            "size_sigma": np.array([acts[2][t, 0, i]], np.float32), "price": int(acts[3][t, 0, i]), "price_offset": int(acts[4][t, 0, i])}
             for i in range(5) if (t + i) % 7}                                              # some agents absent
        cat = np.array([[int(acts[0][t, 0, i]) if (t + i) % 7 else -1 for i in range(5)]], np.int32)
        env.step(d)
        orc.step(cat, acts[1][t], acts[2][t], acts[3][t], acts[4][t])
        la = orc.last_actions(0)
        want = [{"ID": f"agent_{i}", "side": sides[la[i, 1]], "type": types[la[i, 0]], "size": int(la[i, 2]), "price": -1.0 if la[i, 0] == 0 else float(la[i, 3])}
                for i in range(5) if la[i, 1] >= 0]
        assert env.LOB_actions == want, (t, env.LOB_actions, want)
    acc = orc.dump(0)["accounts"]
    for i in range(5):
        a = env.traders[i].acc
        assert (a.cash, a.cash_on_hold, a.nav, a.net_position, a.num_trades) == tuple(int(acc[i, j]) for j in (0, 1, 4, 7, 8))
        assert str(a.nav) == env.step({})[4][f"agent_{i}"]["NAV"] if i == 0 else True
    orc.step(np.full((1, 5), -1, np.int32), acts[1][0], acts[2][0], acts[3][0], acts[4][0])   # (the env.step({}) above: nobody acts, no draw)
    g = env.np_random
    st = orc.dump(0)["rng"]
    assert g.bit_generator.state["state"]["state"] == (int(st[0]) << 64) | int(st[1])
    z = g.standard_normal(3)                                                                  # drawing from the snapshot does not move the env
    assert env.np_random.bit_generator.state["state"]["state"] == (int(st[0]) << 64) | int(st[1]) and z.shape == (3,)
    env.close()
