"""Pins the C oracle to the UNMODIFIED Python reference on identical seeds and actions
(only where /root/reference exists, i.e. this container)."""
import numpy as np
import pytest

from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal

pytestmark = pytest.mark.needs_reference

MIXES = {
    "uniform": np.full(9, 1 / 9), "limit_market": np.array([.10, .15, .30, 0, 0, .15, .30, 0, 0]),
    "modify_heavy": np.array([.05, .05, .15, .30, .05, .05, .15, .15, .05]),
}


def gen(rng, T, A, mix, absent=0.0):
    p = MIXES[mix]
    c = rng.choice(9, size=(T, A), p=p / p.sum()).astype(np.int32)
    if absent:
        c[rng.random((T, A)) < absent] = -1
    return (c, rng.uniform(-1, 1, (T, A)).astype(np.float32), rng.uniform(0, 1, (T, A)).astype(np.float32),
            rng.integers(0, 10, (T, A)).astype(np.int32), rng.integers(0, 3, (T, A)).astype(np.int32))


def run(seed, A, T, mix, extra=None, absent=0.0):
    from oracle.ref_runner import ReferenceMarket
    cfg = dict(num_of_agents=A, init_cash=1_000_000, max_step=T + 5, n_hist=4)
    cfg.update(extra or {})
    ref, orc = ReferenceMarket(cfg), OracleEnv(cfg, 1)
    assert np.array_equal(ref.reset(seed=seed), orc.reset(seeds=[seed])[0])
    acts = gen(np.random.default_rng(seed + 7), T, A, mix, absent)
    for t in range(T):
        ro, rr, rte, rtr = ref.step(*[x[t] for x in acts])
        oo, orw, ote, otr = orc.step(*[x[t][None] for x in acts])
        assert np.abs(ro.astype(np.float64) - oo[0]).max() <= 1e-6
        assert np.abs(rr - orw[0]).max() <= 1e-6
        assert rte == bool(ote[0]) and rtr == bool(otr[0])
        a, b = ref.dump(), orc.dump(0)
        assert_dump_equal(a, b, ctx=f"seed={seed} t={t}")
        assert np.allclose(a["reward_terms"], b["reward_terms"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("A,mix", [(4, "uniform"), (4, "limit_market"), (8, "modify_heavy")])
def test_random_trajectories(seed, A, mix):
    run(seed, A, 200, mix)


def test_low_cash_bankruptcy_and_rejections():
    run(100, 4, 250, "uniform", dict(init_cash=3000))


def test_partial_action_dicts():
    run(7, 5, 150, "uniform", absent=0.3)


@pytest.mark.parametrize("n_hist", [1, 2, 6])
def test_n_hist(n_hist):
    run(300 + n_hist, 3, 80, "limit_market", dict(n_hist=n_hist))


def test_reset_seed_none_continues_stream():
    from oracle.ref_runner import ReferenceMarket
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=1000, n_hist=4)
    ref, orc = ReferenceMarket(cfg), OracleEnv(cfg, 1)
    ref.reset(seed=5); orc.reset(seeds=[5])
    acts = gen(np.random.default_rng(1), 40, 4, "uniform")
    for t in range(20):
        ref.step(*[x[t] for x in acts]); orc.step(*[x[t][None] for x in acts])
    assert np.array_equal(ref.reset(seed=None), orc.reset(seeds=None)[0])
    for t in range(20, 40):
        ro, rr, _, _ = ref.step(*[x[t] for x in acts]); oo, orw, _, _ = orc.step(*[x[t][None] for x in acts])
    assert np.array_equal(ro, oo[0])
    assert_dump_equal(ref.dump(), orc.dump(0))


@pytest.mark.parametrize("case", range(12))
def test_random_configurations(case):
    """Configuration fuzz: agents, tick, size limits, cash, anchor range, history depth, reward coefficients and the
    action mix are all drawn at random (seeded); the oracle must track the reference bit for bit on each."""
    rng = np.random.default_rng(9000 + case)
    A = int(rng.integers(1, 9))
    lo = int(rng.choice([1, 3, 10, 250, 3000]))
    extra = dict(
        tick_size=int(rng.choice([1, 1, 2, 5])), n_hist=int(rng.integers(1, 7)),
        min_size=int(rng.integers(1, 4)), mkt_max_size=int(rng.choice([10, 40, 100])), limit_size_multiple=int(rng.choice([1, 3, 10])),
        init_cash=int(rng.choice([2_000, 50_000, 1_000_000, 80_000_000])),
        initial_price_min=lo, initial_price_max=lo + int(rng.integers(0, 60)),
        order_penalty=float(rng.choice([0.0, 0.1, 0.7])), trade_penalty=float(rng.choice([0.0, 0.05])),
        drawdown_penalty=float(rng.choice([0.0, 0.2, 1.0])), passive_bonus=float(rng.choice([0.0, 0.1])),
        loss_multiplier=float(rng.choice([1.0, 1.5, 3.0])))
    mix = str(rng.choice(["uniform", "limit_market", "modify_heavy"]))
    run(9100 + case, A, 90, mix, extra, absent=float(rng.choice([0.0, 0.0, 0.2])))
