"""GPU parity tests: the CUDA env (through the C-ABI) against the CPU oracle on identical seeds
and action sequences.  Integer state (book, order-map order, ledger, fills, RNG stream) must be
BIT-EXACT; observations (f32) and rewards (f64) within 1e-6 (they are normally bit-equal too).
"""
import os

import numpy as np
import pytest
import torch

from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal

import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

pytestmark = pytest.mark.gpu

OBS_TOL = 1e-6   # north-star tolerance for observations (float32)
REW_TOL = 1e-6   # north-star tolerance for rewards (float64)


def run_pair(cfg, M, T, mix, seed, dump_every=25, dump_markets=(0, 1, 2), absent_p=0.0, order_capacity=0):
    A = cfg["num_of_agents"]
    env = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=64, order_capacity=order_capacity)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(seed)
    o_g = env.reset(seed=seeds).cpu().numpy()
    o_c = orc.reset(seeds=seeds)
    assert np.array_equal(o_g, o_c), "reset observation differs"
    acts = make_actions(seed + 99, T, M, A, mix)
    if absent_p > 0:
        rng = np.random.default_rng(seed + 5)
        acts[0][rng.random(acts[0].shape) < absent_p] = -1
    max_obs = max_rew = 0.0
    for t in range(T):
        step = [torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts]
        og, rg, teg, trg = env.step(*step)
        oc, rc, tec, trc = orc.step(*[a[t] for a in acts], nthreads=4)
        og, rg = og.cpu().numpy(), rg.cpu().numpy()
        d_obs = np.abs(og.astype(np.float64) - oc.astype(np.float64)).max()
        d_rew = np.abs(rg - rc).max()
        max_obs, max_rew = max(max_obs, d_obs), max(max_rew, d_rew)
        assert d_obs <= OBS_TOL, f"step {t}: obs differs by {d_obs}"
        assert d_rew <= REW_TOL, f"step {t}: reward differs by {d_rew}"
        assert np.array_equal(teg.cpu().numpy(), tec) and np.array_equal(trg.cpu().numpy(), trc), f"step {t}: flags"
        if t % dump_every == 0 or t == T - 1:
            for m in dump_markets:
                if m < M:
                    assert_dump_equal(env.dump(m), orc.dump(m), ctx=f"t={t} m={m}")
    st = env.status().cpu().numpy()
    assert (st == 0).all(), f"sticky status bits set: {np.unique(st)}"
    env.close()
    return max_obs, max_rew


def base_cfg(**kw):
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=10_000, n_hist=4)
    cfg.update(kw)
    return cfg


def test_config2_uniform_4x1024_bit_exact():
    """BASELINE config 2 shape: 4 agents x 1024 markets, uniform random actions; every market's
    final book/ledger/fills/RNG compared bit-exactly at the end, 3 markets every 25 steps."""
    cfg = base_cfg()
    M, T = 1024, 96
    env = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=64)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(1000)
    env.reset(seed=seeds); orc.reset(seeds=seeds)
    acts = make_actions(7, T, M, 4, "uniform")
    for t in range(T):
        og, rg, _, _ = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        oc, rc, _, _ = orc.step(*[a[t] for a in acts], nthreads=8)
    assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= OBS_TOL
    assert np.abs(rg.cpu().numpy() - rc).max() <= REW_TOL
    for m in range(M):
        assert_dump_equal(env.dump(m), orc.dump(m), ctx=f"m={m}")
    env.close()


def _full_size(A, M, T, mix, seed):
    """BASELINE-size run: obs / reward / flags compared at EVERY step for EVERY market, and every market's book, order-map order,
    ledger, fill log and RNG state at three points of the run (oracle on all host threads; GPU side parsed from one checkpoint)."""
    cfg = base_cfg(num_of_agents=A)
    env = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=32)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(seed)
    assert np.array_equal(env.reset(seed=seeds).cpu().numpy(), orc.reset(seeds=seeds))
    acts = make_actions(seed + 1, T, M, A, mix)
    nthreads = os.cpu_count() or 8
    for t in range(T):
        og, rg, teg, trg = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        oc, rc, tec, trc = orc.step(*[a[t] for a in acts], nthreads=nthreads)
        assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= OBS_TOL, f"t={t}"
        assert np.abs(rg.cpu().numpy() - rc).max() <= REW_TOL, f"t={t}"
        assert np.array_equal(teg.cpu().numpy(), tec) and np.array_equal(trg.cpu().numpy(), trc), f"t={t}"
        if t in (T // 3, 2 * T // 3, T - 1):
            dumps = env.dump_all()
            for m in range(M):
                assert_dump_equal(dumps[m], orc.dump(m), ctx=f"t={t} m={m}")
    assert (env.status().cpu().numpy() == 0).all()
    env.close()


def test_config3_full_size_4x4096_limit_market_every_market():
    """BASELINE config 3 at its full size (it was only checked at 256 markets)."""
    _full_size(4, 4096, 96, "limit_market", 1000)


def test_config4_full_size_8x8192_modify_heavy_every_market():
    """BASELINE config 4 at its full size."""
    _full_size(8, 8192, 64, "modify_heavy", 3000)


def test_dump_all_equals_dump():
    """dump_all() (host-side parse of one checkpoint through cda_state_layout) == dump() (per-field device gathers)."""
    cfg = base_cfg()
    env = cda.VecCDAEnv(cfg, num_markets=24, fill_capacity=32)
    env.reset(seed=77)
    acts = make_actions(5, 60, 24, 4, "uniform")
    for t in range(60):
        env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
    dumps = env.dump_all()
    for m in range(24):
        assert_dump_equal(dumps[m], env.dump(m), ctx=f"m={m}")
        assert dumps[m]["status"] == 0
    env.close()


@pytest.mark.parametrize("A,M,T", [(4, 1024, 200), (8, 256, 120), (3, 64, 90)])
def test_fused_random_rollout_equals_oracle_policy_twin(A, M, T):
    """cda_rollout_random (T steps in ONE launch, on-device uniform policy) against the oracle's twin of that policy stream
    (oracle/cda_oracle.c orc_rollout_random): last-step obs / reward / flags and every market's book, ledger and RNG state."""
    cfg = base_cfg(num_of_agents=A)
    env = cda.VecCDAEnv(cfg, num_markets=M)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(4242)
    env.reset(seed=seeds); orc.reset(seeds=seeds)
    done = 0
    for chunk in (1, 7, T - 8):                                   # launches of different lengths continue one trajectory
        og, rg, teg, trg = env.rollout_random(chunk, policy_seed=99)
        oc, rc, tec, trc = orc.rollout_random(chunk, policy_seed=99, nthreads=os.cpu_count() or 8)
        done += chunk
        assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= OBS_TOL, f"after {done} steps"
        assert np.abs(rg.cpu().numpy() - rc).max() <= REW_TOL
        assert np.array_equal(teg.cpu().numpy(), tec) and np.array_equal(trg.cpu().numpy(), trc)
        dumps = env.dump_all()
        for m in range(M):
            assert_dump_equal(dumps[m], orc.dump(m), ctx=f"after {done} steps, m={m}", fills=False)
    env.close()


def test_uniform_small_every_step_dump():
    run_pair(base_cfg(), M=8, T=150, mix="uniform", seed=11, dump_every=1, dump_markets=range(8))


def test_config3_limit_market_mix():
    run_pair(base_cfg(), M=256, T=128, mix="limit_market", seed=21)


def test_config4_modify_heavy_8_agents():
    run_pair(base_cfg(num_of_agents=8), M=128, T=160, mix="modify_heavy", seed=31)


def test_low_cash_rejections_and_bankruptcy():
    run_pair(base_cfg(init_cash=3000), M=64, T=200, mix="uniform", seed=41, dump_every=10)


@pytest.mark.parametrize("n_hist", [1, 2, 6, 10])
def test_n_hist_variants(n_hist):
    run_pair(base_cfg(n_hist=n_hist, num_of_agents=3), M=16, T=40, mix="limit_market", seed=50 + n_hist, dump_every=10)


def test_absent_agents_partial_action_dicts():
    run_pair(base_cfg(num_of_agents=5), M=32, T=120, mix="uniform", seed=61, absent_p=0.3, dump_every=10)


def test_16_agents_and_capacity_256():
    run_pair(base_cfg(num_of_agents=16), M=16, T=100, mix="uniform", seed=71, dump_every=20, order_capacity=256)


def test_truncation_flag_lands_on_max_step():
    cfg = base_cfg(max_step=5)
    env = cda.VecCDAEnv(cfg, num_markets=4)
    env.reset(seed=3)
    acts = make_actions(1, 6, 4, 4, "uniform")
    flags = []
    for t in range(6):
        _, _, _, tr = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        flags.append(int(tr[0].item()))
    assert flags == [0, 0, 0, 0, 1, 1]
    env.close()


def test_reset_seed_none_keeps_stream_and_masked_reset():
    cfg = base_cfg()
    M = 8
    env = cda.VecCDAEnv(cfg, num_markets=M)
    orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(500)
    env.reset(seed=seeds); orc.reset(seeds=seeds)
    acts = make_actions(3, 40, M, 4, "uniform")
    for t in range(20):
        env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        orc.step(*[a[t] for a in acts])
    mask = np.array([1, 0, 1, 0, 0, 1, 0, 0], np.uint8)
    og = env.reset(seed=None, mask=mask).cpu().numpy()
    oc = orc.reset(seeds=None, mask=mask)
    assert np.array_equal(og[mask == 1], oc[mask == 1])
    for t in range(20, 40):
        og, rg, _, _ = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        oc, rc, _, _ = orc.step(*[a[t] for a in acts])
    assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= OBS_TOL
    for m in range(M):
        assert_dump_equal(env.dump(m), orc.dump(m), ctx=f"m={m}", fills=False)
    env.close()


def test_step_host_matches_device_step():
    cfg = base_cfg()
    M = 64
    e1 = cda.VecCDAEnv(cfg, num_markets=M)
    e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e1.reset(seed=9); e2.reset(seed=9)
    acts = make_actions(4, 30, M, 4, "uniform")
    for t in range(30):
        o1, r1, _, _ = e1.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        o2, r2, _, _ = e2.step_host(*[a[t] for a in acts])
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
    e1.close(); e2.close()


def test_checkpoint_roundtrip():
    cfg = base_cfg()
    M = 16
    env = cda.VecCDAEnv(cfg, num_markets=M)
    env.reset(seed=77)
    acts = make_actions(8, 40, M, 4, "uniform")
    dev = lambda t: [torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts]
    for t in range(20):
        env.step(*dev(t))
    sd = env.state_dict()
    ref = [env.step(*dev(t))[0].clone() for t in range(20, 40)]
    env.load_state_dict(sd)
    for i, t in enumerate(range(20, 40)):
        assert torch.equal(env.step(*dev(t))[0], ref[i])
    env.close()


def test_host_ring_view_equals_stacked_obs_including_masked_reset():
    """cda_step_host_ring ships only the newest snapshot (mirrored); the strided ring view must equal the
    stacked observation of the ordinary host path bit for bit, also across per-market resets."""
    cfg = base_cfg(n_hist=4)
    M = 96
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    o1 = e1.reset(seed=11).cpu().numpy(); o2 = e2.reset_host_ring(seed=11)
    assert np.array_equal(o1, o2)
    acts = make_actions(6, 40, M, 4, "uniform")
    blk = torch.empty((40, 5, M, 4), dtype=torch.int32, pin_memory=True)
    for f in (0, 3, 4):
        blk[:, f].copy_(torch.from_numpy(acts[f]))
    for f in (1, 2):
        blk[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
    for t in range(40):
        if t == 17:
            mask = (np.arange(M) % 3 == 0).astype(np.uint8)
            a = e1.reset(seed=None, mask=mask).cpu().numpy(); b = e2.reset_host_ring(seed=None, mask=mask)
            assert np.array_equal(a[mask == 1], b[mask == 1])
        o1, r1, te1, tr1 = e1.step_host_block(blk[t])
        o2, r2, te2, tr2 = e2.step_host_ring(blk[t])
        assert o2.shape == (M, 168) and np.array_equal(o1, o2), f"t={t}"
        assert np.array_equal(r1, r2) and np.array_equal(te1, te2) and np.array_equal(tr1, tr2)
    e1.close(); e2.close()


@pytest.mark.parametrize("n_hist", [1, 3])
def test_host_ring_other_history_depths(n_hist):
    cfg = base_cfg(n_hist=n_hist)
    M = 32
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e1.reset(seed=5); e2.reset_host_ring(seed=5)
    acts = make_actions(2, 12, M, 4, "limit_market")
    blk = torch.empty((12, 5, M, 4), dtype=torch.int32, pin_memory=True)
    for f in (0, 3, 4):
        blk[:, f].copy_(torch.from_numpy(acts[f]))
    for f in (1, 2):
        blk[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
    for t in range(12):
        o1 = e1.step_host_block(blk[t])[0]; o2 = e2.step_host_ring(blk[t])[0]
        assert np.array_equal(o1, o2)
    e1.close(); e2.close()


def _pinned_block(acts, T, M, A):
    blk = torch.empty((T, 5, M, A), dtype=torch.int32, pin_memory=True)
    for f in (0, 3, 4):
        blk[:, f].copy_(torch.from_numpy(acts[f]))
    for f in (1, 2):
        blk[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
    return blk


@pytest.mark.parametrize("n_hist,T,A,market_major,zerocopy", [
    (4, 100, 4, False, True), (4, 100, 4, True, True), (1, 70, 4, True, True), (6, 75, 4, False, True),
    (4, 70, 8, True, True),      # 8 agents: the record (2A+2 = 18 words) still rides behind the snapshot
    (4, 40, 24, True, True),     # 24 agents: the record does not fit into a slot -> separate record array; A % 4 == 0 -> TMA-staged actions
    (4, 40, 6, True, True),      # A % 4 != 0: plain action loads from the market-major block
    (4, 70, 4, True, False),     # staged fallback (CDA_ZEROCOPY=0): copies instead of kernel stores into host memory
])
def test_host_window_view_equals_stacked_obs_across_wraps_and_masked_reset(n_hist, T, A, market_major, zerocopy, monkeypatch):
    """cda_step_window ships only the newest snapshot into the next slot of a 32-slot pinned window per market (and the
    result record right behind it); the view of the n_hist latest slots must equal the ordinary host path's stacked
    observation bit for bit, and the record views the ordinary reward / flags — across window restarts (T > 32 steps),
    per-market resets and truncation (max_step = 50), for both action-block layouts."""
    if not zerocopy:
        monkeypatch.setenv("CDA_ZEROCOPY", "0")
    cfg = base_cfg(n_hist=n_hist, num_of_agents=A, max_step=50)
    M = 96
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    o1 = e1.reset(seed=11).cpu().numpy(); o2 = e2.reset_host_window(seed=11)
    assert o2.shape == (M, n_hist * 42) and np.array_equal(o1, o2)
    blk = _pinned_block(make_actions(6, T, M, A, "uniform"), T, M, A)
    blk_mm = blk.permute(0, 2, 1, 3).contiguous().pin_memory()     # [T, M, 5, A]
    seen_trunc = False
    for t in range(T):
        if t in (7, 28, 29, 60):
            mask = (np.arange(M) % 3 == t % 3).astype(np.uint8)
            a = e1.reset(seed=None, mask=mask).cpu().numpy(); b = e2.reset_host_window(seed=None, mask=mask)
            assert np.array_equal(a[mask == 1], b[mask == 1])
            if t > 0:
                assert np.array_equal(prev[mask == 0], b[mask == 0])     # untouched markets keep their stack
        o1, r1, te1, tr1 = e1.step_host_block(blk[t])
        o2, r2, te2, tr2 = e2.step_host_window(blk_mm[t] if market_major else blk[t], market_major=market_major)
        assert np.array_equal(o1, o2), f"t={t}"
        assert r2.shape == (M, A) and np.array_equal(r1, r2), f"t={t}"
        assert np.array_equal(te1, te2) and np.array_equal(tr1, tr2), f"t={t}"
        seen_trunc |= bool(tr2.any())
        prev = o1.copy()
    assert seen_trunc or T < 50
    assert (e2.status().cpu().numpy() == 0).all()
    e1.close(); e2.close()


@pytest.mark.parametrize("n_hist,A,market_major,cell", [(4, 4, True, 64), (4, 4, False, 0), (1, 4, True, 0), (6, 8, True, 64), (3, 24, True, 0), (4, 6, True, 0), (4, 5, True, 0)])
def test_host_planes_equal_stacked_obs_across_ring_wraps_and_masked_reset(n_hist, A, market_major, cell):
    """cda_step_planes stores, per step, every market's newest snapshot + result record into ONE dense plane of a pinned ring;
    the n_hist most recent planes must equal the ordinary host path's stacked observation bit for bit (np.asarray(obs), obs[m],
    obs[m, e]) and the record views its reward / flags — across ring wraps, per-market resets and truncation."""
    cfg = base_cfg(n_hist=n_hist, num_of_agents=A, max_step=50)
    M, T = 96, 70
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e2.PLANE_SLOTS, e2.PLANE_CELL_WORDS = max(8, n_hist + 2), cell
    o1 = e1.reset(seed=11).cpu().numpy(); o2 = e2.reset_host_planes(seed=11)
    assert o2.shape == (M, n_hist * 42) and np.array_equal(o1, np.asarray(o2))
    blk = _pinned_block(make_actions(6, T, M, A, "uniform"), T, M, A)
    blk_mm = blk.permute(0, 2, 1, 3).contiguous().pin_memory()
    seen_trunc = False
    for t in range(T):
        if t in (7, 28, 29, 60):
            mask = (np.arange(M) % 3 == t % 3).astype(np.uint8)
            a = e1.reset(seed=None, mask=mask).cpu().numpy(); b = np.asarray(e2.reset_host_planes(seed=None, mask=mask))
            assert np.array_equal(a[mask == 1], b[mask == 1])
            if t > 0:
                assert np.array_equal(prev[mask == 0], b[mask == 0])
        o1, r1, te1, tr1 = e1.step_host_block(blk[t])
        o2, r2, te2, tr2 = e2.step_host_planes(blk_mm[t] if market_major else blk[t], market_major=market_major)
        assert np.array_equal(o1, np.asarray(o2)), f"t={t}"
        assert np.array_equal(o1[5], o2[5]) and o1[M - 1, n_hist * 42 - 1] == o2[M - 1, n_hist * 42 - 1] and o1[3, 0] == o2[3, 0]
        assert r2.shape == (M, A) and np.array_equal(r1, r2) and np.array_equal(te1, te2) and np.array_equal(tr1, tr2), f"t={t}"
        seen_trunc |= bool(tr2.any())
        prev = o1.copy()
    assert seen_trunc
    assert (e2.status().cpu().numpy() == 0).all()
    e1.close(); e2.close()


def test_host_window_needs_reset_first():
    env = cda.VecCDAEnv(base_cfg(), num_markets=8)
    blk = torch.zeros((5, 8, 4), dtype=torch.int32, pin_memory=True)
    with pytest.raises(RuntimeError):
        env.step_host_window(blk)
    env.close()


@pytest.mark.parametrize("case", range(12))
def test_random_configurations_gpu_vs_oracle(case):
    """The configuration fuzz of tests/test_oracle_vs_reference.py::test_random_configurations, CUDA env against the oracle."""
    rng = np.random.default_rng(9000 + case)
    A = int(rng.integers(1, 9))
    lo = int(rng.choice([1, 3, 10, 250, 3000]))
    cfg = base_cfg(
        num_of_agents=A, tick_size=int(rng.choice([1, 1, 2, 5])), n_hist=int(rng.integers(1, 7)),
        min_size=int(rng.integers(1, 4)), mkt_max_size=int(rng.choice([10, 40, 100])), limit_size_multiple=int(rng.choice([1, 3, 10])),
        init_cash=int(rng.choice([2_000, 50_000, 1_000_000, 80_000_000])),
        initial_price_min=lo, initial_price_max=lo + int(rng.integers(0, 60)),
        order_penalty=float(rng.choice([0.0, 0.1, 0.7])), trade_penalty=float(rng.choice([0.0, 0.05])),
        drawdown_penalty=float(rng.choice([0.0, 0.2, 1.0])), passive_bonus=float(rng.choice([0.0, 0.1])),
        loss_multiplier=float(rng.choice([1.0, 1.5, 3.0])), max_step=95)
    mix = str(rng.choice(["uniform", "limit_market", "modify_heavy"]))
    run_pair(cfg, M=48, T=90, mix=mix, seed=9100 + case, dump_every=15, absent_p=float(rng.choice([0.0, 0.0, 0.2])))
