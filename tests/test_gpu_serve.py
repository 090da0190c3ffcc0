"""Resident step server (cda_serve_*, VecCDAEnv.serve): the step kernel stays on the SMs with every market's book and ledger in shared
memory and is stepped by doorbell messages from the host.  It must produce exactly what the launch-per-step plane path produces — same
planes, same records, same state afterwards — through idle-lease relaunches, masked resets, info / dump calls in between (which retire
the kernel implicitly) and an action block that is REWRITTEN IN PLACE every step (a stale read of host memory would show here)."""
import time

import numpy as np
import pytest
import torch

from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal

import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

pytestmark = pytest.mark.gpu


def base_cfg(**kw):
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=10_000, n_hist=4)
    cfg.update(kw)
    return cfg


def mm_blocks(acts, T, M, A):
    """[T][M][5][A] int32: one market-major action block per step (float fields as bits)."""
    b = np.empty((T, 5, M, A), np.int32)
    b[:, 0], b[:, 3], b[:, 4] = acts[0][:T], acts[3][:T], acts[4][:T]
    b[:, 1] = acts[1][:T].view(np.int32); b[:, 2] = acts[2][:T].view(np.int32)
    return np.ascontiguousarray(b.transpose(0, 2, 1, 3))


@pytest.mark.parametrize("A,n_hist,M,reuse_block,mix", [(4, 4, 257, True, "limit_market"), (8, 2, 96, False, "modify_heavy"), (3, 4, 64, True, "uniform"),
                                                        (4, 1, 33, False, "uniform")])
def test_resident_server_equals_launch_per_step_planes(A, n_hist, M, reuse_block, mix):
    cfg = base_cfg(num_of_agents=A, n_hist=n_hist, max_step=50)
    T = 80
    e1 = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=64); e2 = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=64)
    o1 = e1.reset_host_planes(seed=21); o2 = e2.reset_host_planes(seed=21)
    assert e2.serve(True), "the resident server must be available for this shape"
    assert np.array_equal(np.asarray(o1), np.asarray(o2))
    blocks = mm_blocks(make_actions(8, T, M, A, mix), T, M, A)
    pin_all = torch.from_numpy(blocks).pin_memory()                 # e1 (and e2 unless reuse_block): a distinct pinned block per step
    pin_one = torch.empty((M, 5, A), dtype=torch.int32, pin_memory=True)
    for t in range(T):
        if t in (9, 30, 31, 62):       # per-market resets in between: retire the kernel, reset, relaunch on the next step
            mask = (np.arange(M) % 3 == t % 3).astype(np.uint8)
            a = np.asarray(e1.reset_host_planes(seed=None, mask=mask)); b = np.asarray(e2.reset_host_planes(seed=None, mask=mask))
            assert np.array_equal(a, b), f"t={t}: stacks after a masked reset"
        if t == 20:
            time.sleep(0.02)           # longer than the idle lease: the kernel has given the SMs back and is launched again
        if t == 41:                    # a call that needs the state in HBM
            assert torch.equal(e1.info_all()["nav"], e2.info_all()["nav"])
        if t == 55:
            (f1, c1), (f2, c2) = e1.fills(), e2.fills()
            assert torch.equal(c1, c2)
            for m in range(M):
                assert torch.equal(f1[m, :int(c1[m])], f2[m, :int(c2[m])]), f"fills of market {m}"

        if reuse_block:
            pin_one.numpy()[...] = blocks[t]
        r1 = e1.step_host_planes(pin_all[t])
        r2 = e2.step_host_planes(pin_one if reuse_block else pin_all[t])
        assert np.array_equal(np.asarray(r1[0]), np.asarray(r2[0])), f"t={t}: obs"
        assert np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2]) and np.array_equal(r1[3], r2[3]), f"t={t}: record"
    assert e2.serve_launches >= 6, e2.serve_launches      # first launch + 4 resets + lease + info + fills
    assert e1.serve_launches == 0
    d1, d2 = e1.dump_all(), e2.dump_all()
    for m in range(M):
        assert_dump_equal(d1[m], d2[m], ctx=f"m={m}")
    assert (e2.status().cpu().numpy() == 0).all()
    e1.close(); e2.close()


def test_resident_server_equals_oracle(monkeypatch):
    monkeypatch.setenv("CDA_SERVE_LEASE_US", "500000")   # (read when the env binds its server) a busy test box must not retire the kernel mid-loop
    cfg = base_cfg(max_step=10_000)
    M, T, A = 128, 96, 4
    env = cda.VecCDAEnv(cfg, num_markets=M); orc = OracleEnv(cfg, M)
    seeds = np.arange(M, dtype=np.uint64) + np.uint64(77)
    o = env.reset_host_planes(seed=seeds); oc = orc.reset(seeds=seeds)
    assert env.serve(True)
    assert np.array_equal(np.asarray(o), oc)
    acts = make_actions(3, T, M, A, "limit_market")
    pin = torch.from_numpy(mm_blocks(acts, T, M, A)).pin_memory()
    outs = []
    for t in range(T):                # back to back (the oracle runs afterwards: a pause longer than the lease would retire the kernel)
        og, rg, teg, trg = env.step_host_planes(pin[t])
        outs.append((np.asarray(og), rg.copy(), teg.copy(), trg.copy()))
    assert env.serve_launches == 1, "96 back-to-back steps are served by ONE launch"
    for t in range(T):
        og, rg, teg, trg = outs[t]
        oc, rc, tec, trc = orc.step(*[a[t] for a in acts], nthreads=4)
        assert np.abs(og.astype(np.float64) - oc).max() <= 1e-6 and np.abs(rg - rc).max() <= 1e-6, t
        assert np.array_equal(teg, tec) and np.array_equal(trg, trc), t
    dumps = env.dump_all()
    for m in range(M):
        assert_dump_equal(dumps[m], orc.dump(m), ctx=f"m={m}", fills=False)
    env.close()


def test_resident_server_declines_what_it_cannot_serve():
    e = cda.VecCDAEnv(base_cfg(), num_markets=16, decimal_ledger=True)
    e.reset_host_planes(seed=1)
    assert e.serve(True) is False                      # decimal ledger: launch path
    blk = torch.zeros((16, 5, 4), dtype=torch.int32, pin_memory=True)
    e.step_host_planes(blk)
    e.close()
    big = cda.VecCDAEnv(base_cfg(), num_markets=8192)  # two waves: does not fit one resident grid
    big.reset_host_planes(seed=1)
    assert big.serve(True) is False
    big.step_host_planes(torch.zeros((8192, 5, 4), dtype=torch.int32, pin_memory=True))
    big.close()
    ok = cda.VecCDAEnv(base_cfg(), num_markets=4096)   # BASELINE cfg3: exactly one wave
    ok.reset_host_planes(seed=1)
    assert ok.serve(True) is True
    ok.step_host_planes(torch.zeros((4096, 5, 4), dtype=torch.int32, pin_memory=True))
    assert ok.serve(False) is False
    ok.close()


def test_resident_server_full_size_every_step_matches_the_device_path():
    """BASELINE cfg3 shape (4 x 4096: the one-wave grid the server is built for): 400 served steps, every market's planes and records
    compared with the plain device step's outputs at every step (a completion word that overtook a late output store would show here)."""
    cfg = base_cfg(max_step=150)
    M, A, T = 4096, 4, 400
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e1.reset(seed=5); e2.reset_host_planes(seed=5)
    assert e2.serve(True)
    acts = make_actions(12, T, M, A, "limit_market")
    dev = [torch.from_numpy(a).cuda() for a in acts]
    pin = torch.from_numpy(mm_blocks(acts, T, M, A)).pin_memory()
    one = torch.empty((M, 5, A), dtype=torch.int32, pin_memory=True)
    for t in range(T):
        if t % 150 == 149:      # episodes end by truncation: reset everything (retires and relaunches the server)
            o1, r1, te1, tr1 = e1.step(*[a[t] for a in dev]); one.copy_(pin[t]); o2, r2, te2, tr2 = e2.step_host_planes(one)
            assert tr2.all() and np.array_equal(o1.cpu().numpy(), np.asarray(o2))
            e1.reset(seed=None); e2.reset_host_planes(seed=None)
            continue
        one.copy_(pin[t])
        o1, r1, te1, tr1 = e1.step(*[a[t] for a in dev])
        o2, r2, te2, tr2 = e2.step_host_planes(one)
        assert np.array_equal(o1.cpu().numpy(), np.asarray(o2)), t
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(te1.cpu().numpy(), te2) and np.array_equal(tr1.cpu().numpy(), tr2), t
    assert e2.serve_launches >= 3, e2.serve_launches      # (more when the host was held up for longer than the lease between two steps)
    e1.close(); e2.close()


def test_resident_server_switches_itself_off_when_it_is_relaunched_for_most_steps(monkeypatch):
    """A lease shorter than the host's time between two steps: every step needs a relaunch.  After a window of 64 such steps the server
    declines (CDA_EUNSUPPORTED, nothing stepped), step_host_planes continues on the launch path, and the results stay those of the
    launch path throughout."""
    monkeypatch.setenv("CDA_SERVE_LEASE_US", "1")
    cfg = base_cfg()
    M, A, T = 64, 4, 90
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e1.reset_host_planes(seed=9); e2.reset_host_planes(seed=9)
    assert e2.serve(True)
    pin = torch.from_numpy(mm_blocks(make_actions(4, T, M, A, "uniform"), T, M, A)).pin_memory()
    for t in range(T):
        time.sleep(0.0005)
        r1 = e1.step_host_planes(pin[t]); r2 = e2.step_host_planes(pin[t])
        assert np.array_equal(np.asarray(r1[0]), np.asarray(r2[0])) and np.array_equal(r1[1], r2[1]), t
    assert e2.host_resident is False and 48 <= e2.serve_launches <= 64, e2.serve_launches
    e1.close(); e2.close()
