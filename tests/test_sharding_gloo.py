"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): shard ranges, global-id seeding,
action slicing and the observation all-gather reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_continuousdoubleauction_b200.sharding import all_gather_rows, shard_range, shard_seeds, shard_slice


def test_shard_ranges_partition():
    for M in (0, 1, 7, 4096, 32768):
        for G in (1, 2, 3, 8):
            r = [shard_range(M, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    assert np.array_equal(np.concatenate([shard_seeds(1000, 10, k, 3) for k in range(3)]), np.arange(10, dtype=np.uint64) + 1000)


def _worker(rank, world, port, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the oracle stands in for the GPU env here (CPU test of the sharding logic only)
    from oracle.cda_oracle import OracleEnv
    from gym_continuousdoubleauction_b200.workloads import make_actions
    cfg = dict(num_of_agents=4, init_cash=1_000_000, max_step=1000, n_hist=4)
    lo, hi = shard_range(M, rank, world)
    env = OracleEnv(cfg, hi - lo)
    env.reset(seeds=shard_seeds(1000, M, rank, world))
    acts = make_actions(7, 12, M, 4, "uniform")          # replicated action tensor, sliced locally
    for t in range(12):
        o, r, _, _ = env.step(*[shard_slice(a[t], rank, world) for a in acts])
    obs_all = all_gather_rows(torch.from_numpy(o.copy()), M)
    rew_all = all_gather_rows(torch.from_numpy(r.copy()), M)
    if rank == 0:
        q.put((obs_all.numpy(), rew_all.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,M", [(2, 16), (3, 10)])
def test_sharded_run_equals_single_process(world, M):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(k, world, port, M, q)) for k in range(world)]
    for p in procs:
        p.start()
    obs_all, rew_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.cda_oracle import OracleEnv
    from gym_continuousdoubleauction_b200.workloads import make_actions
    ref = OracleEnv(dict(num_of_agents=4, init_cash=1_000_000, max_step=1000, n_hist=4), M)
    ref.reset(seeds=np.arange(M, dtype=np.uint64) + 1000)
    acts = make_actions(7, 12, M, 4, "uniform")
    for t in range(12):
        o, r, _, _ = ref.step(*[a[t] for a in acts])
    assert np.array_equal(obs_all, o) and np.array_equal(rew_all, r)
