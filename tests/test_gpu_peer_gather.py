"""Fused step + all-gather over NVLink peer memory (needs >= 2 GPUs; skipped on a 1-GPU box).
Every rank's gather buffer must equal the NCCL all-gather of the per-rank outputs, and the sharded
result must equal a single-process run over all markets (global-id seeding)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, M, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import gym_continuousdoubleauction_b200 as cda
    from gym_continuousdoubleauction_b200.workloads import make_actions
    from gym_continuousdoubleauction_b200.sharding import shard_seeds, shard_slice
    cfg = dict(num_of_agents=4, max_step=40)       # truncation flags switch on inside the run
    env = cda.VecCDAEnv(cfg, num_markets=M, device=rank)
    ref = cda.VecCDAEnv(cfg, num_markets=M, device=rank)
    seeds = shard_seeds(1000, world * M, rank, world)
    env.reset(seed=seeds); ref.reset(seed=seeds)
    o0 = env.enable_peer_gather()
    o_all = torch.empty((world * M, env.W), dtype=torch.float32, device="cuda")
    r_all = torch.empty((world * M, 4), dtype=torch.float64, device="cuda")
    te_all = torch.empty(world * M, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(o_all, ref.obs)
    ok = bool(torch.equal(o_all, o0))               # the published initial stacks
    acts = make_actions(3, T, world * M, 4, "uniform")
    for t in range(T):                              # T > 32: the gather windows restart at least once
        a = [torch.from_numpy(np.ascontiguousarray(shard_slice(x[t], rank, world))).cuda() for x in acts]
        g_obs, g_rew, g_term, g_trunc = env.step_gather(*a)      # flag wait included: no NCCL call, no barrier
        o, r, te, tr = ref.step(*a)
        dist.all_gather_into_tensor(o_all, o); dist.all_gather_into_tensor(r_all, r); dist.all_gather_into_tensor(te_all, tr)
        ok &= bool(torch.equal(o_all, g_obs)) and bool(torch.equal(r_all, g_rew)) and bool(torch.equal(te_all, g_trunc))
        if t % 3 == rank % 3:
            torch.cuda.synchronize()                # ranks drift apart on purpose: a rank may be one step ahead of its peers' readers
    if rank == 0:
        q.put((ok, g_obs.cpu().numpy(), g_rew.cpu().numpy()))
    dist.barrier()
    env.close(); ref.close()
    dist.destroy_process_group()


WORLD = int(os.environ.get("CDA_TEST_WORLD", "0")) or min(8, max(2, torch.cuda.device_count()))   # every GPU of the box (2, 4 or 8)


@pytest.mark.skipif(torch.cuda.device_count() < 2 or torch.cuda.device_count() < WORLD, reason="needs >= 2 GPUs (CDA_TEST_WORLD of them)")
def test_peer_gather_equals_nccl_allgather_and_single_process():
    world, M, T = WORLD, 256, 45
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(k, world, port, M, T, q)) for k in range(world)]
    for p in procs:
        p.start()
    ok, obs_all, rew_all = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
    import gym_continuousdoubleauction_b200 as cda
    from gym_continuousdoubleauction_b200.workloads import make_actions
    one = cda.VecCDAEnv(dict(num_of_agents=4, max_step=40), num_markets=world * M, device=0)
    one.reset(seed=np.arange(world * M, dtype=np.uint64) + 1000)
    acts = make_actions(3, T, world * M, 4, "uniform")
    for t in range(T):
        o, r, _, _ = one.step(*[torch.from_numpy(np.ascontiguousarray(x[t])).cuda() for x in acts])
    assert np.array_equal(o.cpu().numpy(), obs_all) and np.array_equal(r.cpu().numpy(), rew_all)
    one.close()
