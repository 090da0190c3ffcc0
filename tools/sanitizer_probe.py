#!/usr/bin/env python
"""Small runs of every shipped step-kernel body for compute-sanitizer (memcheck / racecheck):
   <160,4,0,0> device step (ledger off and on, with Decimal ties being resolved), <160,4,0,1> routed step (host window, host planes with the
   completion doorbell, full-stack host block, fused gather with world = 1), <160,4,1,0> fused rollout; plus reset / info / twin replay kernels.
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitizer_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

M, A, T = 48, 4, 14
acts = make_actions(3, T, M, A, "uniform")
dev = [torch.from_numpy(a).cuda() for a in acts]
pin = torch.empty((T, M, 5, A), dtype=torch.int32, pin_memory=True)
for f in (0, 3, 4): pin[:, :, f].copy_(torch.from_numpy(acts[f]))
for f in (1, 2): pin[:, :, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
fm = pin.permute(0, 2, 1, 3).contiguous().pin_memory()

env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1000), num_markets=M, fill_capacity=16)
env.reset(seed=5)
for t in range(T): env.step(*[d[t] for d in dev])
env.info_all(); env.fills(); env.dump(3)
env.rollout_random(9, policy_seed=2)
env.reset_host_window(seed=None, mask=np.ones(M, np.uint8))
for t in range(T): env.step_host_window(pin[t], market_major=True)
for t in range(4): env.step_host_window(fm[t], market_major=False)
env.attach_host_planes()
for t in range(T): env.step_host_planes(pin[t])
for t in range(3): env.step_host_block(fm[t])
env.reset_host_ring(seed=7)
for t in range(3): env.step_host_ring(fm[t])
torch.cuda.synchronize(); env.close()

# decimal ledger at low cash: journal appends, replay kernel, tie resolution passes (deterministic re-execution)
low = dict(num_of_agents=7, init_cash=3000, max_step=255, n_hist=2, tick_size=3, min_size=1, mkt_max_size=10, limit_size_multiple=3, initial_price_min=3, initial_price_max=21)
e2 = cda.VecCDAEnv(low, num_markets=M, decimal_ledger=True)
e2.reset(seed=np.arange(M, dtype=np.uint64) * 1000 + 61018)
a2 = make_actions(11, 120, M, 7, "modify_heavy")
for t in range(120): e2.step(*[torch.from_numpy(np.ascontiguousarray(x[t])).cuda() for x in a2])
e2.rollout_random(25, policy_seed=4)
e2.decimal_fields([0, 1])
print("tie-resolution passes:", e2._L.cda_debug_restart_count())
torch.cuda.synchronize(); e2.close()

# fused gather epilogue with a one-rank group (replicated-output code path, completion flag, wait kernel)
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("gloo", rank=0, world_size=1)
e3 = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1000), num_markets=M)
e3.reset(seed=9); e3.enable_peer_gather()
for t in range(T): e3.step_gather(*[d[t] for d in dev])
torch.cuda.synchronize(); e3.close(); dist.destroy_process_group()
print("sanitizer probe done")
