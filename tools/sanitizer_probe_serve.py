#!/usr/bin/env python
"""Small run of the resident step server's body (<160,4,1,1,0>: poller CTA + worker warps, per-CTA action copies, claim word, completion
counting) for compute-sanitizer — served steps, an implicit retire (info / masked reset), an explicit stop, a relaunch after the lease, a
partially filled last CTA (M % 4 != 0), the volatile-load action path (A % 4 != 0).
usage: CDA_SERVE_LEASE_US=200000 compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_probe_serve.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

for A, M in ((4, 50), (3, 21)):
    T = 12
    acts = make_actions(3, T, M, A, "uniform")
    pin = torch.empty((T, M, 5, A), dtype=torch.int32, pin_memory=True)
    for f in (0, 3, 4): pin[:, :, f].copy_(torch.from_numpy(acts[f]))
    for f in (1, 2): pin[:, :, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1000), num_markets=M, fill_capacity=16)
    env.reset_host_planes(seed=5)
    assert env.serve(True)
    for t in range(T):
        if t == 4: env.info_all()
        if t == 6: env.reset_host_planes(seed=None, mask=(np.arange(M) % 2).astype(np.uint8))
        if t == 8: env.serve_stop()
        if t == 10: time.sleep(0.5)
        env.step_host_planes(pin[t])
    env.fills(); env.dump(3)
    print(f"A={A} M={M}: {T} served steps, resident launches {env.serve_launches}, status {int(env.status().max().item())}")
    env.close()
print("sanitizer probe (resident step server) done")
