#!/usr/bin/env python
"""Stress check of the host paths' completion protocol (run under gpurun; a script, not a pytest module).
Every warp orders its output stores before its count with a GPU-scope fence; only the last warp makes a system-scope fence before it rings
the pinned completion word.  If that word could overtake another SM's output store the host would read a stale cell: so step M markets T
times through (a) the resident step server and (b) the launch-per-step plane path, and compare EVERY cell of EVERY step with the plain
device step's outputs (read back with an ordinary synchronous copy).  Prints the number of mismatching steps (must be 0).
usage: python tests/serve_stress.py [M=4096] [T=4000]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
A, EP = 4, 400
cfg = dict(num_of_agents=A, max_step=EP)
for mode in ("resident step server", "launch per step"):
    e1 = cda.VecCDAEnv(cfg, num_markets=M); e2 = cda.VecCDAEnv(cfg, num_markets=M)
    e1.reset(seed=3); e2.reset_host_planes(seed=3)
    if mode.startswith("resident"):
        assert e2.serve(True)
    one = torch.empty((M, 5, A), dtype=torch.int32, pin_memory=True)
    bad = 0; t0 = time.time()
    for c0 in range(0, T, EP):
        acts = make_actions(100 + c0, EP, M, A, "limit_market")
        dev = [torch.from_numpy(a).cuda() for a in acts]
        blk = np.empty((EP, 5, M, A), np.int32)
        blk[:, 0], blk[:, 3], blk[:, 4] = acts[0], acts[3], acts[4]
        blk[:, 1] = acts[1].view(np.int32); blk[:, 2] = acts[2].view(np.int32)
        blk = np.ascontiguousarray(blk.transpose(0, 2, 1, 3))
        for t in range(EP):
            one.numpy()[...] = blk[t]
            o1, r1, te1, tr1 = e1.step(*[a[t] for a in dev])
            o2, r2, te2, tr2 = e2.step_host_planes(one)
            # the host reads the planes right after the completion word: no other synchronisation in between
            ok = np.array_equal(np.asarray(o2), o1.cpu().numpy()) and np.array_equal(r2, r1.cpu().numpy()) and \
                np.array_equal(te2, te1.cpu().numpy()) and np.array_equal(tr2, tr1.cpu().numpy())
            bad += not ok
        e1.reset(seed=None); e2.reset_host_planes(seed=None)
    print(f"{mode:22s}: {M} markets x {T} steps, every cell of every step compared with the device step: {bad} mismatching steps; "
          f"resident launches {e2.serve_launches}; status bits {int(e2.status().max().item())}; {time.time() - t0:.0f} s", flush=True)
    e1.close(); e2.close()
