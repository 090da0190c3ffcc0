#!/usr/bin/env python
"""Device-side cost of each piece of the host (end-to-end) path: per-step device time of back-to-back window
steps WITHOUT host synchronisation (CUDA events around 200 steps), with the zero-copy input / zero-copy reward /
copy-engine transfer switched on and off; then the same with a host sync per step.  Run under gpurun."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions
M, A = 4096, 4
slots = int(os.environ.get("WSLOTS", "64"))
cda.VecCDAEnv.WINDOW_SLOTS = slots
env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
env.reset(seed=1000)
acts = make_actions(7, 300, M, A, "limit_market")
dev = [torch.from_numpy(a).cuda() for a in acts]
pin = torch.empty((300, 5, M, A), dtype=torch.int32, pin_memory=True)
for f in (0, 3, 4): pin[:, f].copy_(torch.from_numpy(acts[f]))
for f in (1, 2): pin[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
for i in range(256): env.step(*[d[i] for d in dev])
env.reset_host_window(seed=None)
blocks = [pin[i] for i in range(300)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(30): env.step_host_window(blocks[i], sync=False)
torch.cuda.synchronize(); e0.record()
for i in range(200): env.step_host_window(blocks[(30 + i) %% 300], sync=False)
e1.record(); torch.cuda.synchronize()
dev_us = e0.elapsed_time(e1) * 5
t0 = time.perf_counter()
for i in range(200): env.step_host_window(blocks[(230 + i) %% 300], sync=True)
host_us = (time.perf_counter() - t0) / 200 * 1e6
e0.record()
for i in range(200): env.step(*[d[i] for d in dev])
e1.record(); torch.cuda.synchronize()
print("%%-46s device %%6.1f us/step   host-synced %%6.1f us/step   (plain device step %%5.1f us)" %% (os.environ.get("TAG"), dev_us, host_us, e0.elapsed_time(e1) * 5))
''' % ROOT

for tag, env in (("full path, 16 slots", dict(WSLOTS="16")), ("full path, 32 slots", dict(WSLOTS="32")), ("full path, 64 slots", dict(WSLOTS="64")),
                 ("no input, no output transfer (launch+kernel+sync)", dict(CDA_DEBUG_WINDOW="3"))):
    e = dict(os.environ); e.update(env); e["TAG"] = tag; e.setdefault("WSLOTS", "16")
    subprocess.run([sys.executable, "-c", CHILD], env=e)
