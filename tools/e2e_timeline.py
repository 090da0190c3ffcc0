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
env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M, status_policy="ignore")   # (the no-input mode replays one action block: books overflow)
env.reset(seed=1000)
acts = make_actions(7, 300, M, A, "limit_market")
dev = [torch.from_numpy(a).cuda() for a in acts]
pin = torch.empty((300, 5, M, A), dtype=torch.int32, pin_memory=True)
for f in (0, 3, 4): pin[:, f].copy_(torch.from_numpy(acts[f]))
for f in (1, 2): pin[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
for i in range(256): env.step(*[d[i] for d in dev])
env.reset_host_window(seed=None)
mm = os.environ.get("MM", "1") == "1"          # market-major action block (one bulk copy per CTA)
if os.environ.get("INLINE", "1") != "1":       # timing only: record in the separate record array (two more stores per market)
    env._win_inline = False; env._win_last = slots - 1
    for q in range(slots): env._win_views[q] = env._win_views[q] or env._win_views[env.n_hist - 1]
if mm:
    pin = pin.permute(0, 2, 1, 3).contiguous().pin_memory()
blocks = [pin[i] for i in range(300)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(30): env.step_host_window(blocks[i], sync=False, market_major=mm)
torch.cuda.synchronize(); e0.record()
for i in range(200): env.step_host_window(blocks[(30 + i) %% 300], sync=False, market_major=mm)
e1.record(); torch.cuda.synchronize()
dev_us = e0.elapsed_time(e1) * 5
t0 = time.perf_counter()
for i in range(200): env.step_host_window(blocks[(230 + i) %% 300], sync=True, market_major=mm)
host_us = (time.perf_counter() - t0) / 200 * 1e6
e0.record()
for i in range(200): env.step(*[d[i] for d in dev])
e1.record(); torch.cuda.synchronize()
print("%%-46s device %%6.1f us/step   host-synced %%6.1f us/step   (plain device step %%5.1f us)" %% (os.environ.get("TAG"), dev_us, host_us, e0.elapsed_time(e1) * 5))
''' % ROOT

for tag, env in () if __name__ != "__main__" else (
        ("32 slots, field-major actions, separate record", dict(WSLOTS="32", MM="0", INLINE="0")),
        ("32 slots, market-major actions, separate record", dict(WSLOTS="32", MM="1", INLINE="0")),
        ("32 slots, field-major actions, inline record", dict(WSLOTS="32", MM="0", INLINE="1")),
        ("32 slots, market-major actions, inline record", dict(WSLOTS="32")), ("same, 64 slots", dict(WSLOTS="64")),
        ("no input transfer", dict(WSLOTS="32", CDA_DEBUG_WINDOW="1")), ("no output transfer", dict(WSLOTS="32", CDA_DEBUG_WINDOW="2")),
        ("no input, no output transfer (launch+kernel+sync)", dict(CDA_DEBUG_WINDOW="3"))):
    e = dict(os.environ); e.update(env); e["TAG"] = tag; e.setdefault("WSLOTS", "16")
    subprocess.run([sys.executable, "-c", CHILD], env=e)
