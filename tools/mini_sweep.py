#!/usr/bin/env python
"""Throughput of several prebuilt library variants (tools/bin/libcda_<name>.so) on three workloads: the one-wave headline
(4 x 4096), BASELINE config 4 (8 x 8192 modify-heavy) and a many-wave case (4 x 32768).  Run under gpurun:
    python tools/mini_sweep.py r01m r02b head"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(A, M, mix, steps=50, prewarm=150):
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    p = torch.tensor({"limit_market": [.10,.15,.30,0,0,.15,.30,0,0], "modify_heavy": [.05,.05,.15,.30,.05,.05,.15,.15,.05]}[mix], device="cuda")
    def batch():
        cat = torch.multinomial(p, M * A, replacement=True, generator=g).to(torch.int32).view(M, A)
        return (cat, torch.rand((M, A), device="cuda", generator=g) * 2 - 1, torch.rand((M, A), device="cuda", generator=g),
                torch.randint(0, 10, (M, A), device="cuda", generator=g, dtype=torch.int32), torch.randint(0, 3, (M, A), device="cuda", generator=g, dtype=torch.int32))
    for _ in range(prewarm): env.step(*batch())
    bs = [batch() for _ in range(steps)]
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 255); ev[i][0].record(); env.step(*bs[i]); ev[i][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): env.step(*bs[i])
    e1.record(); torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) / steps
    st = int(env.status().max().item()); env.close()
    return "%%dx%%d %%s: %%6.1f us (%%5.1f M/s) hot %%6.1f us (%%5.1f M/s) st %%d" %% (A, M, mix[:6], ms * 1e3, M / ms / 1e3, hot * 1e3, M / hot / 1e3, st)
def roll(A, M, T):
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    env.rollout_random(128, policy_seed=1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(4): env.rollout_random(T, policy_seed=2 + r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    st = int(env.status().max().item()); env.close()
    return "rollout %%dx%%d T=%%d: %%5.1f M/s st %%d" %% (A, M, T, M * T / ms / 1e3, st)
print("%%-14s" %% os.environ["TAG"], " | ".join(run(*w) for w in ((4, 4096, "limit_market"), (8, 8192, "modify_heavy"), (4, 32768, "limit_market"))),
      "|", roll(4, 4096, 64), "|", roll(4, 32768, 64), flush=True)
''' % ROOT
for spec in sys.argv[1:]:          # name[:ENV=VAL,ENV=VAL]
    name, _, envs = spec.partition(":")
    lib = os.path.join(ROOT, "tools", "bin", f"libcda_{name}.so")
    extra = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CDA_B200_LIB=lib, TAG=spec[:14], **extra))
