#!/usr/bin/env python
"""One-GPU sweep for profiles/: the other BASELINE workloads, a large-M sweep (bandwidth/issue asymptote)
and the fused random-policy rollout.  Run under gpurun; writes gpurun_out/sweep.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

def b_alg(A): return 1986 + 188 * A
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"step_kernel": [], "rollout": []}

def run(A, M, mix, steps=60, prewarm=200, P=64):
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    # prewarm with the fused random rollout is not the same mix; use real batches, generated on device for big M
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    p = torch.tensor({"uniform": [1/9]*9, "limit_market": [.10,.15,.30,0,0,.15,.30,0,0], "modify_heavy": [.05,.05,.15,.30,.05,.05,.15,.15,.05]}[mix], device="cuda")
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
    st = int(env.status().max().item())
    r = dict(agents=A, markets=M, mix=mix, ms_per_step=ms, steps_per_s=M / (ms * 1e-3), hot_ms=hot, hot_steps_per_s=M / (hot * 1e-3),
             achieved_gbs=b_alg(A) * M / (ms * 1e-3) / 1e9, frac=b_alg(A) * M / (ms * 1e-3) / 1e9 / peak, status=st, cap=env.order_capacity)
    print(r, flush=True)
    env.close()
    return r

for A, M, mix in [(4, 1024, "uniform"), (4, 4096, "limit_market"), (8, 8192, "modify_heavy"), (4, 8192, "limit_market"),
                  (4, 16384, "limit_market"), (4, 32768, "limit_market"), (4, 131072, "limit_market"), (4, 524288, "limit_market")]:
    out["step_kernel"].append(run(A, M, mix, steps=40 if M > 100000 else 60, prewarm=200))

# fused random-policy rollout (CDA_rand / RandomRLModule workload): T steps per launch, state stays on chip
for A, M, T in [(4, 4096, 64), (4, 4096, 256), (4, 32768, 64), (8, 8192, 64)]:
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    env.rollout_random(128, policy_seed=1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); reps = 5
    for r in range(reps): env.rollout_random(T, policy_seed=2 + r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rr = dict(agents=A, markets=M, steps_per_launch=T, ms_per_launch=ms, steps_per_s=M * T / (ms * 1e-3), status=int(env.status().max().item()))
    print(rr, flush=True); out["rollout"].append(rr)
    env.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
