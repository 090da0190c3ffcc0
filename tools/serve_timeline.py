#!/usr/bin/env python
"""Where a step of the RESIDENT STEP SERVER spends its time (run under gpurun).
Host side: wall time of VecCDAEnv.step_host_planes per step (serve on), next to the launch-per-step plane path on the same books.
Device side (cda_debug_serve_timeline): per market, the globaltimer of message seen / actions here / step computed / outputs fenced for
the LAST step, plus the poller's read of the host message and the completion ring.
$CDA_SERVE_DEBUG: 1 = actions read from a device copy (no input transfer), 2 = outputs kept on the device, 3 = both.
usage: python tools/serve_timeline.py [M=4096] [steps=300]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gym_continuousdoubleauction_b200 as cda  # noqa: E402
from gym_continuousdoubleauction_b200 import _native  # noqa: E402
from gym_continuousdoubleauction_b200.workloads import make_actions  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
A, mix, PRE = 4, "limit_market", 276
L = _native.lib()
env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
seeds = np.arange(M, dtype=np.uint64) + np.uint64(1000)
acts = make_actions(7, PRE + T, M, A, mix)
dev = [torch.from_numpy(a).cuda() for a in acts]
blk = np.empty((T, 5, M, A), np.int32)
blk[:, 0], blk[:, 3], blk[:, 4] = acts[0][PRE:], acts[3][PRE:], acts[4][PRE:]
blk[:, 1] = acts[1][PRE:].view(np.int32); blk[:, 2] = acts[2][PRE:].view(np.int32)
pin = torch.from_numpy(np.ascontiguousarray(blk.transpose(0, 2, 1, 3))).pin_memory()


REWRITE = int(os.environ.get("ST_REWRITE", "0"))   # 1: the host REWRITES one pinned block before every step (what a policy does) instead of using pre-built blocks
one = torch.empty((M, 5, A), dtype=torch.int32, pin_memory=True)


def run(serve):
    env.reset(seed=seeds)
    for t in range(PRE):
        env.step(*[a[t] for a in dev])
    env.attach_host_planes()
    if serve:
        assert env.serve(True)
    for t in range(20):
        env.step_host_planes(pin[t])
    dt = np.empty(T - 20)
    for t in range(20, T):
        b = pin[t]
        if REWRITE:
            one.copy_(pin[t]); b = one
        t0 = time.perf_counter()
        o, r, te, tr = env.step_host_planes(b)
        _ = float(r[0, 0]) + float(o[M - 1, env.W - 1])
        dt[t - 20] = time.perf_counter() - t0
    if serve:
        env.serve(False)
    return dt * 1e6


buf = L.cda_debug_serve_timeline(M)
d_srv = run(True)
cudart = C.CDLL("libcudart.so.12")
raw = (C.c_uint64 * (16 * (M + 1)))()
cudart.cudaMemcpy(raw, C.c_void_p(buf), C.c_size_t((M + 1) * 128), 2)
L.cda_debug_serve_timeline(0)
d_lps = run(False)
pr = lambda n, d: print(f"{n:42s} mean {d.mean():6.1f}  p10 {np.percentile(d, 10):6.1f}  p50 {np.percentile(d, 50):6.1f}  p90 {np.percentile(d, 90):6.1f} us per step (host wall, L2-hot)")
print(f"M={M} A={A} {mix}, ST_REWRITE={REWRITE}, CDA_SERVE_DEBUG={os.environ.get('CDA_SERVE_DEBUG', '0')}, resident launches {env.serve_launches}")
pr("resident step server", d_srv)
pr("launch per step (cda_step_planes)", d_lps)
a = np.frombuffer(raw, dtype=np.uint64).reshape(M + 1, 16).astype(np.float64)
t0 = a[M, 1]                       # poller read the host's message
q = lambda x: "min %6.2f  p10 %6.2f  p50 %6.2f  p90 %6.2f  max %6.2f" % tuple(np.percentile((x - t0) / 1e3, [0, 10, 50, 90, 100]))
print("device timeline of the last step, us after the poller read the message (globaltimer, 0.25-us ticks):")
print("  message seen by the warp   ", q(a[:M, 0]))
print("  action record here         ", q(a[:M, 1]))
print("  step computed              ", q(a[:M, 2]))
print("  outputs stored and fenced  ", q(a[:M, 3]))
print("  completion word rung        %6.2f" % ((a[M, 0] - t0) / 1e3))
print("  per-warp: wait for actions mean %.2f, compute mean %.2f max %.2f, store+fence mean %.2f max %.2f us" % (
    ((a[:M, 1] - a[:M, 0]) / 1e3).mean(), ((a[:M, 2] - a[:M, 1]) / 1e3).mean(), ((a[:M, 2] - a[:M, 1]) / 1e3).max(),
    ((a[:M, 3] - a[:M, 2]) / 1e3).mean(), ((a[:M, 3] - a[:M, 2]) / 1e3).max()))
print("  host wall of that step %.1f us  => host <-> device hand-shakes + Python: %.1f us" % (d_srv[-1], d_srv[-1] - (a[M, 0] - t0) / 1e3))
print("  slowest 5 steps (host wall):", np.sort(d_srv)[-5:].round(1), " steps above 1.3 x median:", int((d_srv > 1.3 * np.median(d_srv)).sum()), "of", len(d_srv),
      " at", np.nonzero(d_srv > 1.3 * np.median(d_srv))[0][:40])
