#!/usr/bin/env python
"""Per-phase cycle breakdown of the step kernel (debug build with -DCDA_PROFILE_PHASES).
Run under gpurun.  Prints mean cycles per warp per phase at steady state."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_continuousdoubleauction_b200 import _native  # noqa: E402

flags = sys.argv[1].split() if len(sys.argv) > 1 else []
subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + ["-DCDA_PROFILE_PHASES"] + flags +
                      ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC, "-o", _native.SO_PATH,
                       os.path.join(_native.CSRC, "cda_b200.cu")])
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gym_continuousdoubleauction_b200 as cda  # noqa: E402
from gym_continuousdoubleauction_b200.workloads import make_actions  # noqa: E402

A, M, mix = 4, int(os.environ.get("PP_M", 4096)), os.environ.get("PP_MIX", "limit_market")
env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
env.reset(seed=1000)
T = 300
acts = [torch.from_numpy(a).cuda() for a in make_actions(7, T, M, A, mix)]
L = _native.lib()
buf = L.cda_debug_phase_buffer()
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
for t in range(T - 20):
    env.step(*[a[t] for a in acts])
torch.cuda.synchronize()
ctypes.cdll.LoadLibrary("libcudart.so.12") if False else None
import ctypes as C
cudart = C.CDLL("libcudart.so.12")
cudart.cudaMemset(C.c_void_p(buf), 0, C.c_size_t(M * 128))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(T - 20, T):
    env.step(*[a[t] for a in acts])
e1.record(); torch.cuda.synchronize()
raw = (C.c_uint64 * (16 * M))()
cudart.cudaMemcpy(raw, C.c_void_p(buf), C.c_size_t(M * 128), 2)
arr = np.frombuffer(raw, dtype=np.uint64).reshape(M, 16).astype(np.float64)
host = arr.sum(0)
per_mkt = arr[:, :12].sum(1) / 20
print('per-market total cycles/step: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f' % (per_mkt.mean(), np.percentile(per_mkt, 50), np.percentile(per_mkt, 90), np.percentile(per_mkt, 99), per_mkt.max()))
names = ["header load", "decode (after draws)", "shuffle", "wait pool tiles", "do_actions", "mtm+topK", "obs math", "obs+ring write", "reward/done", "state store"]
tot = sum(host[:12])
names += ["wait actions (+accts issue)", "normal draws"]
print(f"M={M} mix={mix}: {e0.elapsed_time(e1)/20*1e3:.1f} us per step (instrumented), mean cycles per warp-step = {tot/(20*M):.0f}")
for i, n in enumerate(names):
    print(f"  {n:30s} {host[i]/(20*M):9.0f} cyc  {100*host[i]/tot:5.1f}%")
subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC,
                       "-o", _native.SO_PATH, os.path.join(_native.CSRC, "cda_b200.cu")])
