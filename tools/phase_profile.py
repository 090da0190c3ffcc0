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
NS = int(os.environ.get('PP_STEPS', 20))
for t in range(T - NS):
    env.step(*[a[t] for a in acts])
torch.cuda.synchronize()
ctypes.cdll.LoadLibrary("libcudart.so.12") if False else None
import ctypes as C
cudart = C.CDLL("libcudart.so.12")
cudart.cudaMemset(C.c_void_p(buf), 0, C.c_size_t(M * 128))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(T - NS, T):
    env.step(*[a[t] for a in acts])
e1.record(); torch.cuda.synchronize()
raw = (C.c_uint64 * (16 * M))()
cudart.cudaMemcpy(raw, C.c_void_p(buf), C.c_size_t(M * 128), 2)
arr = np.frombuffer(raw, dtype=np.uint64).reshape(M, 16).astype(np.float64)
host = arr.sum(0)
per_mkt = arr[:, :12].sum(1) / NS
print('per-market total cycles/step: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f' % (per_mkt.mean(), np.percentile(per_mkt, 50), np.percentile(per_mkt, 90), np.percentile(per_mkt, 99), per_mkt.max()))
st, en = arr[:, 12], arr[:, 13]     # globaltimer (ns) of the LAST instrumented step: warp start / end
t0 = st.min()
print("last step, ns relative to the first warp's start: warp starts p50 %.0f p90 %.0f max %.0f | warp ends min %.0f p50 %.0f p99 %.0f max %.0f | lifetime mean %.0f max %.0f" % (
    np.percentile(st - t0, 50), np.percentile(st - t0, 90), (st - t0).max(), (en - t0).min(), np.percentile(en - t0, 50), np.percentile(en - t0, 99), (en - t0).max(),
    (en - st).mean(), (en - st).max()))
names = ["header load", "decode (after draws)", "shuffle", "wait pool tiles", "do_actions", "mtm+topK", "obs math", "obs+ring write", "reward/done", "state store"]
tot = sum(host[:12])
names += ["wait actions (+accts issue)", "normal draws"]
print(f"M={M} mix={mix}: {e0.elapsed_time(e1)/NS*1e3:.1f} us per step (instrumented), mean cycles per warp-step = {tot/(NS*M):.0f}")
for i, n in enumerate(names):
    print(f"  {n:30s} {host[i]/(NS*M):9.0f} cyc  {100*host[i]/tot:5.1f}%")
order = np.argsort(per_mkt)
slow = order[-max(1, M // 100):]
print("slowest 1%% of markets: mean %.0f cycles/step; phase breakdown (cycles per step) vs all markets:" % per_mkt[slow].mean())
for i, n in enumerate(names):
    print(f"  {n:30s} slow {arr[slow, i].mean()/NS:9.0f}   all {arr[:, i].mean()/NS:9.0f}")
smid = arr[:, 14].astype(int)
cnt = np.bincount(smid, minlength=148)
tot_sm = np.bincount(smid, weights=per_mkt, minlength=148) / np.maximum(cnt, 1)
mx_sm = np.array([per_mkt[smid == s_].max() if cnt[s_] else 0 for s_ in range(len(cnt))])
print("warps per SM: min %d max %d; SMs by warp count: %s" % (cnt[cnt > 0].min(), cnt.max(), dict(zip(*np.unique(cnt, return_counts=True)))))
for c in np.unique(cnt):
    if c: print("   SMs with %d warps: mean warp cycles %.0f, mean of per-SM max %.0f" % (c, tot_sm[cnt == c].mean(), mx_sm[cnt == c].mean()))
dec = np.array_split(np.arange(M), 8)
print("mean cycles by market-index octile:", " ".join("%.0f" % per_mkt[d].mean() for d in dec))
print("ten slowest markets (cycles per phase):")
print("   market  total " + " ".join(f"{n[:9]:>9s}" for n in names))
for mm in order[-10:]:
    print(f"   {mm:6d} sm{smid[mm]:3d} n{cnt[smid[mm]]:2d} {per_mkt[mm]:6.0f} " + " ".join(f"{arr[mm, i]/NS:9.0f}" for i in range(len(names))))
info = env.info_all()
mk = info["market"].cpu().numpy()
print("slow markets: mean trades this step %.2f (all %.2f)" % (info["num_trades_step"].cpu().numpy()[slow].sum(1).mean() / 2, info["num_trades_step"].cpu().numpy().sum(1).mean() / 2))
subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC,
                       "-o", _native.SO_PATH, os.path.join(_native.CSRC, "cda_b200.cu")])
