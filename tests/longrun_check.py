#!/usr/bin/env python
"""Long-run parity check against the oracle (run under gpurun): M markets x T steps, obs compared every
`every` steps, status + every market's book compared at the end.  usage: longrun_check.py [nvcc -D flags]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gym_continuousdoubleauction_b200 import _native
flags = sys.argv[1].split() if len(sys.argv) > 1 else []
subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC, "-o", _native.SO_PATH, os.path.join(_native.CSRC, "cda_b200.cu")])
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions
from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal
A, M, T, mix = 4, int(os.environ.get("LR_M", 2048)), int(os.environ.get("LR_T", 480)), os.environ.get("LR_MIX", "limit_market")
cfg = dict(num_of_agents=A, max_step=1 << 30)
env = cda.VecCDAEnv(cfg, num_markets=M, fill_capacity=0); orc = OracleEnv(cfg, M)
seeds = np.arange(M, dtype=np.uint64) + 1000
env.reset(seed=seeds); orc.reset(seeds=seeds)
acts = make_actions(7, T, M, A, mix)
bad = None
for t in range(T):
    og, rg, _, _ = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
    oc, rc, _, _ = orc.step(*[a[t] for a in acts], nthreads=32)
    if t % 8 == 0 or t == T - 1:
        d = np.abs(og.cpu().numpy().astype(np.float64) - oc).max(axis=1)
        if d.max() > 1e-6:
            m = int(np.argmax(d > 1e-6)); bad = (t, m, float(d[m]), int((d > 1e-6).sum())); break
print("flags", flags, "first obs mismatch (t, market, diff, n_markets):", bad, "status max", int(env.status().max().item()))
if bad:
    t, m = bad[0], bad[1]
    g, c = env.dump(m), orc.dump(m)
    for k in ("bids", "asks"):
        print(k, "gpu n", g[k].shape[0], "cpu n", c[k].shape[0], "equal", np.array_equal(g[k], c[k]))
    print("gpu best", g["best_bid"], g["best_ask"], "cpu best", c["best_bid"], c["best_ask"], "time", g["time"], c["time"])
else:
    for m in range(0, M, 37):
        assert_dump_equal(env.dump(m), orc.dump(m), ctx=f"m={m}", fills=False)
    print("final dumps equal")
subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", _native.CSRC, "-o", _native.SO_PATH, os.path.join(_native.CSRC, "cda_b200.cu")])
