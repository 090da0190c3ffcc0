#!/usr/bin/env python
"""Long-episode parity against the oracle (run under gpurun; a script, not a pytest module): M markets stepped T times (default 1024 x 4096:
the max_step of INTEGRATION.md's example), observations / rewards compared every 64 steps, every market's book / ledger / RNG every 1024 steps
and at the end, the sticky status (pool overflow!) checked at the end, the deepest book side reported.
usage: python tests/longrun_episode.py [M=1024] [T=4096] [mix=limit_market]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions
from oracle.cda_oracle import OracleEnv
from parity_utils import assert_dump_equal

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mix = sys.argv[3] if len(sys.argv) > 3 else "limit_market"
A = 4
cfg = dict(num_of_agents=A, max_step=1 << 30)
env = cda.VecCDAEnv(cfg, num_markets=M, decimal_ledger=True); orc = OracleEnv(cfg, M, decimal_ledger=True, dec128=True)
seeds = np.arange(M, dtype=np.uint64) + np.uint64(1000)
env.reset(seed=seeds); orc.reset(seeds=seeds)
t0 = time.time(); deepest = 0; chunk = 256
for c0 in range(0, T, chunk):
    acts = make_actions(7 + c0, chunk, M, A, mix)
    for t in range(chunk):
        og, rg, _, _ = env.step(*[torch.from_numpy(np.ascontiguousarray(a[t])).cuda() for a in acts])
        oc, rc, _, _ = orc.step(*[a[t] for a in acts], nthreads=os.cpu_count() or 8)
        if (c0 + t) % 64 == 63:
            assert np.abs(og.cpu().numpy().astype(np.float64) - oc).max() <= 1e-6 and np.abs(rg.cpu().numpy() - rc).max() <= 1e-6, c0 + t
    if (c0 + chunk) % 1024 == 0 or c0 + chunk >= T:
        dumps = env.dump_all()
        for m in range(M):
            assert_dump_equal(dumps[m], orc.dump(m), ctx=f"t={c0 + chunk} m={m}", fills=False)
            deepest = max(deepest, len(dumps[m]["bids"]), len(dumps[m]["asks"]))
        print(f"t={c0 + chunk}: all {M} markets equal (books, ledgers, RNG); deepest side so far {deepest} of {env.order_capacity}", flush=True)
st = int(env.status().max().item())
f = env.decimal_fields(range(0, M, 64))
for m in range(0, M, 64):
    d = orc.dump_decimal(m)
    assert f[m]["cash"] == d["cash"] and f[m]["VWAP"] == d["VWAP"] and f[m]["nav"] == d["nav"], m
print(f"== {M} markets x {T} steps ({mix}, decimal_ledger on): parity held, status bits {st}, deepest book side {deepest} / capacity {env.order_capacity}, {time.time() - t0:.0f} s")
assert st == 0
