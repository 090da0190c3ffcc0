#!/usr/bin/env python
"""Where the Decimal twin's time goes (cfg3, 4 x 4096): plain step with the ledger off / on without replays / on with the periodic
replay, and the replay kernel alone after N steps of journalling.  Run under gpurun."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions
M, A = 4096, 4
dec = os.environ.get("DEC", "1") == "1"
env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M, decimal_ledger=dec, status_policy="ignore")
env.reset(seed=1000)
acts = make_actions(7, 600, M, A, "limit_market")
dev = [torch.from_numpy(a).cuda() for a in acts]
for i in range(256): env.step(*[d[i] for d in dev])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): env.step(*[d[256 + i] for d in dev])
e1.record(); torch.cuda.synchronize()
msg = "%%-40s %%6.2f us/step (L2-hot, 200 steps)" %% (os.environ.get("TAG"), e0.elapsed_time(e1) * 5)
if dec and os.environ.get("SYNC"):
    n = int(os.environ["SYNC"])
    env._L.cda_twin_sync(env._h, None, None); torch.cuda.synchronize()
    for i in range(n): env.step(*[d[460 + i] for d in dev])
    torch.cuda.synchronize(); e0.record(); env._L.cda_twin_sync(env._h, None, None); e1.record(); torch.cuda.synchronize()
    msg += "   replay kernel after %%d steps: %%.1f us" %% (n, e0.elapsed_time(e1) * 1e3)
print(msg, "  tie-resolution passes so far:", env._L.cda_debug_restart_count())
''' % ROOT
for tag, ev in (("ledger off", dict(DEC="0")), ("ledger on, no replays", dict(CDA_TWIN_FLUSH="100000")), ("ledger on, replay every 10", dict(SYNC="10")),
                ("ledger on, replay every 5", dict(CDA_TWIN_FLUSH="5", SYNC="5")), ("ledger on, replay every 20", dict(CDA_TWIN_FLUSH="20", SYNC="20"))):
    e = dict(os.environ); e.update(ev); e["TAG"] = tag
    subprocess.run([sys.executable, "-c", CHILD], env=e)
