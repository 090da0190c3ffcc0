import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions
M,A=4096,4
env=cda.VecCDAEnv(dict(num_of_agents=A,max_step=1<<30),num_markets=M,decimal_ledger=True)
env.reset(seed=1000)
acts=make_actions(7,300,M,A,"limit_market")
dev=[torch.from_numpy(a).cuda() for a in acts]
for i in range(290): env.step(*[d[i] for d in dev])
torch.cuda.synchronize()
