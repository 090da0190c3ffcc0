"""The CUDA env (through the C-ABI) against golden trajectories recorded from the UNMODIFIED
Python reference — same checks the oracle has to pass in test_oracle_golden.py."""
import os

import numpy as np
import pytest
import torch

from test_oracle_golden import GOLD, check_backend_against_golden

import gym_continuousdoubleauction_b200 as cda

pytestmark = pytest.mark.gpu


class GpuOne:
    def __init__(self, cfg):
        self.e = cda.VecCDAEnv(cfg, num_markets=1, fill_capacity=64)

    def reset_one(self, seed):
        return self.e.reset(seed=[seed]).cpu().numpy()[0].copy()

    def step_one(self, cat, mean, sigma, price, off):
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)[None]).cuda()
        o, r, te, tr = self.e.step(t(cat, np.int32), t(mean, np.float32), t(sigma, np.float32), t(price, np.int32), t(off, np.int32))
        return o.cpu().numpy()[0].copy(), r.cpu().numpy()[0].copy(), int(te[0]), int(tr[0])

    def dump_one(self):
        return self.e.dump(0)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_gpu_matches_reference_golden(path):
    check_backend_against_golden(GpuOne, path)
