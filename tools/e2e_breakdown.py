#!/usr/bin/env python
"""Where does the end-to-end step time go?  (run under gpurun)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_continuousdoubleauction_b200 as cda
from gym_continuousdoubleauction_b200.workloads import make_actions

M, A = 4096, 4
def t(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

for zc, zi, zf in (("1", "1", "0.25"), ("0", "1", "0"), ("0", "0", "0")):
    os.environ["CDA_ZEROCOPY"] = zc; os.environ["CDA_ZEROCOPY_IN"] = zi; os.environ["CDA_ZC_FRACTION"] = zf
    env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M)
    env.reset(seed=1000)
    acts = make_actions(7, 300, M, A, "limit_market")
    dev = [torch.from_numpy(a).cuda() for a in acts]
    pin = torch.empty((300, 5, M, A), dtype=torch.int32, pin_memory=True)
    for f in (0, 3, 4): pin[:, f].copy_(torch.from_numpy(acts[f]))
    for f in (1, 2): pin[:, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
    for i in range(256): env.step(*[d[i] for d in dev])
    it = [0]
    blocks = [pin[i] for i in range(300)]
    def host_step():
        b = blocks[it[0] % 300]; it[0] += 1
        env.step_host_block(b)
    def dev_step_sync():
        i = it[0] % 300; it[0] += 1
        env.step(*[d[i] for d in dev]); torch.cuda.current_stream().synchronize()
    def host_step_nosync():
        b = blocks[it[0] % 300]; it[0] += 1
        env.step_host_block(b, sync=False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(100): host_step_nosync()
    ev1.record(); torch.cuda.synchronize()
    print(f"   back-to-back device time per step_host_block (no host sync): {ev0.elapsed_time(ev1)*10:.1f} us")
    env.reset_host_window(seed=None)
    def win_step():
        b = blocks[it[0] % 300]; it[0] += 1
        env.step_host_window(b)
    print(f"   step_host_window (newest snapshot only, strided DMA, sync inside the C call): {t(win_step):7.1f} us")
    print(f"zerocopy out={zc} in={zi} fraction={zf}: step_host_block (sync each) {t(host_step):7.1f} us | device step + sync {t(dev_step_sync):7.1f} us")
    env.close()

# raw copies
obs_d = torch.empty(M * 168 + M * A * 2 + M, dtype=torch.float32, device="cuda")
obs_h = torch.empty_like(obs_d, device="cpu").pin_memory()
act_h = torch.empty(5 * M * A, dtype=torch.int32).pin_memory()
act_d = torch.empty(5 * M * A, dtype=torch.int32, device="cuda")
def d2h(): obs_h.copy_(obs_d, non_blocking=True); torch.cuda.current_stream().synchronize()
def h2d(): act_d.copy_(act_h, non_blocking=True); torch.cuda.current_stream().synchronize()
def empty_sync(): torch.cuda.current_stream().synchronize()
print(f"D2H {obs_d.numel()*4/1e6:.2f} MB + sync: {t(d2h):.1f} us ; H2D {act_h.numel()*4/1e3:.0f} KB + sync: {t(h2d):.1f} us ; bare sync {t(empty_sync):.2f} us")
big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
def bigd2h(): big_h.copy_(big_d, non_blocking=True); torch.cuda.current_stream().synchronize()
def bigh2d(): big_d.copy_(big_h, non_blocking=True); torch.cuda.current_stream().synchronize()
print(f"PCIe D2H {256/ (t(bigd2h, 5, 2)/1e6) / 1024:.1f} GiB/s, H2D {256/(t(bigh2d, 5, 2)/1e6)/1024:.1f} GiB/s")
