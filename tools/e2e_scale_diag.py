#!/usr/bin/env python
"""Why does the end-to-end (host buffers) path lose efficiency with more GPUs on one node?

Launch with torchrun, one rank per GPU (`--nproc-per-node 8`).  Inside ONE launch the script times the host path
with only the first n ranks active for n in 1, 2, 4, 8 (the idle ranks sleep in a gloo barrier: no spinning, no GPU
work), and for every n a set of variants that remove one suspected limiter at a time:

    full          step_host_window(sync=True): actions read from pinned memory, outputs stored to pinned memory, sync per step
    nosync        same traffic, but 50 steps are queued before one synchronisation (PCIe + kernel only, no per-step host round trip)
    dev_sync      env.step on device-resident actions + a stream synchronisation per step (launch + sync only, no PCIe payload)
    full_smi      `full` with one `nvidia-smi -lms 100` sampler per active rank (what bench.py round 1 did)
    full_numa     `full` after pinning the process to the cores of its GPU's NUMA node and re-creating the env (pinned
                  buffers then come from that node: first touch)

Prints one line per (n, variant): max over active ranks of us/step, plus the host topology (cores, NUMA nodes,
GPU <-> NUMA affinity).  Run: gpurun --gpus 8 -- 'python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1
--master-port 29511 tools/e2e_scale_diag.py > gpurun_out/e2e_scale_diag.txt 2>&1'"""
import glob
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def gpu_numa(idx):
    try:
        p = torch.cuda.get_device_properties(idx)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        return bdf, int(node)
    except Exception as e:  # noqa: BLE001
        return f"<{e}>", -1


def node_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:  # noqa: BLE001
        return None
    cpus = []
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-"); cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # CPU barriers only: idle ranks must not touch their GPU
    import gym_continuousdoubleauction_b200 as cda
    from gym_continuousdoubleauction_b200.workloads import make_actions

    bdf, node = gpu_numa(local)
    aff = sorted(os.sched_getaffinity(0))
    if rank == 0:
        print("== host:", sh("nproc"), "cpus visible;", sh("lscpu | egrep 'Model name|Socket|NUMA node|Thread|Core' | tr -s ' ' | tr '\\n' ';'"))
        print("== numa nodes:", [os.path.basename(p) for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))])
        print("== nvidia-smi topo -m\n" + sh("nvidia-smi topo -m"))
        sys.stdout.flush()
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(f"== rank {rank}: gpu {local} pci {bdf} numa_node {node}; affinity {len(aff)} cpus [{aff[0]}..{aff[-1]}]", flush=True)
    dist.barrier()

    M, A, T = 4096, 4, 300
    acts = make_actions(7 + rank, T, M, A, "limit_market")

    def make_env(planes=False):
        env = cda.VecCDAEnv(dict(num_of_agents=A, max_step=1 << 30), num_markets=M, device=local, status_policy="ignore")   # (the no-input timing mode replays one action block: books overflow)
        env.reset(seed=np.arange(M, dtype=np.uint64) + np.uint64(1000 + rank * M))
        dev = [torch.from_numpy(a).cuda() for a in acts]
        for i in range(200):
            env.step(*[d[i] for d in dev])
        pin = torch.empty((T, M, 5, A), dtype=torch.int32, pin_memory=True)
        for f in (0, 3, 4):
            pin[:, :, f].copy_(torch.from_numpy(acts[f]))
        for f in (1, 2):
            pin[:, :, f].view(torch.float32).copy_(torch.from_numpy(acts[f]))
        if planes:
            env.attach_host_planes()
        else:
            env.attach_host_window()
        torch.cuda.synchronize()
        return env, dev, [pin[i] for i in range(T)]

    env, dev, blocks = make_env()
    K = 200

    def t_full(env, dev, blocks):
        for i in range(20):
            env.step_host_window(blocks[i], market_major=True)
        t0 = time.perf_counter()
        for i in range(K):
            o, r, te, tr = env.step_host_window(blocks[(20 + i) % T], market_major=True)
            _ = float(r[0, 0]) + float(o[M - 1, env.W - 1])
        return (time.perf_counter() - t0) / K * 1e6

    def t_nosync(env, dev, blocks):
        for i in range(20):
            env.step_host_window(blocks[i], market_major=True, sync=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            env.step_host_window(blocks[(20 + i) % T], market_major=True, sync=(i % 50 == 49))
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / K * 1e6

    def t_planes(env, dev, blocks):
        for i in range(20):
            env.step_host_planes(blocks[i])
        t0 = time.perf_counter()
        for i in range(K):
            o, r, te, tr = env.step_host_planes(blocks[(20 + i) % T])
            _ = float(r[0, 0]) + float(o[M - 1, env.W - 1])
        return (time.perf_counter() - t0) / K * 1e6

    def t_planes_stacked(env, dev, blocks):      # ... plus the consumer's copy into a contiguous [M, 168] array
        buf = np.empty((M, env.W), np.float32)
        for i in range(20):
            env.step_host_planes(blocks[i])
        t0 = time.perf_counter()
        for i in range(K):
            o, r, te, tr = env.step_host_planes(blocks[(20 + i) % T])
            o.stacked(out=buf)
            _ = float(r[0, 0]) + float(buf[M - 1, env.W - 1])
        return (time.perf_counter() - t0) / K * 1e6

    def t_dev_sync(env, dev, blocks):
        st = torch.cuda.current_stream()
        for i in range(20):
            env.step(*[d[i] for d in dev])
        st.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            env.step(*[d[(20 + i) % T] for d in dev]); st.synchronize()
        env.attach_host_window()
        return (time.perf_counter() - t0) / K * 1e6

    def t_full_smi(env, dev, blocks):
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw", "--format=csv,noheader", "-lms", "100", "-i", str(local)],
                             stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.3)
        try:
            return t_full(env, dev, blocks)
        finally:
            p.terminate(); p.wait()

    results = {}

    def run_variant(name, fn, envset, n):
        dist.barrier()
        v = torch.zeros(1, dtype=torch.float64)
        if rank < n:
            v[0] = fn(*envset)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        results[(n, name)] = float(v[0])
        if rank == 0:
            print(f"n_active={n:<2d} {name:<16s} {float(v[0]):8.1f} us/step  -> {n * M / float(v[0]):7.1f} M env-steps/s aggregate", flush=True)

    L = env._L
    quick = os.environ.get("DIAG_SPLIT", "0") == "1"
    ns = [n for n in ((1, 8) if quick else (1, 2, 4, 8)) if n <= world]

    def with_mode(mode, fn):
        def g(*a):
            L.cda_debug_set_window_mode(mode)
            try:
                return fn(*a)
            finally:
                L.cda_debug_set_window_mode(0)
        return g

    if os.environ.get("DIAG_LAYOUT", "0") == "1":
        # output layout: window row pitch (slots) vs the dense plane ring
        ns = [n for n in (1, 2, 4, 8) if n <= world]
        for n in ns:
            run_variant("window32", t_full, (env, dev, blocks), n)
        env.close()
        for slots in (16, 8):
            cda.VecCDAEnv.WINDOW_SLOTS = slots
            e2 = make_env()
            for n in ns:
                run_variant(f"window{slots}", t_full, e2, n)
            e2[0].close()
        cda.VecCDAEnv.WINDOW_SLOTS = 32
        for cell in (64, 52):
            cda.VecCDAEnv.PLANE_CELL_WORDS = cell
            e2 = make_env(planes=True)
            for n in ns:
                run_variant(f"planes{cell}", t_planes, e2, n)
                if cell == 64:
                    run_variant("planes64+copy", t_planes_stacked, e2, n)
            e2[0].close()
    elif quick:
        # which direction is the limiter?  (timing modes of the window path: bit 0 = no input transfer, bit 1 = no output transfer)
        for n in ns:
            for name, fn in (("full", t_full), ("no_in", with_mode(1, t_full)), ("no_out", with_mode(2, t_full)), ("no_io", with_mode(3, t_full)),
                             ("nosync", t_nosync), ("nosync_no_in", with_mode(1, t_nosync)), ("nosync_no_out", with_mode(2, t_nosync))):
                run_variant(name, fn, (env, dev, blocks), n)
        env.close()
        for tag, ev in (("in_dma", {"CDA_ZEROCOPY_IN": "0"}), ("out_dma", {"CDA_ZEROCOPY": "0"}), ("io_dma", {"CDA_ZEROCOPY_IN": "0", "CDA_ZEROCOPY": "0"})):
            os.environ.update(ev)
            e2 = make_env()
            for k in ev:
                del os.environ[k]
            for n in ns:
                run_variant(tag, t_full, e2, n)
                run_variant(tag + "_nosync", t_nosync, e2, n)
            e2[0].close()
    else:
        for n in ns:
            for name, fn in (("full", t_full), ("nosync", t_nosync), ("dev_sync", t_dev_sync), ("full_smi", t_full_smi)):
                run_variant(name, fn, (env, dev, blocks), n)
        # NUMA-local variant: pin to the GPU's node, re-create env + pinned buffers
        env.close()
        cpus = node_cpus(node) if node >= 0 else None
        if cpus:
            os.sched_setaffinity(0, set(cpus) & set(aff) or set(aff))
        env2 = make_env()
        for n in ns:
            run_variant("full_numa", t_full, env2, n)
            run_variant("dev_sync_numa", t_dev_sync, env2, n)
        if rank == 0:
            print("== efficiency of `full` vs n=1:", {n: round(results[(1, "full")] / results[(n, "full")], 3) for n in ns})
            print("== efficiency of `full_numa` vs n=1:", {n: round(results[(1, "full_numa")] / results[(n, "full_numa")], 3) for n in ns})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
