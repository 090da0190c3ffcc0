#!/usr/bin/env python
"""Can several copy engines share one strided D2H transfer?  Fork/join over K streams (events), device time per
iteration measured with CUDA events on the main stream (back to back, no host sync) and host-synced wall time.
Run under gpurun."""
import ctypes as C
import sys
import time

import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaStreamWaitEvent.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
D2H = 2
M = 4096
prop = torch.cuda.get_device_properties(0)
print("device:", prop.name, "| asyncEngineCount via cudart:", end=" ")
v = C.c_int(); rt.cudaDeviceGetAttribute(C.byref(v), 40, 0); print(v.value)   # cudaDevAttrAsyncEngineCount = 40
main = torch.cuda.current_stream()
st = C.c_void_p(main.cuda_stream)
side = [torch.cuda.Stream() for _ in range(8)]
side_p = [C.c_void_p(s.cuda_stream) for s in side]
dev = torch.zeros(M * 168 * 4 + (1 << 20), dtype=torch.uint8, device="cuda")
host = torch.zeros(M * 64 * 168 + (1 << 20), dtype=torch.uint8).pin_memory()
host2 = torch.zeros(1 << 20, dtype=torch.uint8).pin_memory()
ev_fork = torch.cuda.Event(); ev_join = [torch.cuda.Event() for _ in range(9)]
spin = torch.zeros(1 << 16, device="cuda")


def op(K, slots, with_small, width=168, spitch=672):
    """fork: K row-range chunks of the strided copy on K side streams (+ the small reward/flags copy on another)"""
    def f():
        spin.add_(1.0)     # stands in for the step kernel: something on the main stream to fork from
        if K == 0:
            rt.cudaMemcpy2DAsync(host.data_ptr(), slots * 168, dev.data_ptr() + 504, spitch, width, M, D2H, st)
            if with_small:
                rt.cudaMemcpyAsync(host2.data_ptr(), dev.data_ptr(), 139264, D2H, st)
            return
        ev_fork.record(main)
        rows = M // K
        for k in range(K):
            rt.cudaStreamWaitEvent(side_p[k], C.c_void_p(ev_fork.cuda_event), 0)
            rt.cudaMemcpy2DAsync(host.data_ptr() + k * rows * slots * 168, slots * 168, dev.data_ptr() + 504 + k * rows * spitch, spitch, width, rows, D2H, side_p[k])
            ev_join[k].record(side[k])
        if with_small:
            rt.cudaStreamWaitEvent(side_p[K], C.c_void_p(ev_fork.cuda_event), 0)
            rt.cudaMemcpyAsync(host2.data_ptr(), dev.data_ptr(), 139264, D2H, side_p[K])
            ev_join[K].record(side[K])
        for k in range(K + (1 if with_small else 0)):
            rt.cudaStreamWaitEvent(st, C.c_void_p(ev_join[k].cuda_event), 0)
    return f


def measure(f, n=200):
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    d = e0.elapsed_time(e1) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        f(); rt.cudaStreamSynchronize(st)
    return d, (time.perf_counter() - t0) / n * 1e6


d, h = measure(lambda: spin.add_(1.0)); print(f"stand-in kernel alone: device {d:.1f} us, host-synced {h:.1f} us")
for slots in (16, 64):
    for small in (False, True):
        for K in (0, 1, 2, 4, 8):
            d, h = measure(op(K, slots, small))
            print(f"slots={slots:2d} small={int(small)} K={K}: device {d:6.1f} us   host-synced {h:6.1f} us")
# time-major alternative: ONE contiguous copy of newest snapshots + rewards + flags
def one():
    spin.add_(1.0); rt.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), M * 168 + 139264, D2H, st)
d, h = measure(one); print(f"1D contiguous {M*168+139264} B on the main stream: device {d:.1f} us  host-synced {h:.1f} us")
def one_k(K):
    def f():
        spin.add_(1.0); ev_fork.record(main)
        n = (M * 168 + 139264) // K // 16 * 16
        for k in range(K):
            rt.cudaStreamWaitEvent(side_p[k], C.c_void_p(ev_fork.cuda_event), 0)
            rt.cudaMemcpyAsync(host.data_ptr() + k * n, dev.data_ptr() + k * n, n, D2H, side_p[k])
            ev_join[k].record(side[k])
        for k in range(K):
            rt.cudaStreamWaitEvent(st, C.c_void_p(ev_join[k].cuda_event), 0)
    return f
for K in (2, 4):
    d, h = measure(one_k(K)); print(f"1D contiguous split over K={K} streams: device {d:.1f} us  host-synced {h:.1f} us")
