#!/usr/bin/env python
"""How fast does the copy engine move M short rows (one 42-float snapshot per market) device -> pinned host
with cudaMemcpy2DAsync, compared with one contiguous copy of the same bytes?  (run under gpurun)
Decides the layout of the sliding-window host path (cda_step_host_window)."""
import ctypes as C
import sys
import time

import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
D2H = 2
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
dev = torch.zeros(M * 168 * 4, dtype=torch.uint8, device="cuda")
host = torch.zeros(M * 64 * 168 + 4096, dtype=torch.uint8).pin_memory()


def t(fn, n=300, warm=30):
    for _ in range(warm):
        fn()
    rt.cudaStreamSynchronize(st)
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        rt.cudaStreamSynchronize(st)
    return (time.perf_counter() - t0) / n * 1e6


def c1d(nbytes):
    return lambda: rt.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), nbytes, D2H, st)


def c2d(width, spitch, dpitch, doff=0, soff=0):
    return lambda: rt.cudaMemcpy2DAsync(host.data_ptr() + doff, dpitch, dev.data_ptr() + soff, spitch, width, M, D2H, st)


print(f"M={M}; bare sync {t(lambda: None):.1f} us")
print(f"1D {M*168/1e3:.0f} KB: {t(c1d(M*168)):.1f} us ; 1D {M*672/1e3:.0f} KB: {t(c1d(M*672)):.1f} us ; 1D {M*34/1e3:.0f} KB {t(c1d(M*34)):.1f} us")
for name, f in (("2D w168 sp168 dp168 (contiguous both)", c2d(168, 168, 168)),
                ("2D w168 sp672 dp168", c2d(168, 672, 168, 0, 504)),
                ("2D w168 sp168 dp672", c2d(168, 168, 672)),
                ("2D w168 sp168 dp16*168", c2d(168, 168, 16 * 168)),
                ("2D w168 sp168 dp64*168", c2d(168, 168, 64 * 168)),
                ("2D w168 sp672 dp64*168 off 5*168", c2d(168, 672, 64 * 168, 5 * 168, 504)),
                ("2D w176 sp176 dp64*176 (16B-multiple rows)", c2d(176, 176, 64 * 168)),
                ("2D w192 sp192 dp64*192 (64B rows)", c2d(192, 192, 64 * 168)),
                ("2D w256 sp256 dp64*168", c2d(256, 256, 64 * 168)),
                ("2D w672 sp672 dp64*168", c2d(672, 672, 64 * 168))):
    print(f"{name:48s} {t(f):7.1f} us")
