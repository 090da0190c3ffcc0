// pcie_store_bench.cu — how fast can SM stores / TMA bulk stores push the per-step host outputs of the CDA env over
// PCIe?  4096 warps (1024 CTAs x 4 warps), each writing what one market writes per window step: a 168-B snapshot into
// slot `pos` of its 32-slot row (row stride 5376 B) and a 64-B result record.  Patterns differ in how the bytes are
// cut into store instructions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pcie_store_bench tools/pcie_store_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// mode bits: 1 snapshot by 128-B aligned chunks (current kernel); 2 snapshot as whole 128-B lines; 4 snapshot by bulk store
// (16-B aligned superset); 8 record = 4x8-B rewards + separate 2-B flags (current); 16 record = one 40-B store instruction;
// 32 record = one full 64-B store instruction; 64 record by bulk store (64 B); 128 snapshot+record to a dense [M][64 floats] block (256-B rows, 2 full lines)
__global__ void __launch_bounds__(128) k_store(float *win, unsigned char *rec, int M, int pos, int mode, int spin) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = blockIdx.x * 4 + warp;
    if (m >= M) return;
    float *s = sm + warp * 96;
    for (int i = lane; i < 96; i += 32) s[i] = (float)(m + i + pos);
    __syncwarp();
    // stand-in for the matching work: spin so that stores of different warps are not all issued at t = 0
    long long t0 = clock64(); while (clock64() - t0 < (long long)spin * (1 + (m * 2654435761u >> 28))) {}
    float *rg = win + (size_t)m * (32 * 42) + pos * 42;
    if (mode & 1) {
        for (int cc = lane - (int)((reinterpret_cast<size_t>(rg) >> 2) & 31); cc < 42; cc += 32) if (cc >= 0) rg[cc] = s[cc + 2];
    }
    if (mode & 2) {
        const int mis = (int)((reinterpret_cast<size_t>(rg) >> 2) & 31);
        const int lo = -mis, hi = ((42 + mis + 31) & ~31) - mis;              // whole lines: floats [lo, hi) relative to rg
        float *row_end = win + (size_t)(m + 1) * (32 * 42);
        for (int cc = lo + lane; cc < hi; cc += 32) if (rg + cc < row_end && rg + cc >= win + (size_t)m * (32 * 42)) rg[cc] = s[(cc + 34) % 96];
    }
    if (mode & 4) {
        const size_t a = reinterpret_cast<size_t>(rg);
        const size_t a0 = a & ~(size_t)15, a1 = (a + 168 + 15) & ~(size_t)15;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a0), "r"(smem_u32(s)), "r"((unsigned)(a1 - a0)) : "memory");
        }
    }
    if (mode & 128) {
        float *d = win + (size_t)m * 64;
        d[lane] = s[lane]; d[32 + lane] = s[32 + lane];
    }
    double *rw = reinterpret_cast<double *>(rec + (size_t)m * 64);
    if (mode & 8) {
        if (lane < 4) rw[lane] = (double)s[lane];
        if (lane == 0) *reinterpret_cast<unsigned short *>(rec + (size_t)m * 64 + 32) = (unsigned short)pos;
    }
    if (mode & 16) { if (lane < 5) rw[lane] = lane < 4 ? (double)s[lane] : __longlong_as_double((long long)pos); }
    if (mode & 32) { if (lane < 8) rw[lane] = lane < 4 ? (double)s[lane] : __longlong_as_double((long long)pos); }
    if (mode & 64) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rw), "r"(smem_u32(s) + 192u), "r"(64u) : "memory");
    }
    if ((mode & (4 | 64)) && lane == 0) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

int main(int argc, char **argv) {
    const int M = 4096, reps = 200;
    float *hwin; unsigned char *hrec;
    CK(cudaHostAlloc(&hwin, (size_t)M * 32 * 42 * 4, cudaHostAllocMapped));
    CK(cudaHostAlloc(&hrec, (size_t)M * 64, cudaHostAllocMapped));
    float *dwin_h, *dwin_d; unsigned char *drec_h, *drec_d;
    CK(cudaHostGetDevicePointer(&dwin_h, hwin, 0)); CK(cudaHostGetDevicePointer(&drec_h, hrec, 0));
    CK(cudaMalloc(&dwin_d, (size_t)M * 32 * 42 * 4)); CK(cudaMalloc(&drec_d, (size_t)M * 64));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct { const char *name; int mode; } pats[] = {
        {"nothing (launch + spin only)", 0},
        {"snapshot: 128-B aligned chunks (current)", 1}, {"snapshot: whole 128-B lines", 2}, {"snapshot: one bulk store (16-B superset)", 4},
        {"record: 4x8 B + 2-B flags (current)", 8}, {"record: one 40-B store", 16}, {"record: one 64-B store", 32}, {"record: bulk store 64 B", 64},
        {"current total (chunks + 2-store record)", 1 | 8}, {"chunks + 40-B record", 1 | 16}, {"lines + 64-B record", 2 | 32}, {"bulk snapshot + 40-B record", 4 | 16},
        {"bulk snapshot + bulk record", 4 | 64}, {"dense 256-B rows (2 full lines per market, record inside)", 128},
    };
    for (int spin : {0, 1500}) {
        for (int host = 1; host >= 0; --host) {
            printf("---- destination: %s, spin %d cycles x (1..16)\n", host ? "pinned host (PCIe)" : "device HBM", spin);
            for (auto &p : pats) {
                float *w = host ? dwin_h : dwin_d; unsigned char *r = host ? drec_h : drec_d;
                for (int i = 0; i < 10; ++i) k_store<<<M / 4, 128, 4 * 96 * 4>>>(w, r, M, 3 + i % 29, p.mode, spin);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                for (int i = 0; i < reps; ++i) k_store<<<M / 4, 128, 4 * 96 * 4>>>(w, r, M, 3 + i % 29, p.mode, spin);
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("%-62s %7.2f us per launch\n", p.name, ms * 1e3 / reps);
            }
        }
    }
    return 0;
}
