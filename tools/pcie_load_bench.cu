// pcie_load_bench.cu — how should 4096 markets' action records (80 B each, 327 KB in all) be pulled out of pinned host
// memory by the kernel?  Varies the size of the individual read requests (cp.async.bulk global->shared from a mapped host
// pointer, or plain 128-bit loads) at a fixed total.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pcie_load_bench tools/pcie_load_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// each CTA pulls `bytes` contiguous bytes in `pieces` bulk copies, waits for them and writes one word to `sink`
__global__ void k_bulk(const unsigned char *src, int bytes, int pieces, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)bytes) : "memory");
        const int pb = bytes / pieces;
        for (int i = 0; i < pieces; ++i)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm) + i * pb), "l"(src + (size_t)blockIdx.x * bytes + (size_t)i * pb), "r"((unsigned)pb), "r"(b) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
    if (threadIdx.x == 0) sink[blockIdx.x] = reinterpret_cast<unsigned *>(sm)[0];
}
// plain loads: every thread one 16-B load, coalesced
__global__ void k_ldg(const uint4 *src, int n16, unsigned *sink) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) { const uint4 v = src[i]; if (v.x == 0x12345678u) sink[0] = v.y; }
}

int main() {
    const int total = 4096 * 80, reps = 200;
    unsigned char *h, *dh, *dd; unsigned *sink;
    CK(cudaHostAlloc(&h, total, cudaHostAllocMapped)); CK(cudaHostGetDevicePointer(&dh, h, 0));
    CK(cudaMalloc(&dd, total)); CK(cudaMalloc(&sink, 1 << 20));
    for (int i = 0; i < total; ++i) h[i] = (unsigned char)i;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct { int bytes, pieces; } pats[] = {{320, 5}, {320, 1}, {640, 1}, {1280, 1}, {2560, 1}, {5120, 1}, {20480, 1}, {81920, 1}, {163840, 1}, {163840, 8}, {81920, 16}};
    for (int host = 1; host >= 0; --host) {
        const unsigned char *src = host ? dh : dd;
        printf("---- source: %s\n", host ? "pinned host (PCIe)" : "device HBM");
        for (auto &p : pats) {
            const int grid = total / p.bytes;
            for (int i = 0; i < 5; ++i) k_bulk<<<grid, 128, p.bytes>>>(src, p.bytes, p.pieces, sink);
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
            for (int i = 0; i < reps; ++i) k_bulk<<<grid, 128, p.bytes>>>(src, p.bytes, p.pieces, sink);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("%5d CTAs x %6d B in %2d bulk copies each        %7.2f us per launch\n", grid, p.bytes, p.pieces, ms * 1e3 / reps);
        }
        for (int i = 0; i < 5; ++i) k_ldg<<<total / 16 / 128, 128>>>(reinterpret_cast<const uint4 *>(src), total / 16, sink);
        CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) k_ldg<<<total / 16 / 128, 128>>>(reinterpret_cast<const uint4 *>(src), total / 16, sink);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("plain 16-B loads, one per thread, coalesced            %7.2f us per launch\n", ms * 1e3 / reps);
        // copy engine for comparison
        for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(dd, host ? h : dd, total, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, 0));
        CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(dd, host ? h : dd + 0, total, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, 0));
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("cudaMemcpyAsync of the whole block (back to back)      %7.2f us per copy\n", ms * 1e3 / reps);
    }
    return 0;
}
