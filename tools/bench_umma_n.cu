// Developer microbenchmark (not part of the product path): what a tcgen05.mma.kind::tf32 (M = 128, K = 8) costs as a
// function of N and of how many INDEPENDENT accumulators the issue stream rotates over, and what a tcgen05.commit costs
// the issuing thread.  One CTA per SM, one issuing thread, operands in static shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_umma_n.bin tools/bench_umma_n.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    const uint32_t lo = ((smem_addr & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// `iters` groups of 64 MMAs (N = n) rotating over NACC accumulators, one commit + wait per group; `extra_commits` additional
// commits (on other barriers, never waited on) spread through each group.  The issue loop is fully unrolled with the descriptors
// formed by one add each: a single thread issues an instruction every few cycles, so a loop with index arithmetic per MMA
// measures the thread, not the tensor core (first version of this file: 162 cycles per MMA whatever N).
template <int NACC>
__global__ void __launch_bounds__(128, 1) k_bench(int n, int iters, int extra_commits, long long *out)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar, bars2[64];
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars2[i])), "r"(1 << 20) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(128, n);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
        const int every = extra_commits > 0 ? 64 / extra_commits : 1 << 30;
        uint32_t parity = 0;
        const long long t0 = clock64();
        const uint64_t da0 = make_desc_sw128(a0), db0 = make_desc_sw128(b0);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const uint64_t da = da0 + (uint64_t)((((j >> 2) & 1) * 16384 + (j & 3) * 32) >> 4), db = db0 + (uint64_t)(((j & 3) * 32) >> 4);
                mma(tmem + (uint32_t)((j % NACC) * (512 / NACC)), da, db, idesc, (it | j) != 0);
                if (extra_commits > 0 && (j % every) == every - 1) commit(smem_u32(&bars2[j]));
            }
            commit(smem_u32(&bar));
            wait(smem_u32(&bar), parity);
            parity ^= 1;
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 160 * sizeof(long long));
    const int smem = 100 * 1024, iters = 64, grid = 148;
    cudaFuncSetAttribute(k_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto run = [&](int n, int nacc, int extra) {
        if (nacc == 1) k_bench<1><<<grid, 128, smem>>>(n, iters, extra, out);
        else if (nacc == 2) k_bench<2><<<grid, 128, smem>>>(n, iters, extra, out);
        else k_bench<4><<<grid, 128, smem>>>(n, iters, extra, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[160];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += h[i];
        return avg / grid / iters / 64;
    };
    run(128, 1, 0);
    for (int n : {16, 32, 64, 128, 256})
        for (int nacc : {1, 2, 4}) {
            if (n * nacc > 512) continue;
            printf("tf32 M128 N%3d K8, %d accumulator(s) in rotation: %6.1f cycles per MMA\n", n, nacc, run(n, nacc, 0));
        }
    for (int n : {64, 256})
        for (int extra : {0, 2, 4, 16}) {
            const double c0 = run(n, 2, 0), c1 = run(n, 2, extra);
            printf("tf32 N%3d, 2 accumulators, %2d extra commits per 64 MMAs: %6.1f cycles per MMA (%+.1f cycles per commit)\n", n, extra, c1,
                   extra ? (c1 - c0) * 64 / extra : 0.0);
        }
    return 0;
}
