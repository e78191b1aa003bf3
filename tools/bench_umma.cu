// Developer microbenchmark (not part of the product path): cycles per tcgen05.mma for the shapes
// the gathered GEMM can use, issued back to back by one thread from static shared-memory operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_umma.bin tools/bench_umma.cu
//   tools/bench_umma.bin
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
// kind: 0 = tf32 (K = 8), 1 = bf16 (K = 16)
__device__ __forceinline__ uint32_t make_idesc(int kind, int m, int n)
{
    const uint32_t fmt = kind == 0 ? 2u : 1u;     // tf32 = 2, bf16 = 1
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}

// mode 0: all MMAs into one accumulator, fixed descriptors
// mode 1: alternate between two accumulators
// mode 2: one accumulator, descriptors recomputed from a rotating stage offset (like the product kernel)
template <int KIND>
__global__ void __launch_bounds__(128, 1) k_bench(int n, int m, int iters, int mode, long long *out)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
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
        const uint32_t idesc = make_idesc(KIND, m, n);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        const long long t0 = clock64();
        uint32_t parity = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t so = mode == 2 ? (uint32_t)(it & 1) * 32768u : 0u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ko = ks * 32u;
                const uint64_t da = make_desc_sw128(a0 + (mode == 2 ? (so >> 1) : 0u) + ko), db = make_desc_sw128(b0 + so + ko);
                const uint64_t da2 = make_desc_sw128(a0 + 16384u + ko), db2 = make_desc_sw128(b0 + 32768u + ko);
                const uint32_t d0 = tmem, d1 = tmem + (mode == 1 ? 256u : 0u);
                mma<KIND>(d0, da, db2, idesc, (it | ks) != 0);
                mma<KIND>(d1, da2, db, idesc, (it | ks) != 0);
                mma<KIND>(d0, da, db, idesc, 1u);
            }
            if ((it & 7) == 7) {     // commit + wait every 96 MMAs so the queue depth stays bounded
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok)
                                 : "r"(smem_u32(&bar)), "r"(parity)
                                 : "memory");
                parity ^= 1;
            }
        }
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 148 * sizeof(long long));
    const int smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(k_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 512;      // multiple of 8
    for (int kind = 0; kind < 2; ++kind)
        for (int grid : {1, 148})
            for (int mode = 0; mode < 3; ++mode)
                for (int n : {32, 64, 128, 256}) {
                    for (int rep = 0; rep < 2; ++rep) {
                        if (kind == 0) k_bench<0><<<grid, 128, smem>>>(n, 128, iters, mode, out);
                        else k_bench<1><<<grid, 128, smem>>>(n, 128, iters, mode, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    long long h[148];
                    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double avg = 0;
                    for (int i = 0; i < grid; ++i) avg += (double)h[i];
                    avg /= grid;
                    printf("kind %s grid %3d mode %d M 128 N %3d : %.1f cycles per MMA\n", kind == 0 ? "tf32" : "bf16", grid, mode, n, avg / (iters * 12.0));
                }
    return 0;
}
