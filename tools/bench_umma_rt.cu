// Developer microbenchmark (not part of the product path): the MMA issue pattern of the row-tile conv kernel in isolation -
// per 8-wide K step two N = 2*Cpad MMAs (value / rate tile x [W_hi ; W_lo]) and two N = Cpad MMAs (lo tiles x W_hi) into two
// accumulators - as a function of the A tiles' start row (tap shift kx), the row width (128-byte / 64-byte swizzle) and Cpad.
// Warp-converged issue loop (descriptors in uniform registers), one commit + wait per 96 MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_umma_rt.bin tools/bench_umma_rt.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int ROWB>
__global__ void __launch_bounds__(128, 1) k_bench(int cpad, int shift, int iters, int pattern, long long *out)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 120 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
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
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    if (threadIdx.x < 32) {
        const uint64_t top = ROWB == 128 ? ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) : ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29)) << 32);
        const uint32_t idesc_cat = make_idesc(128, 2 * cpad), idesc_hi = make_idesc(128, cpad);
        const uint32_t tile16 = (136u * ROWB) >> 4;
        const uint32_t x_lo = ((smem_u32(smem) & 0x3ffffu) >> 4) | (1u << 16), w_lo = ((smem_u32(smem + 96 * 1024) & 0x3ffffu) >> 4) | (1u << 16);
        const uint32_t dv = tmem, dr = tmem + 2 * cpad;
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const uint32_t xa = x_lo + (uint32_t)(kx * shift) * (ROWB >> 4);
#pragma unroll
                    for (int ks = 0; ks < ROWB / 32; ++ks) {
                        const uint64_t dw = top | (uint64_t)(w_lo + 2u * ks);
                        if (pattern == 0) {
                            mma(dv, top | (uint64_t)(xa + 2u * ks), dw, idesc_cat, (it | kx | ks) != 0);
                            mma(dr, top | (uint64_t)(xa + 2u * tile16 + 2u * ks), dw, idesc_cat, (it | kx | ks) != 0);
                            mma(dv, top | (uint64_t)(xa + tile16 + 2u * ks), dw, idesc_hi, 1u);
                            mma(dr, top | (uint64_t)(xa + 3u * tile16 + 2u * ks), dw, idesc_hi, 1u);
                        } else {                          // the same four A tiles, all N = 2 * cpad
                            mma(dv, top | (uint64_t)(xa + 2u * ks), dw, idesc_cat, (it | kx | ks) != 0);
                            mma(dr, top | (uint64_t)(xa + 2u * tile16 + 2u * ks), dw, idesc_cat, (it | kx | ks) != 0);
                            mma(dv, top | (uint64_t)(xa + tile16 + 2u * ks), dw, idesc_cat, 1u);
                            mma(dr, top | (uint64_t)(xa + 3u * tile16 + 2u * ks), dw, idesc_cat, 1u);
                        }
                    }
                }
                if ((it & 1) == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            if ((it & 1) == 1) {
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
                parity ^= 1;
            }
        }
        if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 160 * sizeof(long long));
    const int smem = 122 * 1024, iters = 512, grid = 148;
    cudaFuncSetAttribute(k_bench<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bench<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto run = [&](int rowb, int cpad, int shift, int pattern) {
        if (rowb == 128) k_bench<128><<<grid, 128, smem>>>(cpad, shift, iters, pattern, out);
        else k_bench<64><<<grid, 128, smem>>>(cpad, shift, iters, pattern, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[160];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += h[i];
        const int mmas = 3 * (rowb / 32) * 4;
        return avg / grid / iters / mmas;
    };
    run(128, 32, 0, 0);
    for (int rowb : {128, 64})
        for (int cpad : {16, 32, 64})
            for (int shift : {0, 1, 8})
                printf("rows of %3d B, Cpad %2d (N = %3d and %3d), taps %d row(s) apart: %6.1f cycles per MMA   | all four N = %3d: %6.1f\n", rowb, cpad, 2 * cpad, cpad, shift,
                       run(rowb, cpad, shift, 0), 2 * cpad, run(rowb, cpad, shift, 1));
    return 0;
}
