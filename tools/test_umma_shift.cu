// Developer test (not part of the product path): can a tcgen05.mma A operand start at an arbitrary ROW of a
// swizzled shared-memory tile?  The row-tile conv kernel (aec_rt.cuh) reads the tap (ky, kx) of a 3x3 window as the
// same converted input row shifted by kx - 1 pixels, i.e. the same tile with a start address moved by whole rows.
// For SWIZZLE_128B (128-byte rows) and SWIZZLE_64B (64-byte rows) this prints the largest error of
//     D[m][n] = sum_k A[shift + m][k] * B[n][k]        M = 128, N = 16, K = 32 (128B) or 16 (64B)
// for shifts 0..9 with the descriptor's base-offset field (bits 49-51) set to 0 and to (start_address >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/test_umma_shift.bin tools/test_umma_shift.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout: 2 = SWIZZLE_128B (SBO 1024), 4 = SWIZZLE_64B (SBO 512)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int layout, uint32_t sbo, uint32_t base_off)
{
    const uint32_t lo = ((addr & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t hi = (sbo >> 4) | (1u << 14) | ((base_off & 7u) << 17) | ((uint32_t)layout << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__device__ __forceinline__ float bval(int n, int k) { return (float)(((n * 5 + k) % 13) - 6); }

// row_bytes = 128 (SW128) or 64 (SW64); swizzle on absolute address bits: 16-byte chunk index ^= (addr >> 7) & mask
__device__ __forceinline__ uint32_t sw_off(int r, int k, int row_bytes)
{
    const uint32_t lin = (uint32_t)r * row_bytes + (uint32_t)k * 4;
    const uint32_t mask = row_bytes == 128 ? 7u : 3u;
    return lin ^ (((lin >> 7) & mask) << 4);
}

__global__ void __launch_bounds__(128, 1) k_test(int row_bytes, int shift, int use_base_off, float *max_err)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int K = row_bytes / 4;
    unsigned char *A = smem, *B = smem + 32 * 1024;
    for (int i = threadIdx.x; i < 160 * K; i += 128) *(float *)(A + sw_off(i / K, i % K, row_bytes)) = aval(i / K, i % K);
    for (int i = threadIdx.x; i < 16 * K; i += 128) *(float *)(B + sw_off(i / K, i % K, row_bytes)) = bval(i / K, i % K);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const int layout = row_bytes == 128 ? 2 : 4;
        const uint32_t sbo = 8u * row_bytes;
        const uint32_t idesc = make_idesc_tf32(128, 16);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t a_addr = smem_u32(A) + (uint32_t)shift * row_bytes + ks * 32;
            const uint32_t b_addr = smem_u32(B) + ks * 32;
            const uint32_t bo = use_base_off ? ((smem_u32(A) + (uint32_t)shift * row_bytes) >> 7) & 7u : 0u;
            const uint64_t da = make_desc(a_addr, layout, sbo, bo), db = make_desc(b_addr, layout, sbo, 0);
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float err = 0.f;
    const int m = threadIdx.x;
    for (int n = 0; n < 16; ++n) {
        float want = 0.f;
        for (int k = 0; k < K; ++k) want += aval(shift + m, k) * bval(n, k);
        err = fmaxf(err, fabsf(__uint_as_float(r[n]) - want));
    }
    atomicMax((int *)max_err, __float_as_int(err));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int main()
{
    float *d_err;
    cudaMalloc(&d_err, 4);
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int row_bytes : {128, 64})
        for (int use_bo = 0; use_bo < 2; ++use_bo)
            for (int shift = 0; shift < 10; ++shift) {
                cudaMemset(d_err, 0, 4);
                k_test<<<1, 128, 64 * 1024>>>(row_bytes, shift, use_bo, d_err);
                cudaError_t e = cudaDeviceSynchronize();
                float err = -1.f;
                cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
                printf("rows of %3d B  base_offset %s  shift %d : max |err| %g  %s\n", row_bytes, use_bo ? "(addr>>7)&7" : "0          ", shift, err,
                       e == cudaSuccess ? (err == 0.f ? "EXACT" : "wrong") : cudaGetErrorString(e));
                if (e != cudaSuccess) return 1;
            }
    return 0;
}
