// Developer microbenchmark (not part of the product path): cycles per tcgen05.mma for the shapes
// the gathered GEMM can use, issued back to back by one thread from static shared-memory operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_umma.bin tools/bench_umma.cu
//   tools/bench_umma.bin
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__global__ void __launch_bounds__(640, 1) k_bench(int n, int m, int iters, int mode, long long *out, int wr_sleep)
{
    __shared__ volatile int s_done;
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) s_done = 0;
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
        s_done = 1;
    }
    if (threadIdx.x >= 128) {      // background shared-memory writers (16 bytes per lane per store)
        float4 *dst = reinterpret_cast<float4 *>(smem + 128 * 1024) + (threadIdx.x - 128);
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        long long cnt = 0;
        while (!s_done) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k * 512] = v;
            v.x += 1.f;
            ++cnt;
            if (wr_sleep) __nanosleep(wr_sleep);
        }
        if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&out[148 + 0], (unsigned long long)cnt);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// Pass-structured issue like the product kernel: pass k waits for the commit of pass k - depth (its stage
// being free), issues 12 MMAs (N = 256) and commits.  `extra_waits` adds that many try_waits on an
// already completed barrier per pass (the product kernel waits on 3 operand barriers per pass).
__global__ void __launch_bounds__(128, 1) k_pass(int n, int passes, int depth, int extra_waits, long long *out, int use_test)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar[8], done_bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");     // phase 0 complete
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
    auto wait = [&](uint32_t b, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    };
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(0, 128, n);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        const long long t0 = clock64();
        for (int k = 0; k < passes; ++k) {
            const int st = k % depth;
            if (k >= depth) wait(smem_u32(&bar[st]), (uint32_t)(k / depth - 1) & 1u);
            for (int e = 0; e < extra_waits; ++e) {
                if (use_test) {
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(&done_bar)), "r"(0u) : "memory");
                } else wait(smem_u32(&done_bar), 0u);
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                const uint32_t so = (uint32_t)(st & 1) * 32768u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ko = ks * 32u;
                    const uint64_t da = make_desc_sw128(a0 + ko), db = make_desc_sw128(b0 + so + ko);
                    const uint64_t da2 = make_desc_sw128(a0 + 16384u + ko), db2 = make_desc_sw128(b0 + so + ko);
                    mma<0>(tmem, da, db2, idesc, (k | ks) != 0);
                    mma<0>(tmem, da2, db, idesc, 1u);
                    mma<0>(tmem, da, db, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[st])) : "memory");
            }
            __syncwarp();
        }
        // drain
        for (int k = passes; k < passes + depth; ++k) wait(smem_u32(&bar[k % depth]), (uint32_t)(k / depth - 1) & 1u);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// Same pass structure, but all mbarrier waits are done by a helper warp that then releases the MMA warp
// through a named barrier (bar.arrive / bar.sync): the MMA warp executes no mbarrier wait at all.
__global__ void __launch_bounds__(128, 1) k_pass_named(int n, int passes, int extra_waits, long long *out)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar[8], done_bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
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
    auto wait = [&](uint32_t b, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    };
    const int depth = 2;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(0, 128, n);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        const long long t0 = clock64();
        for (int k = 0; k < passes; ++k) {
            const int st = k % depth;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (k & 3)) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                const uint32_t so = (uint32_t)(st & 1) * 32768u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ko = ks * 32u;
                    const uint64_t da = make_desc_sw128(a0 + ko), db = make_desc_sw128(b0 + so + ko);
                    const uint64_t da2 = make_desc_sw128(a0 + 16384u + ko), db2 = make_desc_sw128(b0 + so + ko);
                    mma<0>(tmem, da, db2, idesc, (k | ks) != 0);
                    mma<0>(tmem, da2, db, idesc, 1u);
                    mma<0>(tmem, da, db, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[st])) : "memory");
            }
            __syncwarp();
        }
        for (int k = passes; k < passes + depth; ++k) wait(smem_u32(&bar[k % depth]), (uint32_t)(k / depth - 1) & 1u);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    } else if (threadIdx.x < 64) {
        for (int k = 0; k < passes; ++k) {
            const int st = k % depth;
            if (k >= depth) wait(smem_u32(&bar[st]), (uint32_t)(k / depth - 1) & 1u);
            for (int e = 0; e < extra_waits; ++e) wait(smem_u32(&done_bar), 0u);
            asm volatile("bar.arrive %0, 64;" ::"r"(1 + (k & 3)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// Faithful skeleton of the product kernel's pipeline, no real data: MMA warp (named-barrier gated),
// gatekeeper warp (mbarrier waits), a weight loader (32 KB cp.async.bulk per pass), and `nprod` fake
// producer groups of 128 threads that wait for the stage to be free, burn `work` cycles, and arrive.
// `stages` site stages of 2 halves; N columns per MMA.
__global__ void __launch_bounds__(640, 1) k_chain(int n, int passes, int stages, int work, int w_stages, const float *wsrc, long long *out)
{
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t x_full[4][2], x_empty[4], w_full[4], w_empty[4];
    __shared__ uint32_t s_tmem;
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 208 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&x_full[i][0])), "r"(128) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&x_full[i][1])), "r"(128) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&x_empty[i])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&w_full[i])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&w_empty[i])), "r"(1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    auto wait = [&](uint32_t b, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    };
    const uint32_t w_base = smem_u32(smem), x_base = smem_u32(smem + 64 * 1024);      // weights: 2 x 32 KB, sites: stages x (n*256 bytes) <= 144 KB
    const uint32_t x_stage = (uint32_t)n * 256u;       // hi + lo tiles of n rows x 128 bytes
    if (warp == 4) {
        const uint32_t idesc = make_idesc(0, 128, n);
        const long long t0 = clock64();
        for (int k = 0; k < passes; ++k) {
            const uint32_t sx = k % stages, sw = k % w_stages;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (k & 3)) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                const uint32_t xh = x_base + sx * x_stage, xl = xh + x_stage / 2, wh = w_base + sw * 32768u, wl = wh + 16384u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ko = ks * 32u;
                    mma<0>(tmem, make_desc_sw128(wh + ko), make_desc_sw128(xl + ko), idesc, (k | ks) != 0);
                    mma<0>(tmem, make_desc_sw128(wl + ko), make_desc_sw128(xh + ko), idesc, 1u);
                    mma<0>(tmem, make_desc_sw128(wh + ko), make_desc_sw128(xh + ko), idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&w_empty[sw])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&x_empty[sx])) : "memory");
            }
            __syncwarp();
        }
        for (int k = passes; k < passes + stages; ++k) wait(smem_u32(&x_empty[k % stages]), (uint32_t)(k / stages - 1) & 1u);
        const long long t1 = clock64();
        if (tid == 128) out[blockIdx.x] = t1 - t0;
    } else if (warp == 7) {
        for (int k = 0; k < passes; ++k) {
            const uint32_t sx = k % stages, sw = k % w_stages;
            wait(smem_u32(&x_full[sx][0]), (uint32_t)(k / stages) & 1u);
            wait(smem_u32(&x_full[sx][1]), (uint32_t)(k / stages) & 1u);
            wait(smem_u32(&w_full[sw]), (uint32_t)(k / w_stages) & 1u);
            asm volatile("bar.arrive %0, 64;" ::"r"(1 + (k & 3)) : "memory");
        }
    } else if (warp == 5) {
        if (tid == 160)
            for (int k = 0; k < passes; ++k) {
                const uint32_t sw = k % w_stages;
                if (k >= w_stages) wait(smem_u32(&w_empty[sw]), (uint32_t)(k / w_stages - 1) & 1u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&w_full[sw])), "r"(32768u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_base + sw * 32768u),
                             "l"(wsrc + (size_t)(k % 16) * 8192), "r"(32768u), "r"(smem_u32(&w_full[sw]))
                             : "memory");
            }
    } else if (warp >= 8) {
        const int g = (tid - 256) / 128;      // 3 groups; item q = 2*pass + half, group takes q % 3 == g
        for (uint32_t q = g; q < 2u * passes; q += 3) {
            const uint32_t k = q >> 1, h = q & 1, sx = k % stages;
            if (k >= (uint32_t)stages) wait(smem_u32(&x_empty[sx]), (k / stages - 1) & 1u);
            if (work) {
                const long long t0 = clock64();
                while (clock64() - t0 < work) {}
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&x_full[sx][h])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 160 * sizeof(long long));
    const int smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(k_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 512;      // multiple of 8
    const int grid = 148;
    cudaFuncSetAttribute(k_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int use_test : {0, 1})
    for (int depth : {2})
        for (int extra : {0, 1, 3, 6}) {
            long long h[160];
            const int passes = 2048;
            for (int rep = 0; rep < 2; ++rep) {
                k_pass<<<grid, 128, smem>>>(256, passes, depth, extra, out, use_test);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < grid; ++i) avg += (double)h[i];
            printf("pass-structured (test_wait %d) N 256 depth %d extra waits %d : %.1f cycles per pass of 12 MMAs (%.1f per MMA)\n", use_test, depth, extra, avg / grid / passes, avg / grid / passes / 12);
        }
    cudaFuncSetAttribute(k_pass_named, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int extra : {0, 3, 6}) {
        long long h[160];
        const int passes = 2048;
        for (int rep = 0; rep < 2; ++rep) {
            k_pass_named<<<grid, 128, smem>>>(256, passes, extra, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += (double)h[i];
        printf("pass-structured NAMED-barrier helper, extra waits %d : %.1f cycles per pass (%.1f per MMA)\n", extra, avg / grid / passes, avg / grid / passes / 12);
    }
    {
        const int smem2 = 209 * 1024 + 1024;
        cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        float *wsrc;
        cudaMalloc(&wsrc, 16 * 32768);
        cudaMemset(wsrc, 0, 16 * 32768);
        struct Cfg { int n, stages, work, wst; };
        const Cfg cfgs[] = {{256, 2, 0, 2}, {256, 2, 300, 2}, {256, 2, 600, 2}, {256, 2, 900, 2}, {256, 2, 1200, 2},
                            {256, 2, 1500, 2}, {192, 3, 0, 2}, {192, 3, 600, 2}, {192, 3, 1200, 2}, {192, 3, 1500, 2}, {192, 3, 1800, 2}, {192, 3, 2100, 2}, {128, 4, 600, 2}, {128, 4, 1500, 2}, {128, 4, 2100, 2}, {256, 2, 600, 1}};
        for (const Cfg &c : cfgs) {
            long long h[160];
            const int passes = 2048;
            for (int rep = 0; rep < 2; ++rep) {
                k_chain<<<grid, 640, smem2>>>(c.n, passes, c.stages, c.work, c.wst, wsrc, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < grid; ++i) avg += (double)h[i];
            const double per = avg / grid / passes;
            printf("chain N %3d stages %d weight stages %d producer work %4d : %.1f cycles per pass, %.2f cycles per site (ideal %.2f)\n", c.n, c.stages, c.wst,
                   c.work, per, per / (c.n / 2), 12.0 * (c.n / 2) / (c.n / 2));
        }
    }
    if (getenv("SKIP_WRITERS")) return 0;
    for (int writers : {0, 4, 8, 12, 16})
        for (int wr_sleep : {0, 100, 400})
            for (int n : {128, 256}) {
                if (writers == 0 && wr_sleep) continue;
                long long h[160];
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemset(out, 0, 160 * sizeof(long long));
                    k_bench<0><<<grid, 128 + 32 * writers, smem>>>(n, 128, iters, 2, out, wr_sleep);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, out, 160 * sizeof(long long), cudaMemcpyDeviceToHost);
                double avg = 0;
                for (int i = 0; i < grid; ++i) avg += (double)h[i];
                avg /= grid;
                const double stores = (double)h[148] * 4.0 * 32 * 16 / grid;     // bytes written per CTA by the writers
                printf("tf32 N %3d writers %2d warps sleep %3d : %.1f cycles per MMA, writers %.1f B/cycle/SM\n", n, writers, wr_sleep,
                       avg / (iters * 12.0), stores / avg);
            }
    return 0;
}
