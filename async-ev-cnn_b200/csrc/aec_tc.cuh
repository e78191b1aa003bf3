// aec_tc.cuh - the conv re-evaluation as a gathered implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Why tensor cores here: ncu on the SIMT version (profiles/r1a_summary.md) shows the gathered GEMM
// at 1-6 % DRAM, ~40 % FMA pipe, ~65 % L1/shared - a dense contraction bound by the SIMT FMA and
// shared-memory paths, which is the case BASELINE.json's north_star reserves tensor cores for.
//
//   rows   M = 128 per tile = 64 sites x {value row, rate row}   (conv2d.py:118-123: both maps use
//              the same gather addresses and the same weights)
//   K        = kh*kw*Cin ordered (ky,kx,ci), in blocks of 32 (one 128-byte swizzled smem row)
//   N        = Cout, tiles of <= 256 columns (TMEM columns)
//
// Precision: north_star asks for float32 maps within 1e-4 relative, and the oracle's pool ties must
// stay exact, so a bare TF32 product (10-bit mantissa) is not enough.  Every operand is split into
// hi = tf32(x) and lo = x - hi (exact in fp32) and three MMAs are issued per K step:
//   D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo        ("3xTF32", error ~2^-21 per product)
// Each accumulator row sees the same instruction sequence whatever its position in the tile, so
// identical patches still produce identical bits (what keeps exact pool ties exact).
//
// Data movement per K block: the A operand (gathered V = F*slope and R = A*slope of the previous
// layer, through the pool argmax when the previous layer is a pool) goes global -> registers ->
// hi/lo split -> swizzled shared memory; the B operand (weights, pre-split and pre-swizzled on the
// host into the exact shared-memory image) arrives with one bulk async copy (cp.async.bulk, the
// TMA engine's 1-D path) completing on an mbarrier.  MMAs are issued by one thread and signal
// stage reuse / accumulator readiness through tcgen05.commit -> mbarrier.
#pragma once
#include "aec_kernels.cuh"

namespace aec {
namespace tc {

constexpr int kTcThreads = 256;
constexpr int kTileSites = 64;                 // sites per tile -> 128 accumulator rows
constexpr int kBlockK = 32;                    // fp32 elements per K block (128-byte rows)
constexpr int kATileBytes = 128 * 128;         // one A tile (hi or lo): 128 rows x 128 bytes

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must fault the launch (reported by the C ABI), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gmem, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row
// swizzle atoms of 1024 bytes (SBO), descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    const uint32_t lo = ((smem_addr & 0x3ffffu) >> 4) | (1u << 16);              // start address, LBO (unused) = 1
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);                  // SBO = 1024 B, version = 1, SWIZZLE_128B
    return ((uint64_t)hi << 32) | lo;
}

// Instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// hi = x rounded to tf32 (10 explicit mantissa bits, low 13 bits zero), lo = x - hi exactly.
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __fsub_rn(x, hi);
}

struct TcParams {
    const uint32_t *sites;
    const int *counter;
    unsigned long long *accum;
    Src src;
    const float *wimg;     // [n_tiles][KB][2][Ntile][32] pre-split (hi, lo), pre-swizzled weight image
    const float *bias;     // [Npad]
    float *F, *A;
    long long fstride;
    int C, H, W;           // output map
    int K, KB;             // contraction length, number of 32-wide K blocks
    int Ntile, n_tiles;    // columns per tile (multiple of 16, <= 256), tiles along N
    int kh, kw, pad_t, pad_l;
    int stages;            // shared-memory pipeline depth
    int tmem_cols;         // power of two >= max(32, Ntile)
};

// byte offset of 16-byte chunk j of row r inside a 128-byte-swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void store_split(unsigned char *a_hi, unsigned char *a_lo, uint32_t off, const float4 &v)
{
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4 *>(a_hi + off) = h;
    *reinterpret_cast<float4 *>(a_lo + off) = l;
}

// One K block of the A operand: 64 sites x 8 chunks of 4 channels; thread t owns chunk t&7 of
// sites t>>3 and 32 + (t>>3).  Value rows are 0..63, rate rows 64..127.
template <int KIND>
__device__ __forceinline__ void gather_kblock(const TcParams &p, int kb, unsigned char *a_hi, unsigned char *a_lo,
                                              const int *s_str, const int *s_y, const int *s_x)
{
    const int tid = threadIdx.x;
    const int j = tid & 7;
    const int k = kb * kBlockK + 4 * j;
    const int Cin = p.src.C;
    const int tap = k / Cin, c = k - tap * Cin;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    const bool kvalid = k < p.K;
    float4 f[2], a[2];
    bool ok[2];
    long long base[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = (tid >> 3) + 32 * u;
        const int s = s_str[i];
        const int iy = s_y[i] + ky - p.pad_t, ix = s_x[i] + kx - p.pad_l;
        ok[u] = kvalid && s >= 0 && iy >= 0 && iy < p.src.H && ix >= 0 && ix < p.src.W;
        f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[u] = f[u];
        base[u] = 0;
        if (ok[u]) {
            if (KIND == 1) {
                const long long off = (long long)s * p.src.fstride + ((long long)iy * p.src.W + ix) * Cin + c;
                f[u] = __ldg(reinterpret_cast<const float4 *>(p.src.F + off));
                a[u] = __ldg(reinterpret_cast<const float4 *>(p.src.A + off));
            } else {
                base[u] = (long long)s * p.src.istride + ((long long)iy * p.src.W + ix) * Cin + c;
            }
        }
    }
    if (KIND == 2) {
        uchar4 id[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            id[u] = make_uchar4(0, 0, 0, 0);
            if (ok[u]) id[u] = __ldg(reinterpret_cast<const uchar4 *>(p.src.idx + base[u]));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            const int i = (tid >> 3) + 32 * u;
            const int iy = s_y[i] + ky - p.pad_t, ix = s_x[i] + kx - p.pad_l;
            const float *Fb = p.src.F + (long long)s_str[i] * p.src.fstride;
            const float *Ab = p.src.A + (long long)s_str[i] * p.src.fstride;
            const int y0 = iy * p.src.pstride, x0 = ix * p.src.pstride;
            const unsigned char ids[4] = {id[u].x, id[u].y, id[u].z, id[u].w};
            float fv[4], av[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int w = ids[e];
                const int dy = w / p.src.pkw, dx = w - dy * p.src.pkw;
                const long long off = ((long long)(y0 + dy) * p.src.cW + (x0 + dx)) * Cin + c + e;
                fv[e] = __ldg(Fb + off);
                av[e] = __ldg(Ab + off);
            }
            f[u] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            a[u] = make_float4(av[0], av[1], av[2], av[3]);
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = (tid >> 3) + 32 * u;
        float4 v, r;
        float sl;
        sl = slope_of(f[u].x, p.src.alpha); v.x = __fmul_rn(f[u].x, sl); r.x = __fmul_rn(a[u].x, sl);
        sl = slope_of(f[u].y, p.src.alpha); v.y = __fmul_rn(f[u].y, sl); r.y = __fmul_rn(a[u].y, sl);
        sl = slope_of(f[u].z, p.src.alpha); v.z = __fmul_rn(f[u].z, sl); r.z = __fmul_rn(a[u].z, sl);
        sl = slope_of(f[u].w, p.src.alpha); v.w = __fmul_rn(f[u].w, sl); r.w = __fmul_rn(a[u].w, sl);
        store_split(a_hi, a_lo, sw128_off(i, j), v);
        store_split(a_hi, a_lo, sw128_off(kTileSites + i, j), r);
    }
}

// Dynamic shared memory: [pad to 1024][stages x {A_hi 16K, A_lo 16K, B_hi Ntile*128, B_lo Ntile*128}]
template <int KIND>
__global__ void __launch_bounds__(kTcThreads) k_conv_eval_tc(const __grid_constant__ TcParams p)
{
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[8], bar_free[8], bar_acc;
    __shared__ uint32_t s_tmem;
    __shared__ int s_str[kTileSites], s_y[kTileSites], s_x[kTileSites];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_sites = *p.counter;
    if (blockIdx.x == 0 && tid == 0 && n_sites > 0) atomicAdd(p.accum, (unsigned long long)n_sites);
    const int m_tiles = (n_sites + kTileSites - 1) / kTileSites;
    const int total_tiles = m_tiles * p.n_tiles;
    if ((int)blockIdx.x >= total_tiles) return;       // uniform per CTA: nothing allocated yet

    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)p.Ntile * 128u;            // one B tile (hi or lo)
    const uint32_t stage_bytes = 2u * kATileBytes + 2u * b_bytes;

    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&bar_full[i]), 1);
            mbar_init(smem_u32(&bar_free[i]), 1);
        }
        mbar_init(smem_u32(&bar_acc), 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t idesc = make_idesc_tf32(p.Ntile);
    const int HW = p.H * p.W;

    uint32_t it = 0;          // K-block iterations issued by this CTA (drives stage index and parities)
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
        const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
        for (int i = tid; i < kTileSites; i += kTcThreads) {
            const int gi = mt * kTileSites + i;
            if (gi < n_sites) {
                const uint32_t e = p.sites[gi];
                const int s = (int)(e / (uint32_t)HW);
                const int site = (int)(e - (uint32_t)s * (uint32_t)HW);
                s_str[i] = s;
                s_y[i] = site / p.W;
                s_x[i] = site - (site / p.W) * p.W;
            } else {
                s_str[i] = -1;
                s_y[i] = 0;
                s_x[i] = 0;
            }
        }
        __syncthreads();

        const float *wtile = p.wimg + (size_t)nt * p.KB * 2 * p.Ntile * kBlockK;
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
            const int st = (int)(it % (uint32_t)p.stages);
            const uint32_t use = it / (uint32_t)p.stages;
            unsigned char *stage = smem + (size_t)st * stage_bytes;
            if (use > 0) mbar_wait(smem_u32(&bar_free[st]), (use - 1) & 1);     // MMAs that read this stage are done
            if (tid == 0) {
                mbar_expect_tx(smem_u32(&bar_full[st]), 2u * b_bytes);
                bulk_g2s(smem_u32(stage + 2 * kATileBytes), wtile + (size_t)kb * 2 * p.Ntile * kBlockK, 2u * b_bytes,
                         smem_u32(&bar_full[st]));
            }
            gather_kblock<KIND>(p, kb, stage, stage + kATileBytes, s_str, s_y, s_x);
            fence_proxy_async();          // generic-proxy writes of A -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid == 0) {
                mbar_wait(smem_u32(&bar_full[st]), use & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage), a_lo = a_hi + kATileBytes;
                const uint32_t b_hi = a_hi + 2 * kATileBytes, b_lo = b_hi + b_bytes;
#pragma unroll
                for (int ks = 0; ks < kBlockK / 8; ++ks) {
                    const uint32_t ko = (uint32_t)ks * 32u;           // 8 tf32 = 32 bytes along K inside the swizzled row
                    const uint64_t dah = make_desc_sw128(a_hi + ko), dal = make_desc_sw128(a_lo + ko);
                    const uint64_t dbh = make_desc_sw128(b_hi + ko), dbl = make_desc_sw128(b_lo + ko);
                    mma_tf32(tmem_base, dal, dbh, idesc, (kb | ks) != 0 ? 1u : 0u);
                    mma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    mma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                mma_commit(smem_u32(&bar_free[st]));
                if (kb == p.KB - 1) mma_commit(smem_u32(&bar_acc));
            }
        }

        // ---- epilogue: TMEM -> registers -> F (value rows, + bias) / A (rate rows)
        mbar_wait(smem_u32(&bar_acc), tile_iter & 1);
        tc_fence_after();
        {
            const int row = (warp & 3) * 32 + lane;             // TMEM lane == accumulator row
            const bool is_rate = row >= kTileSites;
            const int si = row & (kTileSites - 1);
            const int s = s_str[si];
            float *dst = nullptr;
            if (s >= 0) dst = (is_rate ? p.A : p.F) + (long long)s * p.fstride + ((long long)s_y[si] * p.W + s_x[si]) * p.C;
            const int half = p.Ntile >> 1;                       // warps 0-3: first half of the columns, 4-7: second half
            const int c0 = (warp >> 2) * half;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            const bool vec_ok = (p.C & 3) == 0;
            for (int cc = 0; cc < half; cc += 8) {
                uint32_t r[8];
                tmem_ld8(taddr + (uint32_t)(c0 + cc), r);
                tmem_ld_wait();
                if (dst) {
                    const int n = nt * p.Ntile + c0 + cc;
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[e]);
                    if (!is_rate) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(o[e], __ldg(p.bias + n + e));
                    }
                    if (vec_ok && n + 8 <= p.C) {
                        *reinterpret_cast<float4 *>(dst + n) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4 *>(dst + n + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (n + e < p.C) dst[n + e] = o[e];
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();          // TMEM drained and site list free before the next tile overwrites them
        tc_fence_after();
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

}  // namespace tc
}  // namespace aec
