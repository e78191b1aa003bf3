// aec_tc.cuh - the conv re-evaluation as a gathered implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Why tensor cores here: ncu on the SIMT version (profiles/r1a_summary.md) shows the gathered GEMM
// at 1-6 % DRAM, ~40 % FMA pipe, ~65 % L1/shared - a dense contraction bound by the SIMT FMA and
// shared-memory paths, which is the case BASELINE.json's north_star reserves tensor cores for.
//
//   rows   M = 128 per m-tile = 64 sites x {value row, rate row}   (conv2d.py:118-123: both maps
//              use the same gather addresses and the same weights)
//   K        = kh*kw*Cin ordered (ky,kx,ci), in blocks of 32 (one 128-byte swizzled smem row)
//   N        = Cout in n-tiles of <= 128 columns
//   unit     = up to 4 m-tiles x one n-tile: the 4 accumulators (4 x Ntile TMEM columns) share every
//              weight K block, so weights cross L2 -> shared memory once per 256 sites
//
// Precision: north_star asks for float32 maps within 1e-4 relative, and the oracle's pool ties must
// stay exact, so a bare TF32 product (10-bit mantissa) is not enough.  Every operand is split into
// hi = tf32(x) and lo = x - hi (exact in fp32) and three MMAs are issued per K step:
//   D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        ("3xTF32", error ~2^-21 per product)
// Each accumulator row sees the same instruction sequence whatever its position in a tile, so
// identical patches still produce identical bits (what keeps exact pool ties exact).
//
// Warp roles of the persistent CTA (one per SM, 640 threads):
//   warps 0-3   epilogue: TMEM -> registers -> F (+bias) / A, one accumulator row per thread
//   warp  4     MMA issuer (one lane): waits operand barriers, issues tcgen05.mma, commits
//   warp  5     weight loader (one lane): cp.async.bulk of the pre-split, pre-swizzled weight image
//   warp  6     site decoder: work-list entries -> (stream, y, x) in shared memory, double buffered
//   warps 8-19  gather producers, 3 groups of 4 warps taking (K block, m-tile) items round-robin:
//               global -> registers (V = F*slope, R = A*slope) -> hi/lo split -> swizzled smem
// Pipelines (all mbarrier based): A stages (producers <-> MMA), B stages (loader <-> MMA),
// accumulator buffers (MMA <-> epilogue; two when 2 x 4 x Ntile <= 512 columns), site-info buffers.
#pragma once
#include "aec_kernels.cuh"

namespace aec {
namespace tc {

constexpr int kTcThreads = 640;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4, kLoadWarp = 5, kSiteWarp = 6;
constexpr int kProdWarp0 = 8;
constexpr int kGroups = 3, kGroupThreads = 128;
constexpr int kTileSites = 64;                 // sites per m-tile -> 128 accumulator rows
constexpr int kMT = 4;                         // m-tiles per unit
constexpr int kUnitSites = kMT * kTileSites;
constexpr int kBlockK = 32;                    // fp32 elements per K block (128-byte rows)
constexpr int kATileBytes = 128 * 128;         // one A tile (hi or lo): 128 rows x 128 bytes
constexpr int kAStageBytes = 2 * kATileBytes;
constexpr int kMaxNtile = 128;
constexpr int kMaxStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
    const float *srcF, *srcA;   // previous layer's channel-last (F, A) pair (conv state or pool copy)
    long long src_stride;       // floats per stream
    float alpha;                // previous layer's activation slope
    int Cin, Hin, Win;
    const float *wimg;          // [n_tiles][KB][2][Ntile][32] pre-split (hi, lo), pre-swizzled weight image
    const float *bias;          // [Ntile * n_tiles]
    float *F, *A;
    long long fstride;
    int C, H, W;                // output map
    int K, KB;                  // contraction length, number of 32-wide K blocks
    int Ntile, n_tiles;         // columns per n-tile (multiple of 16, <= 128), n-tiles
    int kh, kw, pad_t, pad_l;
    int a_stages, b_stages;     // shared-memory pipeline depths
    int n_acc;                  // accumulator buffers in TMEM (1 or 2)
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

// One (K block, m-tile) item of the A operand: 64 sites x 8 chunks of 4 channels gathered by one
// producer group; thread t owns chunk t&7 of sites (t>>3) + 16u, u = 0..3.  Value rows are 0..63,
// rate rows 64..127 of the tile.
__device__ __forceinline__ void gather_item(const TcParams &p, int kb, int m, const int *ss, const int *syx, unsigned char *a_hi,
                                            unsigned char *a_lo, int t)
{
    const int j = t & 7, tq = t >> 3;
    const int k = kb * kBlockK + 4 * j;
    const int tap = k / p.Cin, c = k - tap * p.Cin;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    const int dy = ky - p.pad_t, dx = kx - p.pad_l;
    const bool kvalid = k < p.K;
    const int rel = (dy * p.Win + dx) * p.Cin + c;
    float4 f[4], a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = m * kTileSites + tq + 16 * u;
        const int s = ss[i];
        const int yx = syx[i];
        const int y = yx >> 16, x = yx & 0xffff;
        const bool ok = kvalid && s >= 0 && (unsigned)(y + dy) < (unsigned)p.Hin && (unsigned)(x + dx) < (unsigned)p.Win;
        f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[u] = f[u];
        if (ok) {
            const long long off = (long long)s * p.src_stride + ((y * p.Win + x) * p.Cin + rel);
            f[u] = __ldg(reinterpret_cast<const float4 *>(p.srcF + off));
            a[u] = __ldg(reinterpret_cast<const float4 *>(p.srcA + off));
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int r = tq + 16 * u;
        float4 v, w;
        float sl;
        sl = slope_of(f[u].x, p.alpha); v.x = __fmul_rn(f[u].x, sl); w.x = __fmul_rn(a[u].x, sl);
        sl = slope_of(f[u].y, p.alpha); v.y = __fmul_rn(f[u].y, sl); w.y = __fmul_rn(a[u].y, sl);
        sl = slope_of(f[u].z, p.alpha); v.z = __fmul_rn(f[u].z, sl); w.z = __fmul_rn(a[u].z, sl);
        sl = slope_of(f[u].w, p.alpha); v.w = __fmul_rn(f[u].w, sl); w.w = __fmul_rn(a[u].w, sl);
        store_split(a_hi, a_lo, sw128_off(r, j), v);
        store_split(a_hi, a_lo, sw128_off(kTileSites + r, j), w);
    }
}

// Dynamic shared memory: [pad to 1024][a_stages x {A_hi 16K, A_lo 16K}][b_stages x {B_hi, B_lo: Ntile*128 each}]
__global__ void __launch_bounds__(kTcThreads, 1) k_conv_eval_tc(const __grid_constant__ TcParams p)
{
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t bar_a_full[kMaxStages], bar_a_empty[kMaxStages], bar_b_full[kMaxStages], bar_b_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_si_full[2], bar_si_free[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_str[2][kUnitSites], s_yx[2][kUnitSites];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_sites = *p.counter;
    if (blockIdx.x == 0 && tid == 0 && n_sites > 0) atomicAdd(p.accum, (unsigned long long)n_sites);
    const int m_tiles = (n_sites + kTileSites - 1) / kTileSites;
    // m-tiles per unit: 4 when there is enough work to keep every CTA busy, fewer for short lists
    int mte = kMT;
    while (mte > 1 && (long long)((m_tiles + mte - 1) / mte) * p.n_tiles < 4LL * gridDim.x) mte >>= 1;
    const int m_groups = (m_tiles + mte - 1) / mte;
    const int total_units = m_groups * p.n_tiles;
    if ((int)blockIdx.x >= total_units) return;       // uniform per CTA: nothing allocated yet

    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_b = smem + (size_t)p.a_stages * kAStageBytes;
    const uint32_t b_bytes = (uint32_t)p.Ntile * 128u;            // one B tile (hi or lo)

    if (tid == 0) {
        for (int i = 0; i < p.a_stages; ++i) {
            mbar_init(smem_u32(&bar_a_full[i]), kGroupThreads);
            mbar_init(smem_u32(&bar_a_empty[i]), 1);
        }
        for (int i = 0; i < p.b_stages; ++i) {
            mbar_init(smem_u32(&bar_b_full[i]), 1);
            mbar_init(smem_u32(&bar_b_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_acc_full[i]), 1);
            mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps * 32);
            mbar_init(smem_u32(&bar_si_full[i]), 1);
            mbar_init(smem_u32(&bar_si_free[i]), kEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int HW = p.H * p.W;
    const int n_units_cta = (total_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < kEpiWarps) {
        // ===================== epilogue =====================
        const int row = warp * 32 + lane;                   // TMEM lane == accumulator row
        const bool is_rate = row >= kTileSites;
        const int sr = row & (kTileSites - 1);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float *const dst_map = is_rate ? p.A : p.F;
        const bool vec_ok = (p.C & 3) == 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = blockIdx.x + ul * gridDim.x;
            const int mg = unit / p.n_tiles, nt = unit - mg * p.n_tiles;
            const int mt_count = min(mte, m_tiles - mg * mte);
            const int buf = ul & 1, ab = ul % p.n_acc;
            mbar_wait(smem_u32(&bar_si_full[buf]), (uint32_t)(ul >> 1) & 1u);
            mbar_wait(smem_u32(&bar_acc_full[ab]), (uint32_t)(ul / p.n_acc) & 1u);
            tc_fence_after();
            for (int m = 0; m < mt_count; ++m) {
                const int s = s_str[buf][m * kTileSites + sr];
                const int yx = s_yx[buf][m * kTileSites + sr];
                float *dst = nullptr;
                if (s >= 0) dst = dst_map + (long long)s * p.fstride + (long long)((yx >> 16) * p.W + (yx & 0xffff)) * p.C;
                const uint32_t col0 = (uint32_t)(ab * kMT * p.Ntile + m * p.Ntile);
                for (int cc = 0; cc < p.Ntile; cc += 16) {
                    uint32_t r[16];
                    tmem_ld16(lane_addr + col0 + (uint32_t)cc, r);
                    tmem_ld_wait();
                    if (dst) {
                        const int n = nt * p.Ntile + cc;
                        float o[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(r[e]);
                        if (!is_rate) {
#pragma unroll
                            for (int e = 0; e < 16; e += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4 *>(p.bias + n + e));
                                o[e] = __fadd_rn(o[e], bv.x); o[e + 1] = __fadd_rn(o[e + 1], bv.y);
                                o[e + 2] = __fadd_rn(o[e + 2], bv.z); o[e + 3] = __fadd_rn(o[e + 3], bv.w);
                            }
                        }
                        if (vec_ok && n + 16 <= p.C) {
#pragma unroll
                            for (int e = 0; e < 16; e += 4)
                                *reinterpret_cast<float4 *>(dst + n + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (n + e < p.C) dst[n + e] = o[e];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bar_acc_empty[ab]));
            mbar_arrive(smem_u32(&bar_si_free[buf]));
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.Ntile);
            uint32_t q = 0, qb = 0;
            for (int ul = 0; ul < n_units_cta; ++ul) {
                const int unit = blockIdx.x + ul * gridDim.x;
                const int mg = unit / p.n_tiles;
                const int mt_count = min(mte, m_tiles - mg * mte);
                const int ab = ul % p.n_acc;
                const uint32_t ua = (uint32_t)(ul / p.n_acc);
                if (ua > 0) mbar_wait(smem_u32(&bar_acc_empty[ab]), (ua - 1) & 1u);    // epilogue drained this buffer
                tc_fence_after();
                for (int kb = 0; kb < p.KB; ++kb, ++qb) {
                    const int sb = (int)(qb % (uint32_t)p.b_stages);
                    mbar_wait(smem_u32(&bar_b_full[sb]), (qb / (uint32_t)p.b_stages) & 1u);
                    const uint32_t b_hi = smem_u32(smem_b + (size_t)sb * 2 * b_bytes), b_lo = b_hi + b_bytes;
                    for (int m = 0; m < mt_count; ++m, ++q) {
                        const int sa = (int)(q % (uint32_t)p.a_stages);
                        mbar_wait(smem_u32(&bar_a_full[sa]), (q / (uint32_t)p.a_stages) & 1u);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + (size_t)sa * kAStageBytes), a_lo = a_hi + kATileBytes;
                        const uint32_t d = tmem_base + (uint32_t)(ab * kMT * p.Ntile + m * p.Ntile);
#pragma unroll
                        for (int ks = 0; ks < kBlockK / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * 32u;       // 8 tf32 = 32 bytes along K inside the swizzled row
                            const uint64_t dah = make_desc_sw128(a_hi + ko), dal = make_desc_sw128(a_lo + ko);
                            const uint64_t dbh = make_desc_sw128(b_hi + ko), dbl = make_desc_sw128(b_lo + ko);
                            mma_tf32(d, dal, dbh, idesc, (kb | ks) != 0 ? 1u : 0u);
                            mma_tf32(d, dah, dbl, idesc, 1u);
                            mma_tf32(d, dah, dbh, idesc, 1u);
                        }
                        mma_commit(smem_u32(&bar_a_empty[sa]));
                    }
                    mma_commit(smem_u32(&bar_b_empty[sb]));
                }
                mma_commit(smem_u32(&bar_acc_full[ab]));
            }
        }
    } else if (warp == kLoadWarp) {
        // ===================== weight loader =====================
        if (lane == 0) {
            uint32_t qb = 0;
            for (int ul = 0; ul < n_units_cta; ++ul) {
                const int unit = blockIdx.x + ul * gridDim.x;
                const int nt = unit % p.n_tiles;
                const float *wtile = p.wimg + (size_t)nt * p.KB * 2 * p.Ntile * kBlockK;
                for (int kb = 0; kb < p.KB; ++kb, ++qb) {
                    const int sb = (int)(qb % (uint32_t)p.b_stages);
                    const uint32_t use = qb / (uint32_t)p.b_stages;
                    if (use > 0) mbar_wait(smem_u32(&bar_b_empty[sb]), (use - 1) & 1u);
                    mbar_expect_tx(smem_u32(&bar_b_full[sb]), 2u * b_bytes);
                    bulk_g2s(smem_u32(smem_b + (size_t)sb * 2 * b_bytes), wtile + (size_t)kb * 2 * p.Ntile * kBlockK, 2u * b_bytes,
                             smem_u32(&bar_b_full[sb]));
                }
            }
        }
    } else if (warp == kSiteWarp) {
        // ===================== site decoder =====================
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = blockIdx.x + ul * gridDim.x;
            const int mg = unit / p.n_tiles;
            const int mt_count = min(mte, m_tiles - mg * mte);
            const int buf = ul & 1;
            const uint32_t us = (uint32_t)(ul >> 1);
            if (us > 0) mbar_wait(smem_u32(&bar_si_free[buf]), (us - 1) & 1u);
            for (int i = lane; i < kUnitSites; i += 32) {
                const long long gi = (long long)mg * mte * kTileSites + i;
                int s = -1, yx = 0;
                if (i < mt_count * kTileSites && gi < n_sites) {
                    const uint32_t e = p.sites[gi];
                    s = (int)(e / (uint32_t)HW);
                    const int site = (int)(e - (uint32_t)s * (uint32_t)HW);
                    const int y = site / p.W;
                    yx = (y << 16) | (site - y * p.W);
                }
                s_str[buf][i] = s;
                s_yx[buf][i] = yx;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_si_full[buf]));
        }
    } else if (warp >= kProdWarp0) {
        // ===================== gather producers =====================
        const int pt = tid - kProdWarp0 * 32;
        const int g = pt / kGroupThreads, t = pt - g * kGroupThreads;
        uint32_t q = 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = blockIdx.x + ul * gridDim.x;
            const int mg = unit / p.n_tiles;
            const int mt_count = min(mte, m_tiles - mg * mte);
            const int buf = ul & 1;
            bool waited = false;
            for (int kb = 0; kb < p.KB; ++kb) {
                for (int m = 0; m < mt_count; ++m, ++q) {
                    if ((int)(q % (uint32_t)kGroups) != g) continue;
                    if (!waited) {
                        mbar_wait(smem_u32(&bar_si_full[buf]), (uint32_t)(ul >> 1) & 1u);
                        waited = true;
                    }
                    const int sa = (int)(q % (uint32_t)p.a_stages);
                    const uint32_t use = q / (uint32_t)p.a_stages;
                    if (use > 0) mbar_wait(smem_u32(&bar_a_empty[sa]), (use - 1) & 1u);   // MMAs that read this stage are done
                    unsigned char *a_hi = smem + (size_t)sa * kAStageBytes;
                    gather_item(p, kb, m, s_str[buf], s_yx[buf], a_hi, a_hi + kATileBytes, t);
                    fence_proxy_async();      // generic-proxy writes of A -> visible to the tensor core (async proxy)
                    mbar_arrive(smem_u32(&bar_a_full[sa]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace tc
}  // namespace aec
