// aec_tc.cuh - the conv re-evaluation as a gathered implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Why tensor cores here: ncu on the SIMT version (profiles/r1a_summary.md) shows the gathered GEMM
// at 1-6 % DRAM, ~40 % FMA pipe, ~65 % L1/shared - a dense contraction bound by the SIMT FMA and
// shared-memory paths, which is the case BASELINE.json's north_star reserves tensor cores for.
//
// Orientation (round-1c): the WEIGHTS are the M operand and the gathered sites the N operand,
//     D[c, n] = sum_k W[k, c] * X[n, k]        M = 128 output channels (one weight tile)
//                                              N = 256 = 128 sites x {value row, rate row}
//                                              K = kh*kw*Cin ordered (ky,kx,ci), blocks of 32
// (conv2d.py:118-123: the value map and the rate map use the same gather addresses and the same
// weights).  Measured (tools/bench_umma.cu): a tcgen05.mma with M = 128 takes max(64, N/2) cycles, so
// N = 256 instructions carry twice the work of the N <= 128 ones of the sites-as-M orientation at the
// same issue cost, the accumulator comes out channel-per-lane - the epilogue writes 32 consecutive channels
// of one site per warp store (one 128-byte line) - and two weight tiles can share one gather (Cout >= 256).
// A layer with 32 output channels repeats them 4x along M (the rows are free: M is 128 anyway) so that
// all four epilogue warps share its 256 columns.
//
// What bounds it (profiles/r1c_summary.md, r1d_summary.md): per K block the tensor core reads 12 x 12 KB
// of operands, the producers store a 64 KB site stage and the loader 32 KB of weights - 240 KB per 1536
// cycles, the measured shared-memory ceiling (~94 B/cycle of MMA reads + ~64 B/cycle of stores).
// At steady state (profiles/r1e_summary.md) the short-unit layers are additionally bounded by the support roles:
// with gathers, stores and MMAs knocked out conv2 still takes 72 % of its time (decoder ~8.5 k cycles per unit,
// epilogue ~7.4 k, against a 10.9 k-cycle unit).
//
// Narrow layers (Cout <= 64, round 2): `kSM` = SITES are the M operand.  With the weights as M a 32- or 64-channel
// layer fills a quarter or half of the 128 rows the tensor core computes anyway, and the profile of round 1
// (profiles/r2_summary.md) shows the kernel bounded by shared-memory bandwidth, a large part of it the tensor core's
// own operand reads (12 KB per MMA).  Sites-as-M turns that around:
//     D_v[site, n] = sum_k Xv[site, k] * Wcat[n, k]      M = 128 sites (value rows; D_r likewise for the rate rows)
//                                                        N = 2*Cpad: Wcat = [W_hi ; W_lo] stacked along N
//   per 8-wide K step and 128 sites:  Xv_hi x Wcat (N = 2*Cpad)  +  Xv_lo x W_hi (N = Cpad, same columns as the hi.hi part)
//                                     and the same two for the rate rows: 4 MMAs of 64 cycles (N <= 128) = 2 cycles per
//   site instead of 3, 22-28 KB of operand reads instead of 36, and the three products of the 3xTF32 scheme
//   (hi.hi, hi.lo, lo.hi) are still all there.  The accumulator comes out site-per-lane: an epilogue thread owns one
//   site, adds the two column halves (W_hi part + W_lo part) and writes the site's channels as contiguous 16-byte
//   stores.  The gather producers, the stage format and all pipelines are shared with the weights-as-M form; only
//   where a value / rate row lands in the stage differs (tile_v = rows of 128 sites, tile_r behind it).
//
// Precision: north_star asks for float32 maps within 1e-4 relative, and the oracle's pool ties must
// stay exact, so a bare TF32 product (10-bit mantissa) is not enough.  Every operand is split into
// hi = tf32(x) and lo = x - hi (exact in fp32) and three MMAs are issued per K step:
//   D += W_hi.X_lo + W_lo.X_hi + W_hi.X_hi        ("3xTF32", error ~2^-21 per product)
// Each accumulator column sees the same instruction sequence whatever its position in a tile, so
// identical patches still produce identical bits (what keeps exact pool ties exact).
//
// Warp roles of the persistent CTA (one per SM, 640 threads):
//   warps 0-3   epilogue: TMEM -> registers -> F (+bias) / A, one output channel per thread
//   warp  4     MMA issuer: a converged warp, one elected lane issues tcgen05.mma and the commits
//   warp  7     gatekeeper: does every mbarrier wait on the MMA warp's behalf (named-barrier hand-off)
//   warp  5     weight loader (one lane): cp.async.bulk of the pre-split, pre-swizzled weight image
//   warp  6     site decoder: work-list entries -> source pointer, in-map tap bits, output offset (ring of 4 units)
//   warps 8-19  gather producers, 3 groups of 4 warps taking (K block, half) items round-robin:
//               global -> registers (V = F*slope, R = A*slope) -> hi/lo split -> swizzled smem;
//               the next item's loads are issued before the current one is converted (setmaxnreg: 112 regs)
// Pipelines (all mbarrier based): site stages (producers <-> MMA), weight stages (loader <-> MMA),
// accumulator buffers (MMA <-> epilogue), site-info buffers (decoder <-> producers/epilogue).
#pragma once
#include "aec_kernels.cuh"

namespace aec {
namespace tc {

constexpr int kTcThreads = 640;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4, kLoadWarp = 5, kSiteWarp = 6, kGateWarp = 7;
constexpr int kProdWarp0 = 8;
constexpr int kGroups = 3, kGroupThreads = 128; // gather warps 8-19: 3 groups of 4 warps
constexpr int kPairs = 4;                      // (site, chunk) pairs per producer thread and item
constexpr int kItemSites = 64;                 // sites per half stage -> 128 operand rows (value rows, then rate rows)
constexpr int kUnitSites = 128;                // sites per unit -> N = 256
constexpr int kUnitCols = 256;                 // accumulator columns per (unit, weight tile)
constexpr int kBlockK = 32;                    // fp32 elements per K block (128-byte rows)
constexpr int kItemTileBytes = 128 * 128;      // one item's hi (or lo) rows: 128 rows x 128 bytes
constexpr int kSiteTileBytes = 2 * kItemTileBytes;   // hi (or lo) tile of a site stage: 256 rows
constexpr int kSiteStageBytes = 2 * kSiteTileBytes;  // hi + lo
constexpr int kSiteStages = 2;
constexpr int kMaxSiteStages = 4;              // barrier array size
constexpr int kPairSiteStages = 3;             // one-item stages (half units, CTA pairs): three of 32 KB (hi, lo); the rest of the shared memory holds weight stages
constexpr int kMaxWStages = 4;
constexpr int kMaxMtu = 2;                     // weight tiles per unit (2 x 256 accumulator columns)
constexpr int kSiteRing = 4;                   // site-info buffers: producers may run this many units ahead of the epilogue
constexpr int kKtabBlocks = 64;                // K blocks covered by the shared-memory gather table (larger K: computed on the fly)
// Registers per thread after the role split (setmaxnreg): the producers double buffer a whole item
// (2 x 8 float4) in registers so that their global loads are in flight while the previous item is
// converted and while the shared-memory stage is still owned by the tensor core.
constexpr int kRegsEpi = 80, kRegsCtl = 64, kRegsProd = 112;

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
__device__ __forceinline__ uint32_t make_idesc_tf32(int n, int m = 128)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// True in exactly one lane of a converged warp.  The MMA / loader warps run their loops converged on
// warp-uniform values and guard only the issuing instructions with this, so ptxas keeps descriptors
// in uniform registers and emits back-to-back UTCHMMA (a `lane == 0` branch around the whole loop
// costs ~10 bookkeeping instructions and a per-instruction ELECT/BRA loop for every MMA).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
// ---- CTA pair (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
        : "memory");
}
// wait (bounded) that also acquires what a thread of the other CTA released before its remote arrive
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) __trap();
    }
}
// D (M = 256: rows 0-127 in this CTA's TMEM, 128-255 in the peer's) += A (128 rows from each CTA's shared memory) x B (N/2 rows from each)
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of the pair's MMAs -> the mbarrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
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
    const float *srcF;          // previous layer's channel-last F map (conv state or pool copy); A = F + a_minus_f bytes
    long long a_minus_f;        // byte distance from the source F map to the source A map (same layout)
    const char *zero_f;         // the 128 zero bytes in front of the source F map (in front of A after adding a_minus_f)
    long long src_stride;       // floats per stream
    float alpha;                // previous layer's activation slope
    int Cin, Hin, Win;
    const float *wimg;          // [m_tiles][KB][2][Mrows][32] pre-split (hi, lo), pre-swizzled weight image
    const float *bias;          // [Mrows * m_tiles]
    float *F, *A;
    long long fstride;
    int C, H, W;                // output map
    int K, KB;                  // contraction length, number of 32-wide K blocks
    int ks_last;                // 8-wide MMA steps of the last K block: ceil((K - 32*(KB-1)) / 8), the rest is padding
    int Mrows, m_tiles;         // rows per weight tile (multiple of 8, <= 128) = Mch * rep, weight tiles
    int Mch, rep;               // output channels per weight tile; rep (1, 2, 4) copies of them along M so that the
                                // epilogue of a narrow layer spreads over rep x Mch/32 warps (copy g stores the 16-column
                                // chunks with chunk % rep == g; the extra rows cost the tensor core nothing, M is 128 anyway)
    int mtu;                    // weight tiles per unit (1 or 2): the unit's sites are gathered once for all of them
    int kh, kw, pad_t, pad_l;
    SiteCode code;              // work-list entry coding of the output layer
    int w_stages;               // weight pipeline depth
    int w_stages_half;          // ... when the launch runs on half units (three 32 KB site stages instead of 128 KB: more weight stages fit)
    int n_acc;                  // accumulator buffers in TMEM (2 when mtu == 1, else 1)
    // kPool (weights-as-M only): the 2x2 / stride-2 pool behind this layer is evaluated here for the windows that arrive as quads
    // (four consecutive work-list entries, the first flagged with quad_bit: emit_sites_quads) - maxpool.py:130-151, cutils.pyx:161-177
    uint32_t quad_bit;
    const int *site_counter;    // real work-set sites (the list also holds padding entries)
    uint8_t *pool_idx;          // [S][pool_stride] argmax rows
    float *pool_Fp, *pool_Ap;   // [S][pool_stride] (F, A) copies at the argmax
    long long pool_stride;
    uint32_t *pool_flags;       // [S][pHWw] sticky recompute flags: set where a window comes out unstable
    unsigned long long *pool_accum;   // the pool layer's work counter; [32] behind it: windows evaluated here
    int pW, pWw, pHWw;
    float pool_alpha;           // this layer's activation slope (the pool ranks rates R = A * slope(F))
    unsigned long long *timing; // null, or 16 cycle counters accumulated over CTAs (aec_net_tc_timing): see TcTimingSlot
    int half_units;             // weights-as-M: 0 never, 1 = units of 64 sites (N = 128) when full units would leave CTAs idle (few streams)
    int debug;                  // AEC_TC_DEBUG experiment bits (results invalid when non-zero): 1 no gather loads, 2 no operand stores, 4 no MMA, 8 no epilogue stores, 16 no weight copies
};

// Cycle accounting per warp role (measurement only; active when TcParams::timing != nullptr).
enum TcTimingSlot { kTMmaTotal = 0, kTMmaWaitAcc, kTMmaWaitX, kTMmaWaitW, kTProdTotal, kTProdWaitSite, kTProdWaitStage, kTEpiTotal,
                    kTEpiWaitAcc, kTEpiWaitSite, kTLoadTotal, kTLoadWaitW, kTCtas, kTUnits, kTGateWaitX };
__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, bool on, long long &acc)
{
    if (on) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity);
    }
}

// byte offset of 16-byte chunk j of row r inside a 128-byte-swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }

// Gather-table entry of (K block, 16-byte chunk j): where the 4 channels k = 32*kb + 4*j .. +3 of a
// site's patch live relative to the site's own pixel, and which tap they belong to.
struct KEntry {
    int relb;     // byte offset from the site's pixel (y*Win + x)*Cin*4
    int tap;      // tap index ky*kw + kx (< 32), or -1 for the K padding beyond kh*kw*Cin
};
__device__ __forceinline__ KEntry make_kentry(const TcParams &p, int kb, int j)
{
    const int k = kb * kBlockK + 4 * j;
    const int tap = k / p.Cin, c = k - tap * p.Cin;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    KEntry e;
    e.relb = (((ky - p.pad_t) * p.Win + (kx - p.pad_l)) * p.Cin + c) * 4;
    e.tap = k < p.K ? tap : -1;
    return e;
}

// What the site decoder leaves for the producers, per site of a unit: the address of the site's own input
// pixel (channel 0) in the source F map, and one bit per tap (kh*kw <= 32) telling whether the tap is
// inside the map; 0 for the padding sites of a partial unit, so that they gather nothing.
struct __align__(16) SiteSrc {
    const char *ptr;
    uint32_t taps;
    uint32_t pad;
};

// One (K block, half) item of the site operand: 64 sites x 8 chunks of 4 channels gathered by one
// producer group; thread t owns chunk t&7 of sites (t>>3) + 16u, u = 0..3.  Value rows are 0..63,
// rate rows 64..127 of the item's 128 rows.  `ss` points at the item's first site.
// item_load issues the 8 global loads, branch-free: a tap outside the map (or K padding, or a padding
// site) reads the 128 zero bytes every map is allocated with in front of it - the same negative offset
// is a zero line for F and for A - so no register is cleared and no branch is taken.  item_store
// converts (V = F*slope, R = A*slope), splits and writes the swizzled operand rows; the caller
// overlaps the two across consecutive items.
__device__ __forceinline__ void item_load(const TcParams &p, const KEntry e, const SiteSrc *ss, int t, float4 (&f)[kPairs], float4 (&a)[kPairs])
{
    const int tq = t >> 3;
    const uint32_t tapbit = e.tap >= 0 && !(p.debug & 1) ? 1u << e.tap : 0u;
#pragma unroll
    for (int u = 0; u < kPairs; ++u) {
        const SiteSrc q = ss[tq + 16 * u];
        const char *pf = (q.taps & tapbit) ? q.ptr + e.relb : p.zero_f;
        f[u] = __ldg(reinterpret_cast<const float4 *>(pf));
        a[u] = __ldg(reinterpret_cast<const float4 *>(pf + p.a_minus_f));
    }
}

// 16-byte store to a shared-memory address (the generic `*ptr = v` form compiles to ST.E with a 64-bit
// address, and makes the following proxy fence a full MEMBAR that also waits for the prefetched loads).
__device__ __forceinline__ void sts128(uint32_t addr, const float4 &v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// hi = x with the low 13 mantissa bits cleared (a tf32 value), lo = x - hi (exact); two lanes at a time
// on the packed fp32 pipe.  Truncation instead of rounding keeps |lo| < 2^-10 |x|, which the tf32
// product of lo still resolves to 2^-21 |x|.
__device__ __forceinline__ void split2(float x0, float x1, float &h0, float &h1, float &l0, float &l1)
{
    h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u);
    h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
    const float2 l = __ffma2_rn(make_float2(h0, h1), make_float2(-1.f, -1.f), make_float2(x0, x1));
    l0 = l.x;
    l1 = l.y;
}

template <uint32_t kRateOff>
__device__ __forceinline__ void item_store(const TcParams &p, uint32_t x_hi, uint32_t x_lo, int t, const float4 (&f)[kPairs],
                                           const float4 (&a)[kPairs])
{
    if (p.debug & 2) return;
    const int j = t & 7, tq = t >> 3;
#pragma unroll
    for (int u = 0; u < kPairs; ++u) {
        const int r = tq + 16 * u;
        const float2 s01 = make_float2(slope_of(f[u].x, p.alpha), slope_of(f[u].y, p.alpha));
        const float2 s23 = make_float2(slope_of(f[u].z, p.alpha), slope_of(f[u].w, p.alpha));
        const float2 v01 = __fmul2_rn(make_float2(f[u].x, f[u].y), s01), v23 = __fmul2_rn(make_float2(f[u].z, f[u].w), s23);
        const float2 w01 = __fmul2_rn(make_float2(a[u].x, a[u].y), s01), w23 = __fmul2_rn(make_float2(a[u].z, a[u].w), s23);
        float4 h, l;
        split2(v01.x, v01.y, h.x, h.y, l.x, l.y);
        split2(v23.x, v23.y, h.z, h.w, l.z, l.w);
        const uint32_t ov = sw128_off(r, j);
        sts128(x_hi + ov, h);
        sts128(x_lo + ov, l);
        split2(w01.x, w01.y, h.x, h.y, l.x, l.y);
        split2(w23.x, w23.y, h.z, h.w, l.z, l.w);
        const uint32_t ow = ov + kRateOff;      // rate row: weights-as-M value row + 64, sites-as-M the same row of tile_r (same swizzle phase)
        sts128(x_hi + ow, h);
        sts128(x_lo + ow, l);
    }
}

// Dynamic shared memory: [pad to 1024][w_stages x {W_hi, W_lo: Mrows*128 bytes each}][2 x {X_hi 32K, X_lo 32K}].
// The weight stages come first: the A descriptor always spans 128 rows, so with Mrows < 128 it reads
// past the tile into whatever follows (the next stage / the site stages); those rows only feed
// accumulator lanes >= Mrows, which the epilogue never reads.
//
// kPair (weights-as-M, an even number of weight tiles): two CTAs of a cluster (one TPC) work on a unit together with
// tcgen05.mma.cta_group::2, M = 256: CTA r of the pair holds weight tile 2*pg + r (the A rows 128r ..) and the accumulator of
// those channels, and converts only item r (64 of the unit's 128 sites, half of the B rows); the tensor cores read the
// other half from the peer's shared memory.  Per K block and SM: 32 KB of operand stores instead of 64 and 8 KB instead of
// 12 KB of operand reads per MMA - the shared-memory path is what bounds the one-CTA form (header).  Only the leader
// (rank 0) issues MMAs; the peer's gatekeeper relays "my stage, my weights and my accumulator are ready" to a ring of
// mbarriers in the leader, and every commit is multicast to the barriers of both CTAs.  Column c of the accumulator
// sees the same instruction sequence as in the one-CTA form: results are bit-identical (test_pair_units_change_no_bit).
template <bool kFastDecode, bool kSM, bool kPool = false, bool kPair = false>
__global__ void __launch_bounds__(kTcThreads, 1) k_conv_eval_tc(const __grid_constant__ TcParams p)
{
    pdl_enter();
    static_assert(!kPool || (kFastDecode && !kSM), "the fused pool lives in the weights-as-M epilogue and the batched decoder");
    static_assert(!kPair || (kFastDecode && !kSM), "CTA pairs: weights-as-M form with the batched decoder");
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t bar_x_full[kMaxSiteStages], bar_x_empty[kMaxSiteStages], bar_w_full[kMaxWStages], bar_w_empty[kMaxWStages];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_si_full[kSiteRing], bar_si_free[kSiteRing];
    __shared__ __align__(8) uint64_t bar_peer[4];      // kPair, leader: the peer's gatekeeper arrives for pass qw on [qw & 3]
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) SiteSrc s_src[kSiteRing][kUnitSites];
    __shared__ __align__(16) long long s_dst[kSiteRing][kUnitSites];   // byte offset of the site's channel 0 in F (and in A)
    __shared__ KEntry s_ktab[kKtabBlocks * 8];
    // kPool: per group of four entries, the pool window they form (element offset of its channel 0 in idx / Fp / Ap, -1: not a
    // quad) and where its recompute flag lives (word << 5 | bit)
    __shared__ long long s_pool[kPool ? kSiteRing : 1][kPool ? kUnitSites / 4 : 1];
    __shared__ uint32_t s_pflag[kPool ? kSiteRing : 1][kPool ? kUnitSites / 4 : 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_sites = __shfl_sync(0xffffffffu, *p.counter, 0);      // warp-uniform for the compiler's sake
    if (blockIdx.x == 0 && tid == 0 && n_sites > 0) atomicAdd(p.accum, (unsigned long long)(kPool ? *p.site_counter : n_sites));
    const uint32_t crank = kPair ? cluster_ctarank() : 0u;                    // 0 = leader of the pair
    const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // CTA (or pair) index ...
    const int n_workers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;     // ... of this many
    const int n_mgroups = kPair ? p.m_tiles / 2 : (p.m_tiles + p.mtu - 1) / p.mtu;
    // Few sites (few streams, deep layers): with 128-site units fewer units than CTAs - then a unit is ONE 64-site item (MMA
    // N = 128, half the tensor time per K block, twice the CTAs at work); the layout of an item inside stage and accumulator
    // is unchanged.  Uniform over the grid: every CTA reads the same counter.
    const bool half = !kSM && !kPair && p.half_units && ((n_sites + kUnitSites - 1) / kUnitSites) * n_mgroups < (int)gridDim.x;
    const int usites = half ? kItemSites : kUnitSites;          // sites per unit
    const bool one_item = half || kPair;                        // a site stage of this CTA is ONE item (pair: the CTA's half of the unit)
    const int ipk = one_item ? 1 : 2;                           // items per K block (converted by this CTA)
    // Site stages: two of two items each, or - half units - the same four item slots as four one-item stages.  Three producer
    // groups take the items round-robin, so a group's consecutive items are three apart: with fewer than three stages it
    // could reach a stage TWO uses ahead of its consumer, and a parity wait cannot tell "two phases behind" from "done".
    // One-item stages (half units, pairs) use the compact layout: three stages of 32 KB (hi tile, lo tile behind it), and the 32 KB
    // that frees go to the weight ring.  For pairs the third weight stage is what made them pay (conv5 0.60 -> 0.50 ms); for half
    // units (one stream) it is neutral: a lone SM streams its weight image at ~23 bytes per cycle however many bulk copies it
    // has in flight (conv5: 1.18 MB per CTA = the 52 k cycles the launch takes, profiles/r2_summary.md finding 15).
    const uint32_t n_xst = one_item ? (uint32_t)kPairSiteStages : (uint32_t)kSiteStages;
    const int w_stages = (half && !kPair) ? p.w_stages_half : p.w_stages;
    // first byte of site stage sx (its hi tile; the lo tile is kSiteTileBytes behind): item slot (sx >> 1, sx & 1) when a stage is one item
    const uint32_t kLoOff = one_item ? (uint32_t)kItemTileBytes : (uint32_t)kSiteTileBytes;      // from a stage's hi tile to its lo tile
    auto stage_off = [&](uint32_t sx) { return one_item ? sx * 2u * (uint32_t)kItemTileBytes : sx * (uint32_t)kSiteStageBytes; };
    const int n_blocks = (n_sites + usites - 1) / usites;
    const int total_units = n_blocks * n_mgroups;
    if (worker >= total_units) return;                // uniform per CTA (and per pair): nothing allocated yet

    unsigned char *smem_w = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_tile = (uint32_t)p.Mrows * 128u;             // one weight tile (hi or lo)
    unsigned char *smem_x = smem_w + (size_t)w_stages * 2 * w_tile;

    if (tid == 0) {
        for (int i = 0; i < (int)n_xst; ++i) {
            mbar_init(smem_u32(&bar_x_full[i]), (uint32_t)(ipk * kGroupThreads));   // every item of the stage (one producer group each)
            mbar_init(smem_u32(&bar_x_empty[i]), 1);
        }
        for (int i = 0; i < w_stages; ++i) {
            mbar_init(smem_u32(&bar_w_full[i]), 1);
            mbar_init(smem_u32(&bar_w_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_acc_full[i]), 1);
            mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps * 32);
        }
        for (int i = 0; i < kSiteRing; ++i) {
            mbar_init(smem_u32(&bar_si_full[i]), 1);
            mbar_init(smem_u32(&bar_si_free[i]), kEpiWarps * 32);
        }
        if constexpr (kPair)
            for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar_peer[i]), 1);
        fence_barrier_init();
    }
    for (int i = tid; i < min(p.KB, kKtabBlocks) * 8; i += kTcThreads) s_ktab[i] = make_kentry(p, i >> 3, i & 7);
    if (warp == 0) {
        if constexpr (kPair) {      // the same warp of both CTAs
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();      // both CTAs' barriers initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
    const bool timing = p.timing != nullptr;
    const int n_units_cta = (total_units - worker + n_workers - 1) / n_workers;
    // weight tiles of unit (.., mg) that THIS CTA holds: first tile and how many
    auto tile0_of = [&](int mg) { return kPair ? 2 * mg + (int)crank : mg * p.mtu; };
    auto tiles_of = [&](int mg) { return kPair ? 1 : min(p.mtu, p.m_tiles - mg * p.mtu); };

    if (warp < kEpiWarps) {
        // ===================== epilogue =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        // TMEM lane == row of the weight tile == output channel; columns == (site, map) rows.
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool warp_live = warp * 32 < p.Mrows;          // warps whose 32 lanes hold no channel only arrive
        long long tw_acc = 0, tw_si = 0;
        int n_quads = 0;                                     // kPool: windows evaluated here (counted once, by the first weight tile's CTA)
        const long long t_begin = timing ? clock64() : 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = worker + ul * n_workers;
            const int mg = unit % n_mgroups;
            const int mt_count = tiles_of(mg);
            const int buf = ul % kSiteRing, ab = ul % p.n_acc;
            timed_wait(smem_u32(&bar_si_full[buf]), (uint32_t)(ul / kSiteRing) & 1u, timing, tw_si);
            timed_wait(smem_u32(&bar_acc_full[ab]), (uint32_t)(ul / p.n_acc) & 1u, timing, tw_acc);
            tc_fence_after();
            if constexpr (kSM) {
                // sites-as-M: TMEM lane == site of the unit; columns [0, 2*Cpad) of the buffer are D_v (hi.hi + lo.hi products in
                // [0, Cpad), hi.lo products in [Cpad, 2*Cpad)), the next 2*Cpad columns D_r.  One thread = one site: it adds the
                // two halves, adds the bias to the value row and writes 4 channels per 16-byte store.
                const int cpad = p.Mrows;
                const long long dst = s_dst[buf][warp * 32 + lane];
                const bool site_ok = dst >= 0 && !(p.debug & 8);
                const uint32_t tbase = lane_addr + (uint32_t)(ab * 4 * cpad);
                const int nch = (p.C + 15) >> 4;                       // 16-channel chunks holding real channels
                const int total = 2 * nch;                             // value chunks, then rate chunks
                uint32_t ra[16], rb[16], rc[16], rd[16];
                auto issue = [&](int ci, uint32_t(&hi)[16], uint32_t(&lo)[16]) {
                    const int map = ci >= nch ? 1 : 0, c0 = (ci - map * nch) << 4;
                    const uint32_t ta = tbase + (uint32_t)(map * 2 * cpad + c0);
                    tmem_ld16(ta, hi);
                    tmem_ld16(ta + (uint32_t)cpad, lo);
                };
                auto emit = [&](int ci, const uint32_t(&hi)[16], const uint32_t(&lo)[16]) {
                    const int map = ci >= nch ? 1 : 0, c0 = (ci - map * nch) << 4;
                    if (!site_ok) return;
                    char *const out = (char *)(map ? p.A : p.F) + dst + (long long)c0 * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (c0 + 4 * q >= p.C) break;                   // C % 4 == 0: whole float4 groups only
                        float4 o;
                        o.x = __fadd_rn(__uint_as_float(hi[4 * q + 0]), __uint_as_float(lo[4 * q + 0]));
                        o.y = __fadd_rn(__uint_as_float(hi[4 * q + 1]), __uint_as_float(lo[4 * q + 1]));
                        o.z = __fadd_rn(__uint_as_float(hi[4 * q + 2]), __uint_as_float(lo[4 * q + 2]));
                        o.w = __fadd_rn(__uint_as_float(hi[4 * q + 3]), __uint_as_float(lo[4 * q + 3]));
                        if (!map) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + 4 * q));
                            o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
                        }
                        *reinterpret_cast<float4 *>(out + 16 * q) = o;
                    }
                };
                issue(0, ra, rb);
#pragma unroll 1
                for (int ci = 0; ci < total; ci += 2) {
                    tmem_ld_wait();
                    if (ci + 1 < total) issue(ci + 1, rc, rd);
                    emit(ci, ra, rb);
                    if (ci + 1 >= total) break;
                    tmem_ld_wait();
                    if (ci + 2 < total) issue(ci + 2, ra, rb);
                    emit(ci + 1, rc, rd);
                }
            } else if constexpr (kPool) {
              if (warp_live) {
                // Weights-as-M with the pool fused: thread = channel, 8 sites (two possible windows) per step - their value
                // columns and the rate columns 64 further on - so that a quad's four (F, A) pairs are in registers together.
                // The list may hold padding entries anywhere (end of a stream's block): every site is checked by its own offset.
                for (int mt = 0; mt < mt_count; ++mt) {
                    const int c = (tile0_of(mg) + mt) * p.Mch + row;
                    const bool c_ok = row < p.Mrows && c < p.C && !(p.debug & 8);
                    const long long a_minus_f = (const char *)p.A - (const char *)p.F;
                    const float bias = c_ok ? __ldg(p.bias + c) : 0.f;
                    const uint32_t taddr = lane_addr + (uint32_t)((ab * p.mtu + mt) * kUnitCols);
                    char *const baseF = (char *)p.F + (long long)c * 4;
                    uint32_t va[8], ra[8], vb[8], rb[8];
                    auto issue = [&](int g8, uint32_t(&v)[8], uint32_t(&r)[8]) {
                        const uint32_t cc = (uint32_t)((g8 >> 3) * 128 + (g8 & 7) * 8);     // item (g8 >> 3): 64 value columns, then 64 rate columns
                        tmem_ld8(taddr + cc, v);
                        tmem_ld8(taddr + cc + 64u, r);
                    };
                    auto emit = [&](int g8, const uint32_t(&v)[8], const uint32_t(&r)[8]) {
                        const int i0 = g8 * 8;
                        const longlong2 *po = reinterpret_cast<const longlong2 *>(&s_dst[buf][i0]);
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const longlong2 o = po[e];
                            f[2 * e] = __fadd_rn(__uint_as_float(v[2 * e]), bias);
                            f[2 * e + 1] = __fadd_rn(__uint_as_float(v[2 * e + 1]), bias);
                            if (c_ok && o.x >= 0) {
                                *reinterpret_cast<float *>(baseF + o.x) = f[2 * e];
                                *reinterpret_cast<float *>(baseF + a_minus_f + o.x) = __uint_as_float(r[2 * e]);
                            }
                            if (c_ok && o.y >= 0) {
                                *reinterpret_cast<float *>(baseF + o.y) = f[2 * e + 1];
                                *reinterpret_cast<float *>(baseF + a_minus_f + o.y) = __uint_as_float(r[2 * e + 1]);
                            }
                        }
                        // Both possible windows of the group branch-free (the epilogue warp shares its scheduler with three producer
                        // warps and runs one dependent chain: what it needs is independent instructions, not fewer of them).
                        // Winner by tournament with the reference's order (cutils.pyx:166-170: larger F, then smaller rate, then the
                        // earlier row): (0,1), (2,3), then the two winners - the first maximal element, like the sequential scan.
                        long long pw[2];
                        bool uns[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            pw[q] = s_pool[buf][(i0 >> 2) + q];
                            float rr[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) rr[k] = __fmul_rn(__uint_as_float(r[4 * q + k]), slope_of(f[4 * q + k], p.pool_alpha));
                            const bool b1 = f[4 * q + 1] > f[4 * q] || (f[4 * q + 1] == f[4 * q] && rr[1] < rr[0]);
                            const bool b3 = f[4 * q + 3] > f[4 * q + 2] || (f[4 * q + 3] == f[4 * q + 2] && rr[3] < rr[2]);
                            const float fa = b1 ? f[4 * q + 1] : f[4 * q], ra_ = b1 ? rr[1] : rr[0];
                            const float fb = b3 ? f[4 * q + 3] : f[4 * q + 2], rb_ = b3 ? rr[3] : rr[2];
                            const uint32_t aa = b1 ? r[4 * q + 1] : r[4 * q], ab_ = b3 ? r[4 * q + 3] : r[4 * q + 2];
                            const bool bb = fb > fa || (fb == fa && rb_ < ra_);
                            const int row_best = bb ? (b3 ? 3 : 2) : (b1 ? 1 : 0);
                            const float f_best = bb ? fb : fa, r_best = bb ? rb_ : ra_;
                            const uint32_t a_best = bb ? ab_ : aa;
                            const float rlow = fminf(fminf(rr[0], rr[1]), fminf(rr[2], rr[3]));
                            const bool ok = c_ok && pw[q] >= 0;
                            uns[q] = ok && r_best != rlow;                                   // cutils.pyx:173-177 (by value)
                            if (ok) {
                                p.pool_idx[pw[q] + c] = (uint8_t)row_best;
                                p.pool_Fp[pw[q] + c] = f_best;
                                p.pool_Ap[pw[q] + c] = __uint_as_float(a_best);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const unsigned any = __ballot_sync(0xffffffffu, uns[q]);
                            if (any && lane == 0) {
                                const uint32_t code = s_pflag[buf][(i0 >> 2) + q];
                                atomicOr(p.pool_flags + (code >> 5), 1u << (code & 31u));
                            }
                            if (pw[q] >= 0 && tile0_of(mg) + mt == 0 && warp == 0 && lane == 0) ++n_quads;
                        }
                    };
                    issue(0, va, ra);
#pragma unroll 1
                    for (int g8 = 0; g8 < usites / 8; g8 += 2) {
                        tmem_ld_wait();
                        issue(g8 + 1, vb, rb);
                        emit(g8, va, ra);
                        tmem_ld_wait();
                        if (g8 + 2 < usites / 8) issue(g8 + 2, va, ra);
                        emit(g8 + 1, vb, rb);
                    }
                }
              }
            } else
            if (warp_live) {
                for (int mt = 0; mt < mt_count; ++mt) {
                    const int g = p.rep > 1 ? (warp * 32) / p.Mch : 0;    // which copy of the channels this warp holds (Mch % 32 == 0 when rep > 1)
                    const int c = (tile0_of(mg) + mt) * p.Mch + (row - g * p.Mch);
                    const bool c_ok = row < p.Mrows && c < p.C && !(p.debug & 8);
                    const long long a_minus_f = (const char *)p.A - (const char *)p.F;
                    const float bias = c_ok ? __ldg(p.bias + c) : 0.f;
                    const uint32_t taddr = lane_addr + (uint32_t)((ab * p.mtu + mt) * kUnitCols);
                    // 16 columns (= 16 sites of one map) per step, TMEM loads double buffered against the stores
                    uint32_t ra[16], rb[16];
                    auto store16 = [&](const uint32_t(&r)[16], int cc) {
                        const bool is_rate = (cc & 64) != 0;                 // columns 64..127 of each 128-column half
                        const int i0 = (cc >> 7) * kItemSites + (cc & 63);   // first site of these 16 columns
                        const longlong2 *po = reinterpret_cast<const longlong2 *>(&s_dst[buf][i0]);
                        if (po[0].x < 0 || !c_ok) return;                    // valid sites are a prefix of the unit
                        char *const base = (char *)p.F + (long long)c * 4 + (is_rate ? a_minus_f : 0LL);
                        const float add = is_rate ? 0.f : bias;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const longlong2 o = po[e];
                            const float v0 = __uint_as_float(r[2 * e]), v1 = __uint_as_float(r[2 * e + 1]);
                            if (o.x >= 0) *reinterpret_cast<float *>(base + o.x) = is_rate ? v0 : __fadd_rn(v0, add);
                            if (o.y >= 0) *reinterpret_cast<float *>(base + o.y) = is_rate ? v1 : __fadd_rn(v1, add);
                        }
                    };
                    const int cstep = 16 * p.rep;                         // this copy's chunks: cc = 16g, 16g + cstep, ...
                    tmem_ld16(taddr + (uint32_t)(16 * g), ra);
                    const int ncols = 2 * usites;                         // a half unit fills the first item's 128 columns only
#pragma unroll 1
                    for (int cc = 16 * g; cc < ncols; cc += 2 * cstep) {
                        tmem_ld_wait();
                        if (cc + cstep < ncols) tmem_ld16(taddr + (uint32_t)(cc + cstep), rb);
                        store16(ra, cc);
                        if (cc + cstep >= ncols) break;
                        tmem_ld_wait();
                        if (cc + 2 * cstep < ncols) tmem_ld16(taddr + (uint32_t)(cc + 2 * cstep), ra);
                        store16(rb, cc + cstep);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bar_acc_empty[ab]));
            mbar_arrive(smem_u32(&bar_si_free[buf]));
        }
        if constexpr (kPool) {
            if (n_quads > 0) { atomicAdd(p.pool_accum, (unsigned long long)n_quads); atomicAdd(p.pool_accum + 32, (unsigned long long)n_quads); }
        }
        if (timing && tid == 0) {
            atomicAdd(p.timing + kTEpiTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTEpiWaitAcc, (unsigned long long)tw_acc);
            atomicAdd(p.timing + kTEpiWaitSite, (unsigned long long)tw_si);
            atomicAdd(p.timing + kTCtas, 1ULL);
            atomicAdd(p.timing + kTUnits, (unsigned long long)n_units_cta);
        }
    } else if (warp < kProdWarp0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
      if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop converged; one elected lane issues.  It executes NO mbarrier wait:
        // measured (tools/bench_umma.cu), every mbarrier.try_wait in the issuing warp costs ~170 cycles
        // that the tensor pipe idles (the wait queues behind the in-flight MMAs), so all operand waits
        // live in the gatekeeper warp, which releases each pass through a named barrier (ids 1..4).
        const uint32_t idesc = kPair ? make_idesc_tf32(kUnitCols, 256) : make_idesc_tf32(half ? kUnitCols / 2 : kUnitCols);
        const uint32_t idesc_cat = make_idesc_tf32(kSM ? 2 * p.Mrows : kUnitCols), idesc_hi = make_idesc_tf32(kSM ? p.Mrows : kUnitCols);
        const uint32_t x_base = smem_u32(smem_x), w_base = smem_u32(smem_w);
        uint32_t qx = 0, qw = 0;
        long long tw_gate = 0;
        const long long t_begin = timing ? clock64() : 0;
        for (int ul = 0; ul < (kPair && crank != 0 ? 0 : n_units_cta); ++ul) {      // pair: the leader issues for both CTAs
            const int unit = worker + ul * n_workers;
            const int mg = unit % n_mgroups;
            const int mt_count = tiles_of(mg);
            const int ab = ul % p.n_acc;
            for (int kb = 0; kb < p.KB; ++kb, ++qx) {
                const uint32_t sx = qx % n_xst;
                const uint32_t x_hi = x_base + stage_off(sx), x_lo = x_hi + kLoOff;
                for (int mt = 0; mt < mt_count; ++mt, ++qw) {
                    const uint32_t sw = qw % (uint32_t)w_stages;
                    const uint32_t w_hi = w_base + sw * 2u * w_tile, w_lo = w_hi + w_tile;
                    const uint32_t d = tmem_base + (uint32_t)((ab * p.mtu + mt) * kUnitCols);
                    const int ks_n = kb == p.KB - 1 ? p.ks_last : kBlockK / 8;
                    const long long t0 = timing ? clock64() : 0;
                    asm volatile("bar.sync %0, 64;" ::"r"(1u + (qw & 3u)) : "memory");     // operands (and accumulator) ready
                    if (timing) tw_gate += clock64() - t0;
                    tc_fence_after();
                    if (elect_one()) {
                        if constexpr (kSM) {
                            // sites-as-M: A = one 128-row site tile (tile_v, then tile_r 16 KB behind it), B = [W_hi ; W_lo] (2*Cpad rows
                            // of the weight stage; its first Cpad rows alone are W_hi).  X_lo x W_hi accumulates into the same columns
                            // as X_hi x W_hi: the epilogue only has to add the W_lo half.
                            if (!(p.debug & 4)) {
                                const uint32_t dv = tmem_base + (uint32_t)(ab * 4 * p.Mrows), dr = dv + (uint32_t)(2 * p.Mrows);
#pragma unroll
                                for (int ks = 0; ks < kBlockK / 8; ++ks) {
                                    const uint32_t ko = (uint32_t)ks * 32u;
                                    const uint64_t dw = make_desc_sw128(w_hi + ko);
                                    const uint64_t dvh = make_desc_sw128(x_hi + ko), dvl = make_desc_sw128(x_lo + ko);
                                    const uint64_t drh = make_desc_sw128(x_hi + (uint32_t)kItemTileBytes + ko), drl = make_desc_sw128(x_lo + (uint32_t)kItemTileBytes + ko);
                                    if (ks >= ks_n) break;
                                    const uint32_t acc = (kb | ks) != 0 ? 1u : 0u;
                                    mma_tf32(dv, dvh, dw, idesc_cat, acc);
                                    mma_tf32(dr, drh, dw, idesc_cat, acc);
                                    mma_tf32(dv, dvl, dw, idesc_hi, 1u);
                                    mma_tf32(dr, drl, dw, idesc_hi, 1u);
                                }
                            }
                        } else if (!(p.debug & 4)) {
#pragma unroll
                            for (int ks = 0; ks < kBlockK / 8; ++ks) {
                                const uint32_t ko = (uint32_t)ks * 32u;       // 8 tf32 = 32 bytes along K inside the swizzled row
                                const uint64_t dwh = make_desc_sw128(w_hi + ko), dwl = make_desc_sw128(w_lo + ko);
                                const uint64_t dxh = make_desc_sw128(x_hi + ko), dxl = make_desc_sw128(x_lo + ko);
                                if (ks >= ks_n) break;
                                if constexpr (kPair) {
                                    mma_tf32_pair(d, dwh, dxl, idesc, (kb | ks) != 0 ? 1u : 0u);
                                    mma_tf32_pair(d, dwl, dxh, idesc, 1u);
                                    mma_tf32_pair(d, dwh, dxh, idesc, 1u);
                                } else {
                                    mma_tf32(d, dwh, dxl, idesc, (kb | ks) != 0 ? 1u : 0u);
                                    mma_tf32(d, dwl, dxh, idesc, 1u);
                                    mma_tf32(d, dwh, dxh, idesc, 1u);
                                }
                            }
                        }
                        if constexpr (kPair) {      // stages and accumulator of BOTH CTAs
                            mma_commit_pair(smem_u32(&bar_w_empty[sw]));
                            mma_commit_pair(smem_u32(&bar_x_empty[sx]));
                            if (kb == p.KB - 1) mma_commit_pair(smem_u32(&bar_acc_full[ab]));
                        } else {
                            mma_commit(smem_u32(&bar_w_empty[sw]));
                            if (mt == mt_count - 1) mma_commit(smem_u32(&bar_x_empty[sx]));
                            if (mt == mt_count - 1 && kb == p.KB - 1) mma_commit(smem_u32(&bar_acc_full[ab]));
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (timing && lane == 0) {
            atomicAdd(p.timing + kTMmaTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTMmaWaitX, (unsigned long long)tw_gate);
        }
      } else if (warp == kGateWarp) {
        // ===================== gatekeeper =====================
        // Waits for everything pass qw needs - accumulator drained (first pass of a unit), both halves of the
        // site stage, the weight stage - then arrives on named barrier 1 + (qw & 3).  An id is reused every
        // 4 passes; the site stage of pass qw can only be full once the MMA warp has committed the passes
        // two K blocks back, i.e. has passed barrier qw - 4, so arrivals never pile up on one id.
        uint32_t qx = 0, qw = 0;
        long long tw_acc = 0, tw_x = 0, tw_w = 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = worker + ul * n_workers;
            const int mg = unit % n_mgroups;
            const int mt_count = tiles_of(mg);
            const int ab = ul % p.n_acc;
            const uint32_t ua = (uint32_t)(ul / p.n_acc);
            if (ua > 0) timed_wait(smem_u32(&bar_acc_empty[ab]), (ua - 1) & 1u, timing, tw_acc);    // epilogue drained this buffer
            for (int kb = 0; kb < p.KB; ++kb, ++qx) {
                const uint32_t sx = qx % n_xst;
                const uint32_t px = (qx / n_xst) & 1u;
                for (int mt = 0; mt < mt_count; ++mt, ++qw) {
                    const uint32_t sw = qw % (uint32_t)w_stages;
                    // every wait costs ~170 cycles even when the barrier is already complete: the weights (normally early)
                    // first, the site stage (normally the last thing to arrive) last, one barrier for both of its halves
                    timed_wait(smem_u32(&bar_w_full[sw]), (qw / (uint32_t)w_stages) & 1u, timing, tw_w);
                    if (mt == 0) timed_wait(smem_u32(&bar_x_full[sx]), px, timing, tw_x);
                    if constexpr (kPair) {
                        // The peer relays: everything of pass qw on its side is ready -> bar_peer[qw & 3] of the leader.  It can be
                        // at most n_xst = 3 passes ahead of the leader's MMAs (its stage of pass qw + 3 is freed by the commit of
                        // pass qw), so a barrier's previous phase (pass qw - 4) has always been consumed when the next arrival comes.
                        if (crank != 0) {
                            if (lane == 0) mbar_arrive_remote(smem_u32(&bar_peer[qw & 3u]), 0u);
                            __syncwarp();
                            continue;
                        }
                        mbar_wait_cluster(smem_u32(&bar_peer[qw & 3u]), (qw >> 2) & 1u);
                    }
                    asm volatile("bar.arrive %0, 64;" ::"r"(1u + (qw & 3u)) : "memory");
                }
            }
        }
        if (timing && lane == 0) {
            atomicAdd(p.timing + kTMmaWaitAcc, (unsigned long long)tw_acc);
            atomicAdd(p.timing + kTMmaWaitW, (unsigned long long)tw_w);
            atomicAdd(p.timing + kTGateWaitX, (unsigned long long)tw_x);
        }
      } else if (warp == kLoadWarp) {
        // ===================== weight loader =====================
        if (lane == 0) {
            uint32_t qw = 0;
            long long tw_w = 0;
            const long long t_begin = timing ? clock64() : 0;
            const size_t tile_floats = (size_t)2 * p.Mrows * kBlockK;        // hi + lo of one (weight tile, K block)
            for (int ul = 0; ul < n_units_cta; ++ul) {
                const int unit = worker + ul * n_workers;
                const int mg = unit % n_mgroups;
                const int mt_count = tiles_of(mg);
                for (int kb = 0; kb < p.KB; ++kb) {
                    for (int mt = 0; mt < mt_count; ++mt, ++qw) {
                        const int sw = (int)(qw % (uint32_t)w_stages);
                        const uint32_t use = qw / (uint32_t)w_stages;
                        if (use > 0) timed_wait(smem_u32(&bar_w_empty[sw]), (use - 1) & 1u, timing, tw_w);
                        const float *src = p.wimg + ((size_t)(tile0_of(mg) + mt) * p.KB + kb) * tile_floats;
                        if (p.debug & 16) { mbar_arrive(smem_u32(&bar_w_full[sw])); continue; }
                        mbar_expect_tx(smem_u32(&bar_w_full[sw]), 2u * w_tile);
                        bulk_g2s(smem_u32(smem_w + (size_t)sw * 2 * w_tile), src, 2u * w_tile, smem_u32(&bar_w_full[sw]));
                    }
                }
            }
            if (timing) {
                atomicAdd(p.timing + kTLoadTotal, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.timing + kTLoadWaitW, (unsigned long long)tw_w);
            }
        }
      } else if (warp == kSiteWarp) {
        // ===================== site decoder =====================
        // kFastDecode (layers with long units or several weight tiles): measured 3-4 % faster there, but 2-8 % slower for the short
        // units of the first two tensor-core layers, which keep the simple loop (profiles/r1e_summary.md).
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = worker + ul * n_workers;
            const int blk = unit / n_mgroups;
            const int buf = ul % kSiteRing;
            const uint32_t us = (uint32_t)(ul / kSiteRing);
            if constexpr (kFastDecode) {
                // The unit's work-list entries are read before the wait for the ring slot, all four per lane at once (one
                // trip to memory per unit instead of four dependent ones), and decoded with reciprocal multiplications and
                // per-row / per-column validity masks instead of integer divisions and a kh x kw loop: the decoder is a
                // single warp and must not need more cycles per unit than a short (5-pass) unit takes.
                uint32_t ent[kUnitSites / 32];
    #pragma unroll
                for (int k = 0; k < kUnitSites / 32; ++k) {
                    const long long gi = (long long)blk * usites + lane + 32 * k;
                    ent[k] = (lane + 32 * k < usites && gi < n_sites) ? __ldg(p.sites + gi) : 0xffffffffu;
                }
                if (us > 0) mbar_wait(smem_u32(&bar_si_free[buf]), (us - 1) & 1u);
    #pragma unroll
                for (int k = 0; k < kUnitSites / 32; ++k) {
                    const int i = lane + 32 * k;
                    SiteSrc q;
                    q.ptr = p.zero_f; q.taps = 0u; q.pad = 0u;
                    long long dst = -1;
                    if constexpr (kPool) {
                        if ((i & 3) == 0) { s_pool[buf][i >> 2] = -1; s_pflag[buf][i >> 2] = 0u; }
                    }
                    if (ent[k] != 0xffffffffu) {
                        int s, y, x;
                        site_decode(p.code, kPool ? ent[k] & ~p.quad_bit : ent[k], s, y, x);
                        const int site = y * p.W + x;
                        if constexpr (kPool) {
                            if (ent[k] & p.quad_bit) {          // first entry of a complete window: (y, x) is its top-left site
                                const int oy = y >> 1, ox = x >> 1;
                                s_pool[buf][i >> 2] = (long long)s * p.pool_stride + (long long)(oy * p.pW + ox) * p.C;
                                s_pflag[buf][i >> 2] = ((uint32_t)(s * p.pHWw + oy * p.pWw + (ox >> 5)) << 5) | (uint32_t)(ox & 31);
                            }
                        }
                        q.ptr = reinterpret_cast<const char *>(p.srcF + ((long long)s * p.src_stride + (long long)(y * p.Win + x) * p.Cin));
                        uint32_t colbits = 0u;
                        for (int kx = 0; kx < p.kw; ++kx)
                            if ((unsigned)(x + kx - p.pad_l) < (unsigned)p.Win) colbits |= 1u << kx;
                        for (int ky = 0; ky < p.kh; ++ky)
                            if ((unsigned)(y + ky - p.pad_t) < (unsigned)p.Hin) q.taps |= colbits << (ky * p.kw);
                        dst = ((long long)s * p.fstride + (long long)site * p.C) * 4;
                    }
                    s_src[buf][i] = q;
                    s_dst[buf][i] = dst;
                }
            } else {
                if (us > 0) mbar_wait(smem_u32(&bar_si_free[buf]), (us - 1) & 1u);
                for (int i = lane; i < kUnitSites; i += 32) {
                    const long long gi = (long long)blk * usites + i;
                    SiteSrc q;
                    q.ptr = p.zero_f; q.taps = 0u; q.pad = 0u;
                    long long dst = -1;
                    if (i < usites && gi < n_sites) {
                        int s, y, x;
                        site_decode(p.code, p.sites[gi], s, y, x);
                        const int site = y * p.W + x;
                        q.ptr = reinterpret_cast<const char *>(p.srcF + ((long long)s * p.src_stride + (long long)(y * p.Win + x) * p.Cin));
                        for (int ky = 0, tap = 0; ky < p.kh; ++ky)
                            for (int kx = 0; kx < p.kw; ++kx, ++tap)
                                if ((unsigned)(y + ky - p.pad_t) < (unsigned)p.Hin && (unsigned)(x + kx - p.pad_l) < (unsigned)p.Win) q.taps |= 1u << tap;
                        dst = ((long long)s * p.fstride + (long long)site * p.C) * 4;
                    }
                    s_src[buf][i] = q;
                    s_dst[buf][i] = dst;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_si_full[buf]));
        }
      }
    } else {
        // ===================== gather producers =====================
        // Items are numbered q = 2 * (unit-local K block counter) + half over the CTA's whole life; group g
        // takes q = g, g + 3, ...  The loads of the group's next item are issued before the current item is
        // converted and stored, so global latency overlaps the conversion and the wait for the stage.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsProd));
        const int pt = tid - kProdWarp0 * 32;
        const int g = pt / kGroupThreads, t = pt - g * kGroupThreads;
        long long tw_si = 0, tw_x = 0;
        const long long t_begin = timing ? clock64() : 0;
        const uint32_t q_end = (uint32_t)n_units_cta * (uint32_t)p.KB * (uint32_t)ipk;      // items: ipk per (unit, K block)
        struct Pos { int ul, kb, h; uint32_t q; };
        auto advance = [&](Pos &c) {
            c.q += kGroups;
            c.h += kGroups;
            if (ipk == 2) { c.kb += c.h >> 1; c.h &= 1; } else { c.kb += c.h; c.h = 0; }
            while (c.kb >= p.KB) { c.kb -= p.KB; ++c.ul; }
        };
        int ready_ul = -1;                 // units whose site info this thread has already waited for
        auto load = [&](const Pos &c, float4 (&f)[kPairs], float4 (&a)[kPairs]) {
            const int buf = c.ul % kSiteRing;
            if (c.ul > ready_ul) {
                timed_wait(smem_u32(&bar_si_full[buf]), (uint32_t)(c.ul / kSiteRing) & 1u, timing, tw_si);
                ready_ul = c.ul;
            }
            const KEntry e = c.kb < kKtabBlocks ? s_ktab[c.kb * 8 + (t & 7)] : make_kentry(p, c.kb, t & 7);
            item_load(p, e, s_src[buf] + (kPair ? (int)crank : c.h) * kItemSites, t, f, a);
        };
        const uint32_t x_base = smem_u32(smem_x);
        auto store = [&](const Pos &c, const float4 (&f)[kPairs], const float4 (&a)[kPairs]) {
            const uint32_t qx = ipk == 2 ? c.q >> 1 : c.q;
            const uint32_t sx = qx % n_xst, use = qx / n_xst;
            if (use > 0) timed_wait(smem_u32(&bar_x_empty[sx]), (use - 1) & 1u, timing, tw_x);   // MMAs that read this stage are done
            // weights-as-M: an item's 128 rows (64 value rows, then its 64 rate rows) are one half of the 256-row tile;
            // sites-as-M: its value rows are rows 64h.. of tile_v and its rate rows the same rows of tile_r (16 KB behind)
            constexpr uint32_t kHalfStride = kSM ? (uint32_t)(kItemSites / 8) * 1024u : (uint32_t)kItemTileBytes;
            constexpr uint32_t kRateOff = kSM ? (uint32_t)kItemTileBytes : (uint32_t)(kItemSites / 8) * 1024u;
            const uint32_t x_hi = x_base + stage_off(sx) + (uint32_t)c.h * kHalfStride;          // c.h == 0 when a stage is one item
            item_store<kRateOff>(p, x_hi, x_hi + kLoOff, t, f, a);
            fence_proxy_async();      // generic-proxy writes of X -> visible to the tensor core (async proxy)
            mbar_arrive(smem_u32(&bar_x_full[sx]));
        };
        Pos cur;
        cur.q = (uint32_t)g; cur.ul = 0; cur.kb = ipk == 2 ? g >> 1 : g; cur.h = ipk == 2 ? g & 1 : 0;
        while (cur.kb >= p.KB) { cur.kb -= p.KB; ++cur.ul; }
        float4 fa[kPairs], aa[kPairs], fb[kPairs], ab4[kPairs];
        if (cur.q < q_end) {
            load(cur, fa, aa);
            while (true) {
                Pos nxt = cur;
                advance(nxt);
                if (nxt.q < q_end) load(nxt, fb, ab4);
                store(cur, fa, aa);
                if (nxt.q >= q_end) break;
                cur = nxt;
                advance(cur);
                if (cur.q < q_end) load(cur, fa, aa);
                store(nxt, fb, ab4);
                if (cur.q >= q_end) break;
            }
        }
        if (timing && pt == 0) {
            atomicAdd(p.timing + kTProdTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTProdWaitSite, (unsigned long long)tw_si);
            atomicAdd(p.timing + kTProdWaitStage, (unsigned long long)tw_x);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) {
        cluster_sync_all();      // the peer's tensor core reads this CTA's shared memory and writes its TMEM until the last commit
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    } else if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace tc
}  // namespace aec
