// aec_rt.cuh - conv re-evaluation by ROW TILES: the dense form of the gathered GEMM of aec_tc.cuh for the layers
// whose work sets are runs of neighbouring sites (sm_100a, tcgen05).
//
// Why (profiles/r2_summary.md): at the steady state of the workload conv2 ... conv4 re-evaluate 40-60 % of their maps, in
// rows that are active almost end to end.  The gathered kernel treats every site on its own: each input pixel of the
// previous layer is fetched, converted (slope, hi/lo split) and stored to shared memory once per (site, tap) - nine times
// for a 3x3 kernel - and that producer work, above all its shared-memory stores, bounds the kernel: knocking the MMAs out
// of conv2 saves 14 %, halving the tensor work (sites-as-M) nothing.
//
// Here a work unit is R consecutive output rows of one stream (R x slot width = 128 sites: one 112-wide row of conv2, two
// 56-wide rows of conv3, or a 126-site segment of a wider row).  For every kernel row ky (and 32-channel block) the
// producers load the R input rows the unit needs ONCE, convert them and store them as a [pixel][channel] tile:
//     tile row t = rr * SW + ix + pad_l          (rr = row in the unit, ix = input column, SW = slot width, zeros between)
// and the kw taps of that kernel row are the SAME tile read from a start address kx rows further on - a tcgen05 shared
// memory descriptor may start at any row of a swizzled tile because the swizzle is a function of the address bits
// (tools/test_umma_shift.cu checks this for 128- and 64-byte rows).  Per site that is a third of the loads, conversions
// and shared-memory stores of the gathered form for a 3x3 kernel, fully coalesced (rows of the channel-last map).
//   MMA orientation = the sites-as-M form of aec_tc.cuh: D_v[site, n] += X_hi[site + kx, :] x [W_hi ; W_lo] (N = 2 Cpad),
//   then X_lo x W_hi (N = Cpad) into the same columns, and the same two for the rate rows: 2 cycles per site and K step.
//   Sites of the tile that are not in the layer's work set (gaps in a row, the padding slots) are computed and dropped:
//   the epilogue stores only where the work-set bitmap has a bit (a site outside it must keep its leaked value,
//   conv2d.py:115-123).
// Semantics: conv2d.py:118-123,144-181 as in aec_tc.cuh; same 3xTF32 products, accumulated in (ky, channel block, kx,
// channel) order with the same instruction sequence for every site.
#pragma once
#include "aec_tc.cuh"

namespace aec {
namespace rt {

using namespace tc;

constexpr int kRtMaxXStages = 4;
constexpr int kRtMaxWStages = 6;
constexpr int kRtRing = 4;            // unit-info buffers
constexpr int kRtProdThreads = kGroups * kGroupThreads;   // 384
// Warp roles of the row-tile CTA (768 threads): EIGHT epilogue warps - the four TMEM lane quarters are readable by warps
// with the matching (warp % 4) only, and four warps were the slowest role of the first version (profiles/r2_summary.md), so
// every quarter has two, each taking every other 16-channel chunk - then the MMA issuer, weight loader, unit decoder,
// gatekeeper, and 12 producer warps.  Registers after the role split: 256 x 72 + 128 x 48 + 384 x 96 = 768 x 80.
constexpr int kRtThreads = 768;
constexpr int kRtEpiWarps = 8;
constexpr int kRtMmaWarp = 8, kRtLoadWarp = 9, kRtUnitWarp = 10, kRtGateWarp = 11;
constexpr int kRtProdWarp0 = 12;
constexpr int kRtRegsEpi = 72, kRtRegsCtl = 48, kRtRegsProd = 96;
constexpr int kRtMaxPairs = 6;        // (tile row, 16-byte chunk) pairs per producer thread and stage

struct RtParams {
    const uint32_t *units;      // work list: stream << sh_s | row group << sh_y | x segment
    const int *counter;         // number of units
    const int *site_counter;    // number of work-set sites (statistics only)
    unsigned long long *accum_sites, *accum_units;
    const float *srcF;          // previous layer's channel-last F map; A = F + a_minus_f bytes
    long long a_minus_f;
    const char *zero_f;         // 128 zero bytes in front of the source F map (and, + a_minus_f, of A)
    long long src_stride;       // floats per stream
    float alpha;
    int Cin, Hin, Win;
    const float *wimg;          // [(ky*kw + kx)*ncb + cb][2*Cpad rows][CB floats]: rows 0..Cpad-1 W_hi, then W_lo; swizzled like the tiles
    const float *bias;
    float *F, *A;
    long long fstride;
    const uint32_t *nset;       // [S][H*Ww] exact work set of the layer for this step
    int C, H, W, Ww, Cpad;
    int kh, kw, pad_t, pad_l;
    int CB, ncb;                // channels per block (16 or 32) and blocks per pixel
    int row_bytes;              // CB * 4: 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B)
    int R, sw_shift, SEG;       // output rows per unit, log2(slot width), sites per x segment
    SiteCode code;
    int P;                      // tile rows written per stage (128 + kw - 1)
    uint32_t x_tile_bytes;      // one of the four tiles of a stage (V_hi, V_lo, R_hi, R_lo), multiple of 1024
    uint32_t w_tile_bytes;      // 2 * Cpad * row_bytes
    int x_stages, w_stages;     // w_stages: weight tile slots; w_resident: every tile has its own slot, loaded once per CTA
    int w_resident;
    int store32;                // direct epilogue stores as 32-byte st.global.v8 (C % 8 == 0 and 32-byte aligned site rows)
    int debug;
    int prod_groups;            // the 12 producer warps work as 2 or 3 groups; group g fills the stages q = g, g + G, ...
    unsigned long long *timing; // null, or the 16 role cycle counters of aec_net_tc_timing (TcTimingSlot)
};

struct __align__(16) UnitInfo {
    int s, y0, x0, valid;
};

__device__ __forceinline__ uint64_t make_desc_rt(uint32_t smem_addr, int row_bytes)
{
    const uint32_t lo = ((smem_addr & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t hi = row_bytes == 128 ? ((1024u >> 4) | (1u << 14) | (2u << 29))      // SBO 1024, SWIZZLE_128B
                                         : ((512u >> 4) | (1u << 14) | (4u << 29));      // SBO 512,  SWIZZLE_64B
    return ((uint64_t)hi << 32) | lo;
}
// byte offset of 16-byte chunk c of tile row t (tiles start on 1024-byte boundaries: the swizzle uses address bits 7..9 / 7..8)
__device__ __forceinline__ uint32_t rt_off(int t, int c, int row_bytes)
{
    const uint32_t lin = (uint32_t)t * (uint32_t)row_bytes + ((uint32_t)c << 4);
    return lin ^ (((lin >> 7) & (row_bytes == 128 ? 7u : 3u)) << 4);
}

// Dynamic shared memory: [pad to 1024][w_stages x w_tile][x_stages x 4 x x_tile].
// kCB = channels per block (16: 64-byte tile rows, SWIZZLE_64B; 32: 128-byte rows, SWIZZLE_128B) is a template parameter so
// that the MMA issue loop and the producers' index arithmetic are compile-time shapes; kStaged selects the epilogue's store path.
template <int kCB, bool kStaged>
__global__ void __launch_bounds__(kRtThreads, 1) k_conv_rows(const __grid_constant__ RtParams p)
{
    pdl_enter();
    extern __shared__ unsigned char rt_smem_raw[];
    // Weights: a layer whose kh*kw*ncb tiles fit (conv2: 36 KB) keeps them RESIDENT - loaded once per CTA, bar_w_full[0] - because a
    // streamed tile travels commit -> loader -> L2 -> shared memory -> gatekeeper in ~4 k cycles (measured, profiles/r2_summary.md)
    // and a ring of D slots therefore sustains D tiles per 4 k cycles only: less than the tensor core consumes when a tile feeds
    // 8 MMAs.  Larger layers stream tile by tile through bar_w_full / bar_w_empty (conv3: 16 MMAs per tile, 5 slots).
    __shared__ __align__(8) uint64_t bar_x_full[kRtMaxXStages], bar_x_empty[kRtMaxXStages], bar_w_full[kRtMaxWStages], bar_w_empty[kRtMaxWStages];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_u_full[kRtRing], bar_u_free[kRtRing];
    __shared__ uint32_t s_tmem;
    __shared__ UnitInfo s_unit[kRtRing];
    __shared__ __align__(16) float s_etile[kRtEpiWarps][32 * 16];       // epilogue: one 32 sites x 16 channels chunk per warp (transpose buffer)
    __shared__ long long s_edst[kRtEpiWarps][32];                       // epilogue: destination byte offset of the warp's sites (-1: no store)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_units = __shfl_sync(0xffffffffu, *p.counter, 0);
    if (blockIdx.x == 0 && tid == 0 && total_units > 0) {
        atomicAdd(p.accum_sites, (unsigned long long)*p.site_counter);
        atomicAdd(p.accum_units, (unsigned long long)total_units);
    }
    if ((int)blockIdx.x >= total_units) return;

    unsigned char *smem_w = reinterpret_cast<unsigned char *>(((uintptr_t)rt_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_x = smem_w + (size_t)p.w_stages * p.w_tile_bytes;
    const uint32_t x_stage_bytes = 4u * p.x_tile_bytes;

    if (tid == 0) {
        for (int i = 0; i < p.x_stages; ++i) {
            mbar_init(smem_u32(&bar_x_full[i]), 12 / p.prod_groups);      // one arrival per warp of the producer group that fills the stage
            mbar_init(smem_u32(&bar_x_empty[i]), 1);
        }
        for (int i = 0; i < (p.w_resident ? 1 : p.w_stages); ++i) {
            mbar_init(smem_u32(&bar_w_full[i]), 1);
            mbar_init(smem_u32(&bar_w_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_acc_full[i]), 1);
            mbar_init(smem_u32(&bar_acc_empty[i]), kRtEpiWarps * 32);
        }
        for (int i = 0; i < kRtRing; ++i) {
            mbar_init(smem_u32(&bar_u_full[i]), 1);
            mbar_init(smem_u32(&bar_u_free[i]), kRtEpiWarps * 32);          // the epilogue is the last role to finish a unit: its release is enough
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
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
    const int n_units_cta = (total_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_st = p.kh * p.ncb;                       // stages per unit: (kernel row, channel block)
    const int SW = 1 << p.sw_shift;
    const bool timing = p.timing != nullptr;

    if (warp < kRtEpiWarps) {
        // ===================== epilogue: one thread = one site of the tile, two warps per TMEM lane quarter =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRtRegsEpi));
        const int quarter = warp & 3, half = warp >> 2;      // this warp takes the 16-channel chunks ci = half, half + 2, ...
        const int m = quarter * 32 + lane;
        const int rr = m >> p.sw_shift, x = m & (SW - 1);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int cpad = p.Cpad;
        const int nch = (p.C + 15) >> 4;
        const int total = 2 * nch;
        long long tw_acc = 0, tw_si = 0;
        const long long t_begin = timing ? clock64() : 0;
        // Whether a site is in the layer's work set (a site outside it keeps its leaked value: no store) is one bit of a
        // bitmap in global memory: the word of the NEXT unit is fetched while the current one is written out, so that its
        // latency is off the unit's critical path.
        auto fetch = [&](int ul_, UnitInfo &u, uint32_t &word) {
            const int buf = ul_ % kRtRing;
            timed_wait(smem_u32(&bar_u_full[buf]), (uint32_t)(ul_ / kRtRing) & 1u, timing, tw_si);
            u = s_unit[buf];
            const int oy = u.y0 + rr, ox = u.x0 + x;
            u.valid = rr < p.R && x < p.SEG && oy < p.H && ox < p.W;
            word = u.valid ? __ldg(p.nset + ((long long)u.s * p.H + oy) * p.Ww + (ox >> 5)) : 0u;
        };
        UnitInfo u, u_next;
        uint32_t word, word_next = 0u;
        u_next.s = u_next.y0 = u_next.x0 = u_next.valid = 0;
        fetch(0, u, word);
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int buf = ul % kRtRing, ab = ul & 1;
            if (ul + 1 < n_units_cta) fetch(ul + 1, u_next, word_next);
            long long dst = -1;
            {
                const int oy = u.y0 + rr, ox = u.x0 + x;
                if (u.valid && ((word >> (ox & 31)) & 1u)) dst = ((long long)u.s * p.fstride + ((long long)oy * p.W + ox) * p.C) * 4;
            }
            timed_wait(smem_u32(&bar_acc_full[ab]), (uint32_t)(ul >> 1) & 1u, timing, tw_acc);
            tc_fence_after();
            mbar_arrive(smem_u32(&bar_u_free[buf]));          // every stage of this unit is complete: no role needs its slot any more
            const bool site_ok = dst >= 0 && !(p.debug & 8);
            const uint32_t tbase = lane_addr + (uint32_t)(ab * 4 * cpad);
            uint32_t ra[16], rb[16];
            auto issue = [&](int ci, uint32_t(&hi)[16], uint32_t(&lo)[16]) {
                const int map = ci >= nch ? 1 : 0, c0 = (ci - map * nch) << 4;
                const uint32_t ta = tbase + (uint32_t)(map * 2 * cpad + c0);
                tmem_ld16(ta, hi);
                tmem_ld16(ta + (uint32_t)cpad, lo);
            };
            // The accumulator comes out site-per-lane (16 channels = 64 bytes per thread and chunk); stored like that, one store
            // instruction is 32 separate 16-byte requests to 32 different lines, and the requests - not the bytes - bounded the
            // epilogue (5.9 k cycles per unit of conv2 against 4.1 k of MMAs).  So each warp transposes its 32 x 16 chunk through a
            // private 2 KB buffer: lane l then stores piece l % 4 of sites l / 4 + 8 i, four lanes = one contiguous 64-byte run.
            // Buffer rows are 64 bytes; the 16-byte piece index is XORed with (row >> 1) & 3 so that neither the writes (lane =
            // row) nor the reads (4 lanes per row) conflict.
            // Measured (profiles/r2_summary.md): the transpose pays for 64 channels (conv3 0.56 -> 0.50 ms: 8 chunks per unit) and
            // costs for 32 (conv2 0.77 -> 0.87 ms: its two extra warp barriers and the shared-memory round trip per chunk outweigh
            // the better requests), so it is used for C > 32 only.
            constexpr bool staged = kStaged;                 // host: C > 32
            if (staged) {
                s_edst[warp][lane] = site_ok ? dst : -1;
                __syncwarp();
            }
            auto emit = [&](int ci, const uint32_t(&hi)[16], const uint32_t(&lo)[16]) {
                const int map = ci >= nch ? 1 : 0, c0 = (ci - map * nch) << 4;
                if (!staged) {
                    if (!site_ok) return;
                    char *const out = (char *)(map ? p.A : p.F) + dst + (long long)c0 * 4;
                    float4 o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        o[q].x = __fadd_rn(__uint_as_float(hi[4 * q + 0]), __uint_as_float(lo[4 * q + 0]));
                        o[q].y = __fadd_rn(__uint_as_float(hi[4 * q + 1]), __uint_as_float(lo[4 * q + 1]));
                        o[q].z = __fadd_rn(__uint_as_float(hi[4 * q + 2]), __uint_as_float(lo[4 * q + 2]));
                        o[q].w = __fadd_rn(__uint_as_float(hi[4 * q + 3]), __uint_as_float(lo[4 * q + 3]));
                        if (!map && c0 + 4 * q < p.C) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + 4 * q));
                            o[q].x = __fadd_rn(o[q].x, b.x); o[q].y = __fadd_rn(o[q].y, b.y); o[q].z = __fadd_rn(o[q].z, b.z); o[q].w = __fadd_rn(o[q].w, b.w);
                        }
                    }
                    if (p.store32) {
                        // One thread = one site: a 16-byte store per thread is 32 half sectors of 32 different rows per instruction,
                        // and the memory system takes ~0.65 sector requests per cycle and SM whether half or full
                        // (tools/bench_store.cu: 10 B/cycle/SM against 20.5 with st.global.v8 = STG.256, full 32-byte sectors).
#pragma unroll
                        for (int q = 0; q < 4; q += 2) {
                            if (c0 + 4 * q >= p.C) break;                       // C % 8 == 0 on this path
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + 16 * q), "f"(o[q].x), "f"(o[q].y), "f"(o[q].z),
                                         "f"(o[q].w), "f"(o[q + 1].x), "f"(o[q + 1].y), "f"(o[q + 1].z), "f"(o[q + 1].w)
                                         : "memory");
                        }
                        return;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (c0 + 4 * q >= p.C) break;
                        *reinterpret_cast<float4 *>(out + 16 * q) = o[q];
                    }
                    return;
                }
                const uint32_t wbase = smem_u32(&s_etile[warp][0]);
                const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 o;
                    o.x = __fadd_rn(__uint_as_float(hi[4 * q + 0]), __uint_as_float(lo[4 * q + 0]));
                    o.y = __fadd_rn(__uint_as_float(hi[4 * q + 1]), __uint_as_float(lo[4 * q + 1]));
                    o.z = __fadd_rn(__uint_as_float(hi[4 * q + 2]), __uint_as_float(lo[4 * q + 2]));
                    o.w = __fadd_rn(__uint_as_float(hi[4 * q + 3]), __uint_as_float(lo[4 * q + 3]));
                    if (!map && c0 + 4 * q < p.C) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + 4 * q));
                        o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
                    }
                    sts128(wbase + (uint32_t)lane * 64u + (((uint32_t)q ^ sw) << 4), o);
                }
                __syncwarp();
                const int piece = lane & 3;
                char *const outm = (char *)(map ? p.A : p.F) + (long long)(c0 + 4 * piece) * 4;
                const bool piece_ok = c0 + 4 * piece < p.C;                  // C % 4 == 0: whole 16-byte pieces only
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = (lane >> 2) + 8 * i;
                    const long long d = s_edst[warp][row];
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(wbase + (uint32_t)row * 64u + (((uint32_t)piece ^ (uint32_t)((row >> 1) & 3)) << 4)));
                    if (d >= 0 && piece_ok) *reinterpret_cast<float4 *>(outm + d) = v;
                }
                __syncwarp();
            };
#pragma unroll 1
            for (int ci = half; ci < total; ci += 2) {         // the other warp of the quarter takes the chunks in between
                issue(ci, ra, rb);
                tmem_ld_wait();
                emit(ci, ra, rb);
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bar_acc_empty[ab]));
            u = u_next;
            word = word_next;
        }
        if (timing && tid == 0) {
            atomicAdd(p.timing + kTEpiTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTEpiWaitAcc, (unsigned long long)tw_acc);
            atomicAdd(p.timing + kTEpiWaitSite, (unsigned long long)tw_si);
            atomicAdd(p.timing + kTCtas, 1ULL);
            atomicAdd(p.timing + kTUnits, (unsigned long long)n_units_cta);
        }
    } else if (warp < kRtProdWarp0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRtRegsCtl));
      if (warp == kRtMmaWarp) {
        // ===================== MMA issuer (no mbarrier waits here: see aec_tc.cuh) =====================
        // The issue loop is kept as lean as the compiler allows: one thread issues an instruction every few cycles, so a
        // descriptor rebuilt from an address per MMA (shift, mask, or: ~6 uniform-datapath instructions) makes the ISSUE
        // slower than the 64-cycle MMA it feeds (measured: ~104 cycles per MMA).  Descriptors are therefore kept as 32-bit
        // low words (start address >> 4, with the constant LBO bit) that advance by plain adds; the high word is constant.
        const uint32_t idesc_cat = make_idesc_tf32(2 * p.Cpad), idesc_hi = make_idesc_tf32(p.Cpad);
        const uint32_t x_base = smem_u32(smem_x), w_base = smem_u32(smem_w);
        const uint64_t desc_top = make_desc_rt(0u, kCB * 4) & 0xffffffff00000000ull;
        const uint32_t tile16 = p.x_tile_bytes >> 4, wtile16 = p.w_tile_bytes >> 4, xstage16 = x_stage_bytes >> 4;
        const uint32_t x0_lo = ((x_base & 0x3ffffu) >> 4) | (1u << 16), w0_lo = ((w_base & 0x3ffffu) >> 4) | (1u << 16);
        constexpr int ks_n = kCB / 8;
        constexpr uint32_t row16 = (uint32_t)(kCB * 4) >> 4;
        uint32_t sx = 0, sw = 0, qp = 0;
        long long tw_gate = 0;
        const long long t_begin = timing ? clock64() : 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int ab = ul & 1;
            const uint32_t dv = tmem_base + (uint32_t)(ab * 4 * p.Cpad), dr = dv + (uint32_t)(2 * p.Cpad);
            for (int st = 0; st < n_st; ++st, ++qp) {
                // one pass per stage = kernel row: all kw taps behind one hand-off
                const uint32_t xv = x0_lo + sx * xstage16;
                const int ky = st / p.ncb, cb = st - ky * p.ncb;
                const long long t0 = timing ? clock64() : 0;
                asm volatile("bar.sync %0, 64;" ::"r"(1u + (qp & 7u)) : "memory");
                if (timing) tw_gate += clock64() - t0;
                tc_fence_after();
                if (elect_one()) {
                    uint32_t xa = xv, swl = sw;
                    for (int kx = 0; kx < p.kw; ++kx, xa += row16) {         // tap kx: the tile read from kx rows further on
                        const uint32_t slot = p.w_resident ? (uint32_t)((ky * p.kw + kx) * p.ncb + cb) : swl;
                        const uint32_t wl = w0_lo + slot * wtile16;
                        if (!(p.debug & 4)) {
#pragma unroll
                            for (int ks = 0; ks < ks_n; ++ks) {
                                const uint64_t dw = desc_top | (uint64_t)(wl + 2u * ks);
                                const uint32_t acc = (st | kx | ks) != 0 ? 1u : 0u;
                                mma_tf32(dv, desc_top | (uint64_t)(xa + 2u * ks), dw, idesc_cat, acc);
                                mma_tf32(dr, desc_top | (uint64_t)(xa + 2u * tile16 + 2u * ks), dw, idesc_cat, acc);
                                mma_tf32(dv, desc_top | (uint64_t)(xa + tile16 + 2u * ks), dw, idesc_hi, 1u);
                                mma_tf32(dr, desc_top | (uint64_t)(xa + 3u * tile16 + 2u * ks), dw, idesc_hi, 1u);
                            }
                        }
                        if (!p.w_resident) {
                            mma_commit(smem_u32(&bar_w_empty[swl]));         // this tap's weight slot may be refilled
                            if (++swl == (uint32_t)p.w_stages) swl = 0;
                        }
                    }
                    mma_commit(smem_u32(&bar_x_empty[sx]));
                    if (st == n_st - 1) mma_commit(smem_u32(&bar_acc_full[ab]));
                }
                if (!p.w_resident) {
                    sw += (uint32_t)p.kw;
                    if (sw >= (uint32_t)p.w_stages) sw -= (uint32_t)p.w_stages;
                }
                __syncwarp();
                if (++sx == (uint32_t)p.x_stages) sx = 0;
            }
        }
        if (timing && lane == 0) {
            atomicAdd(p.timing + kTMmaTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTMmaWaitX, (unsigned long long)tw_gate);
        }
      } else if (warp == kRtGateWarp) {
        // ===================== gatekeeper: every wait of the MMA warp =====================
        uint32_t qx = 0, qw = 0;
        long long tw_acc = 0, tw_x = 0, tw_w = 0;
        if (p.w_resident) timed_wait(smem_u32(&bar_w_full[0]), 0u, timing, tw_w);       // the resident weights have landed
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int ab = ul & 1;
            const uint32_t ua = (uint32_t)(ul >> 1);
            if (ua > 0) timed_wait(smem_u32(&bar_acc_empty[ab]), (ua - 1) & 1u, timing, tw_acc);
            for (int st = 0; st < n_st; ++st, ++qx) {
                const uint32_t sx = qx % (uint32_t)p.x_stages;
                const uint32_t px = (qx / (uint32_t)p.x_stages) & 1u;
                if (!p.w_resident)
                    for (int kx = 0; kx < p.kw; ++kx, ++qw)
                        timed_wait(smem_u32(&bar_w_full[qw % (uint32_t)p.w_stages]), (qw / (uint32_t)p.w_stages) & 1u, timing, tw_w);
                timed_wait(smem_u32(&bar_x_full[sx]), px, timing, tw_x);
                asm volatile("bar.arrive %0, 64;" ::"r"(1u + (qx & 7u)) : "memory");
            }
        }
        if (timing && lane == 0) {
            atomicAdd(p.timing + kTMmaWaitAcc, (unsigned long long)tw_acc);
            atomicAdd(p.timing + kTMmaWaitW, (unsigned long long)tw_w);
            atomicAdd(p.timing + kTGateWaitX, (unsigned long long)tw_x);
        }
      } else if (warp == kRtLoadWarp) {
        // ===================== weight loader =====================
        if (lane == 0) {
            long long tw_w = 0;
            const long long t_begin = timing ? clock64() : 0;
            const size_t tile_floats = (size_t)p.w_tile_bytes / 4;
            if (p.w_resident) {
                // every tile once, in image order (slot = (ky*kw + kx)*ncb + cb), one barrier for all of them
                const uint32_t n_tiles = (uint32_t)(p.kh * p.kw * p.ncb);
                if (!(p.debug & 16)) {
                    mbar_expect_tx(smem_u32(&bar_w_full[0]), n_tiles * p.w_tile_bytes);
                    for (uint32_t t = 0; t < n_tiles; ++t)
                        bulk_g2s(smem_u32(smem_w + (size_t)t * p.w_tile_bytes), p.wimg + (size_t)t * tile_floats, p.w_tile_bytes, smem_u32(&bar_w_full[0]));
                } else {
                    mbar_arrive(smem_u32(&bar_w_full[0]));
                }
            } else {
                uint32_t qw = 0;
                for (int ul = 0; ul < n_units_cta; ++ul) {
                    for (int st = 0; st < n_st; ++st) {
                        const int ky = st / p.ncb, cb = st - ky * p.ncb;
                        for (int kx = 0; kx < p.kw; ++kx, ++qw) {
                            const int sw = (int)(qw % (uint32_t)p.w_stages);
                            const uint32_t use = qw / (uint32_t)p.w_stages;
                            if (use > 0) timed_wait(smem_u32(&bar_w_empty[sw]), (use - 1) & 1u, timing, tw_w);
                            const float *src = p.wimg + (size_t)((ky * p.kw + kx) * p.ncb + cb) * tile_floats;
                            if (p.debug & 16) { mbar_arrive(smem_u32(&bar_w_full[sw])); continue; }
                            mbar_expect_tx(smem_u32(&bar_w_full[sw]), p.w_tile_bytes);
                            bulk_g2s(smem_u32(smem_w + (size_t)sw * p.w_tile_bytes), src, p.w_tile_bytes, smem_u32(&bar_w_full[sw]));
                        }
                    }
                }
            }
            if (timing) {
                atomicAdd(p.timing + kTLoadTotal, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.timing + kTLoadWaitW, (unsigned long long)tw_w);
            }
        }
      } else if (warp == kRtUnitWarp) {
        // ===================== unit decoder: work-list entry -> (stream, first row, first column) =====================
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int unit = blockIdx.x + ul * gridDim.x;
            const int buf = ul % kRtRing;
            const uint32_t us = (uint32_t)(ul / kRtRing);
            const uint32_t e = __ldg(p.units + unit);
            if (us > 0) mbar_wait(smem_u32(&bar_u_free[buf]), (us - 1) & 1u);
            if (lane == 0) {
                UnitInfo u;
                int yg, xg;
                site_decode(p.code, e, u.s, yg, xg);
                u.y0 = yg * p.R;
                u.x0 = xg * p.SEG;
                u.valid = 1;
                s_unit[buf] = u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_u_full[buf]));
        }
      }
    } else {
        // ===================== producers: G groups of warps, group g fills the stages q = g, g + G, ... =====================
        // Pair i of thread pt is (tile row, chunk) number pt + T i of the stage's P x CB/4 pairs (T = threads of the group):
        // consecutive threads take consecutive 16-byte chunks of consecutive pixels, i.e. consecutive addresses of the
        // channel-last source row.  A stage is one dependent chain - global loads, conversion, shared-memory stores, the proxy
        // fence (whose MEMBAR waits for every access the thread has in flight) and the arrive - of about one memory latency; the
        // groups run their chains side by side, which is what hides that latency (a single group of 12 warps with the next
        // stage's loads in flight across the fence was producer-bound: 2.3 k cycles per stage, profiles/r2_summary.md).
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRtRegsProd));
        const int G = p.prod_groups;
        const int gthreads = (12 / G) * 32;
        const int pidx = tid - kRtProdWarp0 * 32;
        const int g = pidx / gthreads, pt = pidx - g * gthreads;
        constexpr int cshift = kCB == 32 ? 3 : 2;             // chunks per tile row = CB / 4
        const int n_pairs = p.P << cshift;
        const uint32_t x_base = smem_u32(smem_x);
        const uint32_t q_end = (uint32_t)n_units_cta * (uint32_t)n_st;
        long long tw_si = 0, tw_x = 0;
        const long long t_begin = timing ? clock64() : 0;
        int ready_ul = -1;
        UnitInfo cur_u;
        cur_u.s = cur_u.y0 = cur_u.x0 = cur_u.valid = 0;
        int ul = 0, st = g;
        while (st >= n_st) { st -= n_st; ++ul; }
        for (uint32_t q = (uint32_t)g; q < q_end; q += (uint32_t)G) {
            if (ul > ready_ul) {
                const int buf = ul % kRtRing;
                timed_wait(smem_u32(&bar_u_full[buf]), (uint32_t)(ul / kRtRing) & 1u, timing && pt == 0, tw_si);
                cur_u = s_unit[buf];
                ready_ul = ul;
            }
            const int ky = st / p.ncb, cb = st - ky * p.ncb;
            const char *base = reinterpret_cast<const char *>(p.srcF + (long long)cur_u.s * p.src_stride + cb * kCB);
            float4 f[kRtMaxPairs], a[kRtMaxPairs];
#pragma unroll
            for (int i = 0; i < kRtMaxPairs; ++i) {
                const int pr = pt + gthreads * i;
                const int t = pr >> cshift, ch = pr & ((1 << cshift) - 1);
                const int rr = t >> p.sw_shift, j = t & (SW - 1);
                const int iy = cur_u.y0 + rr + ky - p.pad_t, ix = cur_u.x0 + j - p.pad_l;
                const bool ok = pr < n_pairs && rr < p.R && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win && !(p.debug & 1);
                const char *pf = ok ? base + ((long long)(iy * p.Win + ix) * p.Cin + 4 * ch) * 4 : p.zero_f;
                f[i] = __ldg(reinterpret_cast<const float4 *>(pf));
                a[i] = __ldg(reinterpret_cast<const float4 *>(pf + p.a_minus_f));
            }
            const uint32_t sx = q % (uint32_t)p.x_stages, use = q / (uint32_t)p.x_stages;
            if (use > 0) timed_wait(smem_u32(&bar_x_empty[sx]), (use - 1) & 1u, timing && pt == 0, tw_x);
            const uint32_t xs = x_base + sx * x_stage_bytes;
            if (!(p.debug & 2)) {
#pragma unroll
                for (int i = 0; i < kRtMaxPairs; ++i) {
                    const int pr = pt + gthreads * i;
                    if (pr >= n_pairs) break;
                    const int t = pr >> cshift, ch = pr & ((1 << cshift) - 1);
                    const float2 s01 = make_float2(slope_of(f[i].x, p.alpha), slope_of(f[i].y, p.alpha));
                    const float2 s23 = make_float2(slope_of(f[i].z, p.alpha), slope_of(f[i].w, p.alpha));
                    const float2 v01 = __fmul2_rn(make_float2(f[i].x, f[i].y), s01), v23 = __fmul2_rn(make_float2(f[i].z, f[i].w), s23);
                    const float2 w01 = __fmul2_rn(make_float2(a[i].x, a[i].y), s01), w23 = __fmul2_rn(make_float2(a[i].z, a[i].w), s23);
                    float4 h, l;
                    const uint32_t o = xs + rt_off(t, ch, kCB * 4);
                    split2(v01.x, v01.y, h.x, h.y, l.x, l.y);
                    split2(v23.x, v23.y, h.z, h.w, l.z, l.w);
                    sts128(o, h);
                    sts128(o + p.x_tile_bytes, l);
                    split2(w01.x, w01.y, h.x, h.y, l.x, l.y);
                    split2(w23.x, w23.y, h.z, h.w, l.z, l.w);
                    sts128(o + 2u * p.x_tile_bytes, h);
                    sts128(o + 3u * p.x_tile_bytes, l);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_x_full[sx]));
            st += G;
            while (st >= n_st) { st -= n_st; ++ul; }
        }
        if (timing && pidx == 0) {
            atomicAdd(p.timing + kTProdTotal, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.timing + kTProdWaitSite, (unsigned long long)tw_si);
            atomicAdd(p.timing + kTProdWaitStage, (unsigned long long)tw_x);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace rt
}  // namespace aec
