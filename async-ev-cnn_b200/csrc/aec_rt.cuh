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

constexpr int kRtMaxXStages = 3;
constexpr int kRtMaxWStages = 6;
constexpr int kRtRing = 4;            // unit-info buffers
constexpr int kRtProdThreads = kGroups * kGroupThreads;   // 384
constexpr int kRtMaxPairs = 3;        // (tile row, 16-byte chunk) pairs per producer thread and stage

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
    int x_stages, w_stages;
    int debug;
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
__global__ void __launch_bounds__(kTcThreads, 1) k_conv_rows(const __grid_constant__ RtParams p)
{
    extern __shared__ unsigned char rt_smem_raw[];
    __shared__ __align__(8) uint64_t bar_x_full[kRtMaxXStages], bar_x_empty[kRtMaxXStages], bar_w_full[kRtMaxWStages], bar_w_empty[kRtMaxWStages];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_u_full[kRtRing], bar_u_free[kRtRing];
    __shared__ uint32_t s_tmem;
    __shared__ UnitInfo s_unit[kRtRing];

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
            mbar_init(smem_u32(&bar_x_full[i]), kRtProdThreads / 32);      // one arrival per producer warp
            mbar_init(smem_u32(&bar_x_empty[i]), 1);
        }
        for (int i = 0; i < p.w_stages; ++i) {
            mbar_init(smem_u32(&bar_w_full[i]), 1);
            mbar_init(smem_u32(&bar_w_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_acc_full[i]), 1);
            mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps * 32);
        }
        for (int i = 0; i < kRtRing; ++i) {
            mbar_init(smem_u32(&bar_u_full[i]), 1);
            mbar_init(smem_u32(&bar_u_free[i]), kEpiWarps * 32 + kRtProdThreads);   // epilogue and producers both read the unit info
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

    if (warp < kEpiWarps) {
        // ===================== epilogue: one thread = one site of the tile =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        const int m = warp * 32 + lane;
        const int rr = m >> p.sw_shift, x = m & (SW - 1);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int cpad = p.Cpad;
        const int nch = (p.C + 15) >> 4;
        const int total = 2 * nch;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int buf = ul % kRtRing, ab = ul & 1;
            mbar_wait(smem_u32(&bar_u_full[buf]), (uint32_t)(ul / kRtRing) & 1u);
            const UnitInfo u = s_unit[buf];
            // is this site in the layer's work set?  (a site outside it keeps its leaked value: no store)
            long long dst = -1;
            {
                const int oy = u.y0 + rr, ox = u.x0 + x;
                if (u.valid && rr < p.R && x < p.SEG && oy < p.H && ox < p.W) {
                    const uint32_t word = __ldg(p.nset + ((long long)u.s * p.H + oy) * p.Ww + (ox >> 5));
                    if ((word >> (ox & 31)) & 1u) dst = ((long long)u.s * p.fstride + ((long long)oy * p.W + ox) * p.C) * 4;
                }
            }
            mbar_arrive(smem_u32(&bar_u_free[buf]));
            mbar_wait(smem_u32(&bar_acc_full[ab]), (uint32_t)(ul >> 1) & 1u);
            tc_fence_after();
            const bool site_ok = dst >= 0 && !(p.debug & 8);
            const uint32_t tbase = lane_addr + (uint32_t)(ab * 4 * cpad);
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
                    if (c0 + 4 * q >= p.C) break;
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
                issue(ci + 1, rc, rd);
                emit(ci, ra, rb);
                tmem_ld_wait();
                if (ci + 2 < total) issue(ci + 2, ra, rb);
                emit(ci + 1, rc, rd);
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bar_acc_empty[ab]));
        }
    } else if (warp < kProdWarp0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
      if (warp == kMmaWarp) {
        // ===================== MMA issuer (no mbarrier waits here: see aec_tc.cuh) =====================
        const uint32_t idesc_cat = make_idesc_tf32(2 * p.Cpad), idesc_hi = make_idesc_tf32(p.Cpad);
        const uint32_t x_base = smem_u32(smem_x), w_base = smem_u32(smem_w);
        const int ks_n = p.CB / 8;
        uint32_t qx = 0, qw = 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int ab = ul & 1;
            const uint32_t dv = tmem_base + (uint32_t)(ab * 4 * p.Cpad), dr = dv + (uint32_t)(2 * p.Cpad);
            for (int st = 0; st < n_st; ++st, ++qx) {
                const uint32_t sx = qx % (uint32_t)p.x_stages;
                const uint32_t xs = x_base + sx * x_stage_bytes;
                for (int kx = 0; kx < p.kw; ++kx, ++qw) {
                    const uint32_t sw = qw % (uint32_t)p.w_stages;
                    const uint32_t wt = w_base + sw * p.w_tile_bytes;
                    const uint32_t xa = xs + (uint32_t)kx * (uint32_t)p.row_bytes;     // tap kx: the tile read from kx rows further on
                    asm volatile("bar.sync %0, 64;" ::"r"(1u + (qw & 7u)) : "memory");     // ids 1..8: the gatekeeper is at most w_stages <= 6 passes ahead
                    tc_fence_after();
                    if (elect_one()) {
                        if (!(p.debug & 4)) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t ko = (uint32_t)ks * 32u;
                                const uint64_t dw = make_desc_rt(wt + ko, p.row_bytes);
                                const uint64_t dvh = make_desc_rt(xa + ko, p.row_bytes), dvl = make_desc_rt(xa + p.x_tile_bytes + ko, p.row_bytes);
                                const uint64_t drh = make_desc_rt(xa + 2u * p.x_tile_bytes + ko, p.row_bytes), drl = make_desc_rt(xa + 3u * p.x_tile_bytes + ko, p.row_bytes);
                                if (ks >= ks_n) break;
                                const uint32_t acc = (st | kx | ks) != 0 ? 1u : 0u;
                                mma_tf32(dv, dvh, dw, idesc_cat, acc);
                                mma_tf32(dr, drh, dw, idesc_cat, acc);
                                mma_tf32(dv, dvl, dw, idesc_hi, 1u);
                                mma_tf32(dr, drl, dw, idesc_hi, 1u);
                            }
                        }
                        mma_commit(smem_u32(&bar_w_empty[sw]));
                        if (kx == p.kw - 1) mma_commit(smem_u32(&bar_x_empty[sx]));
                        if (kx == p.kw - 1 && st == n_st - 1) mma_commit(smem_u32(&bar_acc_full[ab]));
                    }
                    __syncwarp();
                }
            }
        }
      } else if (warp == kGateWarp) {
        // ===================== gatekeeper: every wait of the MMA warp =====================
        uint32_t qx = 0, qw = 0;
        for (int ul = 0; ul < n_units_cta; ++ul) {
            const int ab = ul & 1;
            const uint32_t ua = (uint32_t)(ul >> 1);
            if (ua > 0) mbar_wait(smem_u32(&bar_acc_empty[ab]), (ua - 1) & 1u);
            for (int st = 0; st < n_st; ++st, ++qx) {
                const uint32_t sx = qx % (uint32_t)p.x_stages;
                const uint32_t px = (qx / (uint32_t)p.x_stages) & 1u;
                for (int kx = 0; kx < p.kw; ++kx, ++qw) {
                    const uint32_t sw = qw % (uint32_t)p.w_stages;
                    mbar_wait(smem_u32(&bar_w_full[sw]), (qw / (uint32_t)p.w_stages) & 1u);
                    if (kx == 0) mbar_wait(smem_u32(&bar_x_full[sx]), px);
                    asm volatile("bar.arrive %0, 64;" ::"r"(1u + (qw & 7u)) : "memory");
                }
            }
        }
      } else if (warp == kLoadWarp) {
        // ===================== weight loader: tile (ky, kx, cb) per pass, in the order the MMA warp consumes them =====================
        if (lane == 0) {
            uint32_t qw = 0;
            const size_t tile_floats = (size_t)p.w_tile_bytes / 4;
            for (int ul = 0; ul < n_units_cta; ++ul) {
                for (int st = 0; st < n_st; ++st) {
                    const int ky = st / p.ncb, cb = st - ky * p.ncb;
                    for (int kx = 0; kx < p.kw; ++kx, ++qw) {
                        const int sw = (int)(qw % (uint32_t)p.w_stages);
                        const uint32_t use = qw / (uint32_t)p.w_stages;
                        if (use > 0) mbar_wait(smem_u32(&bar_w_empty[sw]), (use - 1) & 1u);
                        const float *src = p.wimg + (size_t)((ky * p.kw + kx) * p.ncb + cb) * tile_floats;
                        if (p.debug & 16) { mbar_arrive(smem_u32(&bar_w_full[sw])); continue; }
                        mbar_expect_tx(smem_u32(&bar_w_full[sw]), p.w_tile_bytes);
                        bulk_g2s(smem_u32(smem_w + (size_t)sw * p.w_tile_bytes), src, p.w_tile_bytes, smem_u32(&bar_w_full[sw]));
                    }
                }
            }
        }
      } else if (warp == kSiteWarp) {
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
        // ===================== producers: all 12 warps fill one stage together =====================
        // Pair i of thread pt is (tile row, chunk) number pt + 384 i of the stage's P x CB/4 pairs: consecutive threads take
        // consecutive 16-byte chunks of consecutive pixels, i.e. consecutive addresses of the channel-last source row.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsProd));
        const int pt = tid - kProdWarp0 * 32;
        const int cshift = p.CB == 32 ? 3 : 2;                // chunks per tile row = CB / 4
        const int n_pairs = p.P << cshift;
        const uint32_t x_base = smem_u32(smem_x);
        struct Pos { int ul, st; uint32_t q; };
        const uint32_t q_end = (uint32_t)n_units_cta * (uint32_t)n_st;
        int ready_ul = -1;
        UnitInfo cur_u;
        cur_u.s = cur_u.y0 = cur_u.x0 = cur_u.valid = 0;
        auto load = [&](const Pos &c, float4 (&f)[kRtMaxPairs], float4 (&a)[kRtMaxPairs]) {
            if (c.ul > ready_ul) {
                const int buf = c.ul % kRtRing;
                mbar_wait(smem_u32(&bar_u_full[buf]), (uint32_t)(c.ul / kRtRing) & 1u);
                cur_u = s_unit[buf];
                mbar_arrive(smem_u32(&bar_u_free[buf]));
                ready_ul = c.ul;
            }
            const int ky = c.st / p.ncb, cb = c.st - ky * p.ncb;
            const char *base = reinterpret_cast<const char *>(p.srcF + (long long)cur_u.s * p.src_stride + cb * p.CB);
#pragma unroll
            for (int i = 0; i < kRtMaxPairs; ++i) {
                const int pr = pt + kRtProdThreads * i;
                const int t = pr >> cshift, ch = pr & ((1 << cshift) - 1);
                const int rr = t >> p.sw_shift, j = t & (SW - 1);
                const int iy = cur_u.y0 + rr + ky - p.pad_t, ix = cur_u.x0 + j - p.pad_l;
                const bool ok = pr < n_pairs && rr < p.R && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win && !(p.debug & 1);
                const char *pf = ok ? base + ((long long)(iy * p.Win + ix) * p.Cin + 4 * ch) * 4 : p.zero_f;
                f[i] = __ldg(reinterpret_cast<const float4 *>(pf));
                a[i] = __ldg(reinterpret_cast<const float4 *>(pf + p.a_minus_f));
            }
        };
        auto store = [&](const Pos &c, const float4 (&f)[kRtMaxPairs], const float4 (&a)[kRtMaxPairs]) {
            const uint32_t sx = c.q % (uint32_t)p.x_stages, use = c.q / (uint32_t)p.x_stages;
            if (use > 0) mbar_wait(smem_u32(&bar_x_empty[sx]), (use - 1) & 1u);
            const uint32_t xs = x_base + sx * x_stage_bytes;
            if (!(p.debug & 2)) {
#pragma unroll
                for (int i = 0; i < kRtMaxPairs; ++i) {
                    const int pr = pt + kRtProdThreads * i;
                    if (pr >= n_pairs) break;
                    const int t = pr >> cshift, ch = pr & ((1 << cshift) - 1);
                    const float2 s01 = make_float2(slope_of(f[i].x, p.alpha), slope_of(f[i].y, p.alpha));
                    const float2 s23 = make_float2(slope_of(f[i].z, p.alpha), slope_of(f[i].w, p.alpha));
                    const float2 v01 = __fmul2_rn(make_float2(f[i].x, f[i].y), s01), v23 = __fmul2_rn(make_float2(f[i].z, f[i].w), s23);
                    const float2 w01 = __fmul2_rn(make_float2(a[i].x, a[i].y), s01), w23 = __fmul2_rn(make_float2(a[i].z, a[i].w), s23);
                    float4 h, l;
                    const uint32_t o = xs + rt_off(t, ch, p.row_bytes);
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
        };
        auto advance = [&](Pos &c) {
            ++c.q;
            if (++c.st == n_st) { c.st = 0; ++c.ul; }
        };
        Pos cur;
        cur.ul = 0; cur.st = 0; cur.q = 0;
        float4 fa[kRtMaxPairs], aa[kRtMaxPairs], fb[kRtMaxPairs], ab4[kRtMaxPairs];
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

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace rt
}  // namespace aec
