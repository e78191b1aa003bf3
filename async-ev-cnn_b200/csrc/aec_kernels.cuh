// aec_kernels.cuh - sm_100a kernels of the event-driven EFCN hot path.
//
// Data layout in HBM (per network, S streams; the stream index is always outermost):
//   surface   double [S][H*W]                     leaky integration surface (f64: SURVEY Q2)
//   F_l, A_l  float  [S][pad4(H_l*W_l*C_l)]        conv pre-activation / leak-rate maps, CHANNEL-LAST
//   idx_l     uint8  [S][Ho*Wo*C]                  pool argmax row (ky*kw+kx), channel-last
//   Fp_l,Ap_l float  [S][pad4(Ho*Wo*C)]            pool layers: copy of the previous conv's F / A at the
//                                                  stored argmax (kept equal by the leak sweep, which
//                                                  applies the same arithmetic to the copy, and by the
//                                                  pool evaluation, which refreshes re-evaluated windows);
//                                                  lets every consumer gather 16-byte channel vectors
//   flags_l   uint32 [S][Ho][ceil(Wo/32)]          pool sticky recompute bitmap
//   front_l   uint32 [S][H_l][ceil(W_l/32)]        output-event ("frontier") bitmap of layer l
//   signchg_l uint32 [S][H_l][ceil(W_l/32)]        conv sites whose sign flipped in the leak sweep
//   nzr_l     uint32 [S][H_l][ceil(W_l/32)]        every layer: sites whose leak rate CAN be non-zero.  Layer 0: pixels
//                                                  with S > 0 (R = [S > 0]).  Conv: a re-evaluated site gets the OR of
//                                                  its receptive field's input bits (A = W.patch(R) is exactly 0 when
//                                                  every input rate is 0), other sites keep theirs.  Pool: OR over the
//                                                  window.  Maintained by the frontier kernels; the leak sweep skips
//                                                  sites whose bit is 0 (bit 0 => rate exactly 0 => F unchanged)
//   sites     uint32 [S*max(H_l*W_l)]              gathered work list of the layer being updated, streams batched:
//                                                  entry = stream << sh_s | y << sh_y | x with sh_y = bits(W_l - 1),
//                                                  sh_s = sh_y + bits(H_l - 1) (SiteCode): sorted like the row-major
//                                                  site index, decoded with two shifts instead of two divisions
//
// Semantics follow the reference line by line (citations at each kernel); the work decomposition is
// new: frontiers are bitmaps (dedup and ordering for free), the leak of ALL conv layers is one
// vectorised sweep, and changed sites of all streams are gathered into one list per layer so the
// re-evaluation runs as a dense gathered GEMM over (sites x {value,rate}) x k*k*Cin x Cout.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace aec {

// Programmatic dependent launch: every kernel of the step starts with this.  Launched with the programmatic-serialization
// attribute (aec.cu: launch_k) a kernel may be scheduled while its predecessor in the stream is still draining - its CTAs take
// the SMs as they become free and wait here until the predecessor grid has completed and its writes are visible; then they
// allow the next kernel of the chain to be scheduled the same way.  Without the attribute both instructions do nothing.
// Measured (profiles/r2_summary.md, finding 14): no gain for one stream (0.233 vs 0.236 ms) and a loss for many (1024 streams:
// 3.73 vs 3.66 ms per step) - the step is not bound by launch gaps - so the attribute is off unless AEC_PDL=1.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory");
}

constexpr int kThreads = 256;
constexpr int kMaxConv = 24;           // conv layers per network
constexpr int kMaxSweep = 2 * kMaxConv;  // leak-sweep table: conv maps + pooled copies

// ---------------------------------------------------------------------------------------------
// What a consumer sees of its previous layer: value V = surface*layer_actfn (layer.py:77-81) and
// rate R = conv_actfn.  kind 0: integration (V=(float)S, R=[S>0]; integration.py:33-43),
// kind 1: a channel-last (F, A) map pair: V=F*slope, R=A*slope, slope = F>0 ? 1 : alpha
//         (conv2d.py:83-94).  For a conv layer the pair is its own state; for a pool layer it is the
//         materialised copy of the previous conv's maps at the stored argmax, which is what
//         maxpool.py:42-79 gathers (surface = F[argmax], conv_actfn = (A*slope)[argmax]).
// ---------------------------------------------------------------------------------------------
struct Src {
    int kind;
    int C, H, W;            // shape of the previous layer's output map
    const double *S;        // kind 0
    long long sstride;
    const float *F, *A;     // kind 1
    long long fstride;
    float alpha;
};

// Work-list entry coding of one layer (see `sites` above).
struct SiteCode {
    int sh_y, sh_s;
};
__device__ __forceinline__ uint32_t site_encode(const SiteCode c, int s, int y, int x)
{
    return ((uint32_t)s << c.sh_s) | ((uint32_t)y << c.sh_y) | (uint32_t)x;
}
__device__ __forceinline__ void site_decode(const SiteCode c, uint32_t e, int &s, int &y, int &x)
{
    s = (int)(e >> c.sh_s);
    y = (int)((e >> c.sh_y) & ((1u << (c.sh_s - c.sh_y)) - 1u));
    x = (int)(e & ((1u << c.sh_y) - 1u));
}

__device__ __forceinline__ float slope_of(float f, float alpha) { return f > 0.f ? 1.f : alpha; }

__device__ __forceinline__ void src_fetch(const Src &q, int s, int y, int x, int c, float &v, float &r)
{
    if (q.kind == 0) {
        const double sv = q.S[(long long)s * q.sstride + (long long)y * q.W + x];
        v = __double2float_rn(sv);
        r = sv > 0.0 ? 1.f : 0.f;
        return;
    }
    const long long off = (long long)s * q.fstride + ((long long)y * q.W + x) * q.C + c;
    const float f = q.F[off];
    const float a = q.A[off];
    const float sl = slope_of(f, q.alpha);
    v = __fmul_rn(f, sl);
    r = __fmul_rn(a, sl);
}

// 4 consecutive channels (c % 4 == 0, C % 4 == 0, kind 1).
__device__ __forceinline__ void src_fetch4(const Src &q, int s, int y, int x, int c, float4 &v, float4 &r)
{
    const long long off = (long long)s * q.fstride + ((long long)y * q.W + x) * q.C + c;
    const float4 f = __ldg(reinterpret_cast<const float4 *>(q.F + off));
    const float4 a = __ldg(reinterpret_cast<const float4 *>(q.A + off));
    float sl;
    sl = slope_of(f.x, q.alpha); v.x = __fmul_rn(f.x, sl); r.x = __fmul_rn(a.x, sl);
    sl = slope_of(f.y, q.alpha); v.y = __fmul_rn(f.y, sl); r.y = __fmul_rn(a.y, sl);
    sl = slope_of(f.z, q.alpha); v.z = __fmul_rn(f.z, sl); r.z = __fmul_rn(a.z, sl);
    sl = slope_of(f.w, q.alpha); v.w = __fmul_rn(f.w, sl); r.w = __fmul_rn(a.w, sl);
}

// ---------------------------------------------------------------------------------------------
// small block-level helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// Exclusive prefix sum of one int per thread over the block (kThreads threads); returns the
// exclusive prefix, writes the block total to *total.  `scratch` = 9 ints of shared memory.
__device__ __forceinline__ int block_excl_scan(int v, int *scratch, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v);
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (kThreads / 32) ? scratch[lane] : 0;
        const int wi = warp_incl_scan(w);
        if (lane < (kThreads / 32)) scratch[lane] = wi - w;
        if (lane == (kThreads / 32) - 1) scratch[8] = wi;
    }
    __syncthreads();
    const int res = scratch[wid] + incl - v;
    *total = scratch[8];
    __syncthreads();
    return res;
}

// Word w of a bitmap row shifted by d bit positions towards higher x (d may be negative).
__device__ __forceinline__ uint32_t row_shift(const uint32_t *row, int nwords, int w, int d)
{
    if (d == 0) return row[w];
    if (d > 0) {
        const int ws = d >> 5, b = d & 31;
        const int i = w - ws;
        uint32_t lo = (i >= 0 && i < nwords) ? row[i] : 0u;
        if (b == 0) return lo;
        uint32_t prev = (i - 1 >= 0 && i - 1 < nwords) ? row[i - 1] : 0u;
        return (lo << b) | (prev >> (32 - b));
    }
    const int e = -d, ws = e >> 5, b = e & 31;
    const int i = w + ws;
    uint32_t lo = (i >= 0 && i < nwords) ? row[i] : 0u;
    if (b == 0) return lo;
    uint32_t next = (i + 1 < nwords && i + 1 >= 0) ? row[i + 1] : 0u;
    return (lo >> b) | (next << (32 - b));
}

__device__ __forceinline__ uint32_t compress_even_bits(uint32_t x)
{
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

// Emits the set bits of a [H][Ww] bitmap held in shared memory as list entries
// base_id | y << sh_y | x (base_id = stream << sh_s), row-major, into sites[] at a range reserved with one atomicAdd on *counter.
__device__ __forceinline__ void emit_sites(const uint32_t *bm, int H, int Ww, uint32_t base_id, int sh_y,
                                           uint32_t *sites, int *counter, int *scratch)
{
    const int nwords = H * Ww;
    const int per = (nwords + kThreads - 1) / kThreads;
    const int w0 = threadIdx.x * per;
    const int w1 = min(nwords, w0 + per);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
    int total;
    int off = block_excl_scan(cnt, scratch, &total);
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = total > 0 ? atomicAdd(counter, total) : 0;
    __syncthreads();
    off += s_base;
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = bm[w];
        const int y = w / Ww, xb = (w - y * Ww) * 32;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            sites[off++] = base_id | ((uint32_t)y << sh_y) | (uint32_t)(xb + b);
        }
    }
}

__device__ __forceinline__ uint32_t spread_to_pairs(uint32_t x)      // bit b of the low 16 bits -> bits 2b and 2b + 1
{
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

// Work list of a conv layer whose 2x2 / stride-2 pool is evaluated in the conv kernel's epilogue (aec_tc.cuh, kPool): a pool
// window ALL FOUR sites of which are in the work set (`quads`, a [H/2][pWw] bitmap over the windows) is emitted as four
// consecutive entries in the window's row order (0,0) (0,1) (1,0) (1,1), the first carrying `quad_bit`; the epilogue thread
// that owns a channel then has the window's four values in adjacent accumulator columns.  All quads of the stream come first,
// then the remaining sites row-major, then 0xffffffff entries up to a multiple of four: every stream's block starts on a
// multiple of four, so a quad never straddles two 128-entry units.  *site_counter counts the real sites.
__device__ __forceinline__ void emit_sites_quads(const uint32_t *bm, const uint32_t *quads, int H, int Ww, int pWw, uint32_t base_id, int sh_y,
                                                 uint32_t quad_bit, uint32_t *sites, int *counter, int *site_counter, int *scratch)
{
    const int nwords = H * Ww, pwords = (H >> 1) * pWw;
    const int per = (nwords + kThreads - 1) / kThreads, pper = (pwords + kThreads - 1) / kThreads;
    const int w0 = threadIdx.x * per, w1 = min(nwords, w0 + per);
    const int q0 = threadIdx.x * pper, q1 = min(pwords, q0 + pper);
    auto singles = [&](int w) {
        const int y = w / Ww, xw = w - y * Ww;
        return bm[w] & ~spread_to_pairs(quads[(y >> 1) * pWw + (xw >> 1)] >> ((xw & 1) * 16));
    };
    int cq = 0, cs = 0;
    for (int q = q0; q < q1; ++q) cq += __popc(quads[q]);
    for (int w = w0; w < w1; ++w) cs += __popc(singles(w));
    int total_q, total_s;
    int off_q = block_excl_scan(cq, scratch, &total_q);
    int off_s = block_excl_scan(cs, scratch, &total_s);
    const int pad = (4 - (total_s & 3)) & 3;
    __shared__ int s_qbase;
    if (threadIdx.x == 0) {
        const int total = 4 * total_q + total_s;
        s_qbase = total > 0 ? atomicAdd(counter, total + pad) : 0;
        if (total > 0) atomicAdd(site_counter, total);
    }
    __syncthreads();
    off_q = s_qbase + 4 * off_q;
    off_s += s_qbase + 4 * total_q;
    for (int q = q0; q < q1; ++q) {
        uint32_t bits = quads[q];
        const int oy = q / pWw, xb = (q - oy * pWw) * 32;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const uint32_t e = base_id | ((uint32_t)(2 * oy) << sh_y) | (uint32_t)(2 * (xb + b));
            sites[off_q] = e | quad_bit;
            sites[off_q + 1] = e + 1u;
            sites[off_q + 2] = e + (1u << sh_y);
            sites[off_q + 3] = e + (1u << sh_y) + 1u;
            off_q += 4;
        }
    }
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = singles(w);
        const int y = w / Ww, xb = (w - y * Ww) * 32;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            sites[off_s++] = base_id | ((uint32_t)y << sh_y) | (uint32_t)(xb + b);
        }
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < pad; ++i) sites[s_qbase + 4 * total_q + total_s + i] = 0xffffffffu;
}

// Row-tile layers: emits one entry per unit (rt_rows output rows x one x segment) that holds a work-set bit,
// entry = base_id | row group << sh_y | segment, in (row group, segment) order.
__device__ __forceinline__ void emit_units(const uint32_t *bm, int H, int W, int Ww, int rows, int seg, int nxg, uint32_t base_id, int sh_y,
                                           uint32_t *units, int *counter, int *scratch)
{
    const int nyg = (H + rows - 1) / rows;
    const int n_units = nyg * nxg;
    const int per = (n_units + kThreads - 1) / kThreads;
    const int u0 = threadIdx.x * per, u1 = min(n_units, u0 + per);
    auto any = [&](int u) {
        const int yg = u / nxg, xg = u - yg * nxg;
        const int c0 = xg * seg, c1 = min(W, c0 + seg);
        const int w0 = c0 >> 5, w1 = (c1 - 1) >> 5;
        for (int y = yg * rows; y < min(H, (yg + 1) * rows); ++y)
            for (int w = w0; w <= w1; ++w) {
                uint32_t m = 0xffffffffu;
                if (w == w0) m &= 0xffffffffu << (c0 & 31);
                if (w == w1 && (c1 & 31)) m &= (1u << (c1 & 31)) - 1u;
                if (bm[y * Ww + w] & m) return true;
            }
        return false;
    };
    uint32_t hits = 0u;                     // per <= 32 units per thread for any frame the bitmaps of which fit in shared memory
    int cnt = 0;
    for (int u = u0; u < u1; ++u)
        if (any(u)) { hits |= 1u << (u - u0); ++cnt; }
    int total;
    int off = block_excl_scan(cnt, scratch, &total);
    __shared__ int s_ubase;
    if (threadIdx.x == 0) s_ubase = total > 0 ? atomicAdd(counter, total) : 0;
    __syncthreads();
    off += s_ubase;
    while (hits) {
        const int b = __ffs(hits) - 1;
        hits &= hits - 1;
        const int u = u0 + b, yg = u / nxg;
        units[off++] = base_id | ((uint32_t)yg << sh_y) | (uint32_t)(u - yg * nxg);
    }
}

// ---------------------------------------------------------------------------------------------
// K1: integration surface.  One CTA per stream.   integration.py:53-91
//   t_L = max ts; delta = (t_L - t_prev) * leak; S = max(S - delta, 0);
//   S[p] += 1 - (t_L - ts_j)*leak for the LAST event j on each distinct pixel p (numpy fancy `+=`);
//   clamp again; frontier = {alive before, dead after} U {event pixels}.
// Dynamic shared memory: hash_slots * 8 bytes (last-wins hash) + 2 * H*Ww*4 (frontier and alive bitmaps).
// ---------------------------------------------------------------------------------------------
struct IntegrateParams {
    double *surface;        // [S][HW]
    int *prev_ts;           // [S]
    double *delta;          // [S]
    uint8_t *active;        // [S]
    uint32_t *front;        // [S][H*Ww]
    uint32_t *alive;        // [S][H*Ww] pixels with S > 0 after the step (the layer's non-zero-rate bitmap)
    const int32_t *events;  // [total][3] (y,x,ts)
    const int32_t *offsets; // [S+1] (or [S] first event of every stream when `ends` is given)
    const int32_t *ends;    // null: stream s owns events offsets[s] .. offsets[s+1]; else offsets[s] .. ends[s] (batched recordings)
    int *layer_counts;      // [n_layers] work-list counters, zeroed here for the step
    int *err_flag;
    int n_layers;
    int H, W, Ww;
    double leak;
    int max_events, hash_slots;
};

__global__ void __launch_bounds__(kThreads) k_integrate(IntegrateParams p)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *hkey = reinterpret_cast<int *>(smem_raw);
    int *hval = hkey + p.hash_slots;
    uint32_t *bm = reinterpret_cast<uint32_t *>(hval + p.hash_slots);
    uint32_t *al = bm + p.H * p.Ww;
    __shared__ int s_red[kThreads / 32];
    __shared__ int s_tlast;

    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const int HW = p.H * p.W;
    const int nbm = p.H * p.Ww;
    if (s == 0 && tid < p.n_layers) { p.layer_counts[tid] = 0; p.layer_counts[32 + tid] = 0; }   // [32 + l]: work-set sites of a row-tile layer

    const int e0 = p.offsets[s];
    int n = (p.ends ? p.ends[s] : p.offsets[s + 1]) - e0;
    uint32_t *front = p.front + (long long)s * nbm;
    if (n > p.max_events) {
        if (tid == 0) atomicOr(p.err_flag, 2);
        n = 0;
    }
    if (n <= 0) {   // stream untouched by this step
        if (tid == 0) { p.active[s] = 0; p.delta[s] = 0.0; }
        for (int i = tid; i < nbm; i += kThreads) front[i] = 0u;
        return;
    }
    const int32_t *ev = p.events + (long long)e0 * 3;

    for (int i = tid; i < p.hash_slots; i += kThreads) { hkey[i] = -1; hval[i] = -1; }
    uint32_t *alive = p.alive + (long long)s * nbm;
    for (int i = tid; i < nbm; i += kThreads) { bm[i] = 0u; al[i] = alive[i]; }

    // t_L = max(ts)
    int tmax = INT_MIN;
    for (int j = tid; j < n; j += kThreads) tmax = max(tmax, ev[3 * j + 2]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, d));
    if ((tid & 31) == 0) s_red[tid >> 5] = tmax;
    __syncthreads();
    if (tid == 0) {
        int m = s_red[0];
        for (int i = 1; i < kThreads / 32; ++i) m = max(m, s_red[i]);
        s_tlast = m;
    }
    __syncthreads();
    const int t_last = s_tlast;
    const int dt = (int)((unsigned)t_last - (unsigned)p.prev_ts[s]);   // int32 arithmetic like numpy
    const double delta = __dmul_rn((double)dt, p.leak);

    // last-duplicate-wins: hash pixel -> largest event index
    const unsigned mask = (unsigned)p.hash_slots - 1u;
    for (int j = tid; j < n; j += kThreads) {
        const int y = ev[3 * j], x = ev[3 * j + 1];
        if (y < 0 || y >= p.H || x < 0 || x >= p.W) { atomicOr(p.err_flag, 1); continue; }
        const int pix = y * p.W + x;
        unsigned h = ((unsigned)pix * 2654435761u) & mask;
        while (true) {
            const int k = atomicCAS(&hkey[h], -1, pix);
            if (k == -1 || k == pix) { atomicMax(&hval[h], j); break; }
            h = (h + 1) & mask;
        }
    }

    // leak + clamp (integration.py:63-68).  A pixel with S == 0 stays 0 and emits nothing, so only the pixels
    // alive after the previous step are touched: `al` starts as the previous alive bitmap and is walked one
    // word per warp, lane b owning bit b (coalesced over the set bits; the surface is mostly dead, so this
    // reads a fraction of it), kU words in flight per warp.
    double *surf = p.surface + (long long)s * HW;
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kThreads / 32, kU = 12;        // words in flight per warp: the walk is a chain of load latencies
    for (int w0 = wid; w0 < nbm; w0 += kWarps * kU) {
        uint32_t bits[kU];
        double v[kU];
        int pix[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int wi = w0 + u * kWarps;
            bits[u] = wi < nbm ? al[wi] : 0u;
            const int y = wi / p.Ww;
            pix[u] = y * p.W + (wi - y * p.Ww) * 32 + lane;
            v[u] = 0.0;
            if ((bits[u] >> lane) & 1u) v[u] = surf[pix[u]];
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (bits[u] == 0u) continue;                       // warp-uniform
            const bool mine = (bits[u] >> lane) & 1u;
            const double w = __dsub_rn(v[u], delta);
            const bool dead = mine && w <= 0.0;
            if (mine) surf[pix[u]] = dead ? 0.0 : w;
            const uint32_t died = __ballot_sync(0xffffffffu, dead);
            if (lane == 0) {
                const int wi = w0 + u * kWarps;
                al[wi] = bits[u] & ~died;                       // still alive
                bm[wi] = died;                                  // died now: an output event
            }
        }
    }
    __syncthreads();

    // event increments, last occurrence per pixel only (integration.py:71-74,80)
    for (int j = tid; j < n; j += kThreads) {
        const int y = ev[3 * j], x = ev[3 * j + 1];
        if (y < 0 || y >= p.H || x < 0 || x >= p.W) continue;
        const int pix = y * p.W + x;
        unsigned h = ((unsigned)pix * 2654435761u) & mask;
        while (hkey[h] != pix) h = (h + 1) & mask;
        if (hval[h] != j) continue;
        const int age = (int)((unsigned)t_last - (unsigned)ev[3 * j + 2]);
        const double inc = __dsub_rn(1.0, __dmul_rn((double)age, p.leak));
        double v = __dadd_rn(surf[pix], inc);
        if (v <= 0.0) v = 0.0;
        surf[pix] = v;
        atomicOr(&bm[y * p.Ww + (x >> 5)], 1u << (x & 31));
        if (v > 0.0) atomicOr(&al[y * p.Ww + (x >> 5)], 1u << (x & 31));
        else atomicAnd(&al[y * p.Ww + (x >> 5)], ~(1u << (x & 31)));
    }
    __syncthreads();
    for (int i = tid; i < nbm; i += kThreads) { front[i] = bm[i]; alive[i] = al[i]; }
    if (tid == 0) { p.active[s] = 1; p.delta[s] = delta; p.prev_ts[s] = t_last; }
}

// ---------------------------------------------------------------------------------------------
// K2: leak sweep of all conv layers.   conv2d.py:113-115,126-128
//   F <- (float)((double)F - (double)A * delta)   [NEP-50 arithmetic of `f32 -= f32 * np.float64`]
//   sites where sign(F >= 0) flipped in any channel are recorded in signchg.
// Elements with A == 0 keep F bit-for-bit (x - 0 == x), so F is neither read nor written there, and a
// site whose non-zero-rate bit (nzr) is clear has A == 0 in every channel, so it is not touched at all:
// each CTA compacts the live sites of its chunk of bitmap words into shared memory and then streams
// only their (A, F) vectors, several independent 16-byte loads in flight per thread.
// The table also holds the pool layers' (Fp, Ap) copies (signchg == nullptr): the same arithmetic on
// the same bits keeps each copy equal to the conv element it mirrors.
// Layers whose channel count is not a multiple of 4 (a float4 would straddle sites) take the dense
// path: chunks of kSweepChunk float4, A read everywhere.
// Grid: (chunks per stream over all table entries, S).
// ---------------------------------------------------------------------------------------------
struct SweepLayer {
    float *F, *A;
    uint32_t *signchg;
    const uint32_t *nzr;  // [S][H*Ww] sites that can have a non-zero rate (nullptr: dense path)
    const uint32_t *skip; // [S][H*Ww] sites this step re-evaluates anyway (k_frontier_skip); nullptr: leak every live site
    long long fstride;   // floats per stream (multiple of 4)
    int n4;              // float4 per stream
    int chunk0;          // first chunk index of this layer
    int C, W, Ww, HWw;   // HWw = H*Ww
    int wpc;             // sparse path: bitmap words per chunk (<= kSweepMaxWords)
    int c4, c4_shift;    // C/4 and log2(C/4) (or -1 when C/4 is not a power of two)
};
// The first conv layer's map is swept window by window when a 2x2 pool follows it (k_leak_sweep, fused path): a window that the
// pool layer re-evaluates in this step only because its recompute flag is set (maxpool.py:123-126) and none of whose four conv
// sites is re-evaluated reads exactly the values the sweep has in registers - the leaked F and A of its sites - so its argmax
// is taken here instead of being read back by k_pool_eval.  That holds for the FIRST conv layer only: its work set (dilation
// of the surface's events) is final before the sweep; a deeper layer's may still grow by the sign flips the sweep finds.
struct SweepPool {
    uint8_t *idx;              // [S][pstride] pool argmax rows
    float *Fp, *Ap;            // [S][pstride] copies at the argmax
    long long pstride;
    const uint32_t *flags;     // [S][pHWw] sticky recompute flags as the previous step left them (the frontier kernel updates them later)
    uint32_t *uns;             // [S][pHWw] out: windows evaluated here that came out unstable (consumed and cleared by k_frontier_all)
    unsigned long long *accum; // windows evaluated here, added to the pool layer's work counter
    int pW, pWw, pHWw;         // pool map width, bitmap words per row, words per stream
    int wpc;                   // window-bitmap words per chunk
    float alpha;
};
struct SweepParams {
    SweepLayer L[kMaxSweep];
    int n_layers;
    const double *delta;
    const uint8_t *active;
};
struct SweepWindowsParams {      // k_sweep_windows: one conv map swept by pool windows
    SweepLayer L;
    SweepPool Q;
    const double *delta;
    const uint8_t *active;
};
constexpr int kSweepVec = 4;   // float4 per thread per iteration
constexpr int kSweepChunk = kThreads * kSweepVec;
constexpr int kSweepMaxWords = 64;                     // bitmap words per sparse chunk
constexpr int kSweepUnitsPerChunk = 8192;              // target float4 per sparse chunk at full density

__device__ __forceinline__ float leak1(float f, float a, double delta)
{
    return __double2float_rn(__dsub_rn((double)f, __dmul_rn((double)a, delta)));
}

// Applies the leak to one float4 of F given its rates; returns the flipped-sign lanes (bit e = channel e).
__device__ __forceinline__ unsigned leak4(float4 *Fp, const float4 av, double delta)
{
    const float4 f = *Fp;
    float4 g;
    g.x = leak1(f.x, av.x, delta);
    g.y = leak1(f.y, av.y, delta);
    g.z = leak1(f.z, av.z, delta);
    g.w = leak1(f.w, av.w, delta);
    *Fp = g;
    return ((f.x >= 0.f) != (g.x >= 0.f) ? 1u : 0u) | ((f.y >= 0.f) != (g.y >= 0.f) ? 2u : 0u) |
           ((f.z >= 0.f) != (g.z >= 0.f) ? 4u : 0u) | ((f.w >= 0.f) != (g.w >= 0.f) ? 8u : 0u);
}

// Same with F already in registers.
__device__ __forceinline__ unsigned leak4v(float4 *Fp, const float4 f, const float4 av, double delta)
{
    float4 g;
    g.x = leak1(f.x, av.x, delta);
    g.y = leak1(f.y, av.y, delta);
    g.z = leak1(f.z, av.z, delta);
    g.w = leak1(f.w, av.w, delta);
    *Fp = g;
    return ((f.x >= 0.f) != (g.x >= 0.f) ? 1u : 0u) | ((f.y >= 0.f) != (g.y >= 0.f) ? 2u : 0u) |
           ((f.z >= 0.f) != (g.z >= 0.f) ? 4u : 0u) | ((f.w >= 0.f) != (g.w >= 0.f) ? 8u : 0u);
}

struct PoolBest {
    float f, r, a, rlow;
    int row;
};
__device__ __forceinline__ void pool_first(PoolBest &b, float f, float a, float alpha)
{
    b.f = f; b.a = a; b.r = __fmul_rn(a, slope_of(f, alpha)); b.rlow = b.r; b.row = 0;
}
__device__ __forceinline__ void pool_next(PoolBest &b, int row, float f, float a, float alpha)
{
    const float r = __fmul_rn(a, slope_of(f, alpha));
    if (f > b.f || (f == b.f && r < b.r)) { b.row = row; b.f = f; b.r = r; b.a = a; }   // cutils.pyx:166-170
    if (r < b.rlow) b.rlow = r;                                                          // cutils.pyx:173-174
}


// Fused path of the leak sweep (see SweepPool): one chunk = `wpc` words of the POOL layer's window bitmap.
//   need  = windows holding a site that must be leaked (live and not re-evaluated)  U  sticky windows evaluated here
//   item  = (window, 4 channels): the four sites' (A, F) vectors are loaded together (only where needed), live sites are leaked
//           exactly like the plain sweep (same arithmetic, sign flips into signchg), and a sticky window none of whose sites
//           is re-evaluated gets its argmax / (Fp, Ap) copy / unstable bit from those registers (cutils.pyx:161-177).
// Its own kernel (grid: chunks x S) so that the plain sweep keeps its 48 registers per thread.
__global__ void __launch_bounds__(kThreads, 4) k_sweep_windows(const __grid_constant__ SweepWindowsParams p)
{
    pdl_enter();
    __shared__ int s_win[kSweepMaxWords * 32];
    __shared__ int s_scan[9];
    const int s = blockIdx.y;
    if (!p.active[s]) return;
    const double delta = p.delta[s];
    const SweepLayer &L = p.L;
    const SweepPool &Q = p.Q;
    const int chunk_local = blockIdx.x;
    const float4 *A4 = reinterpret_cast<const float4 *>(L.A + (long long)s * L.fstride);
    float4 *F4 = reinterpret_cast<float4 *>(L.F + (long long)s * L.fstride);
    uint32_t *sc = L.signchg ? L.signchg + (long long)s * L.HWw : nullptr;
    __shared__ uint32_t s_cw[kSweepMaxWords][4];      // per window word: live bits of its conv words (row 0: lo, hi; row 1: lo, hi)
    __shared__ uint32_t s_fz[kSweepMaxWords];         // per window word: windows evaluated here
    __shared__ int s_cnt;
    const int w0 = chunk_local * Q.wpc, w1 = min(Q.pHWw, w0 + Q.wpc);
    const uint32_t *nz = L.nzr + (long long)s * L.HWw;
    const uint32_t *sk = L.skip + (long long)s * L.HWw;
    const uint32_t *fl = Q.flags + (long long)s * Q.pHWw;
    if (threadIdx.x == 0) s_cnt = 0;
    uint32_t need = 0u;
    if ((int)threadIdx.x < w1 - w0) {
        const int w = w0 + threadIdx.x;
        const int py = w / Q.pWw, pw = w - py * Q.pWw;
        uint32_t live[4], nset[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cw = 2 * pw + (k & 1);
            live[k] = 0u;
            if (cw < L.Ww) {
                const int i = (2 * py + (k >> 1)) * L.Ww + cw;
                const uint32_t sv = __ldg(sk + i);
                live[k] = __ldg(nz + i) & ~sv;
                nset[k & 1] |= sv;
            }
            s_cw[threadIdx.x][k] = live[k];
        }
        auto pooled = [](uint32_t lo, uint32_t hi) {
            lo |= lo >> 1;
            hi |= hi >> 1;
            return compress_even_bits(lo) | (compress_even_bits(hi) << 16);
        };
        const uint32_t lastmask = (Q.pW & 31) && pw == Q.pWw - 1 ? ((1u << (Q.pW & 31)) - 1u) : 0xffffffffu;
        const uint32_t fz = __ldg(fl + w) & ~pooled(nset[0], nset[1]) & lastmask;
        s_fz[threadIdx.x] = fz;
        need = (pooled(live[0] | live[2], live[1] | live[3]) | fz) & lastmask;
    }
    if (!__syncthreads_or(need != 0u)) return;
    int total;
    int off = block_excl_scan(__popc(need), s_scan, &total);
    {
        const int wl = threadIdx.x;                   // window word (local)
        while (need) {
            const int b = __ffs(need) - 1;
            need &= need - 1;
            s_win[off++] = (wl << 5) | b;
        }
    }
    __syncthreads();
    const int c4 = L.c4;
    const int items = total * c4;
    const int cW = L.W;
    int fused_here = 0;
    for (int u = threadIdx.x; u < items; u += kThreads) {
        const int wi = L.c4_shift >= 0 ? (u >> L.c4_shift) : (u / c4);
        const int cg = u - wi * c4;
        const int code = s_win[wi], wl = code >> 5, b = code & 31;
        const int w = w0 + wl;
        const int py = w / Q.pWw, px = (w - py * Q.pWw) * 32 + b;
        const bool fused = (s_fz[wl] >> b) & 1u;
        float4 a[4], f[4];
        bool live[4];
        int idx4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                  // k = window row: (dy, dx) = (k >> 1, k & 1)
            const int cbit = 2 * b + (k & 1);          // conv column within the window word's 64 columns
            live[k] = (s_cw[wl][(k >> 1) * 2 + (cbit >> 5)] >> (cbit & 31)) & 1u;
            idx4[k] = ((2 * py + (k >> 1)) * cW + 2 * px + (k & 1)) * c4 + cg;
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            f[k] = a[k];
            if (live[k] || fused) {
                a[k] = A4[idx4[k]];
                f[k] = F4[idx4[k]];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!live[k]) continue;
            const float4 av = a[k];
            if (av.x == 0.f && av.y == 0.f && av.z == 0.f && av.w == 0.f) continue;
            const float4 old = f[k];
            float4 g;
            g.x = leak1(old.x, av.x, delta);
            g.y = leak1(old.y, av.y, delta);
            g.z = leak1(old.z, av.z, delta);
            g.w = leak1(old.w, av.w, delta);
            F4[idx4[k]] = g;
            f[k] = g;
            const bool flip = ((old.x >= 0.f) != (g.x >= 0.f)) || ((old.y >= 0.f) != (g.y >= 0.f)) || ((old.z >= 0.f) != (g.z >= 0.f)) ||
                              ((old.w >= 0.f) != (g.w >= 0.f));
            if (flip && sc) {
                const int y = 2 * py + (k >> 1), x = 2 * px + (k & 1);
                atomicOr(&sc[y * L.Ww + (x >> 5)], 1u << (x & 31));
            }
        }
        if (fused) {
            const float fr[4][4] = {{f[0].x, f[0].y, f[0].z, f[0].w}, {f[1].x, f[1].y, f[1].z, f[1].w}, {f[2].x, f[2].y, f[2].z, f[2].w}, {f[3].x, f[3].y, f[3].z, f[3].w}};
            const float ar[4][4] = {{a[0].x, a[0].y, a[0].z, a[0].w}, {a[1].x, a[1].y, a[1].z, a[1].w}, {a[2].x, a[2].y, a[2].z, a[2].w}, {a[3].x, a[3].y, a[3].z, a[3].w}};
            PoolBest best[4];
            bool unstable = false;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                pool_first(best[v], fr[0][v], ar[0][v], Q.alpha);
#pragma unroll
                for (int row = 1; row < 4; ++row) pool_next(best[v], row, fr[row][v], ar[row][v], Q.alpha);
                unstable |= best[v].r != best[v].rlow;
            }
            const long long o = (long long)s * Q.pstride + ((long long)py * Q.pW + px) * L.C + 4 * cg;
            *reinterpret_cast<uchar4 *>(Q.idx + o) = make_uchar4((unsigned char)best[0].row, (unsigned char)best[1].row,
                                                                  (unsigned char)best[2].row, (unsigned char)best[3].row);
            *reinterpret_cast<float4 *>(Q.Fp + o) = make_float4(best[0].f, best[1].f, best[2].f, best[3].f);
            *reinterpret_cast<float4 *>(Q.Ap + o) = make_float4(best[0].a, best[1].a, best[2].a, best[3].a);
            if (unstable) atomicOr(&Q.uns[(long long)s * Q.pHWw + w], 1u << b);
            if (cg == 0) ++fused_here;
        }
    }
    if (fused_here) atomicAdd(&s_cnt, fused_here);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) {
        atomicAdd(Q.accum, (unsigned long long)s_cnt);
        atomicAdd(Q.accum + 32, (unsigned long long)s_cnt);   // [32 + l]: the share of the pool layer's windows evaluated here
    }
}

__global__ void __launch_bounds__(kThreads) k_leak_sweep(const __grid_constant__ SweepParams p)
{
    pdl_enter();
    __shared__ int s_live[kSweepMaxWords * 32];
    __shared__ int s_scan[9];
    const int s = blockIdx.y;
    if (!p.active[s]) return;
    const double delta = p.delta[s];
    if (delta == 0.0) return;
    int li = 0;
    const int chunk = blockIdx.x;
#pragma unroll 1
    for (int i = 1; i < p.n_layers; ++i)
        if (chunk >= p.L[i].chunk0) li = i;
    const SweepLayer &L = p.L[li];
    const float4 *A4 = reinterpret_cast<const float4 *>(L.A + (long long)s * L.fstride);
    float4 *F4 = reinterpret_cast<float4 *>(L.F + (long long)s * L.fstride);
    uint32_t *sc = L.signchg ? L.signchg + (long long)s * L.HWw : nullptr;

    if (L.nzr) {
        // ---- sparse path: compact the live sites of words [w0, w1) ...
        const int w0 = (chunk - L.chunk0) * L.wpc, w1 = min(L.HWw, w0 + L.wpc);
        const uint32_t *nz = L.nzr + (long long)s * L.HWw;
        uint32_t bits = 0u;
        if ((int)threadIdx.x < w1 - w0) {
            bits = __ldg(nz + w0 + threadIdx.x);
            // a site that this step re-evaluates gets F and A overwritten and is on the layer's frontier whatever its
            // signs do: leaking it first would be wasted traffic (at steady state that is every live site of conv2..conv7)
            if (L.skip) bits &= ~__ldg(L.skip + (long long)s * L.HWw + w0 + threadIdx.x);
        }
        if (!__syncthreads_or(bits != 0u)) return;       // nothing to leak in this chunk (the common case for skipped layers)
        int total;
        int off = block_excl_scan(__popc(bits), s_scan, &total);
        if (bits) {
            const int w = w0 + threadIdx.x;
            const int y = w / L.Ww, xb = (w - y * L.Ww) * 32;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                s_live[off++] = y * L.W + xb + b;
            }
        }
        __syncthreads();
        // ---- ... and stream their channel vectors: unit u = (live site u / c4, float4 u % c4)
        const int units = total * L.c4;
        for (int u0 = threadIdx.x; u0 < units; u0 += kSweepChunk) {
            float4 a[kSweepVec], f[kSweepVec];
            int idx[kSweepVec];
            // F is loaded together with A: at a live site practically every 16-byte group has a non-zero rate
            // (measured: the two fractions coincide), so waiting for A first only adds a DRAM round trip
#pragma unroll
            for (int j = 0; j < kSweepVec; ++j) {
                const int u = u0 + j * kThreads;
                idx[j] = -1;
                a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                f[j] = a[j];
                if (u < units) {
                    const int ls = L.c4_shift >= 0 ? (u >> L.c4_shift) : (u / L.c4);
                    idx[j] = s_live[ls] * L.c4 + (u - ls * L.c4);
                    a[j] = A4[idx[j]];
                    f[j] = F4[idx[j]];
                }
            }
#pragma unroll
            for (int j = 0; j < kSweepVec; ++j) {
                const float4 av = a[j];
                if (av.x == 0.f && av.y == 0.f && av.z == 0.f && av.w == 0.f) continue;
                const unsigned flips = leak4v(F4 + idx[j], f[j], av, delta);
                if (flips && sc) {
                    const int site = L.c4_shift >= 0 ? (idx[j] >> L.c4_shift) : (idx[j] / L.c4);
                    const int y = site / L.W, x = site - y * L.W;
                    atomicOr(&sc[y * L.Ww + (x >> 5)], 1u << (x & 31));
                }
            }
        }
        return;
    }

    // ---- dense path
    const int base = (chunk - L.chunk0) * kSweepChunk;
    float4 a[kSweepVec];
    int idx[kSweepVec];
#pragma unroll
    for (int j = 0; j < kSweepVec; ++j) {
        idx[j] = base + j * kThreads + threadIdx.x;
        a[j] = idx[j] < L.n4 ? A4[idx[j]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kSweepVec; ++j) {
        const float4 av = a[j];
        if (av.x == 0.f && av.y == 0.f && av.z == 0.f && av.w == 0.f) continue;
        const unsigned flips = leak4(F4 + idx[j], av, delta);
        if (flips && sc) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (flips & (1u << e)) {
                    const int site = (idx[j] * 4 + e) / L.C;
                    const int y = site / L.W, x = site - y * L.W;
                    atomicOr(&sc[y * L.Ww + (x >> 5)], 1u << (x & 31));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: conv frontier.  One CTA per stream.   conv2d.py:118-131, cutils.pyx:73-112
//   N = output sites whose receptive field holds an input event: o in [p+pad-k+1, p+pad] clipped
//   (stride 1), i.e. a dilation of the previous layer's frontier bitmap;
//   E = N U signchg  (output events);  N is appended to the cross-stream work list.
// Dynamic shared memory: (Hin*WwIn + Hin*Ww + H*Ww) * 4 bytes.
// ---------------------------------------------------------------------------------------------
struct ConvFrontParams {
    const uint32_t *prev_front;   // [S][Hin*WwIn]
    uint32_t *front;              // [S][H*Ww]
    uint32_t *signchg;            // [S][H*Ww]  (consumed: cleared)
    uint32_t *nzr;                // [S][H*Ww]  this layer's non-zero-rate bits
    const uint32_t *prev_nzr;     // [S][Hin*WwIn] previous layer's
    const uint8_t *active;
    uint32_t *sites;
    int *counter;
    int Hin, Win, WwIn;
    int H, W, Ww;
    int kh, kw, pad_t, pad_l;
    SiteCode code;
};

__global__ void __launch_bounds__(kThreads) k_conv_frontier(ConvFrontParams p)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *P = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *Hd = P + p.Hin * p.WwIn;
    uint32_t *N = Hd + p.Hin * p.Ww;
    __shared__ int scratch[9];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int nout = p.H * p.Ww;
    uint32_t *front = p.front + (long long)s * nout;
    if (!p.active[s]) {
        for (int i = tid; i < nout; i += kThreads) front[i] = 0u;
        return;
    }
    const uint32_t *pf = p.prev_front + (long long)s * p.Hin * p.WwIn;
    for (int i = tid; i < p.Hin * p.WwIn; i += kThreads) P[i] = pf[i];
    __syncthreads();
    const uint32_t lastmask = (p.W & 31) ? ((1u << (p.W & 31)) - 1u) : 0xffffffffu;
    // horizontal: out x = in x + d, d in [pad_l-kw+1, pad_l]
    auto hdilate = [&]() {
        for (int i = tid; i < p.Hin * p.Ww; i += kThreads) {
            const int y = i / p.Ww, w = i - y * p.Ww;
            uint32_t acc = 0u;
            for (int d = p.pad_l - p.kw + 1; d <= p.pad_l; ++d) acc |= row_shift(P + y * p.WwIn, p.WwIn, w, d);
            if (w == p.Ww - 1) acc &= lastmask;
            Hd[i] = acc;
        }
    };
    // vertical: out y = in y + d, d in [pad_t-kh+1, pad_t]
    auto vdilate = [&](int i) {
        const int y = i / p.Ww, w = i - y * p.Ww;
        uint32_t acc = 0u;
        for (int d = p.pad_t - p.kh + 1; d <= p.pad_t; ++d) {
            const int yi = y - d;
            if (yi >= 0 && yi < p.Hin) acc |= Hd[yi * p.Ww + w];
        }
        return acc;
    };
    hdilate();
    __syncthreads();
    uint32_t *sc = p.signchg + (long long)s * nout;
    for (int i = tid; i < nout; i += kThreads) {
        const uint32_t acc = vdilate(i);
        N[i] = acc;
        const uint32_t flips = sc[i];
        if (flips) sc[i] = 0u;
        front[i] = acc | flips;
    }
    __syncthreads();
    // non-zero-rate bits: re-evaluated sites take the OR of their receptive field's input bits
    const uint32_t *pn = p.prev_nzr + (long long)s * p.Hin * p.WwIn;
    for (int i = tid; i < p.Hin * p.WwIn; i += kThreads) P[i] = pn[i];
    __syncthreads();
    hdilate();
    __syncthreads();
    uint32_t *nz = p.nzr + (long long)s * nout;
    for (int i = tid; i < nout; i += kThreads) {
        const uint32_t n = N[i];
        if (n) nz[i] = (nz[i] & ~n) | (n & vdilate(i));
    }
    emit_sites(N, p.H, p.Ww, (uint32_t)s << p.code.sh_s, p.code.sh_y, p.sites, p.counter, scratch);
}

// ---------------------------------------------------------------------------------------------
// K4: pool frontier.  One CTA per stream.   maxpool.py:116-126,153-154
//   hit = windows containing an input event; flags[hit] = False; W = hit U flags (sticky);
//   output events = W (all evaluated windows); W is appended to the work list.
// Dynamic shared memory: (Hin*WwIn + H*Ww) * 4 bytes.
// ---------------------------------------------------------------------------------------------
struct PoolFrontParams {
    const uint32_t *prev_front;   // [S][Hin*WwIn]
    uint32_t *front;              // [S][H*Ww]
    uint32_t *flags;              // [S][H*Ww]
    uint32_t *nzr;                // [S][H*Ww]  non-zero-rate bits of the (Fp, Ap) copy
    const uint32_t *prev_nzr;     // [S][Hin*WwIn] the conv layer's
    const uint8_t *active;
    uint32_t *sites;
    int *counter;
    int Hin, Win, WwIn;
    int H, W, Ww;
    int kh, kw, stride;
    SiteCode code;
};

__global__ void __launch_bounds__(kThreads) k_pool_frontier(PoolFrontParams p)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *P = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *Wb = P + p.Hin * p.WwIn;
    __shared__ int scratch[9];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int nout = p.H * p.Ww;
    uint32_t *front = p.front + (long long)s * nout;
    if (!p.active[s]) {
        for (int i = tid; i < nout; i += kThreads) front[i] = 0u;
        return;
    }
    const uint32_t *pf = p.prev_front + (long long)s * p.Hin * p.WwIn;
    for (int i = tid; i < p.Hin * p.WwIn; i += kThreads) P[i] = pf[i];
    __syncthreads();
    uint32_t *fl = p.flags + (long long)s * nout;
    const uint32_t lastmask = (p.W & 31) ? ((1u << (p.W & 31)) - 1u) : 0xffffffffu;
    // word i of the [H][Ww] bitmap whose bit = OR of the input bitmap P over the window
    auto window_or = [&](int i) {
        const int oy = i / p.Ww, w = i - oy * p.Ww;
        uint32_t hit = 0u;
        if (p.kh == 2 && p.kw == 2 && p.stride == 2) {
            const uint32_t *r0 = P + (2 * oy) * p.WwIn, *r1 = r0 + p.WwIn;
            uint32_t lo = (2 * w < p.WwIn) ? (r0[2 * w] | r1[2 * w]) : 0u;
            uint32_t hi = (2 * w + 1 < p.WwIn) ? (r0[2 * w + 1] | r1[2 * w + 1]) : 0u;
            lo |= lo >> 1;
            hi |= hi >> 1;
            hit = compress_even_bits(lo) | (compress_even_bits(hi) << 16);
        } else {
            for (int b = 0; b < 32; ++b) {
                const int ox = w * 32 + b;
                if (ox >= p.W) break;
                bool any = false;
                for (int dy = 0; dy < p.kh && !any; ++dy)
                    for (int dx = 0; dx < p.kw; ++dx) {
                        const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
                        if ((P[iy * p.WwIn + (ix >> 5)] >> (ix & 31)) & 1u) { any = true; break; }
                    }
                if (any) hit |= 1u << b;
            }
        }
        if (w == p.Ww - 1) hit &= lastmask;
        return hit;
    };
    for (int i = tid; i < nout; i += kThreads) {
        const uint32_t hit = window_or(i);
        const uint32_t f = fl[i] & ~hit;      // maxpool.py:118-120
        const uint32_t wset = hit | f;        // maxpool.py:123-126
        fl[i] = f;
        Wb[i] = wset;
        front[i] = wset;
    }
    __syncthreads();
    // non-zero-rate bits of the (Fp, Ap) copy: a re-evaluated window copies one of its conv sites
    const uint32_t *pn = p.prev_nzr + (long long)s * p.Hin * p.WwIn;
    for (int i = tid; i < p.Hin * p.WwIn; i += kThreads) P[i] = pn[i];
    __syncthreads();
    uint32_t *nz = p.nzr + (long long)s * nout;
    for (int i = tid; i < nout; i += kThreads) {
        const uint32_t wset = Wb[i];
        if (wset) nz[i] = (nz[i] & ~wset) | (wset & window_or(i));
    }
    emit_sites(Wb, p.H, p.Ww, (uint32_t)s << p.code.sh_s, p.code.sh_y, p.sites, p.counter, scratch);
}

// ---------------------------------------------------------------------------------------------
// K3+K4 fused: the frontier of EVERY layer in one launch, one CTA per stream.  Frontiers, sticky flags and
// non-zero-rate bits are pure bitmap logic on the previous layer's bitmaps (the sign flips come from the leak
// sweep, the flags from earlier steps), so the whole chain can run before any evaluation: the previous
// layer's two bitmaps stay in shared memory from layer to layer, and a step issues one frontier launch
// instead of one per layer.  Each layer has its own work list (the evaluations run afterwards, back to back).
// Same arithmetic as k_conv_frontier / k_pool_frontier (kept for the layer-at-a-time interface).
// Dynamic shared memory: 5 * max_words * 4 bytes (max_words = largest H*Ww of any layer, incl. H_in*Ww mixes).
// ---------------------------------------------------------------------------------------------
struct FrontLayer {
    int type;                     // 1 conv, 2 pool (AEC_LAYER_*)
    int Hin, Win, WwIn, H, W, Ww;
    int kh, kw, pad_t, pad_l, stride;
    SiteCode code;
    int rt_rows, rt_seg, rt_nxg;  // row-tile conv layer (aec_rt.cuh): rows per unit (0 = the layer takes a site list), sites per x segment, segments per row
    uint32_t *nset;               // row-tile layer: [S][H*Ww] copy of the exact work set (the evaluation stores only there)
    uint32_t *swp_uns;            // pool layer whose sticky windows the leak sweep evaluates (SweepPool): their unstable bits, else null
    const uint32_t *swp_skip;     // ... and the conv layer's skip bitmap (= its exact work set: first conv layer only)
    int *counter2;                // row-tile layer / conv layer with a fused pool: number of work-set sites (statistics)
    uint32_t quad_bit;            // conv layer whose pool is evaluated in its epilogue (emit_sites_quads): the entry flag; else 0
    int pool_in_conv;             // pool layer: 1 = the windows all four sites of which the conv re-evaluates are evaluated by the conv kernel
    int pWw;                      // conv layer with quad_bit: bitmap words per row of the pool layer behind it
    uint32_t *front, *signchg, *flags, *nzr;
    uint32_t *skip;               // [S][H*Ww] written by k_frontier_skip: a subset of this step's work set, known before the leak sweep
    uint32_t *sites;
    int *counter;
};
struct FrontAllParams {
    const FrontLayer *layers;     // [n_layers], entry 0 unused (integration layer)
    int n_layers;
    const uint32_t *front0, *nzr0;   // layer 0 bitmaps written by k_integrate, [S][H0*Ww0]
    int words0;
    int max_words;
    const uint8_t *active;
};

// The layer table is copied to shared memory once per CTA: read from global memory layer by layer it cost one memory
// round trip per layer on the CTA's critical path (the chain is latency-bound: one CTA walks every layer of its stream).
constexpr int kMaxFrontLayers = 32;
__device__ __forceinline__ void stage_front_table(FrontLayer *dst, const FrontLayer *src, int n_layers)
{
    const int words = n_layers * (int)(sizeof(FrontLayer) / 4);
    for (int i = threadIdx.x; i < words; i += kThreads) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
}

// kMinBlocks = 8 caps the kernel at 32 registers (a few spills): eight CTAs per SM instead of six, i.e. ONE wave for the 1024 streams of the
// benchmark instead of 888 + 136 (0.133 -> 0.096 ms); with few streams the spills only cost (one stream 56 -> 60 us): the host picks.
// kPrefetch (few streams: the chain's LATENCY is the cost, one CTA per stream and SMs to spare): every layer reads two bitmaps
// from global memory that earlier KERNELS wrote (conv: the sweep's sign flips and the rate bits; pool: the sticky flags and the rate
// bits) - a dependent round trip of ~1.3 k cycles per layer, and one per loop iteration where stores separate the loads.  They
// are fetched one layer AHEAD into registers (kFrontPf words of each per thread; longer bitmaps fall back to direct loads).
constexpr int kFrontPf = 5;
template <int kMinBlocks, bool kPrefetch = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_frontier_all(FrontAllParams p)
{
    pdl_enter();
    __shared__ __align__(16) FrontLayer s_layers[kMaxFrontLayers];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *bufA = reinterpret_cast<uint32_t *>(smem_raw);        // previous layer's frontier
    uint32_t *bufB = bufA + p.max_words;                            // previous layer's non-zero-rate bits
    uint32_t *Hd = bufB + p.max_words;                              // scratch (horizontal dilation)
    uint32_t *N = Hd + p.max_words;                                 // this layer's work set
    uint32_t *Z = N + p.max_words;                                  // this layer's non-zero-rate bits
    __shared__ int scratch[9];
    const int s = blockIdx.x, tid = threadIdx.x;
    stage_front_table(s_layers, p.layers, p.n_layers);
    __syncthreads();
    if (!p.active[s]) {
        for (int li = 1; li < p.n_layers; ++li) {
            const FrontLayer &L = s_layers[li];
            uint32_t *front = L.front + (long long)s * L.H * L.Ww;
            for (int i = tid; i < L.H * L.Ww; i += kThreads) front[i] = 0u;
        }
        return;
    }
    // next layer's {sign flips | sticky flags} and rate bits of this thread's words i = tid + k * kThreads
    uint32_t pf_a[kPrefetch ? kFrontPf : 1], pf_b[kPrefetch ? kFrontPf : 1];
    auto prefetch = [&](int lj) {
        if constexpr (kPrefetch) {
            if (lj >= p.n_layers) return;
            const FrontLayer &Ln = s_layers[lj];
            const int n = Ln.H * Ln.Ww;
            const uint32_t *ga = (Ln.type == 1 ? Ln.signchg : Ln.flags) + (long long)s * n;
            const uint32_t *gb = Ln.nzr + (long long)s * n;
#pragma unroll
            for (int k = 0; k < kFrontPf; ++k) {
                const int i = tid + k * kThreads;
                pf_a[k] = i < n ? ga[i] : 0u;
                pf_b[k] = i < n ? gb[i] : 0u;
            }
        }
    };
    prefetch(1);
    for (int i = tid; i < p.words0; i += kThreads) {
        bufA[i] = p.front0[(long long)s * p.words0 + i];
        bufB[i] = p.nzr0[(long long)s * p.words0 + i];
    }
    __syncthreads();
    for (int li = 1; li < p.n_layers; ++li) {
        const FrontLayer &L = s_layers[li];
        const int nout = L.H * L.Ww;
        uint32_t cur_a[kPrefetch ? kFrontPf : 1], cur_b[kPrefetch ? kFrontPf : 1];
        if constexpr (kPrefetch) {
#pragma unroll
            for (int k = 0; k < kFrontPf; ++k) { cur_a[k] = pf_a[k]; cur_b[k] = pf_b[k]; }
            prefetch(li + 1);
        }
        // runs body(i, a, b) for this thread's words: a = the layer's sign-flip / flag word i, b = its rate-bit word i
        auto for_words = [&](const uint32_t *ga, const uint32_t *gb, auto body) {
            if constexpr (kPrefetch) {
#pragma unroll
                for (int k = 0; k < kFrontPf; ++k) {
                    const int i = tid + k * kThreads;
                    if (i < nout) body(i, cur_a[k], cur_b[k]);
                }
                for (int i = tid + kFrontPf * kThreads; i < nout; i += kThreads) body(i, ga[i], gb[i]);
            } else {
                for (int i = tid; i < nout; i += kThreads) body(i, ga[i], gb[i]);
            }
        };
        const uint32_t lastmask = (L.W & 31) ? ((1u << (L.W & 31)) - 1u) : 0xffffffffu;
        uint32_t *front = L.front + (long long)s * nout;
        uint32_t *nz = L.nzr + (long long)s * nout;
        if (L.type == 1) {
            auto hdilate = [&](const uint32_t *P) {
                for (int i = tid; i < L.Hin * L.Ww; i += kThreads) {
                    const int y = i / L.Ww, w = i - y * L.Ww;
                    uint32_t acc = 0u;
                    for (int d = L.pad_l - L.kw + 1; d <= L.pad_l; ++d) acc |= row_shift(P + y * L.WwIn, L.WwIn, w, d);
                    if (w == L.Ww - 1) acc &= lastmask;
                    Hd[i] = acc;
                }
            };
            auto vdilate = [&](int i) {
                const int y = i / L.Ww, w = i - y * L.Ww;
                uint32_t acc = 0u;
                for (int d = L.pad_t - L.kh + 1; d <= L.pad_t; ++d) {
                    const int yi = y - d;
                    if (yi >= 0 && yi < L.Hin) acc |= Hd[yi * L.Ww + w];
                }
                return acc;
            };
            hdilate(bufA);
            __syncthreads();
            uint32_t *sc = L.signchg + (long long)s * nout;
            for (int i = tid; i < nout; i += kThreads) N[i] = vdilate(i);
            __syncthreads();
            hdilate(bufB);
            __syncthreads();
            for_words(sc, nz, [&](int i, uint32_t flips, uint32_t z) {
                const uint32_t n = N[i];
                if (flips) sc[i] = 0u;
                const uint32_t f = n | flips;
                front[i] = f;
                if (n) { z = (z & ~n) | (n & vdilate(i)); nz[i] = z; }
                Z[i] = z;
                bufA[i] = f;          // safe: bufA was last read by the first hdilate, two barriers ago
            });
            __syncthreads();
        } else {
            uint32_t *fl = L.flags + (long long)s * nout;
            auto window_or = [&](const uint32_t *P, int i) {
                const int oy = i / L.Ww, w = i - oy * L.Ww;
                uint32_t hit = 0u;
                if (L.kh == 2 && L.kw == 2 && L.stride == 2) {
                    const uint32_t *r0 = P + (2 * oy) * L.WwIn, *r1 = r0 + L.WwIn;
                    uint32_t lo = (2 * w < L.WwIn) ? (r0[2 * w] | r1[2 * w]) : 0u;
                    uint32_t hi = (2 * w + 1 < L.WwIn) ? (r0[2 * w + 1] | r1[2 * w + 1]) : 0u;
                    lo |= lo >> 1;
                    hi |= hi >> 1;
                    hit = compress_even_bits(lo) | (compress_even_bits(hi) << 16);
                } else {
                    for (int b = 0; b < 32; ++b) {
                        const int ox = w * 32 + b;
                        if (ox >= L.W) break;
                        bool any = false;
                        for (int dy = 0; dy < L.kh && !any; ++dy)
                            for (int dx = 0; dx < L.kw; ++dx) {
                                const int iy = oy * L.stride + dy, ix = ox * L.stride + dx;
                                if ((P[iy * L.WwIn + (ix >> 5)] >> (ix & 31)) & 1u) { any = true; break; }
                            }
                        if (any) hit |= 1u << b;
                    }
                }
                if (w == L.Ww - 1) hit &= lastmask;
                return hit;
            };
            const uint32_t *pskip = L.swp_uns ? L.swp_skip + (long long)s * L.Hin * L.WwIn : nullptr;
            uint32_t *uns = L.swp_uns ? L.swp_uns + (long long)s * nout : nullptr;
            for_words(fl, nz, [&](int i, uint32_t before, uint32_t z) {
                const uint32_t hit = window_or(bufA, i);
                uint32_t f = before & ~hit;           // maxpool.py:118-120
                const uint32_t wset = hit | f;        // maxpool.py:123-126
                uint32_t todo = wset;                 // windows k_pool_eval has to evaluate
                if (L.pool_in_conv) todo &= ~Hd[i];   // complete windows: the conv epilogue evaluates them (Hd = their bitmap, from the conv layer's pass)
                if (uns) {
                    // windows the leak sweep has evaluated already (sticky flag set, none of their conv sites re-evaluated): they
                    // stay output events of the layer but leave the work list; a window among them that was hit after all (by a
                    // sign flip the sweep found) had its flag cleared above and gets it back if it came out unstable
                    const uint32_t done = before & ~window_or(pskip, i);
                    const uint32_t u = uns[i];
                    if (u) uns[i] = 0u;
                    f |= done & hit & u;
                    todo &= ~done;
                }
                fl[i] = f;
                N[i] = todo;
                front[i] = wset;
                if (wset) { z = (z & ~wset) | (wset & window_or(bufB, i)); nz[i] = z; }
                Z[i] = z;
                Hd[i] = wset;                         // Hd is free in the pool branch: the layer's output events for the next layer
            });
            __syncthreads();
            emit_sites(N, L.H, L.Ww, (uint32_t)s << L.code.sh_s, L.code.sh_y, L.sites, L.counter, scratch);
            __syncthreads();
            for (int i = tid; i < nout; i += kThreads) { bufA[i] = Hd[i]; bufB[i] = Z[i]; }
            __syncthreads();
            continue;
        }
        if (L.type == 1 && L.quad_bit) {
            // the pool behind this layer is evaluated in the conv epilogue for complete windows: Hd (free now) = those windows
            const int pwords = (L.H >> 1) * L.pWw;
            for (int i = tid; i < pwords; i += kThreads) {
                const int oy = i / L.pWw, w = i - oy * L.pWw;
                const uint32_t *r0 = N + (2 * oy) * L.Ww, *r1 = r0 + L.Ww;
                uint32_t lo = (2 * w < L.Ww) ? (r0[2 * w] & r1[2 * w]) : 0u;
                uint32_t hi = (2 * w + 1 < L.Ww) ? (r0[2 * w + 1] & r1[2 * w + 1]) : 0u;
                lo &= lo >> 1;
                hi &= hi >> 1;
                Hd[i] = compress_even_bits(lo) | (compress_even_bits(hi) << 16);
            }
            __syncthreads();
            emit_sites_quads(N, Hd, L.H, L.Ww, L.pWw, (uint32_t)s << L.code.sh_s, L.code.sh_y, L.quad_bit, L.sites, L.counter, L.counter2, scratch);
        } else if (L.type == 1 && L.rt_rows > 0) {
            // row-tile conv layer: the evaluation needs the work set itself (it stores only there) and one entry per active unit
            uint32_t *ns = L.nset + (long long)s * nout;
            int cnt = 0;
            for (int i = tid; i < nout; i += kThreads) { const uint32_t w = N[i]; ns[i] = w; cnt += __popc(w); }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            if ((tid & 31) == 0 && cnt) atomicAdd(L.counter2, cnt);
            emit_units(N, L.H, L.W, L.Ww, L.rt_rows, L.rt_seg, L.rt_nxg, (uint32_t)s << L.code.sh_s, L.code.sh_y, L.sites, L.counter, scratch);
        } else {
            emit_sites(N, L.H, L.Ww, (uint32_t)s << L.code.sh_s, L.code.sh_y, L.sites, L.counter, scratch);
        }
        __syncthreads();
        for (int i = tid; i < nout; i += kThreads) bufB[i] = Z[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K3': the part of every layer's work set that is known BEFORE the leak sweep.  The exact frontier of a conv layer
// is dilate(previous frontier) united with the sign flips the sweep finds; this kernel runs the same bitmap chain
// with the flips left out (event pixels and pixel deaths from k_integrate, dilation, pool windows, sticky flags).
// Every operation is monotone, so skip[l] is a subset of the sites / windows layer l re-evaluates in this step -
// which the leak sweep may therefore leave alone (their F, A or (Fp, Ap) are overwritten, and they are on the
// layer's frontier regardless of their signs).  Nothing but skip[] is written.  One CTA per stream.
// Dynamic shared memory: 3 * max_words * 4 bytes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_frontier_skip(FrontAllParams p)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *bufA = reinterpret_cast<uint32_t *>(smem_raw);        // previous layer's (flip-free) frontier
    uint32_t *Hd = bufA + p.max_words;                              // scratch (horizontal dilation)
    uint32_t *N = Hd + p.max_words;                                 // this layer's set
    __shared__ __align__(16) FrontLayer s_layers[kMaxFrontLayers];
    const int s = blockIdx.x, tid = threadIdx.x;
    if (!p.active[s]) return;                                       // the sweep does not touch an idle stream
    stage_front_table(s_layers, p.layers, p.n_layers);
    for (int i = tid; i < p.words0; i += kThreads) bufA[i] = p.front0[(long long)s * p.words0 + i];
    __syncthreads();
    for (int li = 1; li < p.n_layers; ++li) {
        const FrontLayer &L = s_layers[li];
        const int nout = L.H * L.Ww;
        const uint32_t lastmask = (L.W & 31) ? ((1u << (L.W & 31)) - 1u) : 0xffffffffu;
        uint32_t *skip = L.skip + (long long)s * nout;
        if (L.type == 1) {
            for (int i = tid; i < L.Hin * L.Ww; i += kThreads) {
                const int y = i / L.Ww, w = i - y * L.Ww;
                uint32_t acc = 0u;
                for (int d = L.pad_l - L.kw + 1; d <= L.pad_l; ++d) acc |= row_shift(bufA + y * L.WwIn, L.WwIn, w, d);
                if (w == L.Ww - 1) acc &= lastmask;
                Hd[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < nout; i += kThreads) {
                const int y = i / L.Ww, w = i - y * L.Ww;
                uint32_t acc = 0u;
                for (int d = L.pad_t - L.kh + 1; d <= L.pad_t; ++d) {
                    const int yi = y - d;
                    if (yi >= 0 && yi < L.Hin) acc |= Hd[yi * L.Ww + w];
                }
                N[i] = acc;
                skip[i] = acc;
            }
        } else {
            const uint32_t *fl = L.flags + (long long)s * nout;
            for (int i = tid; i < nout; i += kThreads) {
                const int oy = i / L.Ww, w = i - oy * L.Ww;
                uint32_t hit = 0u;
                if (L.kh == 2 && L.kw == 2 && L.stride == 2) {
                    const uint32_t *r0 = bufA + (2 * oy) * L.WwIn, *r1 = r0 + L.WwIn;
                    uint32_t lo = (2 * w < L.WwIn) ? (r0[2 * w] | r1[2 * w]) : 0u;
                    uint32_t hi = (2 * w + 1 < L.WwIn) ? (r0[2 * w + 1] | r1[2 * w + 1]) : 0u;
                    lo |= lo >> 1;
                    hi |= hi >> 1;
                    hit = compress_even_bits(lo) | (compress_even_bits(hi) << 16);
                } else {
                    for (int b = 0; b < 32; ++b) {
                        const int ox = w * 32 + b;
                        if (ox >= L.W) break;
                        bool any = false;
                        for (int dy = 0; dy < L.kh && !any; ++dy)
                            for (int dx = 0; dx < L.kw; ++dx) {
                                const int iy = oy * L.stride + dy, ix = ox * L.stride + dx;
                                if ((bufA[iy * L.WwIn + (ix >> 5)] >> (ix & 31)) & 1u) { any = true; break; }
                            }
                        if (any) hit |= 1u << b;
                    }
                }
                if (w == L.Ww - 1) hit &= lastmask;
                const uint32_t wset = hit | fl[i];        // maxpool.py:118-126: hit windows and still-flagged windows
                N[i] = wset;
                skip[i] = wset;
            }
        }
        __syncthreads();
        for (int i = tid; i < nout; i += kThreads) bufA[i] = N[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K5: pool evaluation over the work list.   maxpool.py:130-151, cutils.pyx:161-177
//   per (window, channel): argmax of F over the window rows (ascending ky*kw+kx), ties -> smaller
//   rate R = A*slope(F), then smaller row; argmin of R (first); unstable = R[argmax] != R[argmin];
//   idx <- argmax; flags[window] |= any-channel unstable; (Fp, Ap) <- (F, A) at the argmax.
// One thread per (window, VEC consecutive channels): VEC = 4 uses 16-byte loads/stores (C % 4 == 0).
// ---------------------------------------------------------------------------------------------
struct PoolEvalParams {
    const uint32_t *sites;
    const int *counter;
    unsigned long long *accum;    // work counter for the roofline
    const float *F, *A;           // previous conv maps
    long long fstride;
    float alpha;
    int cW;                       // previous conv width
    uint8_t *idx;                 // [S][H*W*C]
    float *Fp, *Ap;               // [S][pstride] copies at the argmax
    long long pstride;            // idx bytes == copy floats per stream
    uint32_t *flags;              // [S][H*Ww]
    int C, H, W, Ww;
    int kh, kw, stride;
    SiteCode code;
};

// WIN2: the 2x2 / stride-2 window of every EFCN pool layer, with the eight 16-byte loads of an item issued before
// the first compare (the generic loop has run-time bounds and makes four dependent round trips to memory).
template <int VEC, bool WIN2>
__global__ void __launch_bounds__(kThreads) k_pool_eval(PoolEvalParams p)
{
    pdl_enter();
    const int n = *p.counter;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(p.accum, (unsigned long long)n);
    const int CG = p.C / VEC;
    const long long total = (long long)n * CG;
    for (long long wi = (long long)blockIdx.x * kThreads + threadIdx.x; wi < total; wi += (long long)gridDim.x * kThreads) {
        const uint32_t e = p.sites[wi / CG];
        const int c = (int)(wi % CG) * VEC;
        int s, oy, ox;
        site_decode(p.code, e, s, oy, ox);
        const int site = oy * p.W + ox;
        const float *Fb = p.F + (long long)s * p.fstride;
        const float *Ab = p.A + (long long)s * p.fstride;
        PoolBest best[VEC];
        if constexpr (WIN2 && VEC == 4) {
            const long long o00 = ((long long)(oy * 2) * p.cW + ox * 2) * p.C + c;
            const long long o10 = o00 + (long long)p.cW * p.C;
            const float4 f0 = __ldg(reinterpret_cast<const float4 *>(Fb + o00)), a0 = __ldg(reinterpret_cast<const float4 *>(Ab + o00));
            const float4 f1 = __ldg(reinterpret_cast<const float4 *>(Fb + o00 + p.C)), a1 = __ldg(reinterpret_cast<const float4 *>(Ab + o00 + p.C));
            const float4 f2 = __ldg(reinterpret_cast<const float4 *>(Fb + o10)), a2 = __ldg(reinterpret_cast<const float4 *>(Ab + o10));
            const float4 f3 = __ldg(reinterpret_cast<const float4 *>(Fb + o10 + p.C)), a3 = __ldg(reinterpret_cast<const float4 *>(Ab + o10 + p.C));
            const float fr[4][4] = {{f0.x, f0.y, f0.z, f0.w}, {f1.x, f1.y, f1.z, f1.w}, {f2.x, f2.y, f2.z, f2.w}, {f3.x, f3.y, f3.z, f3.w}};
            const float ar[4][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}, {a2.x, a2.y, a2.z, a2.w}, {a3.x, a3.y, a3.z, a3.w}};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                pool_first(best[v], fr[0][v], ar[0][v], p.alpha);
#pragma unroll
                for (int row = 1; row < 4; ++row) pool_next(best[v], row, fr[row][v], ar[row][v], p.alpha);
            }
        } else {
            int row = 0;
            for (int dy = 0; dy < p.kh; ++dy)
                for (int dx = 0; dx < p.kw; ++dx, ++row) {
                    const long long off = ((long long)(oy * p.stride + dy) * p.cW + (ox * p.stride + dx)) * p.C + c;
                    float f[VEC], a[VEC];
                    if constexpr (VEC == 4) {
                        const float4 f4 = __ldg(reinterpret_cast<const float4 *>(Fb + off));
                        const float4 a4 = __ldg(reinterpret_cast<const float4 *>(Ab + off));
                        f[0] = f4.x; f[1] = f4.y; f[2] = f4.z; f[3] = f4.w;
                        a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
                    } else {
                        f[0] = Fb[off];
                        a[0] = Ab[off];
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (row == 0) pool_first(best[v], f[v], a[v], p.alpha);
                        else pool_next(best[v], row, f[v], a[v], p.alpha);
                    }
                }
        }
        bool unstable = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) unstable |= best[v].r != best[v].rlow;      // cutils.pyx:177 (by value)
        const long long o = (long long)s * p.pstride + (long long)site * p.C + c;
        if constexpr (VEC == 4) {
            *reinterpret_cast<uchar4 *>(p.idx + o) = make_uchar4((unsigned char)best[0].row, (unsigned char)best[1].row,
                                                                  (unsigned char)best[2].row, (unsigned char)best[3].row);
            *reinterpret_cast<float4 *>(p.Fp + o) = make_float4(best[0].f, best[1].f, best[2].f, best[3].f);
            *reinterpret_cast<float4 *>(p.Ap + o) = make_float4(best[0].a, best[1].a, best[2].a, best[3].a);
        } else {
            p.idx[o] = (uint8_t)best[0].row;
            p.Fp[o] = best[0].f;
            p.Ap[o] = best[0].a;
        }
        // one atomic per (warp, window): lanes of the same window elect a leader
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, e);
        const unsigned uns = __ballot_sync(act, unstable);
        if (unstable && (__ffs(peers & uns) - 1) == (int)(threadIdx.x & 31))
            atomicOr(&p.flags[(long long)s * p.H * p.Ww + oy * p.Ww + (ox >> 5)], 1u << (ox & 31));
    }
}

// ---------------------------------------------------------------------------------------------
// K6: conv re-evaluation over the work list as a gathered GEMM.   conv2d.py:118-123,144-181
//   for every listed site o:  F[o,:] = W^T . patch(V_prev, o) + b ;  A[o,:] = W^T . patch(R_prev, o)
//   rows   = 2 per site (value row, rate row: same gather addresses, same weights)
//   K      = kh*kw*Cin ordered (ky,kx,ci) - the HWIO weight layout flattened, channel-last gathers
//   N      = Cout
// Each output is accumulated sequentially over k in ONE thread, so the summation order does not
// depend on where the site sits in a tile (identical patches -> identical bits, which keeps the
// oracle's exact pool ties exact).
// Tile: TS sites (2*TS rows) x BN outputs, K in chunks of BK; 256 threads, TM x TN per thread.
// ---------------------------------------------------------------------------------------------
struct ConvEvalParams {
    const uint32_t *sites;
    const int *counter;
    unsigned long long *accum;
    Src src;
    const float *wgt;      // [Kpad][Npad], zero padded
    const float *bias;     // [Npad]
    float *F, *A;
    long long fstride;
    int C, H, W;           // output
    int K, Kpad, Npad;
    int kh, kw, pad_t, pad_l;
    SiteCode code;
};

template <int BN, int TN, int TM, int BK>
__global__ void __launch_bounds__(kThreads) k_conv_eval(ConvEvalParams p)
{
    pdl_enter();
    constexpr int TX = BN / TN;            // threads along n
    constexpr int TY = kThreads / TX;      // threads along rows
    constexpr int ROWS = TY * TM;
    constexpr int TS = ROWS / 2;           // sites per tile
    constexpr int LDA = ROWS + 4;
    constexpr int LDB = BN + 4;
    static_assert(sizeof(FrontLayer) % 4 == 0, "FrontLayer is copied word by word");
    static_assert(ROWS % 2 == 0 && TM <= TS && TS % TM == 0, "tile shape");
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][LDB];
    __shared__ int s_str[TS], s_y[TS], s_x[TS];

    const int n_sites = *p.counter;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_sites > 0) atomicAdd(p.accum, (unsigned long long)n_sites);
    const int m_tiles = (n_sites + TS - 1) / TS;
    const int n_tiles = p.Npad / BN;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int Cin = p.src.C;
    const bool vec = (Cin % 4 == 0) && (p.src.kind != 0);

    for (int tile = blockIdx.x; tile < m_tiles * n_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int n0 = nt * BN;
        __syncthreads();
        for (int i = tid; i < TS; i += kThreads) {
            const int gi = mt * TS + i;
            if (gi < n_sites) {
                int s, y, x;
                site_decode(p.code, p.sites[gi], s, y, x);
                s_str[i] = s;
                s_y[i] = y;
                s_x[i] = x;
            } else {
                s_str[i] = -1;
            }
        }
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        __syncthreads();

        for (int k0 = 0; k0 < p.Kpad; k0 += BK) {
            // ---- gather the A-operand chunk: As[kk][site] = V, As[kk][TS+site] = R
            if (vec) {
                for (int u = tid; u < TS * (BK / 4); u += kThreads) {
                    const int kg = u % (BK / 4), i = u / (BK / 4);
                    const int k = k0 + 4 * kg;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), r = v;
                    const int s = s_str[i];
                    if (s >= 0 && k < p.K) {
                        const int tap = k / Cin, c = k - tap * Cin;
                        const int iy = s_y[i] + tap / p.kw - p.pad_t, ix = s_x[i] + tap % p.kw - p.pad_l;
                        if (iy >= 0 && iy < p.src.H && ix >= 0 && ix < p.src.W) src_fetch4(p.src, s, iy, ix, c, v, r);
                    }
                    As[4 * kg + 0][i] = v.x; As[4 * kg + 1][i] = v.y; As[4 * kg + 2][i] = v.z; As[4 * kg + 3][i] = v.w;
                    As[4 * kg + 0][TS + i] = r.x; As[4 * kg + 1][TS + i] = r.y;
                    As[4 * kg + 2][TS + i] = r.z; As[4 * kg + 3][TS + i] = r.w;
                }
            } else {
                for (int u = tid; u < TS * BK; u += kThreads) {
                    const int kk = u % BK, i = u / BK;
                    const int k = k0 + kk;
                    float v = 0.f, r = 0.f;
                    const int s = s_str[i];
                    if (s >= 0 && k < p.K) {
                        const int tap = k / Cin, c = k - tap * Cin;
                        const int iy = s_y[i] + tap / p.kw - p.pad_t, ix = s_x[i] + tap % p.kw - p.pad_l;
                        if (iy >= 0 && iy < p.src.H && ix >= 0 && ix < p.src.W) src_fetch(p.src, s, iy, ix, c, v, r);
                    }
                    As[kk][i] = v;
                    As[kk][TS + i] = r;
                }
            }
            // ---- weights chunk (padded on the host: always in bounds, 16-byte aligned)
            for (int u = tid; u < BK * (BN / 4); u += kThreads) {
                const int kk = u / (BN / 4), j = u % (BN / 4);
                const float4 w = __ldg(reinterpret_cast<const float4 *>(p.wgt + (long long)(k0 + kk) * p.Npad + n0) + j);
                *reinterpret_cast<float4 *>(&Bs[kk][4 * j]) = w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(&As[kk][ty * TM + i]);
                    a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
                }
                if constexpr (TN % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < TN; j += 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(&Bs[kk][tx * TN + j]);
                        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < TN; j += 2) {
                        const float2 t = *reinterpret_cast<const float2 *>(&Bs[kk][tx * TN + j]);
                        b[j] = t.x; b[j + 1] = t.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
        // ---- epilogue: value rows get the bias and go to F, rate rows go to A
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int row = ty * TM + i;
            const bool is_rate = row >= TS;
            const int si = is_rate ? row - TS : row;
            const int s = s_str[si];
            if (s < 0) continue;
            float *dst = (is_rate ? p.A : p.F) + (long long)s * p.fstride + ((long long)s_y[si] * p.W + s_x[si]) * p.C;
            const int nb = n0 + tx * TN;
            if (TN % 4 == 0 && (p.C & 3) == 0 && nb + TN <= p.C) {
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    float4 o = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
                    if (!is_rate) {
                        const float4 bv = __ldg(reinterpret_cast<const float4 *>(p.bias + nb + j));
                        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                    }
                    *reinterpret_cast<float4 *>(dst + nb + j) = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int n = nb + j;
                    if (n < p.C) dst[n] = is_rate ? acc[i][j] : (acc[i][j] + p.bias[n]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K6b: re-evaluation of a conv layer that reads the integration surface directly (Cin = 1): a kh x kw stencil,
// not a GEMM.   conv2d.py:118-123 with V = (float)S, R = [S > 0] (integration.py:33-43)
//   thread = (site, 4 output channels); the kh*kw surface values are read once per thread (float64 -> V, R),
//   weights [k][Cout] sit in shared memory; each output is accumulated sequentially over k in one thread
//   (position-independent summation order, like the GEMM kernels).
// ---------------------------------------------------------------------------------------------
struct StencilParams {
    const uint32_t *sites;
    const int *counter;
    unsigned long long *accum;
    const double *S;       // [streams][Hin*Win]
    long long sstride;
    int Hin, Win;
    const float *wgt;      // [Kpad][Npad]
    const float *bias;
    int Npad;
    float *F, *A;
    long long fstride;
    int C, H, W;           // output (C % 4 == 0)
    int kh, kw, pad_t, pad_l;
    SiteCode code;
};
constexpr int kStencilMaxK = 64, kStencilMaxC = 64;

// KH x KW > 0: compile-time window (3x3 for EFCN) - the loop unrolls and all surface loads are in flight together;
// KH = KW = 0: run-time window sizes.
template <int KH, int KW>
__global__ void __launch_bounds__(kThreads) k_conv_stencil(StencilParams p)
{
    pdl_enter();
    const int kh = KH > 0 ? KH : p.kh, kw = KW > 0 ? KW : p.kw;
    __shared__ __align__(16) float w_s[kStencilMaxK * kStencilMaxC];
    __shared__ __align__(16) float b_s[kStencilMaxC];
    const int K = kh * kw;
    for (int i = threadIdx.x; i < K * p.C; i += kThreads) w_s[i] = p.wgt[(i / p.C) * p.Npad + (i % p.C)];
    for (int i = threadIdx.x; i < p.C; i += kThreads) b_s[i] = p.bias[i];
    __syncthreads();
    const int n = *p.counter;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(p.accum, (unsigned long long)n);
    const int CG = p.C >> 2;
    const long long total = (long long)n * CG;
    for (long long wi = (long long)blockIdx.x * kThreads + threadIdx.x; wi < total; wi += (long long)gridDim.x * kThreads) {
        const uint32_t e = p.sites[wi / CG];
        const int c = (int)(wi % CG) * 4;
        int s, y, x;
        site_decode(p.code, e, s, y, x);
        const int site = y * p.W + x;
        const double *Sb = p.S + (long long)s * p.sstride;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f), a = f;
        auto tap = [&](int ky, int kx) {
            const int iy = y + ky - p.pad_t, ix = x + kx - p.pad_l;
            return ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) ? Sb[(long long)iy * p.Win + ix] : 0.0;
        };
        auto acc = [&](int k, double sv) {
            const float v = __double2float_rn(sv), r = sv > 0.0 ? 1.f : 0.f;
            const float4 w = *reinterpret_cast<const float4 *>(&w_s[k * p.C + c]);
            f.x = fmaf(v, w.x, f.x); f.y = fmaf(v, w.y, f.y); f.z = fmaf(v, w.z, f.z); f.w = fmaf(v, w.w, f.w);
            a.x = fmaf(r, w.x, a.x); a.y = fmaf(r, w.y, a.y); a.z = fmaf(r, w.z, a.z); a.w = fmaf(r, w.w, a.w);
        };
        if constexpr (KH > 0 && KW > 0) {
            double sv[KH * KW];
#pragma unroll
            for (int k = 0; k < KH * KW; ++k) sv[k] = tap(k / KW, k % KW);
#pragma unroll
            for (int k = 0; k < KH * KW; ++k) acc(k, sv[k]);
        } else {
            int k = 0;
            for (int ky = 0; ky < kh; ++ky)
                for (int kx = 0; kx < kw; ++kx, ++k) acc(k, tap(ky, kx));
        }
        const float4 b = *reinterpret_cast<const float4 *>(&b_s[c]);
        f.x += b.x; f.y += b.y; f.z += b.z; f.w += b.w;
        const long long o = (long long)s * p.fstride + (long long)site * p.C + c;
        *reinterpret_cast<float4 *>(p.F + o) = f;
        *reinterpret_cast<float4 *>(p.A + o) = a;
    }
}

// ---------------------------------------------------------------------------------------------
// K7: head = featuremap() of the last layer, channel-last (event_numpy.py:79,101).
// ---------------------------------------------------------------------------------------------
struct HeadParams {
    Src src;
    float *out;          // [S][H*W*C]
    int S;
};

__global__ void __launch_bounds__(kThreads) k_head(HeadParams p)
{
    pdl_enter();
    const long long per = (long long)p.src.H * p.src.W * p.src.C;
    const long long total = per * p.S;
    if (p.src.kind == 1 && (per & 1) == 0) {
        // a conv / pool map is channel-last like the head: element i of stream s is F[s * fstride + i]; two per thread
        const long long per2 = per >> 1, total2 = per2 * p.S;
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total2; i += (long long)gridDim.x * kThreads) {
            const long long s = i / per2, r2 = i - s * per2;
            const float2 f = __ldg(reinterpret_cast<const float2 *>(p.src.F + s * p.src.fstride) + r2);
            reinterpret_cast<float2 *>(p.out + s * per)[r2] = make_float2(__fmul_rn(f.x, slope_of(f.x, p.src.alpha)), __fmul_rn(f.y, slope_of(f.y, p.src.alpha)));
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int s = (int)(i / per);
        const long long rem = i - (long long)s * per;
        const int c = (int)(rem % p.src.C);
        const int site = (int)(rem / p.src.C);
        float v, r;
        src_fetch(p.src, s, site / p.src.W, site % p.src.W, c, v, r);
        p.out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K8: one stream's layer accessors, channel-last float32 (layer.py:53-81):
//   surface (pre-activation, pool: gathered through argmax), layer_actfn (slope), conv_actfn (R),
//   featuremap (V).  Any output pointer may be null.  Used by the Python Layer mirror only.
// ---------------------------------------------------------------------------------------------
struct ViewParams {
    Src src;
    int stream;
    float *surface, *layer_actfn, *conv_actfn, *featuremap;
};

__global__ void __launch_bounds__(kThreads) k_layer_view(ViewParams p)
{
    const Src &q = p.src;
    const long long per = (long long)q.H * q.W * q.C;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < per; i += (long long)gridDim.x * kThreads) {
        const float f = q.F[(long long)p.stream * q.fstride + i];
        const float a = q.A[(long long)p.stream * q.fstride + i];
        const float sl = slope_of(f, q.alpha);
        if (p.surface) p.surface[i] = f;
        if (p.layer_actfn) p.layer_actfn[i] = sl;
        if (p.conv_actfn) p.conv_actfn[i] = __fmul_rn(a, sl);
        if (p.featuremap) p.featuremap[i] = __fmul_rn(f, sl);
    }
}

// ---------------------------------------------------------------------------------------------
// K9: measurement helper - number of float4 groups of the leak-rate maps A holding a non-zero
// (the groups for which the leak sweep must read and write F).  Not on the hot path.
// ---------------------------------------------------------------------------------------------
// The caller passes a table whose chunk0 values are dense (kSweepChunk float4 per chunk).
__global__ void __launch_bounds__(kThreads) k_count_nz4(const __grid_constant__ SweepParams p, unsigned long long *out)
{
    const int s = blockIdx.y;
    int li = 0;
    const int chunk = blockIdx.x;
#pragma unroll 1
    for (int i = 1; i < p.n_layers; ++i)
        if (chunk >= p.L[i].chunk0) li = i;
    const SweepLayer &L = p.L[li];
    const int base = (chunk - L.chunk0) * kSweepChunk;
    const float4 *A4 = reinterpret_cast<const float4 *>(L.A + (long long)s * L.fstride);
    int cnt = 0;
    for (int j = 0; j < kSweepVec; ++j) {
        const int idx = base + j * kThreads + threadIdx.x;
        if (idx < L.n4) {
            const float4 a = A4[idx];
            cnt += (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) ? 1 : 0;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

// K9b: measurement helper - number of set bits of a [S][words] bitmap (live sites of a layer).
// With `notmask`: bits of bm that are clear in notmask (live sites the leak sweep does not skip).
__global__ void __launch_bounds__(kThreads) k_count_bits(const uint32_t *bm, const uint32_t *notmask, long long words, unsigned long long *out)
{
    int cnt = 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < words; i += (long long)gridDim.x * kThreads)
        cnt += __popc(notmask ? bm[i] & ~notmask[i] : bm[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

// ---------------------------------------------------------------------------------------------
// utility kernels: reset / init plumbing
// ---------------------------------------------------------------------------------------------
// dst[s][i] = src[i] (or 0 when src == nullptr) for every stream with mask[s] != 0 (mask nullptr = all); 4-byte words.
__global__ void __launch_bounds__(kThreads) k_broadcast_words(uint32_t *dst, const uint32_t *src, long long words,
                                                               long long stride_words, const uint8_t *mask)
{
    const int s = blockIdx.y;
    if (mask && !mask[s]) return;
    uint32_t *d = dst + (long long)s * stride_words;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < words; i += (long long)gridDim.x * kThreads)
        d[i] = src ? src[i] : 0u;
}

__global__ void __launch_bounds__(kThreads) k_reset_scalars(int *prev_ts, double *delta, uint8_t *active, const uint8_t *mask, int S)
{
    const int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= S || (mask && !mask[s])) return;
    prev_ts[s] = 0;
    delta[s] = 0.0;
    active[s] = 0;
}

// every site of stream 0 -> work list (used once, to evaluate the initial state)
__global__ void __launch_bounds__(kThreads) k_all_sites(uint32_t *sites, int *counter, int H, int W, SiteCode code)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < H * W) sites[i] = site_encode(code, 0, i / W, i % W);
    if (i == 0) *counter = H * W;
}

}  // namespace aec
