// aec_frontend.cuh - the steps either side of the hot path, on the device:
//   k_ndata_decode  N-MNIST / N-Caltech101 5-byte event records -> (y, x, ts) int32 triples, with the runner's
//                   per-sample transform (zero-based timestamps, centre crop) fused in;
//   k_decode_head   YOLO decode of the [h_cells, w_cells, C + 5B] head into pixel boxes, confidences, labels.
#pragma once
#include "aec_kernels.cuh"

namespace aec {

// ---------------------------------------------------------------------------------------------
// One CTA per recording.   src/readers/file_reader.py:36-58, src/libs/runner.py:24-33, src/libs/utils.py:4-28
//   record i = 5 bytes: x, y, (p << 7 | ts[22:16]), ts[15:8], ts[7:0];  a record with y == 240 is a timestamp
//   overflow marker: it adds 2^13 to every LATER timestamp and is dropped.
//   zero_base: ts -= ts of the first kept event.
//   crop: new_top = (x_max - x_min - new_w) // 2, new_left = (y_max - y_min - new_h) // 2 (the reference's own
//         naming), keep x in [new_left, new_left + new_w) and y in [new_top, new_top + new_h), then shift the
//         kept events so that their minimum x and y are 0.
// Three passes over the recording (L2-resident after the first): extents, kept extents, ordered compaction.
// ---------------------------------------------------------------------------------------------
struct NdataParams {
    const uint8_t *raw;            // all recordings, packed
    const long long *byte_off;     // [R+1]
    int32_t *events;               // (y, x, ts) triples; recording r starts at event index byte_off[r] / 5
    int32_t *polarity;             // same indexing (may be null)
    int32_t *counts;               // [R] events kept
    int zero_base, crop, new_h, new_w;
};

__device__ __forceinline__ int floordiv2(int v) { return v >> 1; }      // python // 2 (arithmetic shift floors)

__global__ void __launch_bounds__(kThreads) k_ndata_decode(NdataParams p)
{
    __shared__ int scratch[9];
    __shared__ int s_red[4][kThreads / 32];
    __shared__ int s_val[8];
    const int r = blockIdx.x, tid = threadIdx.x;
    const long long b0 = p.byte_off[r];
    const int n = (int)((p.byte_off[r + 1] - b0) / 5);
    const uint8_t *raw = p.raw + b0;
    const long long out0 = b0 / 5;

    auto block_minmax = [&](int xmin, int xmax, int ymin, int ymax, int slot) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, d));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, d));
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, d));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, d));
        }
        __syncthreads();
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = xmin; s_red[1][tid >> 5] = xmax; s_red[2][tid >> 5] = ymin; s_red[3][tid >> 5] = ymax; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < kThreads / 32; ++i) {
                s_red[0][0] = min(s_red[0][0], s_red[0][i]); s_red[1][0] = max(s_red[1][0], s_red[1][i]);
                s_red[2][0] = min(s_red[2][0], s_red[2][i]); s_red[3][0] = max(s_red[3][0], s_red[3][i]);
            }
            s_val[slot] = s_red[0][0]; s_val[slot + 1] = s_red[1][0]; s_val[slot + 2] = s_red[2][0]; s_val[slot + 3] = s_red[3][0];
        }
        __syncthreads();
    };

    // ---- pass 1: extents of the TD events (overflow markers excluded)
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    for (int i = tid; i < n; i += kThreads) {
        const int x = raw[5 * (long long)i], y = raw[5 * (long long)i + 1];
        if (y != 240) { xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y); }
    }
    block_minmax(xmin, xmax, ymin, ymax, 0);
    const bool any_td = s_val[0] != INT_MAX;
    const bool crop = p.crop && any_td;
    const int new_top = crop ? floordiv2(s_val[1] - s_val[0] - p.new_w) : 0;
    const int new_left = crop ? floordiv2(s_val[3] - s_val[2] - p.new_h) : 0;
    auto inside = [&](int x, int y) {
        return y != 240 && (!crop || (x >= new_left && x < new_left + p.new_w && y >= new_top && y < new_top + p.new_h));
    };
    // ---- pass 2: extents of the kept events (their minimum becomes the origin)
    int shift_x = 0, shift_y = 0;
    if (crop) {
        xmin = INT_MAX; xmax = INT_MIN; ymin = INT_MAX; ymax = INT_MIN;
        for (int i = tid; i < n; i += kThreads) {
            const int x = raw[5 * (long long)i], y = raw[5 * (long long)i + 1];
            if (inside(x, y)) { xmin = min(xmin, x); ymin = min(ymin, y); }
        }
        block_minmax(xmin, xmax, ymin, ymax, 4);
        if (s_val[4] != INT_MAX) { shift_x = s_val[4]; shift_y = s_val[6]; }
    }
    // ---- pass 3: ordered compaction; running counts of overflow markers and of kept events
    int ovf_before = 0, kept_before = 0, ts_first = 0;
    bool have_first = false;
    for (int base = 0; base < n; base += kThreads) {
        const int i = base + tid;
        int x = 0, y = 0, pol = 0, ts = 0;
        bool is_ovf = false, td = false;
        if (i < n) {
            const uint8_t *q = raw + 5 * (long long)i;
            x = q[0]; y = q[1];
            pol = q[2] >> 7;
            ts = ((q[2] & 127) << 16) | (q[3] << 8) | q[4];
            is_ovf = y == 240;
            td = !is_ovf;
        }
        int tot_ovf, tot_td, tot_keep;
        const int ovf_here = block_excl_scan(is_ovf ? 1 : 0, scratch, &tot_ovf);
        ts += (ovf_before + ovf_here) << 13;                       // markers strictly before this record
        // zero base: ts of the first TD event of the recording (before the crop, runner.py:26)
        const int td_here = block_excl_scan(td ? 1 : 0, scratch, &tot_td);
        if (!have_first && tot_td > 0) {
            if (td && td_here == 0) s_val[7] = ts;
            __syncthreads();
            ts_first = s_val[7];
            have_first = true;
        }
        const bool keep = i < n && inside(x, y);
        const int pos = block_excl_scan(keep ? 1 : 0, scratch, &tot_keep);
        if (keep) {
            const long long o = out0 + kept_before + pos;
            p.events[3 * o] = y - shift_y;
            p.events[3 * o + 1] = x - shift_x;
            p.events[3 * o + 2] = p.zero_base ? ts - ts_first : ts;
            if (p.polarity) p.polarity[o] = pol;
        }
        ovf_before += tot_ovf;
        kept_before += tot_keep;
    }
    if (tid == 0) p.counts[r] = kept_before;
}

// ---------------------------------------------------------------------------------------------
// Batching of a recording's events into steps.   src/libs/runner.py:65-72 (intended semantics)
//   count mode (usec == 0): n = max(ceil(N / size), 1) chunks as np.array_split(events, n) cuts them: the first N % n
//                           chunks hold N / n + 1 events, the others N / n;
//   time mode  (usec > 0):  bins = arange(0, ts[N-1], usec); id_j = digitize(ts_j, bins) = min(ts_j / usec + 1, len(bins))
//                           (0 for a negative ts); a new chunk starts wherever id_j != id_(j-1).
// One CTA per recording; the chunk offsets (relative to the recording's first event) go to
// chunk_off[rec_start[r] + 2 r + k], k = 0 .. n_chunks[r] (a recording of N events has at most max(N, 1) chunks).
// ---------------------------------------------------------------------------------------------
struct SplitParams {
    const int32_t *events;        // (y, x, ts) triples
    const long long *rec_start;   // [R] index of the recording's first event
    const int32_t *counts;        // [R] events in the recording
    int32_t *chunk_off;           // [total + 2 R]
    int32_t *n_chunks;            // [R]
    int size, usec;
};

__global__ void __launch_bounds__(kThreads) k_split_batches(SplitParams p)
{
    __shared__ int scratch[9];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int N = p.counts[r];
    const long long e0 = p.rec_start[r];
    int32_t *out = p.chunk_off + e0 + 2 * (long long)r;
    if (p.usec <= 0) {
        const int n = max((N + p.size - 1) / p.size, 1);
        const int q = N / n, rem = N - q * n;
        for (int i = tid; i <= n; i += kThreads) out[i] = i * q + min(i, rem);
        if (tid == 0) p.n_chunks[r] = n;
        return;
    }
    if (N == 0) {
        if (tid == 0) { out[0] = 0; out[1] = 0; p.n_chunks[r] = 1; }
        return;
    }
    const int32_t *ts = p.events + 3 * e0 + 2;
    const int last = ts[3 * (long long)(N - 1)];
    const int nb = last > 0 ? (last + p.usec - 1) / p.usec : 0;
    auto bin = [&](int t) { return t < 0 ? 0 : min(t / p.usec + 1, nb); };
    if (tid == 0) out[0] = 0;
    int cuts_before = 0;
    for (int base = 1; base < N; base += kThreads) {
        const int j = base + tid;
        const bool cut = j < N && bin(ts[3 * (long long)j]) != bin(ts[3 * (long long)(j - 1)]);
        int tot;
        const int pos = block_excl_scan(cut ? 1 : 0, scratch, &tot);
        if (cut) out[1 + cuts_before + pos] = j;
        cuts_before += tot;
    }
    if (tid == 0) { out[1 + cuts_before] = N; p.n_chunks[r] = cuts_before + 1; }
}

// Event range of every stream for step `b` of a batched run: stream s consumes chunk b of recording s (nothing once its
// recording is exhausted).  begin / end index the packed event array.
__global__ void __launch_bounds__(kThreads) k_chunk_ranges(const long long *rec_start, const int32_t *chunk_off, const int32_t *n_chunks,
                                                           int b, int R, int32_t *begin, int32_t *end)
{
    const int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= R) return;
    const long long e0 = rec_start[s];
    const int32_t *co = chunk_off + e0 + 2 * (long long)s;
    const bool live = b < n_chunks[s];
    begin[s] = (int32_t)(e0 + (live ? co[b] : 0));
    end[s] = (int32_t)(e0 + (live ? co[b + 1] : 0));
}

// ---------------------------------------------------------------------------------------------
// YOLO decode.   src/libs/viz.py:27-46 (convert_bboxes, sqrt = True), :131-148,165 (draw_bboxes)
//   head [S][gh][gw][C + 5B]: C class scores, then B x (x, y, w, h, conf);
//   box  x = ((bx + col) / gw) * w_img,  y = ((by + row) / gh) * h_img,  w = bw^2 * w_img,  h = bh^2 * h_img
//   conf = box confidence, valid = conf > threshold, label = argmax_c(class_c * conf) (first maximum).
// One thread per (stream, cell, box); float32 arithmetic in the reference's order (bit-exact).
// ---------------------------------------------------------------------------------------------
struct DecodeParams {
    const float *head;
    float *boxes;        // [S][cells*B][4]
    float *conf;         // [S][cells*B]
    int32_t *label;      // [S][cells*B]
    uint8_t *valid;      // [S][cells*B]
    int S, gh, gw, C, B, h_img, w_img;
    float thr;
};

__global__ void __launch_bounds__(kThreads) k_decode_head(DecodeParams p)
{
    const long long total = (long long)p.S * p.gh * p.gw * p.B;
    const int D = p.C + 5 * p.B;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int b = (int)(i % p.B);
        const long long cell = i / p.B;                  // s * gh * gw + row * gw + col
        const int col = (int)(cell % p.gw), row = (int)((cell / p.gw) % p.gh);
        const float *h = p.head + cell * D;
        const float *bb = h + p.C + 5 * b;
        const float cf = bb[4];
        p.boxes[4 * i + 0] = __fmul_rn(__fdiv_rn(__fadd_rn(bb[0], (float)col), (float)p.gw), (float)p.w_img);
        p.boxes[4 * i + 1] = __fmul_rn(__fdiv_rn(__fadd_rn(bb[1], (float)row), (float)p.gh), (float)p.h_img);
        p.boxes[4 * i + 2] = __fmul_rn(__fmul_rn(bb[2], bb[2]), (float)p.w_img);
        p.boxes[4 * i + 3] = __fmul_rn(__fmul_rn(bb[3], bb[3]), (float)p.h_img);
        p.conf[i] = cf;
        p.valid[i] = cf > p.thr ? 1 : 0;
        int best = 0;
        float bv = __fmul_rn(h[0], cf);
        for (int c = 1; c < p.C; ++c) {
            const float v = __fmul_rn(h[c], cf);
            if (v > bv) { bv = v; best = c; }            // np.argmax: first maximum (NaN-free inputs)
        }
        p.label[i] = best;
    }
}

}  // namespace aec
