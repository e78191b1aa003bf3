/*
 * TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's two native functions.
 *
 * Restates, in plain C, the algorithm of the reference's Cython module `src/libs/cutils.pyx`:
 *   - oracle_im2col_event  follows cutils.pyx:29-134  (event-driven im2col with first-touch dedup)
 *   - oracle_min_argmax    follows cutils.pyx:139-179 (argmax with tie-break on a second matrix)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (async-ev-cnn_b200/) never does.
 *
 * Parity pinned: tests/test_oracle_cutils.py checks both functions bit-for-bit against the compiled
 * reference module (oracle/_ref/cutils.so) when /root/reference is present, and against the
 * committed golden vectors (tests/golden/) everywhere.
 *
 * Build: gcc -O2 -fPIC -shared oracle/cutils_port.c -o oracle/libcutils_port.so   (see oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/*
 * image      : float32 [C][H][W], C-contiguous.
 * ev_y, ev_x : int32 [n_ev] input-event coordinates in image space, visited in the given order.
 * out_cols   : caller-allocated column-major ("F-order") matrix with `rows` rows where
 *                rows = C*kh*kw (chan_as_cols == 0)  -> one column per output site,
 *                rows = kh*kw   (chan_as_cols != 0)  -> C columns per output site, column = site*C + c
 *              sized for the worst case (all Hout*Wout sites).
 * out_y/out_x: caller-allocated int32 [Hout*Wout]; receive the output-site coordinates in
 *              FIRST-TOUCH order (cutils.pyx:108-112).
 * returns the number of distinct output sites, or -1 for a stride that is neither 1 nor the
 * kernel size (cutils.pyx:88-89 raises NotImplementedError).
 */
int oracle_im2col_event(const float *image, int C, int H, int W,
                        const int32_t *ev_y, const int32_t *ev_x, int n_ev,
                        int kh, int kw, int stride, int chan_as_cols,
                        float *out_cols, int32_t *out_y, int32_t *out_x)
{
    const int Hout = (H - kh) / stride + 1;
    const int Wout = (W - kw) / stride + 1;
    const int ksz = kh * kw;
    unsigned char *seen;
    int n_sites = 0;

    if (!(stride == 1 || (stride == kw && stride == kh)))
        return -1;
    seen = (unsigned char *)calloc((size_t)Hout * Wout, 1);
    if (!seen)
        return -2;

    for (int e = 0; e < n_ev; ++e) {
        const int y = ev_y[e], x = ev_x[e];
        int y0, y1, x0, x1; /* span of image rows/cols covered by all windows containing (y,x) */
        if (stride == 1) {   /* cutils.pyx:78-82 */
            y0 = imax(0, y - (kh - 1));
            y1 = imin(H, y + kh);
            x0 = imax(0, x - (kw - 1));
            x1 = imin(W, x + kw);
        } else {             /* stride == kernel: the single window holding the event, cutils.pyx:83-87 */
            y0 = (y / stride) * kh;
            y1 = y0 + kh;
            x0 = (x / stride) * kw;
            x1 = x0 + kw;
        }
        /* number of window positions along each axis (cutils.pyx:92-93) */
        const int ny = (y1 - y0 - kh) / stride + 1;
        const int nx = (x1 - x0 - kw) / stride + 1;

        for (int dy = 0; dy < ny; dy += stride) {      /* cutils.pyx:97-99: step == stride */
            for (int dx = 0; dx < nx; dx += stride) {
                const int top = y0 + dy, left = x0 + dx;
                const int oy = top / stride, ox = left / stride;
                if (seen[oy * Wout + ox])
                    continue;
                seen[oy * Wout + ox] = 1;
                out_y[n_sites] = oy;
                out_x[n_sites] = ox;
                for (int c = 0; c < C; ++c) {
                    const float *plane = image + (size_t)c * H * W;
                    for (int ry = 0; ry < kh; ++ry) {
                        const float *src = plane + (size_t)(top + ry) * W + left;
                        if (chan_as_cols) {
                            float *dst = out_cols + ((size_t)n_sites * C + c) * ksz + ry * kw;
                            for (int rx = 0; rx < kw; ++rx) dst[rx] = src[rx];
                        } else {
                            float *dst = out_cols + (size_t)n_sites * C * ksz + (size_t)c * ksz + ry * kw;
                            for (int rx = 0; rx < kw; ++rx) dst[rx] = src[rx];
                        }
                    }
                }
                ++n_sites;
            }
        }
    }
    free(seen);
    return n_sites;
}

/*
 * max_arg, min_arg : column-major float32 [rows][cols] (element (r,c) at c*rows + r).
 * argmax[c]     : row of the maximum of max_arg[:,c]; rows are scanned ascending, a strictly larger
 *                 value wins, an EQUAL value wins only if its min_arg entry is strictly smaller
 *                 (cutils.pyx:166-170).
 * not_argmin[c] : 1 if min_arg at the chosen row differs BY VALUE from the column minimum of
 *                 min_arg (first minimum, strict <) (cutils.pyx:173-177).
 */
void oracle_min_argmax(const float *max_arg, const float *min_arg, int rows, int cols,
                       int32_t *argmax, int32_t *not_argmin)
{
    for (int c = 0; c < cols; ++c) {
        const float *mx = max_arg + (size_t)c * rows;
        const float *mn = min_arg + (size_t)c * rows;
        int best = 0, low = 0;
        for (int r = 0; r < rows; ++r) {
            if (mx[r] > mx[best]) {
                best = r;
            } else if (r > 0 && mx[r] == mx[best]) {
                if (mn[r] < mn[best])
                    best = r;
            }
            if (mn[r] < mn[low])
                low = r;
        }
        argmax[c] = best;
        not_argmin[c] = (mn[best] != mn[low]) ? 1 : 0;
    }
}
