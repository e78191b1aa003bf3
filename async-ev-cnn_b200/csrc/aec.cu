// aec.cu - C-ABI host side of libaec_b200.so (see include/aec.h).
//
// Owns device memory, turns the reference's layer chain (src/models/event_numpy.py:53-73) into a
// static launch schedule and issues the kernels of aec_kernels.cuh.  No torch types, no CPU
// fallback: every compute entry point launches sm_100a kernels or fails.
#include "../../include/aec.h"
#include "aec_kernels.cuh"
#include "aec_tc.cuh"
#include "aec_rt.cuh"
#include "aec_frontend.cuh"

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace aec;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(AEC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static inline long long pad4(long long v) { return (v + 3) & ~3LL; }

struct HostLayer {
    int type = 0;
    int C = 0, H = 0, W = 0, Ww = 0;
    int Cin = 0, Hin = 0, Win = 0;
    int kh = 0, kw = 0, stride = 1, pad_t = 0, pad_l = 0;
    SiteCode code = {0, 0};     // work-list entry coding of this layer's sites (stream << sh_s | y << sh_y | x)
    float alpha = 0.f;
    int K = 0, Kpad = 0, Npad = 0, BN = 0;
    long long fstride = 0;      // conv: floats per stream; pool: idx bytes per stream
    std::vector<float> h_w, h_b;   // padded host copies until finalize
    float *F = nullptr, *A = nullptr, *initF = nullptr;
    uint8_t *idx = nullptr, *initIdx = nullptr;
    float *Fp = nullptr, *Ap = nullptr, *initFp = nullptr;   // pool: copy of the conv maps at the argmax
    uint32_t *flags = nullptr, *front = nullptr, *signchg = nullptr, *nzr = nullptr, *skip = nullptr;
    uint32_t *sites = nullptr;      // this layer's work list (region of aec_net::sites)
    float *wgt = nullptr, *bias = nullptr;
    // tensor-core path (aec_tc.cuh): pre-split, pre-swizzled weight image and tile geometry
    bool tc = false;
    int KB = 0, Mrows = 0, Mch = 0, rep = 1, m_tiles = 0, mtu = 1, w_stages = 0, n_acc = 1, tc_blocks = 0;
    bool pool_fuse = false;        // conv (gathered, weights-as-M): complete windows of the 2x2 pool behind it are evaluated in its epilogue
    bool pool_in_conv = false;     // pool: ... by the conv layer in front of it
    bool tc_fast_decode = false;   // which site-decoder variant of k_conv_eval_tc the layer runs (fixed at finalize)
    bool tc_sm = false;            // sites-as-M form of the kernel (Cout <= 64, multiple of 4): aec_tc.cuh
    bool tc_pair = false;          // CTA pairs (tcgen05 cta_group::2, M = 256): an even number of weight tiles and many streams (finalize)
    // row-tile form (aec_rt.cuh) of a sites-as-M layer: units of rt_R output rows x one x segment instead of single sites
    bool rt = false;
    int rt_R = 0, rt_sw_shift = 0, rt_SEG = 0, rt_nxg = 0, rt_CB = 0, rt_ncb = 0, rt_P = 0, rt_xst = 0, rt_wst = 0;
    size_t rt_xtile = 0, rt_wtile = 0, rt_smem = 0;
    bool rt_wres = false;          // weights resident in shared memory for the whole launch
    int rt_groups = 3;             // producer groups
    std::vector<float> h_rtimg;
    float *rtimg = nullptr;
    uint32_t *nset = nullptr;      // [S][H*Ww] exact work set of the step (written by the frontier kernel)
    // pool layer behind the FIRST conv layer: its sticky windows are evaluated by the leak sweep (SweepPool, aec_kernels.cuh)
    bool swp_fused = false;
    uint32_t *swp_uns = nullptr;
    size_t tc_smem = 0;
    std::vector<float> h_wimg;
    float *wimg = nullptr;
    unsigned long long *tc_timing = nullptr;   // 16 counters, used while aec_net_tc_timing is enabled
};

struct aec_net {
    int device = 0, S = 0, H = 0, W = 0;
    double leak = 0;
    int max_events = 2048, hash_slots = 4096;
    bool finalized = false;
    std::vector<HostLayer> L;
    std::vector<void *> allocs;
    size_t dev_bytes = 0, per_stream_bytes = 0;
    double *surface = nullptr, *delta = nullptr;
    int *prev_ts = nullptr;
    uint8_t *active = nullptr, *mask = nullptr;
    uint32_t *sites = nullptr;
    bool sweep_skip = true;              // the leak sweep leaves the sites alone that the step re-evaluates (AEC_SWEEP_SKIP=0: leak every live site)
    bool tc_half = true, rt_store32 = true;   // developer switches read at finalize (AEC_TC_HALF, AEC_RT_STORE32)
    int tc_pair_mode = -1;                    // AEC_TC_PAIR: 0 never, 1 wherever the layer allows it, unset (-1): from 32 streams on
    bool pdl = false;                         // AEC_PDL=1: programmatic dependent launch for the kernels of the step (measured slower: 3.73 vs 3.66 ms)
    FrontLayer *front_table = nullptr;   // device copy of the per-layer frontier descriptors (k_frontier_all)
    int front_max_words = 0;
    int *counts = nullptr, *err_flag = nullptr;
    int pending_err = 0, last_err_bits = 0;   // event-error bits latched across an internal sync of the pipelined path
    unsigned long long *accum = nullptr;
    float *head = nullptr;
    size_t head_per_stream = 0;
    int32_t *ev_dev = nullptr, *off_dev = nullptr;
    const int32_t *ev_ends = nullptr;   // batched recordings (aec_net_run_ndata): stream s owns events off[s] .. ev_ends[s]; null otherwise
    size_t ev_cap = 0;
    // pipelined host stepping (aec_net_step_host_async): two slots of event staging / head buffers, copy streams
    struct HostSlot {
        int32_t *ev = nullptr, *off = nullptr;
        size_t ev_cap = 0;
        float *head = nullptr;
        cudaEvent_t ev_ready = nullptr, k_done = nullptr, d2h_done = nullptr;
        bool used = false;
    } slot[2];
    cudaStream_t h2d = nullptr, d2h = nullptr;
    unsigned long long async_calls = 0;
    float *head_cur = nullptr;     // where k_head writes (n->head, or a slot's buffer)
    float *head_last = nullptr;    // where the LAST step wrote the head (aec_net_head_device / aec_net_decode_head read it)
    // CUDA graph of one step (17 launches for EFCN replayed with one call; only the event pointers of the surface kernel and
    // the output pointer of the head kernel change between replays and are patched in place)
    struct StepGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t integ = nullptr, head = nullptr;
        cudaKernelNodeParams integ_kp, head_kp;
        IntegrateParams ip;
        HeadParams hp;
        unsigned long long launches = 0;
        bool failed = false;
    } sg;
    cudaStream_t cap = nullptr;
    int use_graph = -1;            // -1 unknown, 0 off (AEC_GRAPH=0), 1 on
    float *dec_boxes = nullptr, *dec_conf = nullptr;   // aec_net_decode_head scratch
    int32_t *dec_label = nullptr;
    uint8_t *dec_valid = nullptr;
    size_t dec_cap = 0;
    int num_sms = 148;
    int sweep_chunks = 0, sweep_nconv = 0, sweep_conv_chunks = 0;
    SweepParams sweep_all;
    SweepWindowsParams sweep_win;        // the first conv layer's map when the leak sweep also evaluates the pool behind it
    int sweep_win_chunks = 0;
    unsigned long long launches = 0, steps = 0;
    int conv_eval_blocks[4] = {0, 0, 0, 0};
    // per-launch timing (aec_net_profile): one CUDA event between consecutive launches of a step
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;   // events of the current step (n_slots + 1)
    std::vector<double> prof_ms;            // accumulated ms per slot
    std::vector<std::string> prof_names;    // name of the launch each slot timed ("surface", "L3.eval", ...), recorded on the first profiled step
    int prof_slot = 0;
    unsigned long long prof_steps = 0;
    bool tc_timing_on = false;
    int tc_debug = 0;           // AEC_TC_DEBUG knock-out bits, read once at finalize (measurement only)
    float *view = nullptr;      // 4 x max(H*W*C) scratch for aec_net_read_view
    size_t view_elems = 0;
};

template <typename T>
static int dev_alloc(aec_net *n, T **out, size_t count, bool per_stream)
{
    void *p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(AEC_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return fail(AEC_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    n->allocs.push_back(p);
    n->dev_bytes += bytes;
    if (per_stream) n->per_stream_bytes += bytes / (size_t)n->S;
    *out = static_cast<T *>(p);
    return AEC_OK;
}

// Float maps gathered by the tensor-core conv kernel (F, A, Fp, Ap) get kMapGuardFloats zero floats in
// front of them: an out-of-map tap reads that line instead of branching (aec_tc.cuh: item_load).
static const int kMapGuardFloats = 32;
static int dev_alloc_map(aec_net *n, float **out, size_t count)
{
    float *base = nullptr;
    int rc = dev_alloc(n, &base, count + kMapGuardFloats, true);
    if (rc) return rc;
    *out = base + kMapGuardFloats;
    return AEC_OK;
}

extern "C" const char *aec_last_error(void) { return g_err.c_str(); }
extern "C" int aec_version(void) { return 1000; }

extern "C" int aec_net_create(aec_net **out, int device, int n_streams, int height, int width, double leak,
                              int max_events_per_step)
{
    if (!out) return fail(AEC_EINVAL, "out is NULL");
    if (n_streams < 1 || n_streams > 65535) return fail(AEC_EINVAL, "n_streams must be in [1, 65535], got %d", n_streams);
    if (height < 1 || width < 1) return fail(AEC_EINVAL, "bad surface size %dx%d", height, width);
    if ((long long)n_streams * height * width >= (1LL << 32))
        return fail(AEC_EINVAL, "n_streams*H*W must be < 2^32 (work-list entries are 32-bit)");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(AEC_EINVAL, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    aec_net *n = new aec_net();
    n->device = device;
    n->S = n_streams;
    n->H = height;
    n->W = width;
    n->leak = leak;
    if (max_events_per_step > 0) n->max_events = max_events_per_step;
    int slots = 64;
    while (slots < 2 * n->max_events) slots <<= 1;
    n->hash_slots = slots;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    n->num_sms = prop.multiProcessorCount;
    HostLayer l0;
    l0.type = AEC_LAYER_INTEGRATION;
    l0.C = 1;
    l0.H = height;
    l0.W = width;
    l0.Ww = (width + 31) / 32;
    n->L.push_back(l0);
    *out = n;
    return AEC_OK;
}

static int bits_for(int n)          // bits needed for values 0 .. n-1 (at least 1)
{
    int b = 1;
    while ((1 << b) < n) ++b;
    return b;
}

// Work-list entries are 32-bit: stream << sh_s | y << sh_y | x must fit.
static int set_site_code(const aec_net *n, HostLayer &l)
{
    l.code.sh_y = bits_for(l.W);
    l.code.sh_s = l.code.sh_y + bits_for(l.H);
    if (l.code.sh_s >= 32 || ((unsigned long long)n->S << l.code.sh_s) > (1ULL << 32))
        return fail(AEC_EINVAL, "%d streams of %dx%d sites do not fit the 32-bit work-list entries (stream << %d | y << %d | x)", n->S, l.H, l.W,
                    l.code.sh_s, l.code.sh_y);
    return AEC_OK;
}

static void same_pad(int size, int k, int stride, int *before)
{
    int total = (size % stride == 0) ? (k - stride > 0 ? k - stride : 0) : (k - size % stride > 0 ? k - size % stride : 0);
    *before = total / 2;
}

// The gathered GEMM runs on the tensor cores (aec_tc.cuh) when the previous layer is a conv or pool
// layer with a multiple of 4 channels (16-byte gathers); the first conv (Cin = 1, K = kh*kw, read
// from the float64 surface) is a 9-term stencil and stays on the SIMT kernel.  AEC_CONV_PATH=simt
// forces the SIMT kernel everywhere (A/B measurements only).
static void split_tf32_host(float x, float *hi, float *lo)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(hi, &u, 4);
    *lo = x - *hi;
}

static void build_tc_image(HostLayer &l, bool prev_is_map, const float *kernel_hwio)
{
    const char *force = getenv("AEC_CONV_PATH");
    l.tc = prev_is_map && (l.Cin % 4 == 0) && l.kh * l.kw <= 32 && !(force && strcmp(force, "simt") == 0);
    if (!l.tc) return;
    const int c8 = (l.C + 7) / 8 * 8;
    l.m_tiles = (c8 + 127) / 128;
    l.Mch = ((c8 + l.m_tiles - 1) / l.m_tiles + 7) / 8 * 8;        // output channels per weight tile (<= 128)
    l.rep = 1;                                                      // copies of the channels along M (narrow layers: parallel epilogue)
    if (l.m_tiles == 1 && l.Mch == 32) l.rep = 4;                   // measured: pays for 32 channels (one live epilogue warp otherwise); for 64
                                                                    // the larger weight stages (2 instead of 4 in flight) cost more than they give
    // Narrow layers take the sites-as-M form (AEC_TC_SM=0 keeps the weights-as-M form for A/B runs): the weight tile is
    // [W_hi ; W_lo] with Cpad = channels rounded up to 16 rows each (MMA N must be a multiple of 16), no row copies.
    const char *sm_env = getenv("AEC_TC_SM");
    l.tc_sm = l.C <= 64 && l.C % 4 == 0 && !(sm_env && atoi(sm_env) == 0);
    if (l.tc_sm) { l.Mch = (l.C + 15) / 16 * 16; l.rep = 1; }
    l.Mrows = l.Mch * l.rep;
    // Weight tiles per unit.  2 = one gathered site stage feeds two weight tiles (half the gather work per output), but a
    // single-buffered accumulator and units twice as long; 1 (default) = every (site block, weight tile) is its own unit:
    // measured on conv5 (Cout 256, 896 long units over 148 CTAs = 7 rounds for 6.05 units of work) 0.58 -> 0.54 ms.
    const char *mtu_env = getenv("AEC_TC_MTU");
    l.mtu = std::min(l.m_tiles, mtu_env ? std::max(1, std::min(atoi(mtu_env), tc::kMaxMtu)) : 1);
    l.n_acc = l.mtu == 1 ? 2 : 1;
    l.KB = (l.K + tc::kBlockK - 1) / tc::kBlockK;
    const size_t w_stage = 2 * (size_t)l.Mrows * 128, x_bytes = (size_t)tc::kSiteStages * tc::kSiteStageBytes;
    const size_t budget = 210 * 1024;        // 227 KB per CTA minus ~15 KB static shared memory and the 1 KB alignment slack
    l.w_stages = (int)std::min<size_t>(tc::kMaxWStages, std::max<size_t>(2, (budget - x_bytes) / w_stage));
    l.tc_smem = x_bytes + (size_t)l.w_stages * w_stage + 1024;
    if ((size_t)l.h_b.size() < (size_t)l.Mch * l.m_tiles) l.h_b.resize((size_t)l.Mch * l.m_tiles, 0.f);
    // image: [weight tile][K block][hi | lo][row = channel within the tile][32 floats, 16-byte chunks XOR-swizzled by row & 7]
    l.h_wimg.assign((size_t)l.m_tiles * l.KB * 2 * l.Mrows * tc::kBlockK, 0.f);
    for (int mt = 0; mt < l.m_tiles; ++mt)
        for (int kb = 0; kb < l.KB; ++kb)
            for (int r = 0; r < l.Mrows; ++r)
                for (int j = 0; j < 8; ++j)
                    for (int e = 0; e < 4; ++e) {
                        const int k = kb * tc::kBlockK + 4 * j + e, col = mt * l.Mch + r % l.Mch;
                        const float w = (k < l.K && col < l.C) ? kernel_hwio[(size_t)k * l.C + col] : 0.f;
                        float hi, lo;
                        split_tf32_host(w, &hi, &lo);
                        const size_t base = ((size_t)(mt * l.KB + kb) * 2) * l.Mrows * tc::kBlockK;
                        const size_t off = (size_t)r * tc::kBlockK + (size_t)((j ^ (r & 7)) * 4) + e;
                        l.h_wimg[base + off] = hi;
                        l.h_wimg[base + (size_t)l.Mrows * tc::kBlockK + off] = lo;
                    }
}

// Row-tile form (aec_rt.cuh): for a sites-as-M layer with a real window (kh*kw > 1) whose input pixel splits into 64- or
// 128-byte channel blocks.  AEC_RT=0 keeps the gathered kernel (A/B runs).
static void build_rt_image(HostLayer &l, const float *kernel_hwio)
{
    const char *e = getenv("AEC_RT");
    if (!l.tc || !l.tc_sm || l.kh * l.kw <= 1 || l.kw > 8 || (e && atoi(e) == 0)) return;
    if (!(l.Cin == 16 || l.Cin % 32 == 0)) return;
    l.rt_CB = l.Cin == 16 ? 16 : 32;
    l.rt_ncb = l.Cin / l.rt_CB;
    const int row_bytes = l.rt_CB * 4;
    const int need = l.W + l.kw - 1;                 // slot width a whole output row needs (its input columns + the window's halo)
    if (need <= 64) {
        int sw = 16;
        l.rt_sw_shift = 4;
        while (sw < need) { sw <<= 1; ++l.rt_sw_shift; }
        l.rt_R = 128 / sw;
        l.rt_SEG = l.W;
        l.rt_nxg = 1;
    } else {
        l.rt_sw_shift = 7;
        l.rt_R = 1;
        l.rt_SEG = 128 - (l.kw - 1);
        l.rt_nxg = (l.W + l.rt_SEG - 1) / l.rt_SEG;
    }
    const int nyg = (l.H + l.rt_R - 1) / l.rt_R;
    if ((long long)nyg * l.rt_nxg > 32LL * kThreads) return;          // emit_units: at most 32 units per thread of the frontier CTA
    l.rt_P = 128 + l.kw - 1;
    const int n_pairs = l.rt_P * (l.rt_CB / 4);
    l.rt_groups = (n_pairs + 127) / 128 <= rt::kRtMaxPairs ? 3 : 2;      // producer groups of 4 or 6 warps (aec_rt.cuh)
    if ((n_pairs + (12 / l.rt_groups) * 32 - 1) / ((12 / l.rt_groups) * 32) > rt::kRtMaxPairs) return;
    const int cpad = l.Mrows;                                          // channels rounded up to 16 (sites-as-M)
    l.rt_xtile = ((size_t)l.rt_P * row_bytes + 1023) / 1024 * 1024;
    l.rt_wtile = (size_t)2 * cpad * row_bytes;
    const size_t budget = 208 * 1024;                // 227 KB per CTA minus the alignment slack and the kernel's ~18 KB of static shared memory
    const size_t all_w = (size_t)l.kh * l.kw * l.rt_ncb * l.rt_wtile;
    l.rt_wres = all_w <= 48 * 1024 && all_w + 2 * 4 * l.rt_xtile <= budget;   // resident weights (aec_rt.cuh): no streaming latency
    if (l.rt_wres) {
        l.rt_wst = l.kh * l.kw * l.rt_ncb;
        l.rt_xst = all_w + 3 * 4 * l.rt_xtile <= budget ? 3 : 2;
        // a fourth site stage and producer group where they fit (64-byte rows with resident weights: EFCN conv2)
        const char *g4 = getenv("AEC_RT_GROUPS4");
        if (!(g4 && atoi(g4) == 0) && l.rt_groups == 3 && all_w + 4 * 4 * l.rt_xtile <= budget && (n_pairs + 95) / 96 <= rt::kRtMaxPairs) {
            l.rt_xst = 4;
            l.rt_groups = 4;
        }
    } else {
        l.rt_xst = 4 * l.rt_xtile * 3 + (size_t)(l.kw + 2) * l.rt_wtile <= budget ? 3 : 2;
        l.rt_wst = (int)std::min<size_t>(rt::kRtMaxWStages, (budget - (size_t)l.rt_xst * 4 * l.rt_xtile) / l.rt_wtile);
        if (l.rt_wst < l.kw + 1) return;             // the kw tiles of a kernel row plus at least one prefetched tile of the next
    }
    if (l.rt_xst < l.rt_groups) l.rt_groups = l.rt_xst;              // a group per site stage at most
    if ((n_pairs + (12 / l.rt_groups) * 32 - 1) / ((12 / l.rt_groups) * 32) > rt::kRtMaxPairs) return;
    l.rt_smem = (size_t)l.rt_xst * 4 * l.rt_xtile + (size_t)l.rt_wst * l.rt_wtile + 1024;
    l.rt = true;
    // image: [(ky*kw + kx)*ncb + cb][row: W_hi of channel r (r < Cpad), then W_lo][CB floats], 16-byte chunks swizzled on the
    // address bits like the site tiles (tiles are multiples of 1024 bytes)
    const size_t tile_floats = l.rt_wtile / 4;
    l.h_rtimg.assign((size_t)l.kh * l.kw * l.rt_ncb * tile_floats, 0.f);
    for (int ky = 0; ky < l.kh; ++ky)
        for (int kx = 0; kx < l.kw; ++kx)
            for (int cb = 0; cb < l.rt_ncb; ++cb) {
                float *tile = l.h_rtimg.data() + (size_t)((ky * l.kw + kx) * l.rt_ncb + cb) * tile_floats;
                for (int r = 0; r < 2 * cpad; ++r)
                    for (int j = 0; j < l.rt_CB; ++j) {
                        const int col = r % cpad;
                        const int k = (ky * l.kw + kx) * l.Cin + cb * l.rt_CB + j;
                        const float w = col < l.C ? kernel_hwio[(size_t)k * l.C + col] : 0.f;
                        float hi, lo;
                        split_tf32_host(w, &hi, &lo);
                        uint32_t lin = (uint32_t)r * row_bytes + (uint32_t)j * 4;
                        lin ^= ((lin >> 7) & (row_bytes == 128 ? 7u : 3u)) << 4;
                        tile[lin / 4] = r < cpad ? hi : lo;
                    }
            }
}

extern "C" int aec_net_add_conv(aec_net *n, int k_h, int k_w, int c_in, int c_out, const float *kernel_hwio,
                                const float *bias, int stride, float alpha, int padding)
{
    if (!n || n->finalized) return fail(AEC_ESTATE, "add_conv after finalize (or NULL net)");
    if (stride != 1) return fail(AEC_EINVAL, "only stride 1 convolutions are supported (event_numpy.py:64 always passes 1)");
    if (k_h < 1 || k_w < 1 || k_h > 31 || k_w > 31) return fail(AEC_EINVAL, "kernel size %dx%d unsupported", k_h, k_w);
    if (!kernel_hwio || !bias) return fail(AEC_EINVAL, "kernel/bias is NULL");
    if ((int)n->L.size() >= 31) return fail(AEC_EINVAL, "too many layers");
    const HostLayer &p = n->L.back();
    if (c_in != p.C) return fail(AEC_EINVAL, "conv expects %d input channels but previous layer has %d", c_in, p.C);
    int nconv = 0;
    for (auto &l : n->L) nconv += l.type == AEC_LAYER_CONV;
    if (nconv >= kMaxConv) return fail(AEC_EINVAL, "too many conv layers");
    HostLayer l;
    l.type = AEC_LAYER_CONV;
    l.Cin = p.C; l.Hin = p.H; l.Win = p.W;
    l.kh = k_h; l.kw = k_w; l.stride = 1; l.alpha = alpha;
    if (padding == AEC_PAD_SAME) {            // conv2d.py:38-54
        l.H = p.H; l.W = p.W;
        same_pad(p.H, k_h, 1, &l.pad_t);
        same_pad(p.W, k_w, 1, &l.pad_l);
    } else if (padding == AEC_PAD_VALID) {    // conv2d.py:34-37
        l.H = p.H - k_h + 1; l.W = p.W - k_w + 1;
        if (l.H < 1 || l.W < 1) return fail(AEC_EINVAL, "VALID conv larger than its input");
    } else {
        return fail(AEC_EINVAL, "'padding' must be either 'SAME' or 'VALID'");
    }
    l.C = c_out;
    l.Ww = (l.W + 31) / 32;
    l.K = k_h * k_w * c_in;
    l.BN = c_out <= 16 ? 16 : c_out <= 32 ? 32 : c_out <= 64 ? 64 : 128;
    l.Npad = (c_out + l.BN - 1) / l.BN * l.BN;
    l.Kpad = (l.K + 15) / 16 * 16;
    l.fstride = pad4((long long)l.H * l.W * l.C);
    l.h_w.assign((size_t)l.Kpad * l.Npad, 0.f);
    l.h_b.assign((size_t)l.Npad, 0.f);
    for (int k = 0; k < l.K; ++k)          // HWIO flattened is already [k = (ky,kx,ci)][co]
        for (int c = 0; c < c_out; ++c) l.h_w[(size_t)k * l.Npad + c] = kernel_hwio[(size_t)k * c_out + c];
    for (int c = 0; c < c_out; ++c) l.h_b[c] = bias[c];
    { int rc = set_site_code(n, l); if (rc) return rc; }
    build_tc_image(l, p.type != AEC_LAYER_INTEGRATION, kernel_hwio);
    build_rt_image(l, kernel_hwio);
    n->L.push_back(std::move(l));
    return (int)n->L.size() - 1;
}

extern "C" int aec_net_add_pool(aec_net *n, int k_h, int k_w, int stride)
{
    if (!n || n->finalized) return fail(AEC_ESTATE, "add_pool after finalize (or NULL net)");
    if ((int)n->L.size() >= 31) return fail(AEC_EINVAL, "too many layers");
    const HostLayer &p = n->L.back();
    if (p.type != AEC_LAYER_CONV) return fail(AEC_EINVAL, "a pool layer must follow a conv layer");
    if (!(stride == k_h && stride == k_w))   // cutils.pyx:88-89
        return fail(AEC_EINVAL, "This method only support stride equal to 1 or to the kernel's dimensions.");
    if (k_h * k_w > 255) return fail(AEC_EINVAL, "pool window too large");
    if (p.H % stride || p.W % stride)
        return fail(AEC_EINVAL, "pool input %dx%d is not a multiple of the stride %d (the reference indexes out of range, SURVEY Q6)",
                    p.H, p.W, stride);
    HostLayer l;
    l.type = AEC_LAYER_POOL;
    l.Cin = p.C; l.Hin = p.H; l.Win = p.W;
    l.kh = k_h; l.kw = k_w; l.stride = stride;
    l.C = p.C;
    l.H = (p.H - k_h) / stride + 1;    // maxpool.py:27-28
    l.W = (p.W - k_w) / stride + 1;
    l.Ww = (l.W + 31) / 32;
    l.fstride = pad4((long long)l.H * l.W * l.C);
    { int rc = set_site_code(n, l); if (rc) return rc; }
    n->L.push_back(std::move(l));
    return (int)n->L.size() - 1;
}

// ------------------------------------------------------------------------------------------------
static Src make_src(const aec_net *n, int li)
{
    const HostLayer &l = n->L[li];
    Src q;
    memset(&q, 0, sizeof q);
    q.C = l.C; q.H = l.H; q.W = l.W;
    if (l.type == AEC_LAYER_INTEGRATION) {
        q.kind = 0;
        q.S = n->surface;
        q.sstride = (long long)l.H * l.W;
    } else if (l.type == AEC_LAYER_CONV) {
        q.kind = 1;
        q.F = l.F; q.A = l.A; q.fstride = l.fstride; q.alpha = l.alpha;
    } else {
        q.kind = 1;
        q.F = l.Fp; q.A = l.Ap; q.fstride = l.fstride; q.alpha = n->L[li - 1].alpha;
    }
    return q;
}

// Every kernel of the step goes through here: with n->pdl the launch carries the programmatic-serialization attribute
// (the kernels start with pdl_enter(), aec_kernels.cuh), `cluster` > 1 launches thread-block clusters.  Errors are picked
// up by launch_check (cudaGetLastError).
template <typename P>
static void launch_k(const aec_net *n, void (*fn)(P), dim3 grid, unsigned block, size_t smem, cudaStream_t st, const P &p, int cluster = 1)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned na = 0;
    if (n->pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = (unsigned)cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    (void)cudaLaunchKernelEx(&cfg, fn, p);
}

static int launch_check(aec_net *n, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AEC_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    n->launches++;
    return AEC_OK;
}

// Profiling: an event is recorded on the launching stream after every launch of a step, so the
// elapsed time between consecutive events is that kernel's duration inside the real step.
static int prof_mark(aec_net *n, cudaStream_t st, const char *done = nullptr, int layer = -1)
{
    if (!n->profiling) return AEC_OK;
    if (done && n->prof_slot >= 1 && (int)n->prof_names.size() < n->prof_slot) {     // name of the launch that ends at this mark
        char nm[64];
        if (layer >= 0) snprintf(nm, sizeof nm, "L%d.%s", layer, done);
        else snprintf(nm, sizeof nm, "%s", done);
        n->prof_names.resize(n->prof_slot - 1);
        n->prof_names.push_back(nm);
    }
    if (n->prof_slot >= (int)n->prof_events.size()) {
        cudaEvent_t ev;
        CU(cudaEventCreate(&ev));
        n->prof_events.push_back(ev);
    }
    CU(cudaEventRecord(n->prof_events[n->prof_slot++], st));
    return AEC_OK;
}

static int prof_collect(aec_net *n, cudaStream_t st)
{
    if (!n->profiling || n->prof_slot < 2) return AEC_OK;
    CU(cudaStreamSynchronize(st));
    if ((int)n->prof_ms.size() < n->prof_slot - 1) n->prof_ms.resize(n->prof_slot - 1, 0.0);
    for (int i = 0; i + 1 < n->prof_slot; ++i) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, n->prof_events[i], n->prof_events[i + 1]));
        n->prof_ms[i] += ms;
    }
    n->prof_steps++;
    n->prof_slot = 0;
    return AEC_OK;
}

static IntegrateParams integrate_params(aec_net *n, const int32_t *ev, const int32_t *off)
{
    const HostLayer &l = n->L[0];
    IntegrateParams p;
    p.surface = n->surface; p.prev_ts = n->prev_ts; p.delta = n->delta; p.active = n->active;
    p.front = l.front; p.alive = l.nzr; p.events = ev; p.offsets = off; p.ends = n->ev_ends; p.layer_counts = n->counts; p.err_flag = n->err_flag;
    p.n_layers = (int)n->L.size();
    p.H = l.H; p.W = l.W; p.Ww = l.Ww; p.leak = n->leak; p.max_events = n->max_events; p.hash_slots = n->hash_slots;
    return p;
}

static int run_integrate(aec_net *n, const int32_t *ev, const int32_t *off, cudaStream_t st)
{
    const HostLayer &l = n->L[0];
    const IntegrateParams p = integrate_params(n, ev, off);
    const size_t smem = (size_t)n->hash_slots * 8 + (size_t)l.H * l.Ww * 8;
    int rc;
    if ((rc = prof_mark(n, st))) return rc;
    launch_k(n, k_integrate, n->S, kThreads, smem, st, p);
    if ((rc = launch_check(n, "k_integrate"))) return rc;
    return prof_mark(n, st, "surface");
}

// Fills one leak-sweep table entry; returns the number of chunks (grid.x slots) the layer takes.
static int fill_sweep_layer(const HostLayer &l, SweepLayer &o, int chunk0, bool dense_only = false, bool use_skip = false)
{
    if (l.type == AEC_LAYER_POOL) { o.F = l.Fp; o.A = l.Ap; o.signchg = nullptr; }
    else { o.F = l.F; o.A = l.A; o.signchg = l.signchg; }
    o.fstride = l.fstride;
    o.n4 = (int)(l.fstride / 4); o.chunk0 = chunk0;
    o.C = l.C; o.W = l.W; o.Ww = l.Ww; o.HWw = l.H * l.Ww;
    o.nzr = (!dense_only && l.C % 4 == 0) ? l.nzr : nullptr;     // a float4 of the sweep must not straddle sites
    o.skip = use_skip ? l.skip : nullptr;
    o.c4 = l.C / 4;
    o.c4_shift = -1;
    o.wpc = 1;
    if (!o.nzr) return (o.n4 + kSweepChunk - 1) / kSweepChunk;
    for (int b = 0; b < 30; ++b)
        if ((1 << b) == o.c4) o.c4_shift = b;
    o.wpc = std::max(1, std::min(kSweepMaxWords, kSweepUnitsPerChunk / (32 * o.c4)));
    return (o.HWw + o.wpc - 1) / o.wpc;
}

static int run_sweep(aec_net *n, int only_layer, cudaStream_t st)
{
    if (only_layer < 0) {
        int rc = AEC_OK;
        if (n->sweep_win_chunks > 0) {
            dim3 gridw(n->sweep_win_chunks, n->S);
            launch_k(n, k_sweep_windows, gridw, kThreads, 0, st, n->sweep_win);
            if ((rc = launch_check(n, "k_sweep_windows"))) return rc;
            if ((rc = prof_mark(n, st, "window_sweep"))) return rc;
        }
        if (n->sweep_all.n_layers == 0) return AEC_OK;
        dim3 grid(n->sweep_chunks, n->S);
        launch_k(n, k_leak_sweep, grid, kThreads, 0, st, n->sweep_all);
        rc = launch_check(n, "k_leak_sweep");
        return rc ? rc : prof_mark(n, st, "leak_sweep");
    }
    SweepParams p;
    memset(&p, 0, sizeof p);
    const int chunks = fill_sweep_layer(n->L[only_layer], p.L[0], 0);
    p.n_layers = 1; p.delta = n->delta; p.active = n->active;
    dim3 grid(chunks, n->S);
    launch_k(n, k_leak_sweep, grid, kThreads, 0, st, p);
    return launch_check(n, "k_leak_sweep");
}

static int run_conv_eval_tc(aec_net *n, int li, cudaStream_t st, bool fused_step)
{
    HostLayer &l = n->L[li];
    const Src src = make_src(n, li - 1);
    tc::TcParams p;
    p.sites = l.sites; p.counter = n->counts + li; p.accum = n->accum + li;
    p.srcF = src.F; p.a_minus_f = (const char *)src.A - (const char *)src.F; p.zero_f = (const char *)src.F - kMapGuardFloats * 4;
    p.src_stride = src.fstride; p.alpha = src.alpha;
    p.Cin = src.C; p.Hin = src.H; p.Win = src.W;
    p.wimg = l.wimg; p.bias = l.bias; p.F = l.F; p.A = l.A; p.fstride = l.fstride;
    p.C = l.C; p.H = l.H; p.W = l.W; p.K = l.K; p.KB = l.KB; p.ks_last = (l.K - tc::kBlockK * (l.KB - 1) + 7) / 8; p.Mrows = l.Mrows; p.Mch = l.Mch; p.rep = l.rep; p.m_tiles = l.m_tiles; p.mtu = l.mtu;
    p.kh = l.kh; p.kw = l.kw; p.pad_t = l.pad_t; p.pad_l = l.pad_l; p.code = l.code;
    p.w_stages = l.w_stages; p.n_acc = l.n_acc;
    {
        // half units: 96 KB of site stages; what the launch's shared memory holds beyond them is weight stages
        const size_t w_stage = 2 * (size_t)l.Mrows * 128, x_half = (size_t)tc::kPairSiteStages * 2 * tc::kItemTileBytes;
        p.w_stages_half = (int)std::min<size_t>(tc::kMaxWStages, (l.tc_smem - 1024 - x_half) / w_stage);
    }
    p.half_units = n->tc_half ? 1 : 0;
    p.quad_bit = 0u; p.site_counter = nullptr; p.pool_idx = nullptr; p.pool_Fp = p.pool_Ap = nullptr; p.pool_stride = 0; p.pool_flags = nullptr;
    p.pool_accum = nullptr; p.pW = p.pWw = p.pHWw = 0; p.pool_alpha = 1.f;

    p.debug = n->tc_debug;
    p.timing = n->tc_timing_on ? l.tc_timing : nullptr;
    // the work list holds quads only when the fused step's frontier kernel wrote it (the layer-at-a-time interface and the
    // dense initial forward emit plain lists and evaluate every pool window with k_pool_eval)
    const bool pool = l.pool_fuse && fused_step;
    if (pool) {
        const HostLayer &pl = n->L[li + 1];
        p.quad_bit = 0x80000000u; p.site_counter = n->counts + 32 + li;
        p.pool_idx = pl.idx; p.pool_Fp = pl.Fp; p.pool_Ap = pl.Ap; p.pool_stride = pl.fstride; p.pool_flags = pl.flags;
        p.pool_accum = n->accum + (li + 1); p.pW = pl.W; p.pWw = pl.Ww; p.pHWw = pl.H * pl.Ww; p.pool_alpha = l.alpha;
        if (l.tc_pair) launch_k(n, tc::k_conv_eval_tc<true, false, true, true>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p, 2);
        else launch_k(n, tc::k_conv_eval_tc<true, false, true>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p);
    } else
    if (l.tc_pair) launch_k(n, tc::k_conv_eval_tc<true, false, false, true>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p, 2);
    else if (l.tc_sm) launch_k(n, tc::k_conv_eval_tc<true, true>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p);
    else if (l.tc_fast_decode) launch_k(n, tc::k_conv_eval_tc<true, false>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p);
    else launch_k(n, tc::k_conv_eval_tc<false, false>, l.tc_blocks, tc::kTcThreads, l.tc_smem, st, p);
    int rc = launch_check(n, "k_conv_eval_tc");
    return rc ? rc : prof_mark(n, st, "eval", li);
}

// Row-tile evaluation: only inside the fused step, whose frontier kernel wrote the unit list and the work-set bitmap.
static int run_conv_rows(aec_net *n, int li, cudaStream_t st)
{
    HostLayer &l = n->L[li];
    const Src src = make_src(n, li - 1);
    rt::RtParams p;
    memset(&p, 0, sizeof p);
    p.units = l.sites; p.counter = n->counts + li; p.site_counter = n->counts + 32 + li;
    p.accum_sites = n->accum + li; p.accum_units = n->accum + 32 + li;
    p.srcF = src.F; p.a_minus_f = (const char *)src.A - (const char *)src.F; p.zero_f = (const char *)src.F - kMapGuardFloats * 4;
    p.src_stride = src.fstride; p.alpha = src.alpha;
    p.Cin = src.C; p.Hin = src.H; p.Win = src.W;
    p.wimg = l.rtimg; p.bias = l.bias; p.F = l.F; p.A = l.A; p.fstride = l.fstride; p.nset = l.nset;
    p.C = l.C; p.H = l.H; p.W = l.W; p.Ww = l.Ww; p.Cpad = l.Mrows;
    p.kh = l.kh; p.kw = l.kw; p.pad_t = l.pad_t; p.pad_l = l.pad_l;
    p.CB = l.rt_CB; p.ncb = l.rt_ncb; p.row_bytes = l.rt_CB * 4;
    p.R = l.rt_R; p.sw_shift = l.rt_sw_shift; p.SEG = l.rt_SEG; p.code = l.code; p.P = l.rt_P;
    {
        p.store32 = (n->rt_store32 && l.C % 8 == 0 && l.fstride % 8 == 0 && ((uintptr_t)l.F % 32) == 0 && ((uintptr_t)l.A % 32) == 0) ? 1 : 0;
    }
    p.x_tile_bytes = (uint32_t)l.rt_xtile; p.w_tile_bytes = (uint32_t)l.rt_wtile; p.x_stages = l.rt_xst; p.w_stages = l.rt_wst; p.w_resident = l.rt_wres ? 1 : 0;
    p.debug = n->tc_debug;
    p.prod_groups = l.rt_groups;
    p.timing = n->tc_timing_on ? l.tc_timing : nullptr;
    static const int staged_min = getenv("AEC_RT_STAGED_MIN_C") ? atoi(getenv("AEC_RT_STAGED_MIN_C")) : 33;
    const bool staged = l.C >= staged_min;   // epilogue stores through the per-warp transpose buffer (aec_rt.cuh)
    if (l.rt_CB == 32 && staged) launch_k(n, rt::k_conv_rows<32, true>, n->num_sms, rt::kRtThreads, l.rt_smem, st, p);
    else if (l.rt_CB == 32) launch_k(n, rt::k_conv_rows<32, false>, n->num_sms, rt::kRtThreads, l.rt_smem, st, p);
    else if (staged) launch_k(n, rt::k_conv_rows<16, true>, n->num_sms, rt::kRtThreads, l.rt_smem, st, p);
    else launch_k(n, rt::k_conv_rows<16, false>, n->num_sms, rt::kRtThreads, l.rt_smem, st, p);
    int rc = launch_check(n, "k_conv_rows");
    return rc ? rc : prof_mark(n, st, "eval", li);
}

static int run_conv_eval(aec_net *n, int li, cudaStream_t st, bool fused_step = false)
{
    HostLayer &l = n->L[li];
    if (l.rt && fused_step) return run_conv_rows(n, li, st);
    if (l.tc) return run_conv_eval_tc(n, li, st, fused_step);
    static const bool no_stencil = getenv("AEC_CONV_PATH") && strcmp(getenv("AEC_CONV_PATH"), "simt") == 0;
    if (n->L[li - 1].type == AEC_LAYER_INTEGRATION && l.C % 4 == 0 && l.C <= kStencilMaxC && l.kh * l.kw <= kStencilMaxK && !no_stencil) {
        StencilParams p;
        p.sites = l.sites; p.counter = n->counts + li; p.accum = n->accum + li;
        p.S = n->surface; p.sstride = (long long)n->L[0].H * n->L[0].W; p.Hin = n->L[0].H; p.Win = n->L[0].W;
        p.wgt = l.wgt; p.bias = l.bias; p.Npad = l.Npad; p.F = l.F; p.A = l.A; p.fstride = l.fstride;
        p.C = l.C; p.H = l.H; p.W = l.W; p.kh = l.kh; p.kw = l.kw; p.pad_t = l.pad_t; p.pad_l = l.pad_l; p.code = l.code;
        if (l.kh == 3 && l.kw == 3) launch_k(n, k_conv_stencil<3, 3>, n->num_sms * 8, kThreads, 0, st, p);
        else launch_k(n, k_conv_stencil<0, 0>, n->num_sms * 8, kThreads, 0, st, p);
        int rc = launch_check(n, "k_conv_stencil");
        return rc ? rc : prof_mark(n, st, "eval", li);
    }
    ConvEvalParams p;
    p.sites = l.sites; p.counter = n->counts + li; p.accum = n->accum + li;
    p.src = make_src(n, li - 1);
    p.wgt = l.wgt; p.bias = l.bias; p.F = l.F; p.A = l.A; p.fstride = l.fstride;
    p.C = l.C; p.H = l.H; p.W = l.W; p.K = l.K; p.Kpad = l.Kpad; p.Npad = l.Npad;
    p.kh = l.kh; p.kw = l.kw; p.pad_t = l.pad_t; p.pad_l = l.pad_l; p.code = l.code;
    switch (l.BN) {
    case 16: launch_k(n, k_conv_eval<16, 2, 4, 16>, n->conv_eval_blocks[0], kThreads, 0, st, p); break;
    case 32: launch_k(n, k_conv_eval<32, 4, 4, 16>, n->conv_eval_blocks[1], kThreads, 0, st, p); break;
    case 64: launch_k(n, k_conv_eval<64, 4, 8, 16>, n->conv_eval_blocks[2], kThreads, 0, st, p); break;
    default: launch_k(n, k_conv_eval<128, 8, 8, 16>, n->conv_eval_blocks[3], kThreads, 0, st, p); break;
    }
    int rc = launch_check(n, "k_conv_eval");
    return rc ? rc : prof_mark(n, st, "eval", li);
}

static int run_pool_eval(aec_net *n, int li, cudaStream_t st)
{
    HostLayer &l = n->L[li];
    const HostLayer &c = n->L[li - 1];
    PoolEvalParams p;
    p.sites = l.sites; p.counter = n->counts + li; p.accum = n->accum + li;
    p.F = c.F; p.A = c.A; p.fstride = c.fstride; p.alpha = c.alpha; p.cW = c.W;
    p.idx = l.idx; p.Fp = l.Fp; p.Ap = l.Ap; p.pstride = l.fstride; p.flags = l.flags;
    p.C = l.C; p.H = l.H; p.W = l.W; p.Ww = l.Ww; p.kh = l.kh; p.kw = l.kw; p.stride = l.stride; p.code = l.code;
    if (l.C % 4 == 0 && l.kh == 2 && l.kw == 2 && l.stride == 2) launch_k(n, k_pool_eval<4, true>, n->num_sms * 8, kThreads, 0, st, p);
    else if (l.C % 4 == 0) launch_k(n, k_pool_eval<4, false>, n->num_sms * 8, kThreads, 0, st, p);
    else launch_k(n, k_pool_eval<1, false>, n->num_sms * 8, kThreads, 0, st, p);
    int rc = launch_check(n, "k_pool_eval");
    return rc ? rc : prof_mark(n, st, "eval", li);
}

static int run_layer(aec_net *n, int li, bool with_sweep, cudaStream_t st)
{
    HostLayer &l = n->L[li];
    const HostLayer &pv = n->L[li - 1];
    int rc;
    if (with_sweep) CU(cudaMemsetAsync(n->counts + li, 0, sizeof(int), st));   // layer-at-a-time: the call may be repeated within a step
    if (l.type == AEC_LAYER_CONV) {
        if (with_sweep && (rc = run_sweep(n, li, st))) return rc;
        ConvFrontParams p;
        p.prev_front = pv.front; p.front = l.front; p.signchg = l.signchg; p.nzr = l.nzr; p.prev_nzr = pv.nzr; p.active = n->active;
        p.sites = l.sites; p.counter = n->counts + li;
        p.Hin = pv.H; p.Win = pv.W; p.WwIn = pv.Ww; p.H = l.H; p.W = l.W; p.Ww = l.Ww;
        p.kh = l.kh; p.kw = l.kw; p.pad_t = l.pad_t; p.pad_l = l.pad_l; p.code = l.code;
        const size_t smem = ((size_t)pv.H * pv.Ww + (size_t)pv.H * l.Ww + (size_t)l.H * l.Ww) * 4;
        launch_k(n, k_conv_frontier, n->S, kThreads, smem, st, p);
        if ((rc = launch_check(n, "k_conv_frontier"))) return rc;
        if ((rc = prof_mark(n, st, "frontier", li))) return rc;
        return run_conv_eval(n, li, st);
    }
    if (with_sweep && (rc = run_sweep(n, li, st))) return rc;     // the (Fp, Ap) copy leaks before it is refreshed
    PoolFrontParams p;
    p.prev_front = pv.front; p.front = l.front; p.flags = l.flags; p.nzr = l.nzr; p.prev_nzr = pv.nzr; p.active = n->active;
    p.sites = l.sites; p.counter = n->counts + li;
    p.Hin = pv.H; p.Win = pv.W; p.WwIn = pv.Ww; p.H = l.H; p.W = l.W; p.Ww = l.Ww;
    p.kh = l.kh; p.kw = l.kw; p.stride = l.stride; p.code = l.code;
    const size_t smem = ((size_t)pv.H * pv.Ww + (size_t)l.H * l.Ww) * 4;
    launch_k(n, k_pool_frontier, n->S, kThreads, smem, st, p);
    if ((rc = launch_check(n, "k_pool_frontier"))) return rc;
    if ((rc = prof_mark(n, st, "frontier", li))) return rc;
    return run_pool_eval(n, li, st);
}

static int run_eval(aec_net *n, int li, cudaStream_t st)      // evaluation of layer li inside the fused step
{
    return n->L[li].type == AEC_LAYER_CONV ? run_conv_eval(n, li, st, true) : run_pool_eval(n, li, st);
}

static int run_frontier_skip(aec_net *n, cudaStream_t st)
{
    if (!n->sweep_skip) return AEC_OK;
    FrontAllParams p;
    p.layers = n->front_table; p.n_layers = (int)n->L.size();
    p.front0 = n->L[0].front; p.nzr0 = n->L[0].nzr; p.words0 = n->L[0].H * n->L[0].Ww;
    p.max_words = n->front_max_words; p.active = n->active;
    launch_k(n, k_frontier_skip, n->S, kThreads, (size_t)3 * n->front_max_words * 4, st, p);
    int rc = launch_check(n, "k_frontier_skip");
    return rc ? rc : prof_mark(n, st, "skip.frontier");
}

static int run_frontier_all(aec_net *n, cudaStream_t st)
{
    FrontAllParams p;
    p.layers = n->front_table; p.n_layers = (int)n->L.size();
    p.front0 = n->L[0].front; p.nzr0 = n->L[0].nzr; p.words0 = n->L[0].H * n->L[0].Ww;
    p.max_words = n->front_max_words; p.active = n->active;
    // more streams than fit one wave at the kernel's natural register count (6 CTAs per SM): the 32-register build (8 per SM)
    if (n->S > 6 * n->num_sms) launch_k(n, k_frontier_all<8>, n->S, kThreads, (size_t)5 * n->front_max_words * 4, st, p);
    else if (n->S > 4 * n->num_sms) launch_k(n, k_frontier_all<6>, n->S, kThreads, (size_t)5 * n->front_max_words * 4, st, p);
    else launch_k(n, k_frontier_all<4, true>, n->S, kThreads, (size_t)5 * n->front_max_words * 4, st, p);      // few streams: bitmaps prefetched a layer ahead
    int rc = launch_check(n, "k_frontier_all");
    return rc ? rc : prof_mark(n, st, "all.frontier");
}

static HeadParams head_params(aec_net *n, float *out)
{
    HeadParams p;
    p.src = make_src(n, (int)n->L.size() - 1);
    p.out = out;
    p.S = n->S;
    return p;
}

static int run_head(aec_net *n, cudaStream_t st)
{
    const HeadParams p = head_params(n, n->head_cur ? n->head_cur : n->head);
    n->head_last = p.out;
    long long total = (long long)n->head_per_stream * n->S;
    int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, (long long)n->num_sms * 8);
    launch_k(n, k_head, blocks, kThreads, 0, st, p);
    int rc = launch_check(n, "k_head");
    return rc ? rc : prof_mark(n, st, "head");
}

static int broadcast(aec_net *n, void *dst, const void *src, long long bytes_per_stream, long long stride_bytes,
                     const uint8_t *mask, cudaStream_t st)
{
    const long long words = bytes_per_stream / 4;
    int bx = (int)std::min<long long>((words + kThreads - 1) / kThreads, 64);
    if (bx < 1) bx = 1;
    dim3 grid(bx, n->S);
    k_broadcast_words<<<grid, kThreads, 0, st>>>((uint32_t *)dst, (const uint32_t *)src, words, stride_bytes / 4, mask);
    return launch_check(n, "k_broadcast_words");
}

static int reset_streams(aec_net *n, const uint8_t *mask_dev, cudaStream_t st)
{
    int rc;
    const HostLayer &l0 = n->L[0];
    if ((rc = broadcast(n, n->surface, nullptr, (long long)l0.H * l0.W * 8, (long long)l0.H * l0.W * 8, mask_dev, st))) return rc;
    k_reset_scalars<<<(n->S + kThreads - 1) / kThreads, kThreads, 0, st>>>(n->prev_ts, n->delta, n->active, mask_dev, n->S);
    if ((rc = launch_check(n, "k_reset_scalars"))) return rc;
    for (auto &l : n->L) {
        const long long bm = (long long)l.H * l.Ww * 4;
        if ((rc = broadcast(n, l.front, nullptr, bm, bm, mask_dev, st))) return rc;
        if ((rc = broadcast(n, l.nzr, nullptr, bm, bm, mask_dev, st))) return rc;
        if (l.type == AEC_LAYER_CONV) {
            if ((rc = broadcast(n, l.F, l.initF, l.fstride * 4, l.fstride * 4, mask_dev, st))) return rc;
            if ((rc = broadcast(n, l.A, nullptr, l.fstride * 4, l.fstride * 4, mask_dev, st))) return rc;
            if ((rc = broadcast(n, l.signchg, nullptr, bm, bm, mask_dev, st))) return rc;
        } else if (l.type == AEC_LAYER_POOL) {
            if ((rc = broadcast(n, l.idx, l.initIdx, l.fstride, l.fstride, mask_dev, st))) return rc;
            if ((rc = broadcast(n, l.flags, nullptr, bm, bm, mask_dev, st))) return rc;
            if ((rc = broadcast(n, l.Fp, l.initFp, l.fstride * 4, l.fstride * 4, mask_dev, st))) return rc;
            if ((rc = broadcast(n, l.Ap, nullptr, l.fstride * 4, l.fstride * 4, mask_dev, st))) return rc;
        }
    }
    return AEC_OK;
}

extern "C" int aec_net_finalize(aec_net *n)
{
    if (!n || n->finalized) return fail(AEC_ESTATE, "finalize called twice (or NULL net)");
    if (n->L.size() < 2) return fail(AEC_EINVAL, "network has no layers after the integration surface");
    CU(cudaSetDevice(n->device));
    int rc;
    const size_t S = (size_t)n->S;
    size_t maxHW = 0;
    // The leak sweep skips what the step re-evaluates (k_frontier_skip runs the frontier chain once more, before the sweep).  That
    // pays when the sweep moves gigabytes; for a few streams the extra kernel's latency (~15 us) is all it adds: on from 32 streams
    // (one stream: 0.231 -> 0.214 ms per step without it; the results are bit-identical either way, test_sweep_skipping_changes_no_bit)
    { const char *e = getenv("AEC_SWEEP_SKIP"); n->sweep_skip = e ? atoi(e) != 0 : n->S >= 32; }
    {
        // a 2x2 pool directly behind the FIRST conv layer: the leak sweep evaluates its sticky windows (AEC_SWEEP_POOL=0: never)
        const char *e = getenv("AEC_SWEEP_POOL");
        if (n->sweep_skip && !(e && atoi(e) == 0) && n->L.size() > 2 && n->L[1].type == AEC_LAYER_CONV && n->L[2].type == AEC_LAYER_POOL) {
            HostLayer &pl = n->L[2];
            pl.swp_fused = pl.kh == 2 && pl.kw == 2 && pl.stride == 2 && pl.C % 4 == 0;
        }
    }
    {
        // a 2x2 / stride-2 pool behind a conv layer on the gathered weights-as-M kernel: windows all four sites of which are
        // re-evaluated are reduced in the conv epilogue (AEC_POOL_FUSE=0: never).  Needs a free bit in the work-list entries.
        const char *e = getenv("AEC_POOL_FUSE");
        for (size_t li = 1; li + 1 < n->L.size(); ++li) {
            HostLayer &c = n->L[li], &pl = n->L[li + 1];
            if (e && atoi(e) == 0) break;
            if (c.type != AEC_LAYER_CONV || pl.type != AEC_LAYER_POOL || !c.tc || c.tc_sm || c.rt || c.rep != 1) continue;
            if (!(pl.kh == 2 && pl.kw == 2 && pl.stride == 2) || pl.swp_fused) continue;
            if (((unsigned long long)n->S << c.code.sh_s) > (1ULL << 31)) continue;
            if ((unsigned long long)n->S * pl.H * pl.Ww >= (1ULL << 27)) continue;      // flag word index << 5 | bit in 32 bits
            c.pool_fuse = true;
            pl.pool_in_conv = true;
        }
    }
    if ((rc = dev_alloc(n, &n->surface, S * n->H * n->W, true))) return rc;
    if ((rc = dev_alloc(n, &n->delta, S, true))) return rc;
    if ((rc = dev_alloc(n, &n->prev_ts, S, true))) return rc;
    if ((rc = dev_alloc(n, &n->active, S, true))) return rc;
    if ((rc = dev_alloc(n, &n->mask, S, true))) return rc;
    for (auto &l : n->L) {
        const size_t bm = (size_t)l.H * l.Ww;
        if ((rc = dev_alloc(n, &l.front, S * bm, true))) return rc;
        if ((rc = dev_alloc(n, &l.nzr, S * bm, true))) return rc;
        if ((rc = dev_alloc(n, &l.skip, S * bm, true))) return rc;
        if (l.type == AEC_LAYER_CONV) {
            if ((rc = dev_alloc_map(n, &l.F, S * l.fstride))) return rc;
            if ((rc = dev_alloc_map(n, &l.A, S * l.fstride))) return rc;
            if ((rc = dev_alloc(n, &l.signchg, S * bm, true))) return rc;
            if ((rc = dev_alloc(n, &l.initF, (size_t)l.fstride, false))) return rc;
            if ((rc = dev_alloc(n, &l.wgt, l.h_w.size(), false))) return rc;
            if ((rc = dev_alloc(n, &l.bias, l.h_b.size(), false))) return rc;
            CU(cudaMemcpy(l.wgt, l.h_w.data(), l.h_w.size() * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(l.bias, l.h_b.data(), l.h_b.size() * 4, cudaMemcpyHostToDevice));
            l.h_w.clear(); l.h_w.shrink_to_fit();
            if (l.tc) {
                if ((rc = dev_alloc(n, &l.wimg, l.h_wimg.size(), false))) return rc;
                if ((rc = dev_alloc(n, &l.tc_timing, 16, false))) return rc;
                CU(cudaMemcpy(l.wimg, l.h_wimg.data(), l.h_wimg.size() * 4, cudaMemcpyHostToDevice));
                l.h_wimg.clear(); l.h_wimg.shrink_to_fit();
            }
            if (l.rt) {
                if ((rc = dev_alloc(n, &l.rtimg, l.h_rtimg.size(), false))) return rc;
                CU(cudaMemcpy(l.rtimg, l.h_rtimg.data(), l.h_rtimg.size() * 4, cudaMemcpyHostToDevice));
                l.h_rtimg.clear(); l.h_rtimg.shrink_to_fit();
                if ((rc = dev_alloc(n, &l.nset, S * bm, true))) return rc;
            }
        } else if (l.type == AEC_LAYER_POOL) {
            if ((rc = dev_alloc(n, &l.idx, S * l.fstride, true))) return rc;
            if ((rc = dev_alloc(n, &l.flags, S * bm, true))) return rc;
            if ((rc = dev_alloc(n, &l.initIdx, (size_t)l.fstride, false))) return rc;
            if ((rc = dev_alloc_map(n, &l.Fp, S * l.fstride))) return rc;
            if ((rc = dev_alloc_map(n, &l.Ap, S * l.fstride))) return rc;
            if ((rc = dev_alloc(n, &l.initFp, (size_t)l.fstride, false))) return rc;
            if (l.swp_fused && (rc = dev_alloc(n, &l.swp_uns, S * bm, true))) return rc;
        }
        if (l.type != AEC_LAYER_INTEGRATION) maxHW = std::max(maxHW, (size_t)l.H * l.W);
    }
    {
        size_t total_sites = 0;
        for (auto &l : n->L)
            if (l.type != AEC_LAYER_INTEGRATION) total_sites += S * (size_t)l.H * l.W;
        (void)maxHW;
        if ((rc = dev_alloc(n, &n->sites, total_sites, true))) return rc;
        size_t off = 0;
        for (auto &l : n->L)
            if (l.type != AEC_LAYER_INTEGRATION) { l.sites = n->sites + off; off += S * (size_t)l.H * l.W; }
    }
    for (auto &l : n->L)
        if (l.type != AEC_LAYER_INTEGRATION) n->view_elems = std::max(n->view_elems, (size_t)l.H * l.W * l.C);
    if ((rc = dev_alloc(n, &n->view, 4 * n->view_elems, false))) return rc;
    if ((rc = dev_alloc(n, &n->counts, 64, false))) return rc;      // [l]: work-list entries of layer l; [32 + l]: work-set sites of a row-tile layer
    if ((rc = dev_alloc(n, &n->accum, 64, false))) return rc;       // [l]: sites evaluated; [32 + l]: row-tile units evaluated; [31]: scratch
    if ((rc = dev_alloc(n, &n->err_flag, 1, false))) return rc;
    if ((rc = dev_alloc(n, &n->off_dev, S + 1, false))) return rc;
    const HostLayer &last = n->L.back();
    n->head_per_stream = (size_t)last.H * last.W * last.C;
    if ((rc = dev_alloc(n, &n->head, S * n->head_per_stream, true))) return rc;

    // frontier table of the fused all-layer frontier kernel
    {
        std::vector<FrontLayer> tab(n->L.size());
        memset(tab.data(), 0, tab.size() * sizeof(FrontLayer));
        int mw = n->L[0].H * n->L[0].Ww;
        for (size_t li = 1; li < n->L.size(); ++li) {
            const HostLayer &l = n->L[li], &pv = n->L[li - 1];
            FrontLayer &f = tab[li];
            f.type = l.type; f.Hin = pv.H; f.Win = pv.W; f.WwIn = pv.Ww; f.H = l.H; f.W = l.W; f.Ww = l.Ww;
            f.kh = l.kh; f.kw = l.kw; f.pad_t = l.pad_t; f.pad_l = l.pad_l; f.stride = l.stride; f.code = l.code;
            f.front = l.front; f.signchg = l.signchg; f.flags = l.flags; f.nzr = l.nzr; f.skip = l.skip; f.sites = l.sites; f.counter = n->counts + li;
            f.rt_rows = l.rt ? l.rt_R : 0; f.rt_seg = l.rt_SEG; f.rt_nxg = l.rt_nxg; f.nset = l.nset; f.counter2 = n->counts + 32 + li;
            f.swp_uns = l.swp_fused ? l.swp_uns : nullptr; f.swp_skip = pv.skip;
            f.quad_bit = l.pool_fuse ? 0x80000000u : 0u; f.pool_in_conv = l.pool_in_conv ? 1 : 0;
            f.pWw = l.pool_fuse ? n->L[li + 1].Ww : 0;
            mw = std::max(mw, std::max(pv.H * pv.Ww, std::max(pv.H * l.Ww, l.H * l.Ww)));
        }
        n->front_max_words = mw;
        if ((size_t)5 * mw * 4 > 200 * 1024) return fail(AEC_EINVAL, "frame too large for the frontier kernel's shared memory");
        if ((rc = dev_alloc(n, &n->front_table, tab.size(), false))) return rc;
        CU(cudaMemcpy(n->front_table, tab.data(), tab.size() * sizeof(FrontLayer), cudaMemcpyHostToDevice));
        CU(cudaFuncSetAttribute(k_frontier_skip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>((size_t)3 * mw * 4, 48 * 1024)));
        CU(cudaFuncSetAttribute(k_frontier_all<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>((size_t)5 * mw * 4, 48 * 1024)));
        CU(cudaFuncSetAttribute(k_frontier_all<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>((size_t)5 * mw * 4, 48 * 1024)));
        CU(cudaFuncSetAttribute(k_frontier_all<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>((size_t)5 * mw * 4, 48 * 1024)));
    }

    // leak-sweep table: every conv layer's (F, A), then every pool layer's (Fp, Ap) copy
    { const char *e = getenv("AEC_TC_DEBUG"); n->tc_debug = e ? atoi(e) : 0; }
    { const char *e = getenv("AEC_TC_HALF"); n->tc_half = !(e && atoi(e) == 0); }            // 64-site units for small work lists (aec_tc.cuh)
    { const char *e = getenv("AEC_PDL"); n->pdl = e && atoi(e) != 0; }
    { const char *e = getenv("AEC_TC_PAIR"); n->tc_pair_mode = e ? (atoi(e) != 0) : -1; }      // CTA pairs for layers with 2, 4, ... weight tiles
    { const char *e = getenv("AEC_RT_STORE32"); n->rt_store32 = !(e && atoi(e) == 0); }      // 32-byte epilogue stores (aec_rt.cuh)
    memset(&n->sweep_all, 0, sizeof n->sweep_all);
    int chunk0 = 0, nc = 0;
    n->sweep_win_chunks = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (size_t li = 1; li < n->L.size(); ++li) {
            HostLayer &l = n->L[li];
            if (l.type != (pass == 0 ? AEC_LAYER_CONV : AEC_LAYER_POOL)) continue;
            if (pass == 0 && li + 1 < n->L.size() && n->L[li + 1].swp_fused && l.C % 4 == 0) {
                // this map is swept window by window by k_sweep_windows, which evaluates the pool layer's sticky windows on the way
                const HostLayer &pl = n->L[li + 1];
                SweepWindowsParams &w = n->sweep_win;
                memset(&w, 0, sizeof w);
                fill_sweep_layer(l, w.L, 0, false, true);
                SweepPool &q = w.Q;
                q.idx = pl.idx; q.Fp = pl.Fp; q.Ap = pl.Ap; q.pstride = pl.fstride; q.flags = pl.flags; q.uns = pl.swp_uns;
                q.accum = n->accum + (li + 1);
                q.pW = pl.W; q.pWw = pl.Ww; q.pHWw = pl.H * pl.Ww; q.alpha = l.alpha;
                q.wpc = std::max(1, std::min(kSweepMaxWords, kSweepUnitsPerChunk / (32 * 4 * w.L.c4)));   // a window = four sites
                w.delta = n->delta; w.active = n->active;
                n->sweep_win_chunks = (q.pHWw + q.wpc - 1) / q.wpc;
                continue;
            }
            chunk0 += fill_sweep_layer(l, n->sweep_all.L[nc], chunk0, false, n->sweep_skip);
            ++nc;
        }
        if (pass == 0) { n->sweep_nconv = nc; n->sweep_conv_chunks = chunk0; }
    }
    n->sweep_all.n_layers = nc;
    n->sweep_all.delta = n->delta;
    n->sweep_all.active = n->active;
    n->sweep_chunks = chunk0;

    // shared-memory opt-ins and persistent grid sizes
    {
        size_t need = (size_t)n->hash_slots * 8 + (size_t)n->L[0].H * n->L[0].Ww * 8;
        if (need > 200 * 1024) return fail(AEC_EINVAL, "max_events_per_step/surface too large for the surface kernel's shared memory");
        CU(cudaFuncSetAttribute(k_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(need, 48 * 1024)));
        size_t fc = 0, fp = 0;
        for (size_t i = 1; i < n->L.size(); ++i) {
            const HostLayer &l = n->L[i], &pv = n->L[i - 1];
            if (l.type == AEC_LAYER_CONV) fc = std::max(fc, ((size_t)pv.H * pv.Ww + (size_t)pv.H * l.Ww + (size_t)l.H * l.Ww) * 4);
            else fp = std::max(fp, ((size_t)pv.H * pv.Ww + (size_t)l.H * l.Ww) * 4);
        }
        if (fc > 200 * 1024 || fp > 200 * 1024) return fail(AEC_EINVAL, "frame too large for the frontier kernels' shared memory");
        CU(cudaFuncSetAttribute(k_conv_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fc, 48 * 1024)));
        CU(cudaFuncSetAttribute(k_pool_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fp, 48 * 1024)));
        int b = 0;
        size_t tc_max = 0;
        for (auto &l : n->L)
            if (l.type == AEC_LAYER_CONV && l.tc) {
                tc_max = std::max(tc_max, l.tc_smem);
                l.tc_blocks = n->num_sms;              // persistent: one warp-specialised CTA per SM (all 512 TMEM columns)
                // decoder variant per layer (profiles/r1e_summary.md): the batched decoder pays for long units or several
                // weight tiles; the environment is read here, once, not at every launch
                const char *fk = getenv("AEC_TC_FASTDEC_KB");
                l.tc_fast_decode = l.KB >= (fk ? atoi(fk) : 10) || (l.m_tiles > 1 && !fk) || l.pool_fuse;
                // Two CTAs per unit (aec_tc.cuh, kPair) where the layer has an even number of weight tiles.  With few streams the
                // work lists are short and 64-site units on single CTAs spread them over more SMs: pairs from 32 streams on.
                l.tc_pair = !l.tc_sm && l.m_tiles % 2 == 0 && l.mtu == 1 && n->num_sms >= 2 &&
                            (n->tc_pair_mode == 1 || (n->tc_pair_mode < 0 && n->S >= 32));
                if (l.tc_pair) {
                    l.tc_fast_decode = true;
                    l.tc_blocks = n->num_sms & ~1;
                    // a CTA of a pair converts half of a unit's sites: three 32 KB site stages, the rest goes to the weight ring
                    const size_t w_stage = 2 * (size_t)l.Mrows * 128, x_bytes = (size_t)tc::kPairSiteStages * 2 * tc::kItemTileBytes;
                    l.w_stages = (int)std::min<size_t>(tc::kMaxWStages, std::max<size_t>(2, (210 * 1024 - x_bytes) / w_stage));
                    l.tc_smem = x_bytes + (size_t)l.w_stages * w_stage + 1024;
                    tc_max = std::max(tc_max, l.tc_smem);
                }
            }
        if (tc_max > 227 * 1024) return fail(AEC_EINVAL, "tensor-core conv tile needs %zu bytes of shared memory", tc_max);
        if (tc_max) {
            const void *variants[6] = {(const void *)tc::k_conv_eval_tc<false, false>, (const void *)tc::k_conv_eval_tc<true, false>,
                                       (const void *)tc::k_conv_eval_tc<true, true>, (const void *)tc::k_conv_eval_tc<true, false, true>,
                                       (const void *)tc::k_conv_eval_tc<true, false, false, true>, (const void *)tc::k_conv_eval_tc<true, false, true, true>};
            for (const void *fn : variants) {
                cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_max);
                if (e != cudaSuccess) return fail(AEC_ECUDA, "cannot opt in to %zu bytes of dynamic shared memory for the tensor-core conv kernel: %s", tc_max, cudaGetErrorString(e));
                // the role split (setmaxnreg) must fit the register pool the CTA is launched with, or the
                // producers' increase would wait forever
                cudaFuncAttributes fa;
                CU(cudaFuncGetAttributes(&fa, fn));
                const int pool = fa.numRegs * tc::kTcThreads;
                const int want = 128 * tc::kRegsEpi + 128 * tc::kRegsCtl + 384 * tc::kRegsProd;
                if (want > pool || fa.numRegs > tc::kRegsProd || fa.numRegs < tc::kRegsEpi)
                    return fail(AEC_EINVAL, "tensor-core conv kernel was built with %d registers/thread; the role split needs %d of %d", fa.numRegs, want, pool);
                if (tc_max + fa.sharedSizeBytes > 227 * 1024)
                    return fail(AEC_EINVAL, "tensor-core conv kernel needs %zu + %zu bytes of shared memory", tc_max, (size_t)fa.sharedSizeBytes);
            }
        }
        size_t rt_max = 0;
        for (auto &l : n->L)
            if (l.rt) rt_max = std::max(rt_max, l.rt_smem);
        const void *rt_variants[4] = {(const void *)rt::k_conv_rows<16, false>, (const void *)rt::k_conv_rows<16, true>,
                                      (const void *)rt::k_conv_rows<32, false>, (const void *)rt::k_conv_rows<32, true>};
        for (const void *fn : rt_variants) {
            if (!rt_max) break;
            cudaFuncAttributes fa;
            CU(cudaFuncGetAttributes(&fa, fn));
            if (rt_max + fa.sharedSizeBytes > 227 * 1024)
                return fail(AEC_EINVAL, "row-tile conv kernel needs %zu + %zu bytes of shared memory", rt_max, (size_t)fa.sharedSizeBytes);
            CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rt_max));
            const int pool = fa.numRegs * rt::kRtThreads;
            const int want = rt::kRtEpiWarps * 32 * rt::kRtRegsEpi + 128 * rt::kRtRegsCtl + 384 * rt::kRtRegsProd;
            if (want > pool || fa.numRegs > rt::kRtRegsProd || fa.numRegs < rt::kRtRegsEpi)
                return fail(AEC_EINVAL, "row-tile conv kernel was built with %d registers/thread; the role split needs %d of %d", fa.numRegs, want, pool);
        }
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_conv_eval<16, 2, 4, 16>, kThreads, 0));
        n->conv_eval_blocks[0] = std::max(1, b) * n->num_sms;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_conv_eval<32, 4, 4, 16>, kThreads, 0));
        n->conv_eval_blocks[1] = std::max(1, b) * n->num_sms;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_conv_eval<64, 4, 8, 16>, kThreads, 0));
        n->conv_eval_blocks[2] = std::max(1, b) * n->num_sms;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_conv_eval<128, 8, 8, 16>, kThreads, 0));
        n->conv_eval_blocks[3] = std::max(1, b) * n->num_sms;
    }

    // initial state = the chain evaluated on the all-zero surface (conv2d.py:59-63, maxpool.py:31-36),
    // computed once on stream 0 with "every site" work lists, then broadcast by reset.
    cudaStream_t st = 0;
    for (size_t li = 1; li < n->L.size(); ++li) {
        HostLayer &l = n->L[li];
        const int HW = l.H * l.W;
        k_all_sites<<<(HW + kThreads - 1) / kThreads, kThreads, 0, st>>>(l.sites, n->counts + li, l.H, l.W, l.code);
        if ((rc = launch_check(n, "k_all_sites"))) return rc;
        if (l.type == AEC_LAYER_CONV) {
            if ((rc = run_conv_eval(n, (int)li, st))) return rc;
            CU(cudaMemcpyAsync(l.initF, l.F, (size_t)l.fstride * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemsetAsync(l.A, 0, (size_t)l.fstride * 4, st));   // `_init_conv_actfn = zeros` (conv2d.py:62)
        } else {
            if ((rc = run_pool_eval(n, (int)li, st))) return rc;
            CU(cudaMemcpyAsync(l.initIdx, l.idx, (size_t)l.fstride, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(l.initFp, l.Fp, (size_t)l.fstride * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemsetAsync(l.Ap, 0, (size_t)l.fstride * 4, st));
            CU(cudaMemsetAsync(l.flags, 0, (size_t)l.H * l.Ww * 4, st));
        }
    }
    CU(cudaMemsetAsync(n->accum, 0, 64 * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(n->counts, 0, 64 * sizeof(int), st));
    if ((rc = reset_streams(n, nullptr, st))) return rc;
    CU(cudaStreamSynchronize(st));
    n->finalized = true;
    n->steps = 0;
    return AEC_OK;
}

extern "C" void aec_net_destroy(aec_net *n)
{
    if (!n) return;
    cudaSetDevice(n->device);
    cudaDeviceSynchronize();
    for (void *p : n->allocs) cudaFree(p);
    if (n->ev_dev) cudaFree(n->ev_dev);
    for (auto &sl : n->slot) {
        if (sl.ev) cudaFree(sl.ev);
        if (sl.off) cudaFree(sl.off);
        if (sl.head) cudaFree(sl.head);
        if (sl.ev_ready) cudaEventDestroy(sl.ev_ready);
        if (sl.k_done) cudaEventDestroy(sl.k_done);
        if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
    }
    if (n->dec_boxes) { cudaFree(n->dec_boxes); cudaFree(n->dec_conf); cudaFree(n->dec_label); cudaFree(n->dec_valid); }
    if (n->sg.exec) cudaGraphExecDestroy(n->sg.exec);
    if (n->sg.graph) cudaGraphDestroy(n->sg.graph);
    if (n->cap) cudaStreamDestroy(n->cap);
    if (n->h2d) cudaStreamDestroy(n->h2d);
    if (n->d2h) cudaStreamDestroy(n->d2h);
    delete n;
}

extern "C" int aec_net_num_layers(const aec_net *n) { return n ? (int)n->L.size() : 0; }
extern "C" int aec_net_num_streams(const aec_net *n) { return n ? n->S : 0; }
extern "C" size_t aec_net_state_bytes_per_stream(const aec_net *n) { return n ? n->per_stream_bytes : 0; }
extern "C" size_t aec_net_device_bytes(const aec_net *n) { return n ? n->dev_bytes : 0; }
extern "C" unsigned long long aec_net_launch_count(const aec_net *n) { return n ? n->launches : 0; }

extern "C" int aec_net_layer_info(const aec_net *n, int layer, aec_layer_info *info)
{
    if (!n || !info || layer < 0 || layer >= (int)n->L.size()) return fail(AEC_EINVAL, "bad layer index %d", layer);
    const HostLayer &l = n->L[layer];
    info->type = l.type; info->channels = l.C; info->height = l.H; info->width = l.W;
    info->k_h = l.kh; info->k_w = l.kw; info->stride = l.stride; info->pad_top = l.pad_t; info->pad_left = l.pad_l;
    info->in_channels = l.Cin; info->frontier_words_per_row = l.Ww;
    return AEC_OK;
}

#define NEED_FINAL(n)                                                                   \
    do {                                                                                \
        if (!(n) || !(n)->finalized) return fail(AEC_ESTATE, "network is not finalized"); \
        CU(cudaSetDevice((n)->device));                                                 \
    } while (0)

extern "C" int aec_net_reset(aec_net *n, const uint8_t *stream_mask, void *cuda_stream)
{
    NEED_FINAL(n);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint8_t *m = nullptr;
    if (stream_mask) {
        CU(cudaMemcpyAsync(n->mask, stream_mask, (size_t)n->S, cudaMemcpyHostToDevice, st));
        m = n->mask;
    }
    return reset_streams(n, m, st);
}

static int step_body(aec_net *n, const int32_t *ev, const int32_t *off, cudaStream_t st)
{
    int rc;
    if ((rc = run_integrate(n, ev, off, st))) return rc;
    if ((rc = run_frontier_skip(n, st))) return rc;       // which sites the step re-evaluates anyway: the sweep skips them
    if ((rc = run_sweep(n, -1, st))) return rc;
    if ((rc = run_frontier_all(n, st))) return rc;
    for (int li = 1; li < (int)n->L.size(); ++li)
        if ((rc = run_eval(n, li, st))) return rc;
    return run_head(n, st);
}

// Captures one step into a CUDA graph (on an internal stream: the caller's may be the legacy default stream, which
// cannot capture) and remembers the two kernel nodes whose pointer arguments change between steps.
static int build_step_graph(aec_net *n, const int32_t *ev, const int32_t *off)
{
    aec_net::StepGraph &g = n->sg;
    if (!n->cap) CU(cudaStreamCreateWithFlags(&n->cap, cudaStreamNonBlocking));
    const unsigned long long l0 = n->launches;
    CU(cudaStreamBeginCapture(n->cap, cudaStreamCaptureModeThreadLocal));
    int rc = step_body(n, ev, off, n->cap);
    cudaError_t e = cudaStreamEndCapture(n->cap, &g.graph);
    g.launches = n->launches - l0;
    n->launches = l0;
    if (rc) return rc;
    if (e != cudaSuccess) return fail(AEC_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    CU(cudaGraphInstantiate(&g.exec, g.graph, 0));
    size_t num = 0;
    CU(cudaGraphGetNodes(g.graph, nullptr, &num));
    std::vector<cudaGraphNode_t> nodes(num);
    CU(cudaGraphGetNodes(g.graph, nodes.data(), &num));
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        CU(cudaGraphNodeGetType(nd, &ty));
        if (ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        CU(cudaGraphKernelNodeGetParams(nd, &kp));
        if (kp.func == (void *)k_integrate) { g.integ = nd; g.integ_kp = kp; }
        if (kp.func == (void *)k_head) { g.head = nd; g.head_kp = kp; }
    }
    if (!g.integ || !g.head) return fail(AEC_ECUDA, "step graph: surface / head kernel node not found");
    g.ip = integrate_params(n, ev, off);
    g.hp = head_params(n, n->head_cur ? n->head_cur : n->head);
    return AEC_OK;
}

extern "C" int aec_net_step_device(aec_net *n, const int32_t *ev, const int32_t *off, int total, void *cuda_stream)
{
    NEED_FINAL(n);
    (void)total;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc;
    if (n->use_graph < 0) {
        const char *g = getenv("AEC_GRAPH");
        n->use_graph = (g && atoi(g) == 0) ? 0 : 1;
    }
    if (n->use_graph && !n->profiling && !n->tc_timing_on && !n->sg.failed) {
        aec_net::StepGraph &g = n->sg;
        if (!g.exec && (rc = build_step_graph(n, ev, off))) {
            g.failed = true;                  // fall back to plain launches (same kernels, same order)
            cudaGetLastError();
        }
        if (g.exec) {
            float *out = n->head_cur ? n->head_cur : n->head;
            if (g.ip.events != ev || g.ip.offsets != off || g.ip.ends != n->ev_ends) {
                g.ip.events = ev; g.ip.offsets = off; g.ip.ends = n->ev_ends;
                cudaKernelNodeParams kp = g.integ_kp;
                void *args[1] = {&g.ip};
                kp.kernelParams = args;
                CU(cudaGraphExecKernelNodeSetParams(g.exec, g.integ, &kp));
            }
            if (g.hp.out != out) {
                g.hp.out = out;
                cudaKernelNodeParams kp = g.head_kp;
                void *args[1] = {&g.hp};
                kp.kernelParams = args;
                CU(cudaGraphExecKernelNodeSetParams(g.exec, g.head, &kp));
            }
            CU(cudaGraphLaunch(g.exec, st));
            n->head_last = out;
            n->launches += g.launches;
            n->steps++;
            return AEC_OK;
        }
    }
    if ((rc = step_body(n, ev, off, st))) return rc;
    n->steps++;
    return prof_collect(n, st);
}

// Host-side check of a packed event batch: offsets start at 0, never decrease and end at `total` - the surface
// kernel indexes events[off[s] .. off[s+1]) without further checks, and the device buffer is sized from `total`.
static int check_offsets(const aec_net *n, const int32_t *ev, const int32_t *off, int total)
{
    if (total < 0 || !off || (total > 0 && !ev)) return fail(AEC_EINVAL, "bad event buffers");
    if (off[0] != 0) return fail(AEC_EINVAL, "offsets[0] must be 0, got %d", off[0]);
    for (int s = 0; s < n->S; ++s)
        if (off[s + 1] < off[s]) return fail(AEC_EINVAL, "offsets decrease at stream %d (%d -> %d)", s, off[s], off[s + 1]);
    if (off[n->S] != total) return fail(AEC_EINVAL, "offsets[n_streams] = %d does not match total = %d", off[n->S], total);
    return AEC_OK;
}

static int upload_events(aec_net *n, const int32_t *ev, const int32_t *off, int total, cudaStream_t st)
{
    int rc0 = check_offsets(n, ev, off, total);
    if (rc0) return rc0;
    if ((size_t)total > n->ev_cap) {
        CU(cudaStreamSynchronize(st));
        if (n->ev_dev) { CU(cudaFree(n->ev_dev)); n->ev_dev = nullptr; }
        size_t cap = std::max<size_t>((size_t)total * 3 / 2, 1024);
        CU(cudaMalloc(&n->ev_dev, cap * 3 * sizeof(int32_t)));
        n->ev_cap = cap;
    }
    if (total > 0) CU(cudaMemcpyAsync(n->ev_dev, ev, (size_t)total * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(n->off_dev, off, ((size_t)n->S + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    return AEC_OK;
}

static int check_err_flag(aec_net *n, cudaStream_t st)
{
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, n->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flag) CU(cudaMemsetAsync(n->err_flag, 0, sizeof(int), st));
    flag |= n->pending_err;            // an event error of earlier pipelined steps that no caller has seen yet
    n->pending_err = 0;
    if (flag) {
        n->last_err_bits = flag;
        return fail(AEC_EEVENTS, "%s%s", (flag & 1) ? "event coordinates out of range (event skipped). " : "",
                    (flag & 2) ? "a stream exceeded max_events_per_step (stream skipped)." : "");
    }
    return AEC_OK;
}

extern "C" int aec_net_step_host(aec_net *n, const int32_t *ev, const int32_t *off, int total, float *head_out, void *cuda_stream)
{
    NEED_FINAL(n);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc;
    if ((rc = upload_events(n, ev, off, total, st))) return rc;
    if ((rc = aec_net_step_device(n, n->ev_dev, n->off_dev, total, cuda_stream))) return rc;
    if (head_out)
        CU(cudaMemcpyAsync(head_out, n->head, (size_t)n->S * n->head_per_stream * sizeof(float), cudaMemcpyDeviceToHost, st));
    return check_err_flag(n, st);
}

extern "C" int aec_net_host_sync(aec_net *n, void *cuda_stream)
{
    NEED_FINAL(n);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (n->h2d) CU(cudaStreamSynchronize(n->h2d));
    CU(cudaStreamSynchronize(st));
    if (n->d2h) CU(cudaStreamSynchronize(n->d2h));
    return check_err_flag(n, st);
}

extern "C" int aec_net_step_host_async(aec_net *n, const int32_t *ev, const int32_t *off, int total, float *head_out, void *cuda_stream)
{
    NEED_FINAL(n);
    { int rc0 = check_offsets(n, ev, off, total); if (rc0) return rc0; }
    if (n->profiling) return fail(AEC_ESTATE, "per-launch profiling is not available on the pipelined host path");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (!n->h2d) {
        CU(cudaStreamCreateWithFlags(&n->h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&n->d2h, cudaStreamNonBlocking));
    }
    for (auto &s2 : n->slot) {                 // both slots are set up on the first call (allocation synchronises)
        if (!s2.ev_ready) {
            CU(cudaEventCreateWithFlags(&s2.ev_ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s2.k_done, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s2.d2h_done, cudaEventDisableTiming));
            CU(cudaMalloc(&s2.off, ((size_t)n->S + 1) * sizeof(int32_t)));
            CU(cudaMalloc(&s2.head, (size_t)n->S * n->head_per_stream * sizeof(float)));
        }
        if ((size_t)total > s2.ev_cap) {       // grow the staging buffers: nothing may still be reading them
            int rc = aec_net_host_sync(n, cuda_stream);
            if (rc == AEC_EEVENTS) n->pending_err |= n->last_err_bits;      // reported by the caller's next host_sync
            else if (rc) return rc;
            if (s2.ev) { CU(cudaFree(s2.ev)); s2.ev = nullptr; }
            const size_t cap = std::max<size_t>((size_t)total * 3 / 2, 1024);
            CU(cudaMalloc(&s2.ev, cap * 3 * sizeof(int32_t)));
            s2.ev_cap = cap;
        }
    }
    aec_net::HostSlot &sl = n->slot[n->async_calls & 1];
    // copy-in stream: the kernels of this slot's previous step must have consumed its events
    if (sl.used) CU(cudaStreamWaitEvent(n->h2d, sl.k_done, 0));
    if (total > 0) CU(cudaMemcpyAsync(sl.ev, ev, (size_t)total * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, n->h2d));
    CU(cudaMemcpyAsync(sl.off, off, ((size_t)n->S + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, n->h2d));
    CU(cudaEventRecord(sl.ev_ready, n->h2d));
    // compute stream: events in place, this slot's head buffer read back
    CU(cudaStreamWaitEvent(st, sl.ev_ready, 0));
    if (sl.used) CU(cudaStreamWaitEvent(st, sl.d2h_done, 0));
    n->head_cur = sl.head;
    int rc = aec_net_step_device(n, sl.ev, sl.off, total, cuda_stream);
    n->head_cur = nullptr;
    if (rc) return rc;
    CU(cudaEventRecord(sl.k_done, st));
    // copy-out stream
    CU(cudaStreamWaitEvent(n->d2h, sl.k_done, 0));
    if (head_out)
        CU(cudaMemcpyAsync(head_out, sl.head, (size_t)n->S * n->head_per_stream * sizeof(float), cudaMemcpyDeviceToHost, n->d2h));
    CU(cudaEventRecord(sl.d2h_done, n->d2h));
    sl.used = true;
    n->async_calls++;
    return AEC_OK;
}

extern "C" const float *aec_net_head_device(const aec_net *n) { return n ? (n->head_last ? n->head_last : n->head) : nullptr; }
extern "C" size_t aec_net_head_elems_per_stream(const aec_net *n) { return n ? n->head_per_stream : 0; }

extern "C" int aec_net_read_head(aec_net *n, int first_stream, int count, float *host_out, void *cuda_stream)
{
    NEED_FINAL(n);
    if (first_stream < 0 || count < 0 || first_stream + count > n->S) return fail(AEC_EINVAL, "read_head: streams [%d, %d) out of range", first_stream, first_stream + count);
    if (count == 0) return AEC_OK;
    if (!host_out) return fail(AEC_EINVAL, "read_head: host_out is NULL");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const float *src = (n->head_last ? n->head_last : n->head) + (size_t)first_stream * n->head_per_stream;
    CU(cudaMemcpyAsync(host_out, src, (size_t)count * n->head_per_stream * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AEC_OK;
}

extern "C" int aec_net_begin_step(aec_net *n, const int32_t *ev, const int32_t *off, int total, void *cuda_stream)
{
    NEED_FINAL(n);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc;
    if ((rc = upload_events(n, ev, off, total, st))) return rc;
    if ((rc = run_integrate(n, n->ev_dev, n->off_dev, st))) return rc;
    n->steps++;
    return check_err_flag(n, st);
}

extern "C" int aec_net_layer_compute(aec_net *n, int layer, void *cuda_stream)
{
    NEED_FINAL(n);
    if (layer < 1 || layer >= (int)n->L.size()) return fail(AEC_EINVAL, "layer_compute: layer %d out of range", layer);
    return run_layer(n, layer, true, (cudaStream_t)cuda_stream);
}

extern "C" int aec_net_compute_head(aec_net *n, void *cuda_stream)
{
    NEED_FINAL(n);
    return run_head(n, (cudaStream_t)cuda_stream);
}

extern "C" long long aec_net_read_size(const aec_net *n, int layer, int what)
{
    if (!n || layer < 0 || layer >= (int)n->L.size()) return fail(AEC_EINVAL, "bad layer index %d", layer);
    const HostLayer &l = n->L[layer];
    const long long hwc = (long long)l.H * l.W * l.C, bm = (long long)l.H * l.Ww * 4;
    switch (what) {
    case AEC_READ_SURFACE: return l.type == AEC_LAYER_INTEGRATION ? hwc * 8 : fail(AEC_EINVAL, "layer %d has no surface", layer);
    case AEC_READ_F: case AEC_READ_A: case AEC_READ_INIT_F:
        return l.type == AEC_LAYER_CONV ? hwc * 4 : fail(AEC_EINVAL, "layer %d is not a conv layer", layer);
    case AEC_READ_IDX: case AEC_READ_INIT_IDX:
        return l.type == AEC_LAYER_POOL ? hwc : fail(AEC_EINVAL, "layer %d is not a pool layer", layer);
    case AEC_READ_FLAGS: return l.type == AEC_LAYER_POOL ? bm : fail(AEC_EINVAL, "layer %d is not a pool layer", layer);
    case AEC_READ_FRONTIER: return bm;
    default: return fail(AEC_EINVAL, "unknown read selector %d", what);
    }
}

extern "C" int aec_net_read(aec_net *n, int layer, int what, int stream, void *host_out, size_t bytes)
{
    NEED_FINAL(n);
    if (stream < 0 || stream >= n->S) return fail(AEC_EINVAL, "stream %d out of range", stream);
    const long long need = aec_net_read_size(n, layer, what);
    if (need < 0) return (int)need;
    if ((size_t)need != bytes) return fail(AEC_EINVAL, "read: buffer is %zu bytes, need %lld", bytes, need);
    const HostLayer &l = n->L[layer];
    const size_t s = (size_t)stream;
    const void *src = nullptr;
    switch (what) {
    case AEC_READ_SURFACE: src = n->surface + s * l.H * l.W; break;
    case AEC_READ_F: src = l.F + s * l.fstride; break;
    case AEC_READ_A: src = l.A + s * l.fstride; break;
    case AEC_READ_INIT_F: src = l.initF; break;
    case AEC_READ_IDX: src = l.idx + s * l.fstride; break;
    case AEC_READ_INIT_IDX: src = l.initIdx; break;
    case AEC_READ_FLAGS: src = l.flags + s * l.H * l.Ww; break;
    case AEC_READ_FRONTIER: src = l.front + s * l.H * l.Ww; break;
    }
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost));
    return AEC_OK;
}

extern "C" int aec_net_read_step_info(aec_net *n, double *delta_out, uint8_t *active_out)
{
    NEED_FINAL(n);
    CU(cudaDeviceSynchronize());
    if (delta_out) CU(cudaMemcpy(delta_out, n->delta, (size_t)n->S * sizeof(double), cudaMemcpyDeviceToHost));
    if (active_out) CU(cudaMemcpy(active_out, n->active, (size_t)n->S, cudaMemcpyDeviceToHost));
    return AEC_OK;
}

extern "C" int aec_net_read_counters(aec_net *n, unsigned long long *sites, int n_layers, unsigned long long *steps, int reset)
{
    NEED_FINAL(n);
    CU(cudaDeviceSynchronize());
    unsigned long long tmp[64];
    CU(cudaMemcpy(tmp, n->accum, sizeof tmp, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_layers && i < 31; ++i)
        if (sites) sites[i] = i < (int)n->L.size() ? tmp[i] : 0ULL;
    if (steps) *steps = n->steps;
    if (reset) {
        CU(cudaMemset(n->accum, 0, sizeof tmp));
        n->steps = 0;
    }
    return AEC_OK;
}

extern "C" int aec_net_read_unit_counters(aec_net *n, unsigned long long *units, int n_layers)
{
    NEED_FINAL(n);
    if (!units) return fail(AEC_EINVAL, "read_unit_counters: units is NULL");
    CU(cudaDeviceSynchronize());
    unsigned long long tmp[64];
    CU(cudaMemcpy(tmp, n->accum, sizeof tmp, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_layers && i < 31; ++i) units[i] = (i < (int)n->L.size() && (n->L[i].rt || n->L[i].swp_fused || n->L[i].pool_in_conv)) ? tmp[32 + i] : 0ULL;
    return AEC_OK;
}

extern "C" int aec_net_read_view(aec_net *n, int layer, int stream, float *surface, float *layer_actfn, float *conv_actfn,
                                 float *featuremap)
{
    NEED_FINAL(n);
    if (layer < 1 || layer >= (int)n->L.size()) return fail(AEC_EINVAL, "read_view: layer %d is not a conv/pool layer", layer);
    if (stream < 0 || stream >= n->S) return fail(AEC_EINVAL, "stream %d out of range", stream);
    const HostLayer &l = n->L[layer];
    const size_t per = (size_t)l.H * l.W * l.C;
    ViewParams p;
    p.src = make_src(n, layer);
    p.stream = stream;
    p.surface = surface ? n->view : nullptr;
    p.layer_actfn = layer_actfn ? n->view + n->view_elems : nullptr;
    p.conv_actfn = conv_actfn ? n->view + 2 * n->view_elems : nullptr;
    p.featuremap = featuremap ? n->view + 3 * n->view_elems : nullptr;
    CU(cudaDeviceSynchronize());
    int blocks = (int)std::min<size_t>((per + kThreads - 1) / kThreads, (size_t)n->num_sms * 8);
    k_layer_view<<<blocks, kThreads>>>(p);
    int rc = launch_check(n, "k_layer_view");
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    if (surface) CU(cudaMemcpy(surface, p.surface, per * 4, cudaMemcpyDeviceToHost));
    if (layer_actfn) CU(cudaMemcpy(layer_actfn, p.layer_actfn, per * 4, cudaMemcpyDeviceToHost));
    if (conv_actfn) CU(cudaMemcpy(conv_actfn, p.conv_actfn, per * 4, cudaMemcpyDeviceToHost));
    if (featuremap) CU(cudaMemcpy(featuremap, p.featuremap, per * 4, cudaMemcpyDeviceToHost));
    return AEC_OK;
}

extern "C" int aec_net_profile(aec_net *n, int enable)
{
    NEED_FINAL(n);
    n->profiling = enable != 0;
    n->prof_slot = 0;
    n->prof_steps = 0;
    n->prof_names.clear();
    std::fill(n->prof_ms.begin(), n->prof_ms.end(), 0.0);
    return AEC_OK;
}

extern "C" int aec_net_read_profile(aec_net *n, double *ms_per_slot, int n_slots, unsigned long long *steps)
{
    NEED_FINAL(n);
    for (int i = 0; i < n_slots; ++i) ms_per_slot[i] = i < (int)n->prof_ms.size() ? n->prof_ms[i] : 0.0;
    if (steps) *steps = n->prof_steps;
    return (int)n->prof_ms.size();
}

extern "C" int aec_net_profile_slot_name(aec_net *n, int slot, char *buf, int cap)
{
    if (!n || !buf || cap < 1) return fail(AEC_EINVAL, "profile_slot_name: bad arguments");
    buf[0] = 0;
    if (slot < 0 || slot >= (int)n->prof_names.size()) return 0;
    snprintf(buf, (size_t)cap, "%s", n->prof_names[slot].c_str());
    return (int)strlen(buf);
}

extern "C" int aec_net_count_nonzero_rate_groups(aec_net *n, unsigned long long *nz_groups, unsigned long long *total_groups)
{
    NEED_FINAL(n);
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(n->accum + 31, 0, sizeof(unsigned long long)));
    unsigned long long total = 0;
    for (auto &l : n->L)
        if (l.type == AEC_LAYER_CONV) total += (unsigned long long)(l.fstride / 4) * n->S;
    if (total) {
        SweepParams conv_only;
        memset(&conv_only, 0, sizeof conv_only);
        int nc = 0, chunk0 = 0;
        for (auto &l : n->L)
            if (l.type == AEC_LAYER_CONV) {
                chunk0 += fill_sweep_layer(l, conv_only.L[nc], chunk0, true);
                ++nc;
            }
        conv_only.n_layers = nc;
        dim3 grid(chunk0, n->S);
        k_count_nz4<<<grid, kThreads>>>(conv_only, n->accum + 31);
        int rc = launch_check(n, "k_count_nz4");
        if (rc) return rc;
    }
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(nz_groups, n->accum + 31, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (total_groups) *total_groups = total;
    return AEC_OK;
}

extern "C" int aec_net_tc_timing(aec_net *n, int enable, int layer, unsigned long long *out16)
{
    NEED_FINAL(n);
    CU(cudaDeviceSynchronize());
    if (out16) {
        if (layer < 1 || layer >= (int)n->L.size() || !n->L[layer].tc_timing)
            return fail(AEC_EINVAL, "tc_timing: layer %d is not a tensor-core conv layer", layer);
        CU(cudaMemcpy(out16, n->L[layer].tc_timing, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return AEC_OK;
    }
    n->tc_timing_on = enable != 0;
    for (auto &l : n->L)
        if (l.tc_timing) CU(cudaMemset(l.tc_timing, 0, 16 * sizeof(unsigned long long)));
    return AEC_OK;
}

extern "C" int aec_net_tc_geometry(const aec_net *n, int layer, long long *out8)
{
    if (!n || !out8 || layer < 0 || layer >= (int)n->L.size()) return fail(AEC_EINVAL, "tc_geometry: bad layer index %d", layer);
    const HostLayer &l = n->L[layer];
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (l.type != AEC_LAYER_CONV || !l.tc) return AEC_OK;
    const long long k8 = (l.K + 7) / 8;
    out8[0] = 1;
    out8[1] = tc::kUnitSites;
    out8[5] = l.m_tiles;
    out8[3] = k8;
    if (l.rt) {
        // per unit (128 tile sites): kh * ncb stages x kw taps x CB/8 K steps x {value, rate} x {N = 2 Cpad, N = Cpad}
        const long long steps8 = (long long)l.kh * l.rt_ncb * l.kw * (l.rt_CB / 8);
        out8[3] = steps8;
        out8[4] = 4;
        out8[2] = steps8 * 2LL * (2LL * 128 * (2 * l.Mrows) * 8 + 2LL * 128 * l.Mrows * 8);
        out8[6] = 3;
        out8[7] = 1;                                               // units are counted by aec_net_read_unit_counters
    } else if (l.tc_sm) {
        out8[4] = 4;                                               // {value, rate} x {X_hi.[W_hi;W_lo] (N = 2 Cpad), X_lo.W_hi (N = Cpad)}
        out8[2] = k8 * 2LL * (2LL * 128 * (2 * l.Mrows) * 8 + 2LL * 128 * l.Mrows * 8);
        out8[6] = 2;
    } else {
        out8[4] = 3;                                               // W_hi.X_lo + W_lo.X_hi + W_hi.X_hi
        out8[2] = k8 * out8[4] * l.m_tiles * 2LL * 128 * tc::kUnitCols * 8;   // every MMA is M128 x N256 x K8
        out8[6] = l.tc_pair ? 4 : l.tc_fast_decode ? 1 : 0;        // pairs: one M = 256 MMA covers two weight tiles - the same FLOPs per unit
    }
    return AEC_OK;
}

extern "C" int aec_net_sweep_stats(aec_net *n, unsigned long long *out8)
{
    NEED_FINAL(n);
    if (!out8) return fail(AEC_EINVAL, "sweep_stats: out is NULL");
    unsigned long long nz = 0, tot = 0;
    int rc = aec_net_count_nonzero_rate_groups(n, &nz, &tot);
    if (rc) return rc;
    out8[0] = nz;
    out8[1] = tot;
    unsigned long long live_conv = 0, all_conv = 0, live_pool = 0, all_pool = 0, swept_conv = 0, swept_pool = 0;
    for (auto &l : n->L) {
        if (l.type == AEC_LAYER_INTEGRATION) continue;
        const long long words = (long long)n->S * l.H * l.Ww;
        const int blocks = (int)std::min<long long>((words + kThreads - 1) / kThreads, (long long)n->num_sms * 8);
        unsigned long long bits[2] = {0, 0};
        for (int pass = 0; pass < 2; ++pass) {          // 0: live sites; 1: live sites the sweep does not skip (skip[] of the last step)
            CU(cudaMemset(n->accum + 31, 0, sizeof(unsigned long long)));
            k_count_bits<<<blocks, kThreads>>>(l.nzr, pass == 1 && n->sweep_skip ? l.skip : nullptr, words, n->accum + 31);
            if ((rc = launch_check(n, "k_count_bits"))) return rc;
            CU(cudaMemcpy(&bits[pass], n->accum + 31, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        }
        const unsigned long long all = (unsigned long long)n->S * l.H * l.W * l.C;
        if (l.C % 4) bits[0] = bits[1] = (unsigned long long)n->S * l.H * l.W;      // dense path of the sweep: every site is read
        if (l.type == AEC_LAYER_CONV) { live_conv += bits[0] * l.C; swept_conv += bits[1] * l.C; all_conv += all; }
        else { live_pool += bits[0] * l.C; swept_pool += bits[1] * l.C; all_pool += all; }
    }
    CU(cudaMemset(n->accum + 31, 0, sizeof(unsigned long long)));
    out8[2] = live_conv; out8[3] = all_conv; out8[4] = live_pool; out8[5] = all_pool; out8[6] = swept_conv; out8[7] = swept_pool;
    return AEC_OK;
}

// ------------------------------------------------------------------------------------------------
// front end / back end of the path
extern "C" int aec_net_decode_head(aec_net *n, int num_classes, int num_bbox, int h_cells, int w_cells, int h_image, int w_image,
                                   float conf_threshold, float *boxes_out, float *conf_out, int32_t *label_out, uint8_t *valid_out,
                                   void *cuda_stream)
{
    NEED_FINAL(n);
    if (num_classes < 1 || num_bbox < 1 || h_cells < 1 || w_cells < 1)
        return fail(AEC_EINVAL, "decode_head: bad grid / class / box counts");
    if ((size_t)h_cells * w_cells * (num_classes + 5 * num_bbox) != n->head_per_stream)
        return fail(AEC_EINVAL, "decode_head: %dx%dx(%d+5*%d) does not match the head of %zu values per stream", h_cells, w_cells,
                    num_classes, num_bbox, n->head_per_stream);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t nb = (size_t)n->S * h_cells * w_cells * num_bbox;
    if (nb > n->dec_cap) {
        CU(cudaStreamSynchronize(st));
        if (n->dec_boxes) { cudaFree(n->dec_boxes); cudaFree(n->dec_conf); cudaFree(n->dec_label); cudaFree(n->dec_valid); }
        CU(cudaMalloc(&n->dec_boxes, nb * 4 * sizeof(float)));
        CU(cudaMalloc(&n->dec_conf, nb * sizeof(float)));
        CU(cudaMalloc(&n->dec_label, nb * sizeof(int32_t)));
        CU(cudaMalloc(&n->dec_valid, nb));
        n->dec_cap = nb;
    }
    DecodeParams p;
    p.head = n->head_last ? n->head_last : n->head; p.boxes = n->dec_boxes; p.conf = n->dec_conf; p.label = n->dec_label; p.valid = n->dec_valid;
    p.S = n->S; p.gh = h_cells; p.gw = w_cells; p.C = num_classes; p.B = num_bbox; p.h_img = h_image; p.w_img = w_image;
    p.thr = conf_threshold;
    const int blocks = (int)std::min<size_t>((nb + kThreads - 1) / kThreads, (size_t)n->num_sms * 8);
    k_decode_head<<<blocks, kThreads, 0, st>>>(p);
    int rc = launch_check(n, "k_decode_head");
    if (rc) return rc;
    if (boxes_out) CU(cudaMemcpyAsync(boxes_out, n->dec_boxes, nb * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (conf_out) CU(cudaMemcpyAsync(conf_out, n->dec_conf, nb * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (label_out) CU(cudaMemcpyAsync(label_out, n->dec_label, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (valid_out) CU(cudaMemcpyAsync(valid_out, n->dec_valid, nb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AEC_OK;
}

// f1 on the device, stand-alone form (host in / host out) for parity tests against the runner's split.
extern "C" int aec_split_batches(int device, const int32_t *events_yxt, const long long *rec_offsets, int n_recordings,
                                 int batch_event_size, int batch_event_usec, int32_t *chunk_offsets_out, int32_t *n_chunks_out)
{
    if (!rec_offsets || n_recordings < 0 || !chunk_offsets_out || !n_chunks_out) return fail(AEC_EINVAL, "split_batches: bad arguments");
    if (batch_event_usec <= 0 && batch_event_size < 1) return fail(AEC_EINVAL, "split_batches: batch_event_size must be >= 1");
    if (n_recordings == 0) return AEC_OK;
    if (rec_offsets[0] != 0) return fail(AEC_EINVAL, "split_batches: rec_offsets[0] must be 0");
    for (int r = 0; r < n_recordings; ++r)
        if (rec_offsets[r + 1] < rec_offsets[r]) return fail(AEC_EINVAL, "split_batches: rec_offsets decrease at %d", r);
    const long long total = rec_offsets[n_recordings];
    if (total > 0 && !events_yxt) return fail(AEC_EINVAL, "split_batches: NULL events");
    if (total + 2LL * n_recordings >= (1LL << 31)) return fail(AEC_EINVAL, "split_batches: too many events");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(AEC_EINVAL, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    int32_t *d_ev = nullptr, *d_cnt = nullptr, *d_co = nullptr, *d_nc = nullptr;
    long long *d_start = nullptr;
    auto cleanup = [&]() { cudaFree(d_ev); cudaFree(d_cnt); cudaFree(d_co); cudaFree(d_nc); cudaFree(d_start); };
#define CUF(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { cleanup(); return fail(AEC_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    std::vector<int32_t> cnt(n_recordings);
    for (int r = 0; r < n_recordings; ++r) cnt[r] = (int32_t)(rec_offsets[r + 1] - rec_offsets[r]);
    const size_t n_co = (size_t)total + 2 * (size_t)n_recordings;
    CUF(cudaMalloc(&d_ev, std::max<long long>(total, 1) * 3 * sizeof(int32_t)));
    CUF(cudaMalloc(&d_cnt, (size_t)n_recordings * sizeof(int32_t)));
    CUF(cudaMalloc(&d_nc, (size_t)n_recordings * sizeof(int32_t)));
    CUF(cudaMalloc(&d_co, n_co * sizeof(int32_t)));
    CUF(cudaMalloc(&d_start, (size_t)n_recordings * sizeof(long long)));
    if (total) CUF(cudaMemcpy(d_ev, events_yxt, (size_t)total * 3 * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUF(cudaMemcpy(d_cnt, cnt.data(), (size_t)n_recordings * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUF(cudaMemcpy(d_start, rec_offsets, (size_t)n_recordings * sizeof(long long), cudaMemcpyHostToDevice));
    CUF(cudaMemset(d_co, 0, n_co * sizeof(int32_t)));
    SplitParams p;
    p.events = d_ev; p.rec_start = d_start; p.counts = d_cnt; p.chunk_off = d_co; p.n_chunks = d_nc;
    p.size = batch_event_size; p.usec = batch_event_usec;
    k_split_batches<<<n_recordings, kThreads>>>(p);
    CUF(cudaGetLastError());
    CUF(cudaDeviceSynchronize());
    CUF(cudaMemcpy(chunk_offsets_out, d_co, n_co * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUF(cudaMemcpy(n_chunks_out, d_nc, (size_t)n_recordings * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef CUF
    cleanup();
    return AEC_OK;
}

// Raw recordings -> detections without leaving the device (f3 + f1 in front of the path): decode + the
// runner's transform (k_ndata_decode), batching (k_split_batches), then one step per batch index with every stream
// consuming its own batch (k_chunk_ranges feeds the surface kernel), reset on the first batch only (runner.py:64,101).
extern "C" int aec_net_run_ndata(aec_net *n, const uint8_t *raw, const long long *byte_offsets, int zero_base_ts, int crop, int new_h,
                                 int new_w, int batch_event_size, int batch_event_usec, int reset_first, float *head_out,
                                 int32_t *steps_out, int32_t *event_counts_out, void *cuda_stream)
{
    NEED_FINAL(n);
    if (!byte_offsets) return fail(AEC_EINVAL, "run_ndata: byte_offsets is NULL");
    if (batch_event_usec <= 0 && batch_event_size < 1) return fail(AEC_EINVAL, "run_ndata: batch_event_size must be >= 1");
    const int R = n->S;
    if (byte_offsets[0] != 0) return fail(AEC_EINVAL, "run_ndata: byte_offsets[0] must be 0");
    for (int r = 0; r < R; ++r)
        if (byte_offsets[r + 1] < byte_offsets[r] || byte_offsets[r + 1] % 5)
            return fail(AEC_EINVAL, "run_ndata: recording %d is not a whole number of 5-byte records", r);
    const long long total_bytes = byte_offsets[R];
    if (total_bytes > 0 && !raw) return fail(AEC_EINVAL, "run_ndata: NULL recordings");
    const long long n_ev = total_bytes / 5;
    if (n_ev + 2LL * R >= (1LL << 31)) return fail(AEC_EINVAL, "run_ndata: too many events");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint8_t *d_raw = nullptr;
    long long *d_off = nullptr, *d_start = nullptr;
    int32_t *d_ev = nullptr, *d_cnt = nullptr, *d_co = nullptr, *d_nc = nullptr, *d_begin = nullptr, *d_end = nullptr;
    auto cleanup = [&]() {
        n->ev_ends = nullptr;
        cudaFree(d_raw); cudaFree(d_off); cudaFree(d_start); cudaFree(d_ev); cudaFree(d_cnt); cudaFree(d_co); cudaFree(d_nc);
        cudaFree(d_begin); cudaFree(d_end);
    };
#define CUF(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { cleanup(); return fail(AEC_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    const size_t n_co = (size_t)n_ev + 2 * (size_t)R;
    std::vector<long long> start(R);
    for (int r = 0; r < R; ++r) start[r] = byte_offsets[r] / 5;
    CUF(cudaMalloc(&d_raw, std::max<long long>(total_bytes, 1)));
    CUF(cudaMalloc(&d_off, ((size_t)R + 1) * sizeof(long long)));
    CUF(cudaMalloc(&d_start, (size_t)R * sizeof(long long)));
    CUF(cudaMalloc(&d_ev, std::max<long long>(n_ev, 1) * 3 * sizeof(int32_t)));
    CUF(cudaMalloc(&d_cnt, (size_t)R * sizeof(int32_t)));
    CUF(cudaMalloc(&d_nc, (size_t)R * sizeof(int32_t)));
    CUF(cudaMalloc(&d_co, n_co * sizeof(int32_t)));
    CUF(cudaMalloc(&d_begin, (size_t)R * sizeof(int32_t)));
    CUF(cudaMalloc(&d_end, (size_t)R * sizeof(int32_t)));
    if (total_bytes) CUF(cudaMemcpyAsync(d_raw, raw, (size_t)total_bytes, cudaMemcpyHostToDevice, st));
    CUF(cudaMemcpyAsync(d_off, byte_offsets, ((size_t)R + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUF(cudaMemcpyAsync(d_start, start.data(), (size_t)R * sizeof(long long), cudaMemcpyHostToDevice, st));
    NdataParams dp;
    dp.raw = d_raw; dp.byte_off = d_off; dp.events = d_ev; dp.polarity = nullptr; dp.counts = d_cnt;
    dp.zero_base = zero_base_ts; dp.crop = crop; dp.new_h = new_h; dp.new_w = new_w;
    k_ndata_decode<<<R, kThreads, 0, st>>>(dp);
    CUF(cudaGetLastError());
    SplitParams sp;
    sp.events = d_ev; sp.rec_start = d_start; sp.counts = d_cnt; sp.chunk_off = d_co; sp.n_chunks = d_nc;
    sp.size = batch_event_size; sp.usec = batch_event_usec;
    k_split_batches<<<R, kThreads, 0, st>>>(sp);
    CUF(cudaGetLastError());
    n->launches += 2;
    std::vector<int32_t> nc(R);
    CUF(cudaMemcpyAsync(nc.data(), d_nc, (size_t)R * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (event_counts_out) CUF(cudaMemcpyAsync(event_counts_out, d_cnt, (size_t)R * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUF(cudaStreamSynchronize(st));              // the number of steps is data dependent: one read-back of R integers
    int steps = 0;
    for (int r = 0; r < R; ++r) steps = std::max(steps, (int)nc[r]);
    int rc = AEC_OK;
    if (reset_first && (rc = reset_streams(n, nullptr, st))) { cleanup(); return rc; }
    n->ev_ends = d_end;
    for (int b = 0; b < steps && rc == AEC_OK; ++b) {
        k_chunk_ranges<<<(R + kThreads - 1) / kThreads, kThreads, 0, st>>>(d_start, d_co, d_nc, b, R, d_begin, d_end);
        if ((rc = launch_check(n, "k_chunk_ranges"))) break;
        rc = aec_net_step_device(n, d_ev, d_begin, (int)n_ev, cuda_stream);
    }
    if (rc == AEC_OK && head_out)
        if (cudaMemcpyAsync(head_out, n->head_last ? n->head_last : n->head, (size_t)n->S * n->head_per_stream * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = fail(AEC_ECUDA, "run_ndata: copying the head failed");
    if (rc == AEC_OK) rc = check_err_flag(n, st);
    else cudaStreamSynchronize(st);
    if (steps_out) *steps_out = steps;
#undef CUF
    cleanup();
    return rc;
}

extern "C" int aec_decode_ndata(int device, const uint8_t *raw, const long long *byte_offsets, int n_recordings, int zero_base_ts,
                                int crop, int new_h, int new_w, int32_t *events_yxt_out, int32_t *polarity_out, int32_t *counts_out)
{
    if (!byte_offsets || n_recordings < 0 || !counts_out) return fail(AEC_EINVAL, "decode_ndata: bad arguments");
    if (n_recordings == 0) return AEC_OK;
    for (int r = 0; r < n_recordings; ++r)
        if (byte_offsets[r + 1] < byte_offsets[r] || byte_offsets[r] % 5 || byte_offsets[r + 1] % 5)
            return fail(AEC_EINVAL, "decode_ndata: recording %d is not a whole number of 5-byte records", r);
    const long long total_bytes = byte_offsets[n_recordings];
    if (total_bytes > 0 && (!raw || !events_yxt_out)) return fail(AEC_EINVAL, "decode_ndata: NULL buffers");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(AEC_EINVAL, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    const long long n_ev = total_bytes / 5;
    uint8_t *d_raw = nullptr;
    long long *d_off = nullptr;
    int32_t *d_ev = nullptr, *d_pol = nullptr, *d_cnt = nullptr;
    int rc = AEC_OK;
    auto cleanup = [&]() { cudaFree(d_raw); cudaFree(d_off); cudaFree(d_ev); cudaFree(d_pol); cudaFree(d_cnt); };
#define CUF(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { cleanup(); return fail(AEC_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    CUF(cudaMalloc(&d_raw, std::max<long long>(total_bytes, 1)));
    CUF(cudaMalloc(&d_off, ((size_t)n_recordings + 1) * sizeof(long long)));
    CUF(cudaMalloc(&d_ev, std::max<long long>(n_ev, 1) * 3 * sizeof(int32_t)));
    CUF(cudaMalloc(&d_pol, std::max<long long>(n_ev, 1) * sizeof(int32_t)));
    CUF(cudaMalloc(&d_cnt, (size_t)n_recordings * sizeof(int32_t)));
    if (total_bytes) CUF(cudaMemcpy(d_raw, raw, (size_t)total_bytes, cudaMemcpyHostToDevice));
    CUF(cudaMemcpy(d_off, byte_offsets, ((size_t)n_recordings + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    NdataParams p;
    p.raw = d_raw; p.byte_off = d_off; p.events = d_ev; p.polarity = polarity_out ? d_pol : nullptr; p.counts = d_cnt;
    p.zero_base = zero_base_ts; p.crop = crop; p.new_h = new_h; p.new_w = new_w;
    k_ndata_decode<<<n_recordings, kThreads>>>(p);
    CUF(cudaGetLastError());
    CUF(cudaDeviceSynchronize());
    if (n_ev) CUF(cudaMemcpy(events_yxt_out, d_ev, (size_t)n_ev * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (n_ev && polarity_out) CUF(cudaMemcpy(polarity_out, d_pol, (size_t)n_ev * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUF(cudaMemcpy(counts_out, d_cnt, (size_t)n_recordings * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef CUF
    cleanup();
    return rc;
}
