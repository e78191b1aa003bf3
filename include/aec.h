/*
 * aec.h - C ABI of the B200-native event-driven EFCN hot path (libaec_b200.so).
 *
 * This is the drop-in boundary for the reference's event-mode inference path
 * (marcocannici/async-ev-cnn).  The reference's only native interface is the Cython module
 * src/libs/cutils.pyx (im2col_event :29-30, min_argmax :139-140), called from
 * src/layers/conv2d.py:172 and src/layers/maxpool.py:130-139.  Materialising im2col columns is
 * exactly what the GPU design avoids, so this ABI replaces the whole stateful chain those two
 * functions serve - IntegrationLayer.compute (src/layers/integration.py:53-91),
 * Conv2DLayer.compute (src/layers/conv2d.py:105-137), MaxPoolLayer.compute
 * (src/layers/maxpool.py:105-161) and the model graph (src/models/event_numpy.py:53-105) - for
 * MANY independent event streams at once (one reference network object == one stream).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative AEC_E* code, and
 *     aec_last_error() returns a human-readable message for the calling thread's last failure;
 *   - `cuda_stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); work is
 *     enqueued asynchronously on it unless the function is documented as synchronising;
 *   - the library owns all per-stream network state (device memory); the caller owns every buffer
 *     it passes in;  one host thread per network object (not thread-safe, like the reference);
 *   - events are int32 triples (y, x, ts) - the layout of src/libs/runner.py:32 - packed for all
 *     streams, with `offsets[s] .. offsets[s+1]` delimiting stream s.  A stream with no events in a
 *     step is left untouched by that step.
 *   - there is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef AEC_H
#define AEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEC_OK 0
#define AEC_EINVAL (-1)   /* bad argument / unsupported layer configuration            */
#define AEC_ECUDA (-2)    /* a CUDA runtime call failed (message has the CUDA error)   */
#define AEC_ESTATE (-3)   /* call not valid in the network's current state             */
#define AEC_EEVENTS (-4)  /* event coordinates out of range / too many events per step */
#define AEC_ENOMEM (-5)

#define AEC_LAYER_INTEGRATION 0
#define AEC_LAYER_CONV 1
#define AEC_LAYER_POOL 2

#define AEC_PAD_VALID 0
#define AEC_PAD_SAME 1

/* selectors for aec_net_read() */
#define AEC_READ_SURFACE 0  /* integration: float64 [H][W]                   (integration.py:25, f64 per SURVEY Q2) */
#define AEC_READ_F 1        /* conv: float32 [H][W][C] pre-activation map    (conv2d.py:61 `_featuremap`, channel-last) */
#define AEC_READ_A 2        /* conv: float32 [H][W][C] leak-rate map         (conv2d.py:63 `_conv_actfn`, channel-last) */
#define AEC_READ_IDX 3      /* pool: uint8  [Ho][Wo][C] argmax row ky*kw+kx  (maxpool.py:33-35 `_idx_max[0]`) */
#define AEC_READ_FLAGS 4    /* pool: uint32 [Ho][ceil(Wo/32)] sticky recompute bitmap (maxpool.py:36 `_recompute_coords`) */
#define AEC_READ_FRONTIER 5 /* any : uint32 [H][ceil(W/32)] output-event bitmap of the last step (the `new_events` each compute() returns) */
#define AEC_READ_INIT_F 6   /* conv: float32 [H][W][C] state after construction/reset (conv2d.py:59-60 `_init_fm`) */
#define AEC_READ_INIT_IDX 7 /* pool: uint8 [Ho][Wo][C]                        (maxpool.py:33 `_init_idx_max`) */

typedef struct aec_net aec_net;

typedef struct aec_layer_info {
    int type;               /* AEC_LAYER_* */
    int channels, height, width;   /* output shape (out_shape(): layer.py:71-75) */
    int k_h, k_w, stride;
    int pad_top, pad_left;
    int in_channels;
    int frontier_words_per_row;    /* ceil(width/32) */
} aec_layer_info;

/* Thread-local message of the last failing call. */
const char *aec_last_error(void);
/* Library/ABI version (major*1000 + minor). */
int aec_version(void);

/*
 * Creates a network holding `n_streams` independent streams on CUDA device `device`; layer 0 is
 * the leaky integration surface (IntegrationLayer(leak, h, w): integration.py:12-26).
 * `max_events_per_step` bounds the events ONE stream may receive in one step (sizes the
 * last-duplicate-wins hash of the surface kernel); 0 selects the default (2048).
 */
int aec_net_create(aec_net **out, int device, int n_streams, int height, int width, double leak,
                   int max_events_per_step);

/*
 * Appends an event convolution (Conv2DLayer(prev, kernel, bias, stride, alpha, padding):
 * conv2d.py:15-66).  `kernel_hwio` is float32 [k_h][k_w][c_in][c_out] (the checkpoint layout,
 * event_numpy.py:64), `bias` float32 [c_out]; both are HOST pointers, copied.  Only stride 1 is
 * supported (the reference model builder always passes 1, event_numpy.py:64).
 * Returns the new layer's index (>= 1) or a negative error.
 */
int aec_net_add_conv(aec_net *net, int k_h, int k_w, int c_in, int c_out, const float *kernel_hwio,
                     const float *bias, int stride, float alpha, int padding);

/*
 * Appends an event max-pool (MaxPoolLayer(prev, [k_h,k_w], stride): maxpool.py:14-40).  Like the
 * reference's im2col_event (cutils.pyx:83-89) the stride must equal the kernel size; the previous
 * layer must be a convolution and its height/width must be multiples of the stride.
 * Returns the new layer's index or a negative error.
 */
int aec_net_add_pool(aec_net *net, int k_h, int k_w, int stride);

/*
 * Allocates per-stream state, evaluates the initial state of every layer on the all-zero surface
 * (what each reference constructor does: conv2d.py:59-63, maxpool.py:31-36) and resets every
 * stream to it.  Synchronises.  Must be called once, after the last add_* and before any step.
 */
int aec_net_finalize(aec_net *net);

void aec_net_destroy(aec_net *net);

int aec_net_num_layers(const aec_net *net);
int aec_net_num_streams(const aec_net *net);
int aec_net_layer_info(const aec_net *net, int layer, aec_layer_info *info);
/* Bytes of device memory held per stream / in total. */
size_t aec_net_state_bytes_per_stream(const aec_net *net);
size_t aec_net_device_bytes(const aec_net *net);

/*
 * reset() of every layer (integration.py:48-51, conv2d.py:99-103, maxpool.py:84-90) for the streams
 * whose byte in `stream_mask` (HOST, uint8 [n_streams]) is non-zero; NULL resets all streams.
 */
int aec_net_reset(aec_net *net, const uint8_t *stream_mask, void *cuda_stream);

/*
 * One step of graph(events, reset) (event_numpy.py:94-103) for every stream, events already in
 * DEVICE memory: `events_yxt` int32 [total][3], `offsets` int32 [n_streams+1] (both device).
 * Runs integration, the leak sweep of all conv layers, then every layer's frontier update in order.
 * Asynchronous.  The head (`[n_streams][H_last][W_last][C_last]` float32, the reference's
 * `featuremap().transpose(1,2,0)`, event_numpy.py:79) is left in the device buffer returned by
 * aec_net_head_device().
 */
int aec_net_step_device(aec_net *net, const int32_t *events_yxt, const int32_t *offsets, int total_events,
                        void *cuda_stream);

/*
 * Same step with HOST buffers, end to end: copies events and offsets host->device, runs the step,
 * copies the head device->host into `head_out` (float32 [n_streams][H][W][C], may be NULL) and
 * synchronises the stream.  Pinned host buffers make the copies asynchronous; pageable ones work.
 * Returns AEC_EEVENTS if any stream had out-of-range coordinates or more than max_events_per_step
 * events (such events/streams are skipped).
 */
int aec_net_step_host(aec_net *net, const int32_t *events_yxt, const int32_t *offsets, int total_events,
                      float *head_out, void *cuda_stream);

/*
 * Pipelined form of aec_net_step_host for throughput: enqueues the host->device copy of the events on an
 * internal copy-in stream, the step on `cuda_stream` and the device->host copy of the head on an internal
 * copy-out stream, and returns without waiting.  Two steps can be in flight (two staging slots), so the
 * copies of step t+1 / t-1 overlap the kernels of step t.  The host buffers of a call (events, offsets,
 * head_out - use pinned memory) must stay valid and unmodified, and head_out unread, until
 * aec_net_host_sync() returns; consecutive calls must use different head_out buffers.
 * aec_net_host_sync waits for everything enqueued and returns AEC_EEVENTS if any of those steps skipped
 * out-of-range events or over-long streams.
 */
int aec_net_step_host_async(aec_net *net, const int32_t *events_yxt, const int32_t *offsets, int total_events,
                            float *head_out, void *cuda_stream);
int aec_net_host_sync(aec_net *net, void *cuda_stream);

/* Device pointer / element count of the head buffer written by the last step. */
const float *aec_net_head_device(const aec_net *net);
size_t aec_net_head_elems_per_stream(const aec_net *net);
/*
 * Copies the last step's head of streams [first_stream, first_stream + n) to HOST memory
 * (float32 [n][H_last][W_last][C_last]) after waiting for `cuda_stream`: the read-back that goes with
 * aec_net_step_device (graph()'s return value, event_numpy.py:101-103, for a subset of the streams).
 */
int aec_net_read_head(aec_net *net, int first_stream, int n, float *host_out, void *cuda_stream);

/*
 * Layer-at-a-time interface mirroring Layer.compute(events, delta_leak) (layer.py:38-44), used by
 * the Python layer mirror and by test_correctness-style scripts:
 *   aec_net_begin_step  uploads host events and runs ONLY the integration layer (layer 0);
 *   aec_net_layer_compute(l) runs layer l (its leak pass if conv, its frontier update) consuming
 *                        the output events of layer l-1 left by the previous call.
 * Calling begin_step then layer_compute(1..L-1) in order is bit-identical to aec_net_step_*.
 */
int aec_net_begin_step(aec_net *net, const int32_t *events_yxt_host, const int32_t *offsets_host,
                       int total_events, void *cuda_stream);
int aec_net_layer_compute(aec_net *net, int layer, void *cuda_stream);
/* Writes the head buffer from the last layer's current state. */
int aec_net_compute_head(aec_net *net, void *cuda_stream);

/*
 * Synchronising read-back of one stream's state for one layer into HOST memory (`what` = AEC_READ_*;
 * layouts above).  `bytes` must equal the exact size; use aec_net_read_size() to query it.
 */
long long aec_net_read_size(const aec_net *net, int layer, int what);
int aec_net_read(aec_net *net, int layer, int what, int stream, void *host_out, size_t bytes);

/*
 * One stream's layer accessors for a conv or pool layer, evaluated on the device and copied to HOST
 * float32 [H][W][C] buffers (any may be NULL): surface() (layer.py:53-57; pool: previous surface read
 * through the stored argmax, maxpool.py:42-53), layer_actfn() (conv2d.py:83-88), conv_actfn()
 * (conv2d.py:90-94), featuremap() = surface*layer_actfn (layer.py:77-81).  Synchronises.
 */
int aec_net_read_view(aec_net *net, int layer, int stream, float *surface, float *layer_actfn, float *conv_actfn,
                      float *featuremap);

/* delta_leak (float64) and active flag (1 = stream had events) of the last step, HOST arrays [n_streams]. */
int aec_net_read_step_info(aec_net *net, double *delta_out, uint8_t *active_out);

/*
 * Work counters accumulated since the last call with reset != 0: `sites[l]` = number of sites
 * (conv) / windows (pool) re-evaluated by layer l summed over streams and steps, `steps` = steps
 * issued.  Used for the roofline's algorithmic-bytes figure.  Synchronises the device.
 */
int aec_net_read_counters(aec_net *net, unsigned long long *sites, int n_layers, unsigned long long *steps,
                          int reset);

/*
 * Per-launch timing inside the real step.  While enabled, aec_net_step_device/_host record a CUDA
 * event on the launching stream after every kernel launch and synchronise at the end of each step;
 * slot i of aec_net_read_profile() is the accumulated milliseconds of the i-th launch of a step
 * (order: surface, leak sweep, then per layer frontier + evaluation, head).  Returns the number of
 * slots.  Enabling/disabling clears the accumulators.
 */
int aec_net_profile(aec_net *net, int enable);
int aec_net_read_profile(aec_net *net, double *ms_per_slot, int n_slots, unsigned long long *steps);
/*
 * Name of the launch that profile slot `slot` timed, as recorded on the first profiled step: "surface",
 * "skip.frontier", "window_sweep", "leak_sweep", "all.frontier", "L<layer index>.eval" / "L<layer index>.frontier", "head".
 * Writes a NUL-terminated string into buf (capacity cap); returns its length, 0 for a slot that was never recorded.
 */
int aec_net_profile_slot_name(aec_net *net, int slot, char *buf, int cap);


/*
 * Measurement helper: counts the 16-byte groups of the conv leak-rate maps that hold a non-zero
 * rate (for which the leak sweep must read and write F) and the total number of groups.  Synchronises.
 */
int aec_net_count_nonzero_rate_groups(aec_net *net, unsigned long long *nz_groups, unsigned long long *total_groups);

/*
 * The step after the path: YOLO decode of the head left by the last step
 * (src/libs/viz.py:27-46 convert_bboxes with sqrt = True, and the decode half of draw_bboxes :131-148,165).
 * The head of every stream is read as [h_cells][w_cells][num_classes + 5*num_bbox].  HOST outputs, any may be
 * NULL: boxes float32 [n_streams][h_cells*w_cells*num_bbox][4] = (x, y, w, h) in pixels of an h_image x w_image
 * frame, conf float32 [..], label int32 [..] = argmax over classes of class*conf, valid uint8 [..] =
 * conf > conf_threshold.  Float32 arithmetic in the reference's order (bit-exact).  Synchronises.
 * Non-maximum suppression (src/libs/utils.py:38-118) stays on the host.
 */
int aec_net_decode_head(aec_net *net, int num_classes, int num_bbox, int h_cells, int w_cells, int h_image, int w_image,
                        float conf_threshold, float *boxes_out, float *conf_out, int32_t *label_out, uint8_t *valid_out,
                        void *cuda_stream);

/*
 * The step before the path (SURVEY 8f, f3 + f1): decodes N-MNIST / N-Caltech101 recordings - 5 bytes per event,
 * x, y, polarity bit + 23-bit timestamp, records with y == 240 are timestamp-overflow markers
 * (src/readers/file_reader.py:36-58) - and applies the runner's per-sample transform (src/libs/runner.py:24-33):
 * zero-based timestamps (zero_base_ts != 0) and the centre crop of src/libs/utils.py:4-28 to new_h x new_w
 * (crop != 0).  `raw` holds all recordings back to back, recording r = bytes [byte_offsets[r], byte_offsets[r+1]).
 * Outputs (HOST): events (y, x, ts) int32 triples and polarity (may be NULL); recording r's events start at
 * event index byte_offsets[r] / 5 and counts_out[r] of them are valid.  Stand-alone (no network object);
 * synchronises the device.
 */
int aec_decode_ndata(int device, const uint8_t *raw, const long long *byte_offsets, int n_recordings, int zero_base_ts,
                     int crop, int new_h, int new_w, int32_t *events_yxt_out, int32_t *polarity_out, int32_t *counts_out);

/*
 * The batching step of the runner on the device (src/libs/runner.py:65-72 with the intended semantics
 * of SURVEY Q5): splits every recording's events into the chunks one network step consumes.
 *   batch_event_usec == 0: max(ceil(N / batch_event_size), 1) chunks cut like np.array_split (runner.py:71-72);
 *   batch_event_usec  > 0: fixed-duration bins, np.digitize(ts, arange(0, ts[-1], usec)), a new chunk where the bin
 *                          changes (runner.py:66-69).
 * events_yxt (HOST, int32 [total][3]) holds the recordings back to back, recording r = events
 * [rec_offsets[r], rec_offsets[r+1]).  Outputs (HOST): n_chunks_out[r], and the chunk boundaries relative to the
 * recording's first event at chunk_offsets_out[rec_offsets[r] + 2 r + k], k = 0 .. n_chunks_out[r]
 * (chunk_offsets_out has total + 2 * n_recordings entries).  Stand-alone; synchronises the device.
 */
int aec_split_batches(int device, const int32_t *events_yxt, const long long *rec_offsets, int n_recordings,
                      int batch_event_size, int batch_event_usec, int32_t *chunk_offsets_out, int32_t *n_chunks_out);

/*
 * Raw recordings -> detections without a host round trip: what Runner.run does for one sample per stream
 * (runner.py:55-101) - data_transform (aec_decode_ndata's kernel), the batching above, then graph(events_batch, reset)
 * for every batch with reset only on the first (runner.py:64,101) - entirely on the device.  `raw` / `byte_offsets`
 * (HOST) hold exactly n_streams recordings, recording s feeds stream s; a stream whose recording has fewer batches
 * idles for the remaining steps.  head_out (HOST, may be NULL) receives the head after the last step
 * ([n_streams][H_last][W_last][C_last]); steps_out the number of steps run (the largest batch count);
 * event_counts_out (may be NULL) the events kept per recording.  Returns AEC_EEVENTS if a batch exceeded
 * max_events_per_step or an event fell outside the frame.  Synchronises `cuda_stream`.
 */
int aec_net_run_ndata(aec_net *net, const uint8_t *raw, const long long *byte_offsets, int zero_base_ts, int crop, int new_h,
                      int new_w, int batch_event_size, int batch_event_usec, int reset_first, float *head_out,
                      int32_t *steps_out, int32_t *event_counts_out, void *cuda_stream);

/*
 * Measurement helper for the leak sweep's roofline.  out8 = { 16-byte groups of the conv rate maps holding a
 * non-zero rate, all such groups, conv-map elements at sites whose non-zero-rate bit is set (live elements),
 * all conv-map elements, the same two for the pool layers' (Fp, Ap) copies, and the live conv / pool-copy
 * elements the sweep really touches: live sites minus the sites the last step re-evaluated anyway (the sweep
 * skips those: their F, A are overwritten in the same step, conv2d.py:118-123) }.  Synchronises.
 */
int aec_net_sweep_stats(aec_net *net, unsigned long long *out8);

/*
 * Measurement helper for the tensor-core conv kernel: with out16 == NULL, enables (enable != 0) or
 * disables per-role cycle accounting and clears the counters; with out16 != NULL copies the 16
 * counters of conv layer `layer` (summed over CTAs and launches since enabling: MMA-warp total /
 * waiting for accumulator, sites, weights; producer total / waiting for site info, stage; epilogue
 * total / waiting for accumulator, site info; loader total / waiting; CTAs; units).  Synchronises.
 */
int aec_net_tc_timing(aec_net *net, int enable, int layer, unsigned long long *out16);

/*
 * Measurement helper: how conv layer `layer` is mapped onto the tensor cores, for the "issued FLOPs" column of the
 * roofline table.  out8 = { 1 if the layer runs on the tcgen05 kernel else 0, sites per work unit,
 * tensor FLOPs ISSUED per unit over all weight tiles (every tcgen05.mma counted as 2*M*N*K with its padding rows
 * and all split-precision products), 8-wide K steps, MMAs per K step and weight tile, weight tiles, kernel variant
 * id (0 / 1 gathered kernel with the simple / batched decoder, 2 its sites-as-M form, 3 row tiles, 4 gathered kernel on
 * CTA pairs), 1 if units are counted by aec_net_read_unit_counters (row-tile kernel) }.  Returns 0, or a negative code for a bad layer index.
 */
int aec_net_tc_geometry(const aec_net *net, int layer, long long *out8);

/*
 * Measurement helper: work units evaluated per layer since the last counter reset, for the layers that run on the
 * row-tile kernel (aec_net_tc_geometry out8[7] == 1; a unit = 128 tile sites, work-set sites and gaps alike); for a
 * pool layer whose windows are partly evaluated inside the leak sweep (k_sweep_windows) the number of windows evaluated
 * there (they are part of aec_net_read_counters' count for that layer); 0 for every other layer.  Reset together with aec_net_read_counters(reset = 1).
 */
int aec_net_read_unit_counters(aec_net *net, unsigned long long *units, int n_layers);

/* Number of kernels this library has launched since creation of `net` (for bench `gpu_launches`). */
unsigned long long aec_net_launch_count(const aec_net *net);

#ifdef __cplusplus
}
#endif
#endif /* AEC_H */
