"""EventNetCuda - the multi-stream GPU engine behind the reference's event-mode model.

One reference network object is one stream (src/models/event_numpy.py:53-105); this engine holds
`n_streams` of them in one set of device arrays and advances all of them per call.  It is a thin
Python veneer over the C ABI (include/aec.h): numpy in, numpy out, no arithmetic on the host.
"""
import ctypes
import os

import numpy as np

from . import _native as N
from .streams import parse_layers


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def pack_events(per_stream):
    """list of int32 [B_s,3] arrays (None/empty = no events) -> (packed int32 [total,3], offsets int32 [S+1])."""
    lens = [0 if e is None else len(e) for e in per_stream]
    off = np.zeros(len(per_stream) + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    if off[-1] == 0:
        return np.zeros((0, 3), np.int32), off
    ev = np.concatenate([np.asarray(e, np.int32).reshape(-1, 3) for e in per_stream if e is not None and len(e)])
    return np.ascontiguousarray(ev), off


class EventNetCuda:
    """Integration -> (Conv | Pool)* chain for many streams on one GPU.

    layers   'conv1=3,3,1,16 pool1=2,2 ...' or an OrderedDict name -> sizes (src/scripts/config.py:6-12)
    weights  dict with 'w_<name>' [kh,kw,ci,co] and 'b_<name>' [co] (event_numpy.py:64)
    """

    def __init__(self, height, width, layers, weights, leak, alpha=0.1, padding="SAME", n_streams=1, device=0,
                 max_events_per_step=0):
        if padding not in ("SAME", "VALID"):
            raise ValueError("'padding' must be either 'SAME' or 'VALID', but %s has been provided." % padding)
        spec = []
        for name, size in parse_layers(layers).items():
            if "conv" in name:
                spec.append(("conv", name, weights["w_" + name], weights["b_" + name], alpha, padding))
            elif "pool" in name:
                spec.append(("pool", name, int(size[0]), int(size[1]), int(size[0])))
            else:
                raise NotImplementedError("non-event layer %r (the EFCN configs have none)" % name)
        self._build(height, width, leak, spec, n_streams, device, max_events_per_step)

    @classmethod
    def from_spec(cls, height, width, leak, spec, n_streams=1, device=0, max_events_per_step=0):
        """spec: list of ("conv", name, kernel_hwio, bias, alpha, padding) | ("pool", name, kh, kw, stride)."""
        self = cls.__new__(cls)
        self._build(height, width, leak, spec, n_streams, device, max_events_per_step)
        return self

    def _build(self, height, width, leak, spec, n_streams, device, max_events_per_step):
        self._lib = N.lib()
        self._h = ctypes.c_void_p()
        self.height, self.width = int(height), int(width)
        self.n_streams = int(n_streams)
        self.leak = float(leak)
        N.check(self._lib.aec_net_create(ctypes.byref(self._h), int(device), self.n_streams, self.height, self.width,
                                         self.leak, int(max_events_per_step)))
        self.names = ["intgr"]
        for item in spec:
            if item[0] == "conv":
                _, name, w, b, alpha, padding = item
                if padding not in ("SAME", "VALID"):
                    raise ValueError("'padding' must be either 'SAME' or 'VALID', but %s has been provided." % padding)
                w = np.ascontiguousarray(w, dtype=np.float32)
                b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
                kh, kw, ci, co = w.shape
                N.check(self._lib.aec_net_add_conv(self._h, kh, kw, ci, co, _ptr(w), _ptr(b), 1, float(alpha),
                                                   N.AEC_PAD_SAME if padding == "SAME" else N.AEC_PAD_VALID))
            else:
                _, name, kh, kw, stride = item
                N.check(self._lib.aec_net_add_pool(self._h, int(kh), int(kw), int(stride)))
            self.names.append(name)
        N.check(self._lib.aec_net_finalize(self._h))
        self.infos = []
        for i in range(len(self.names)):
            info = N.LayerInfo()
            N.check(self._lib.aec_net_layer_info(self._h, i, ctypes.byref(info)))
            self.infos.append(info)
        last = self.infos[-1]
        self.head_shape = (last.height, last.width, last.channels)
        self._head = np.empty((self.n_streams,) + self.head_shape, np.float32)

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.aec_net_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def shapes(self):
        return [[i.channels, i.height, i.width] for i in self.infos]

    def state_bytes_per_stream(self):
        return int(self._lib.aec_net_state_bytes_per_stream(self._h))

    def device_bytes(self):
        return int(self._lib.aec_net_device_bytes(self._h))

    def launch_count(self):
        return int(self._lib.aec_net_launch_count(self._h))

    # -- stepping -----------------------------------------------------------------------------
    def reset(self, stream_mask=None, cuda_stream=None):
        """Back to the initial state (conv2d.py:99-103, maxpool.py:84-90, integration.py:48-51): every stream, or the
        streams whose entry of `stream_mask` (length n_streams) is non-zero."""
        m = None
        if stream_mask is not None:
            m = np.ascontiguousarray(stream_mask, dtype=np.uint8)
            if m.shape != (self.n_streams,):        # the library copies n_streams bytes from this pointer
                raise ValueError("reset mask must have shape (%d,), got %s" % (self.n_streams, m.shape))
        N.check(self._lib.aec_net_reset(self._h, _ptr(m), cuda_stream))

    def _check_out(self, out):
        want = (self.n_streams,) + self.head_shape
        if not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != want:
            raise ValueError("`out` must be a C-contiguous float32 array of shape %s" % (want,))
        return out

    def _raise_events(self, err):
        if err.code == N.AEC_EEVENTS:
            raise IndexError(str(err)) from None
        raise err

    def step_packed(self, events, offsets, out=None, cuda_stream=None):
        """Host->device->host step: events int32 [total,3], offsets int32 [S+1] -> head [S,H,W,C] float32."""
        events = np.ascontiguousarray(events, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if offsets.shape != (self.n_streams + 1,):
            raise ValueError("offsets must have shape (%d,), got %s" % (self.n_streams + 1, offsets.shape))
        out = self._head if out is None else self._check_out(out)
        try:
            N.check(self._lib.aec_net_step_host(self._h, _ptr(events), _ptr(offsets), int(offsets[-1]), _ptr(out), cuda_stream))
        except N.AecError as e:
            self._raise_events(e)
        return out

    def step_packed_async(self, events, offsets, out, cuda_stream=None):
        """Pipelined host step (aec_net_step_host_async): returns once enqueued.  `events`, `offsets`, `out` must be
        C-contiguous int32 / int32 / float32 arrays (pinned for real overlap) that stay untouched until host_sync()."""
        if events.dtype != np.int32 or offsets.dtype != np.int32 or not events.flags.c_contiguous or not offsets.flags.c_contiguous:
            raise ValueError("events / offsets must be C-contiguous int32 arrays")
        if offsets.shape != (self.n_streams + 1,):
            raise ValueError("offsets must have shape (%d,), got %s" % (self.n_streams + 1, offsets.shape))
        self._check_out(out)
        N.check(self._lib.aec_net_step_host_async(self._h, _ptr(events), _ptr(offsets), int(offsets[-1]), _ptr(out), cuda_stream))

    def host_sync(self, cuda_stream=None):
        try:
            N.check(self._lib.aec_net_host_sync(self._h, cuda_stream))
        except N.AecError as e:
            self._raise_events(e)

    def step(self, per_stream_events, out=None):
        """per_stream_events: list (length n_streams) of int32 [B_s,3] arrays, or one array when n_streams == 1."""
        if isinstance(per_stream_events, np.ndarray) and per_stream_events.ndim == 2:
            per_stream_events = [per_stream_events]
        ev, off = pack_events(per_stream_events)
        return self.step_packed(ev, off, out)

    def step_device(self, events_ptr, offsets_ptr, total, cuda_stream=None):
        """Asynchronous step on device-resident events (raw device pointers as ints)."""
        N.check(self._lib.aec_net_step_device(self._h, ctypes.c_void_p(events_ptr), ctypes.c_void_p(offsets_ptr),
                                              int(total), cuda_stream))

    def decode_head(self, num_classes, num_bbox, h_cells, w_cells, conf_threshold=0.1, h_image=None, w_image=None):
        """YOLO decode of the last step's head on the device (viz.py:27-46,131-148): returns boxes [S, cells*B, 4]
        (x, y, w, h in pixels), conf [S, cells*B], valid bool, label int32 - what viz.draw_bboxes draws from."""
        nb = h_cells * w_cells * num_bbox
        boxes = np.empty((self.n_streams, nb, 4), np.float32)
        conf = np.empty((self.n_streams, nb), np.float32)
        label = np.empty((self.n_streams, nb), np.int32)
        valid = np.empty((self.n_streams, nb), np.uint8)
        N.check(self._lib.aec_net_decode_head(self._h, int(num_classes), int(num_bbox), int(h_cells), int(w_cells),
                                              int(h_image or self.height), int(w_image or self.width), float(conf_threshold),
                                              _ptr(boxes), _ptr(conf), _ptr(label), _ptr(valid), None))
        return boxes, conf, valid.astype(bool), label

    def head_device_ptr(self):
        return int(self._lib.aec_net_head_device(self._h))

    def read_head(self, first_stream=0, n=None, cuda_stream=None):
        """Host copy of the last step's head for streams [first_stream, first_stream + n) (after step_device)."""
        n = self.n_streams - first_stream if n is None else int(n)
        out = np.empty((n,) + self.head_shape, np.float32)
        N.check(self._lib.aec_net_read_head(self._h, int(first_stream), n, _ptr(out), cuda_stream))
        return out

    def begin_step(self, per_stream_events):
        if isinstance(per_stream_events, np.ndarray) and per_stream_events.ndim == 2:
            per_stream_events = [per_stream_events]
        ev, off = pack_events(per_stream_events)
        try:
            N.check(self._lib.aec_net_begin_step(self._h, _ptr(ev), _ptr(off), int(off[-1]), None))
        except N.AecError as e:
            self._raise_events(e)

    def layer_compute(self, layer):
        N.check(self._lib.aec_net_layer_compute(self._h, int(layer), None))

    def compute_head(self, stream=None):
        N.check(self._lib.aec_net_compute_head(self._h, None))

    # -- read-back (synchronising; parity tests and the Layer mirror) ------------------------------
    def read(self, layer, what, stream=0):
        size = N.check(self._lib.aec_net_read_size(self._h, layer, what))
        info = self.infos[layer]
        h, w, c, ww = info.height, info.width, info.channels, info.frontier_words_per_row
        if what == N.AEC_READ_SURFACE:
            buf = np.empty((h, w), np.float64)
        elif what in (N.AEC_READ_F, N.AEC_READ_A, N.AEC_READ_INIT_F):
            buf = np.empty((h, w, c), np.float32)
        elif what in (N.AEC_READ_IDX, N.AEC_READ_INIT_IDX):
            buf = np.empty((h, w, c), np.uint8)
        else:
            buf = np.empty((h, ww), np.uint32)
        assert buf.nbytes == size
        N.check(self._lib.aec_net_read(self._h, layer, what, stream, _ptr(buf), buf.nbytes))
        return buf

    def _bitmap_to_mask(self, bm, width):
        bits = np.unpackbits(bm.view(np.uint8), axis=1, bitorder="little")
        return bits[:, :width].astype(bool)

    def frontier(self, layer, stream=0):
        """bool [H,W] mask of the layer's output events of the last step."""
        return self._bitmap_to_mask(self.read(layer, N.AEC_READ_FRONTIER, stream), self.infos[layer].width)

    def state(self, layer, stream=0):
        """State in the reference's layouts: {'S'} | {'F','A'} [C,H,W] | {'idx' [C,Ho,Wo], 'flags' [Ho,Wo]}."""
        t = self.infos[layer].type
        if t == N.AEC_LAYER_INTEGRATION:
            return {"S": self.read(layer, N.AEC_READ_SURFACE, stream)}
        if t == N.AEC_LAYER_CONV:
            return {"F": np.ascontiguousarray(self.read(layer, N.AEC_READ_F, stream).transpose(2, 0, 1)),
                    "A": np.ascontiguousarray(self.read(layer, N.AEC_READ_A, stream).transpose(2, 0, 1))}
        return {"idx": np.ascontiguousarray(self.read(layer, N.AEC_READ_IDX, stream).transpose(2, 0, 1)),
                "flags": self._bitmap_to_mask(self.read(layer, N.AEC_READ_FLAGS, stream), self.infos[layer].width)}

    def init_state(self, layer):
        t = self.infos[layer].type
        if t == N.AEC_LAYER_CONV:
            return {"F": np.ascontiguousarray(self.read(layer, N.AEC_READ_INIT_F, 0).transpose(2, 0, 1))}
        if t == N.AEC_LAYER_POOL:
            return {"idx": np.ascontiguousarray(self.read(layer, N.AEC_READ_INIT_IDX, 0).transpose(2, 0, 1))}
        return {}

    def view(self, layer, stream=0, which=("surface", "layer_actfn", "conv_actfn", "featuremap")):
        """Layer accessors evaluated on the device, returned [C,H,W] float32 (layer.py:53-81)."""
        info = self.infos[layer]
        bufs = {k: np.empty((info.height, info.width, info.channels), np.float32) for k in which}
        N.check(self._lib.aec_net_read_view(self._h, layer, stream, _ptr(bufs.get("surface")), _ptr(bufs.get("layer_actfn")),
                                            _ptr(bufs.get("conv_actfn")), _ptr(bufs.get("featuremap"))))
        return {k: np.ascontiguousarray(v.transpose(2, 0, 1)) for k, v in bufs.items()}

    def step_info(self):
        delta = np.empty(self.n_streams, np.float64)
        active = np.empty(self.n_streams, np.uint8)
        N.check(self._lib.aec_net_read_step_info(self._h, _ptr(delta), _ptr(active)))
        return delta, active

    def counters(self, reset=False):
        """(sites re-evaluated per layer summed over streams and steps, steps issued) since the last reset."""
        sites = np.zeros(len(self.names), np.uint64)
        steps = ctypes.c_ulonglong(0)
        N.check(self._lib.aec_net_read_counters(self._h, _ptr(sites), len(self.names), ctypes.byref(steps), 1 if reset else 0))
        return sites, int(steps.value)


    # -- measurement ---------------------------------------------------------------------------
    def slot_names(self, n_slots=64):
        """Names of the launches of the last profiled step, in order, as the library recorded them
        (aec_net_profile_slot_name): "L<i>.eval" becomes "<layer name>.eval"."""
        out = []
        buf = ctypes.create_string_buffer(64)
        for i in range(n_slots):
            if N.check(self._lib.aec_net_profile_slot_name(self._h, i, buf, 64)) == 0:
                break
            nm = buf.value.decode()
            if nm.startswith("L") and "." in nm and nm[1:nm.index(".")].isdigit():
                nm = self.names[int(nm[1:nm.index(".")])] + nm[nm.index("."):]
            out.append(nm)
        return out

    def profile(self, enable=True):
        N.check(self._lib.aec_net_profile(self._h, 1 if enable else 0))

    def read_profile(self):
        """-> (dict slot name -> mean ms per step, steps profiled)."""
        names = self.slot_names()
        # slot i = time between mark i and mark i+1; mark 0 precedes the surface kernel
        ms = np.zeros(len(names) + 4, np.float64)
        steps = ctypes.c_ulonglong(0)
        n = N.check(self._lib.aec_net_read_profile(self._h, _ptr(ms), len(ms), ctypes.byref(steps)))
        k = max(1, int(steps.value))
        return {names[i]: ms[i] / k for i in range(min(n, len(names)))}, int(steps.value)

    def sweep_stats(self):
        """Leak-sweep work of the current state: dict with the fraction of conv-map 16-byte groups whose rate is
        non-zero, the conv / pool-copy elements at live (non-zero-rate bit set) sites, and the live elements the sweep
        really touches (`swept_*`: live sites minus those the last step re-evaluated anyway, which the sweep skips)."""
        buf = np.zeros(8, np.uint64)
        N.check(self._lib.aec_net_sweep_stats(self._h, _ptr(buf)))
        v = [int(x) for x in buf]
        return {"nz_groups": v[0], "groups": v[1], "live_conv_elems": v[2], "conv_elems": v[3], "live_pool_elems": v[4],
                "pool_elems": v[5], "swept_conv_elems": v[6], "swept_pool_elems": v[7]}

    def tc_layers(self):
        """Indices of the conv layers that run on the tensor-core kernel (previous layer a map with C % 4 == 0, kh*kw <= 32)."""
        out = []
        for i, info in enumerate(self.infos):
            if info.type == N.AEC_LAYER_CONV and i > 1 and info.in_channels % 4 == 0 and info.k_h * info.k_w <= 32:
                out.append(i)
        return out

    def tc_geometry(self, layer):
        """How conv layer `layer` maps onto the tensor cores (aec_net_tc_geometry), or None for a SIMT layer."""
        buf = np.zeros(8, np.int64)
        N.check(self._lib.aec_net_tc_geometry(self._h, int(layer), _ptr(buf)))
        if not buf[0]:
            return None
        return {"unit_sites": int(buf[1]), "mma_flops_per_unit": float(buf[2]), "m_groups": 1, "k8_steps": int(buf[3]),
                "mma_per_kstep": int(buf[4]), "weight_tiles": int(buf[5]),
                "units_counted": bool(buf[7]),
                "kernel": ("k_conv_eval_tc<simple decode, weights-as-M>", "k_conv_eval_tc<batched decode, weights-as-M>",
                           "k_conv_eval_tc<sites-as-M>", "k_conv_rows (row tiles, sites-as-M)",
                           "k_conv_eval_tc<CTA pairs, cta_group::2 M = 256, weights-as-M>")[int(buf[6])]}

    def unit_counters(self):
        """Work units evaluated per layer by the row-tile kernel since the last counters(reset=True)."""
        units = np.zeros(len(self.names), np.uint64)
        N.check(self._lib.aec_net_read_unit_counters(self._h, _ptr(units), len(self.names)))
        return units

    TC_TIMING_SLOTS = ("mma_total", "mma_wait_acc", "mma_wait_sites", "mma_wait_weights", "prod_total", "prod_wait_siteinfo",
                       "prod_wait_stage", "epi_total", "epi_wait_acc", "epi_wait_siteinfo", "load_total", "load_wait", "ctas", "units",
                       "gate_wait_sites", "mma_section")

    def tc_timing(self, enable=True):
        N.check(self._lib.aec_net_tc_timing(self._h, 1 if enable else 0, 0, None))

    def read_tc_timing(self):
        """-> {layer name: {slot: cycles summed over CTAs and launches}} for the tensor-core conv layers."""
        out = {}
        for i, nm in enumerate(self.names):
            buf = np.zeros(16, np.uint64)
            rc = self._lib.aec_net_tc_timing(self._h, 0, i, _ptr(buf))
            if rc == 0:
                out[nm] = {k: int(buf[j]) for j, k in enumerate(self.TC_TIMING_SLOTS)}
        return out

    def nonzero_rate_fraction(self):
        nz, tot = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        N.check(self._lib.aec_net_count_nonzero_rate_groups(self._h, ctypes.byref(nz), ctypes.byref(tot)))
        return nz.value / max(1, tot.value), int(tot.value)


class CudaAdapter:
    """tests/parity.py adapter over one stream of an EventNetCuda (other streams get the same events
    when `mirror` is set, to exercise the multi-stream path)."""

    def __init__(self, net, stream=0, mirror=True):
        self.net, self.stream, self.mirror = net, stream, mirror
        self.names = net.names

    def step(self, events):
        if self.mirror:
            per = [events] * self.net.n_streams
        else:
            per = [None] * self.net.n_streams
            per[self.stream] = events
        return self.net.step(per)[self.stream].copy()

    def delta(self):
        return self.net.step_info()[0][self.stream]

    def frontier(self, i):
        return self.net.frontier(i, self.stream)

    def state(self, i):
        return self.net.state(i, self.stream)
