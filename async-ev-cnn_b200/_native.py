"""ctypes binding of libaec_b200.so (include/aec.h).  There is NO fallback: if the library is missing
or a call fails, an exception is raised - the product path never computes on the CPU."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libaec_b200.so")

AEC_LAYER_INTEGRATION, AEC_LAYER_CONV, AEC_LAYER_POOL = 0, 1, 2
AEC_PAD_VALID, AEC_PAD_SAME = 0, 1
(AEC_READ_SURFACE, AEC_READ_F, AEC_READ_A, AEC_READ_IDX, AEC_READ_FLAGS, AEC_READ_FRONTIER, AEC_READ_INIT_F,
 AEC_READ_INIT_IDX) = range(8)
AEC_EINVAL, AEC_ECUDA, AEC_ESTATE, AEC_EEVENTS, AEC_ENOMEM = -1, -2, -3, -4, -5

# every symbol include/aec.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "aec_last_error", "aec_version", "aec_net_create", "aec_net_add_conv", "aec_net_add_pool", "aec_net_finalize",
    "aec_net_destroy", "aec_net_num_layers", "aec_net_num_streams", "aec_net_layer_info",
    "aec_net_state_bytes_per_stream", "aec_net_device_bytes", "aec_net_reset", "aec_net_step_device",
    "aec_net_step_host", "aec_net_head_device", "aec_net_head_elems_per_stream", "aec_net_read_head", "aec_net_begin_step",
    "aec_net_layer_compute", "aec_net_compute_head", "aec_net_read_size", "aec_net_read", "aec_net_read_step_info",
    "aec_net_read_counters", "aec_net_launch_count", "aec_net_read_view", "aec_net_profile", "aec_net_read_profile", "aec_net_profile_slot_name",
    "aec_net_count_nonzero_rate_groups", "aec_net_tc_timing", "aec_net_tc_geometry", "aec_net_read_unit_counters", "aec_net_sweep_stats", "aec_net_step_host_async", "aec_net_host_sync", "aec_net_decode_head", "aec_decode_ndata", "aec_split_batches", "aec_net_run_ndata",
]


class AecError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libaec_b200 error %d: %s" % (code, message))
        self.code = code


class LayerInfo(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("channels", ctypes.c_int), ("height", ctypes.c_int), ("width", ctypes.c_int),
                ("k_h", ctypes.c_int), ("k_w", ctypes.c_int), ("stride", ctypes.c_int), ("pad_top", ctypes.c_int),
                ("pad_left", ctypes.c_int), ("in_channels", ctypes.c_int), ("frontier_words_per_row", ctypes.c_int)]


_lib = None


def lib():
    """Loads libaec_b200.so (building is `python -m async_ev_cnn_b200.build`); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError("%s not found: build it with `python -m async_ev_cnn_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i, sz, ull = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_ulonglong
    L.aec_last_error.restype = ctypes.c_char_p
    L.aec_last_error.argtypes = []
    L.aec_version.restype = i
    L.aec_net_create.restype = i
    L.aec_net_create.argtypes = [ctypes.POINTER(vp), i, i, i, i, ctypes.c_double, i]
    L.aec_net_add_conv.restype = i
    L.aec_net_add_conv.argtypes = [vp, i, i, i, i, vp, vp, i, ctypes.c_float, i]
    L.aec_net_add_pool.restype = i
    L.aec_net_add_pool.argtypes = [vp, i, i, i]
    L.aec_net_finalize.restype = i
    L.aec_net_finalize.argtypes = [vp]
    L.aec_net_destroy.restype = None
    L.aec_net_destroy.argtypes = [vp]
    L.aec_net_num_layers.restype = i
    L.aec_net_num_layers.argtypes = [vp]
    L.aec_net_num_streams.restype = i
    L.aec_net_num_streams.argtypes = [vp]
    L.aec_net_layer_info.restype = i
    L.aec_net_layer_info.argtypes = [vp, i, ctypes.POINTER(LayerInfo)]
    L.aec_net_state_bytes_per_stream.restype = sz
    L.aec_net_state_bytes_per_stream.argtypes = [vp]
    L.aec_net_device_bytes.restype = sz
    L.aec_net_device_bytes.argtypes = [vp]
    L.aec_net_reset.restype = i
    L.aec_net_reset.argtypes = [vp, vp, vp]
    L.aec_net_step_device.restype = i
    L.aec_net_step_device.argtypes = [vp, vp, vp, i, vp]
    L.aec_net_step_host.restype = i
    L.aec_net_step_host.argtypes = [vp, vp, vp, i, vp, vp]
    L.aec_net_step_host_async.restype = i
    L.aec_net_step_host_async.argtypes = [vp, vp, vp, i, vp, vp]
    L.aec_net_host_sync.restype = i
    L.aec_net_host_sync.argtypes = [vp, vp]
    L.aec_net_decode_head.restype = i
    L.aec_net_decode_head.argtypes = [vp, i, i, i, i, i, i, ctypes.c_float, vp, vp, vp, vp, vp]
    L.aec_decode_ndata.restype = i
    L.aec_decode_ndata.argtypes = [i, vp, vp, i, i, i, i, i, vp, vp, vp]
    L.aec_split_batches.restype = i
    L.aec_split_batches.argtypes = [i, vp, vp, i, i, i, vp, vp]
    L.aec_net_run_ndata.restype = i
    L.aec_net_run_ndata.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, vp, vp, vp, vp]
    L.aec_net_head_device.restype = vp
    L.aec_net_head_device.argtypes = [vp]
    L.aec_net_head_elems_per_stream.restype = sz
    L.aec_net_head_elems_per_stream.argtypes = [vp]
    L.aec_net_read_head.restype = i
    L.aec_net_read_head.argtypes = [vp, i, i, vp, vp]
    L.aec_net_begin_step.restype = i
    L.aec_net_begin_step.argtypes = [vp, vp, vp, i, vp]
    L.aec_net_layer_compute.restype = i
    L.aec_net_layer_compute.argtypes = [vp, i, vp]
    L.aec_net_compute_head.restype = i
    L.aec_net_compute_head.argtypes = [vp, vp]
    L.aec_net_read_size.restype = ctypes.c_longlong
    L.aec_net_read_size.argtypes = [vp, i, i]
    L.aec_net_read.restype = i
    L.aec_net_read.argtypes = [vp, i, i, i, vp, sz]
    L.aec_net_read_step_info.restype = i
    L.aec_net_read_step_info.argtypes = [vp, vp, vp]
    L.aec_net_read_counters.restype = i
    L.aec_net_read_counters.argtypes = [vp, vp, i, ctypes.POINTER(ull), i]
    L.aec_net_read_view.restype = i
    L.aec_net_read_view.argtypes = [vp, i, i, vp, vp, vp, vp]
    L.aec_net_profile.restype = i
    L.aec_net_profile.argtypes = [vp, i]
    L.aec_net_read_profile.restype = i
    L.aec_net_read_profile.argtypes = [vp, vp, i, ctypes.POINTER(ull)]
    L.aec_net_profile_slot_name.restype = i
    L.aec_net_profile_slot_name.argtypes = [vp, i, ctypes.c_char_p, i]
    L.aec_net_count_nonzero_rate_groups.restype = i
    L.aec_net_count_nonzero_rate_groups.argtypes = [vp, ctypes.POINTER(ull), ctypes.POINTER(ull)]
    L.aec_net_tc_timing.restype = i
    L.aec_net_tc_timing.argtypes = [vp, i, i, vp]
    L.aec_net_tc_geometry.restype = i
    L.aec_net_tc_geometry.argtypes = [vp, i, vp]
    L.aec_net_read_unit_counters.restype = i
    L.aec_net_read_unit_counters.argtypes = [vp, vp, i]
    L.aec_net_sweep_stats.restype = i
    L.aec_net_sweep_stats.argtypes = [vp, vp]
    L.aec_net_launch_count.restype = ull
    L.aec_net_launch_count.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    """Raises AecError for a negative return code, passes non-negative values through."""
    if rc < 0:
        msg = lib().aec_last_error()
        raise AecError(int(rc), msg.decode("utf-8", "replace") if msg else "")
    return rc
