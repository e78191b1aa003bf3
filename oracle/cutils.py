"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/cutils_port.c (the C restatement of the
reference's Cython `src/libs/cutils.pyx`).  Same call shapes and return values as the reference's
`im2col_event` (cutils.pyx:29-30) and `min_argmax` (cutils.pyx:139-140) so the oracle layers read
like the reference's.  Never imported by the product path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcutils_port.so")
_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "cutils_port.c")
    if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", src, "-o", _SO])
    lib = ctypes.CDLL(_SO)
    f32p = ctypes.POINTER(ctypes.c_float)
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.oracle_im2col_event.restype = ctypes.c_int
    lib.oracle_im2col_event.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, i32p, i32p]
    lib.oracle_min_argmax.restype = None
    lib.oracle_min_argmax.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, i32p, i32p]
    _lib = lib
    return lib


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def im2col_event(image, event_y, event_x, k_height, k_width, stride, chan_as_cols=0):
    """-> (cols float32 F-order [C*k*k, n] or [k*k, C*n], (out_y int32[n], out_x int32[n]))."""
    lib = _load()
    image = np.ascontiguousarray(image, dtype=np.float32)
    event_y = np.ascontiguousarray(event_y, dtype=np.int32)
    event_x = np.ascontiguousarray(event_x, dtype=np.int32)
    chans, height, width = image.shape
    out_h = (height - k_height) // stride + 1
    out_w = (width - k_width) // stride + 1
    ksz = k_height * k_width
    if chan_as_cols:
        cols = np.empty((ksz, chans * out_h * out_w), dtype=np.float32, order="F")
    else:
        cols = np.empty((chans * ksz, out_h * out_w), dtype=np.float32, order="F")
    oy = np.empty(out_h * out_w, dtype=np.int32)
    ox = np.empty(out_h * out_w, dtype=np.int32)
    n = lib.oracle_im2col_event(_f32(image), chans, height, width, _i32(event_y), _i32(event_x), event_y.shape[0],
                                k_height, k_width, stride, 1 if chan_as_cols else 0, _f32(cols), _i32(oy), _i32(ox))
    if n == -1:
        raise NotImplementedError("This method only support stride equal to 1 or to the kernel's dimensions.")
    if n < 0:
        raise MemoryError("oracle_im2col_event: allocation failed")
    ncols = n * chans if chan_as_cols else n
    return cols[:, :ncols], (oy[:n].copy(), ox[:n].copy())


def min_argmax(max_arg, min_arg):
    """-> (argmax int32 [cols], not_argmin int32 [cols]) for two F-order float32 [rows, cols] matrices."""
    lib = _load()
    max_arg = np.asfortranarray(max_arg, dtype=np.float32)
    min_arg = np.asfortranarray(min_arg, dtype=np.float32)
    rows, cols = max_arg.shape
    amax = np.empty(cols, dtype=np.int32)
    nmin = np.empty(cols, dtype=np.int32)
    lib.oracle_min_argmax(_f32(max_arg), _f32(min_arg), rows, cols, _i32(amax), _i32(nmin))
    return amax, nmin
