"""Device versions of the steps in front of the path (f3 + f1): decoding N-MNIST / N-Caltech101
recordings (src/readers/file_reader.py:30-58) and the runner's per-sample transform (src/libs/runner.py:24-33,
src/libs/utils.py:4-28), for a batch of recordings at once (one CTA per recording)."""
import ctypes

import numpy as np

from . import _native as N


def decode_ndata(recordings, zero_base_ts=True, crop_to=None, device=0, with_polarity=False):
    """recordings: list of uint8 arrays (raw file contents, 5 bytes per event) or file names.
    crop_to: (frame_h, frame_w) for the reference's centre crop, or None.
    Returns a list of int32 [n_r, 3] (y, x, ts) arrays, one per recording (and the polarities if asked)."""
    raws = []
    for r in recordings:
        if isinstance(r, (str, bytes)) and not isinstance(r, np.ndarray):
            r = np.fromfile(r, dtype=np.uint8)
        r = np.ascontiguousarray(r, dtype=np.uint8).reshape(-1)
        if r.size % 5:
            raise ValueError("a recording is not a whole number of 5-byte records")
        raws.append(r)
    off = np.zeros(len(raws) + 1, np.int64)
    np.cumsum([r.size for r in raws], out=off[1:])
    raw = np.concatenate(raws) if raws and off[-1] else np.zeros(0, np.uint8)
    n_ev = int(off[-1] // 5)
    ev = np.empty((max(n_ev, 1), 3), np.int32)
    pol = np.empty(max(n_ev, 1), np.int32) if with_polarity else None
    cnt = np.zeros(max(len(raws), 1), np.int32)
    new_h, new_w = (int(crop_to[0]), int(crop_to[1])) if crop_to is not None else (0, 0)
    lib = N.lib()
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
    N.check(lib.aec_decode_ndata(int(device), ptr(raw) if raw.size else None, ptr(off), len(raws), 1 if zero_base_ts else 0,
                                 1 if crop_to is not None else 0, new_h, new_w, ptr(ev), ptr(pol), ptr(cnt)))
    out, pols = [], []
    for r in range(len(raws)):
        a = int(off[r] // 5)
        out.append(ev[a:a + int(cnt[r])].copy())
        if with_polarity:
            pols.append(pol[a:a + int(cnt[r])].copy())
    return (out, pols) if with_polarity else out
