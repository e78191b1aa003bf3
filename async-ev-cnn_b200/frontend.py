"""Device versions of the steps in front of the path (f3 + f1): decoding N-MNIST / N-Caltech101
recordings (src/readers/file_reader.py:30-58), the runner's per-sample transform (src/libs/runner.py:24-33,
src/libs/utils.py:4-28) and its batching of a sample into steps (runner.py:65-72), for a batch of recordings at
once (one CTA per recording); `run_recordings` chains them with the network steps without leaving the device."""
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


def _pack_raw(recordings):
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
    return raw, off


def split_batches(samples, batch_event_size=1, batch_event_usec=None, device=0):
    """Device version of runner.split_event_batches for a list of samples (int32 [N_r, 3] (y, x, ts) arrays):
    returns, per sample, the int32 array of chunk boundaries [0, ..., N_r] (chunk k = events[b[k]:b[k+1]])."""
    samples = [np.ascontiguousarray(np.asarray(e, np.int32).reshape(-1, 3)) for e in samples]
    off = np.zeros(len(samples) + 1, np.int64)
    np.cumsum([len(e) for e in samples], out=off[1:])
    total = int(off[-1])
    ev = np.concatenate(samples) if total else np.zeros((0, 3), np.int32)
    co = np.zeros(total + 2 * len(samples) + 1, np.int32)
    nc = np.zeros(max(len(samples), 1), np.int32)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
    N.check(N.lib().aec_split_batches(int(device), ptr(ev) if total else None, ptr(off), len(samples), int(batch_event_size),
                                      int(batch_event_usec or 0), ptr(co), ptr(nc)))
    return [co[int(off[r]) + 2 * r:int(off[r]) + 2 * r + int(nc[r]) + 1].copy() for r in range(len(samples))]


def run_recordings(net, recordings, crop_to=None, zero_base_ts=True, batch_event_size=1, batch_event_usec=None, reset=True,
                   cuda_stream=None):
    """Raw N-data recordings (one per stream of `net`, an EventNetCuda) -> (heads [S, H, W, C], steps run, events kept
    per recording): decode + transform + batching + one network step per batch, all on the device
    (aec_net_run_ndata; the counterpart of Runner.run's inner loops for one sample per stream)."""
    if len(recordings) != net.n_streams:
        raise ValueError("need exactly one recording per stream (%d), got %d" % (net.n_streams, len(recordings)))
    raw, off = _pack_raw(recordings)
    heads = np.empty((net.n_streams,) + net.head_shape, np.float32)
    steps = ctypes.c_int32(0)
    cnt = np.zeros(net.n_streams, np.int32)
    new_h, new_w = (int(crop_to[0]), int(crop_to[1])) if crop_to is not None else (0, 0)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
    try:
        N.check(N.lib().aec_net_run_ndata(net.handle, ptr(raw) if raw.size else None, ptr(off), 1 if zero_base_ts else 0,
                                          1 if crop_to is not None else 0, new_h, new_w, int(batch_event_size),
                                          int(batch_event_usec or 0), 1 if reset else 0, ptr(heads), ctypes.byref(steps), ptr(cnt),
                                          cuda_stream))
    except N.AecError as e:
        if e.code == N.AEC_EEVENTS:
            raise IndexError(str(e)) from None
        raise
    return heads, int(steps.value), cnt
