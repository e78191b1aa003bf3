"""Seeded synthetic event streams and random-init EFCN weights.

There is no dataset and no checkpoint in the build environment, so both the GPU path and the CPU
oracle are driven by these generators.  Streams have the reference's event layout
`int32 [B,3] = (y, x, ts)` (src/libs/runner.py:32) with timestamps in microseconds, sorted.
"""
from collections import OrderedDict

import numpy as np

EFCN_LAYERS = ("conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,32 pool2=2,2 conv3=3,3,32,64 pool3=2,2 "
               "conv4=3,3,64,128 pool4=2,2 conv5=3,3,128,256 pool5=2,2 conv6=1,1,256,512 conv7=1,1,512,110")


def parse_layers(text):
    """'conv1=3,3,1,16 pool1=2,2 ...' -> OrderedDict name -> [ints]  (src/scripts/config.py:6-12)."""
    if not isinstance(text, str):
        return text
    return OrderedDict((tok.split("=")[0], [int(v) for v in tok.split("=")[1].split(",")])
                       for tok in text.split(" ") if tok)


def xavier_weights(layers, seed=0, bias=0.1, exact=False):
    """Random-init weights keyed like the reference checkpoint: `w_<name>` [kh,kw,ci,co], `b_<name>` [co]
    (src/models/event_numpy.py:64).  xavier-uniform + constant bias is the TF model's initialiser
    (src/models/frame_tf.py:76-78).

    exact=True draws weights from {-1,0,1} and bias 1.0; with a power-of-two leak and alpha every
    partial sum is exactly representable in float32, so any summation order gives the same bits and
    parity tests can demand bit-equality of the float maps too.
    """
    layers = parse_layers(layers)
    rng = np.random.default_rng(seed)
    out = {}
    for name, size in layers.items():
        if "conv" not in name:
            continue
        kh, kw, ci, co = size
        if exact:
            w = rng.integers(-1, 2, size=(kh, kw, ci, co)).astype(np.float32)
            b = np.full(co, 1.0, np.float32)
        else:
            lim = np.sqrt(6.0 / (kh * kw * (ci + co)))
            w = rng.uniform(-lim, lim, size=(kh, kw, ci, co)).astype(np.float32)
            b = np.full(co, bias, np.float32)
        out["w_" + name], out["b_" + name] = w, b
    return out


def synthetic_events(kind, n_streams, n_steps, batch, height, width, seed=0, rate=0.4, dt_int=None):
    """-> int32 [n_streams, n_steps, batch, 3] of (y, x, ts).

    kind 'uniform': y,x uniform over the frame (saturates the frontier).
    kind 'edge'   : a moving vertical edge - x within +-3 px of a column that sweeps 2 px per step,
                    y within +-30 rows of the centre (the realistic, clustered case).
    Timestamps: exponential gaps with mean 1/rate microseconds (rate ~0.4 ev/us ~ 1e5 events per
    300 ms N-Caltech101 recording), cumulated, floored to int32, so they are sorted within and
    across steps.  dt_int=(lo,hi) instead draws integer gaps in [lo,hi) (test_correctness.py:166).
    """
    rng = np.random.default_rng(seed)
    n = n_steps * batch
    if dt_int is None:
        gaps = rng.exponential(1.0 / rate, size=(n_streams, n))
        ts = np.floor(np.cumsum(gaps, axis=1)).astype(np.int64)
    else:
        gaps = rng.integers(dt_int[0], dt_int[1], size=(n_streams, n))
        ts = np.cumsum(gaps, axis=1).astype(np.int64)
    if ts.max() >= 2 ** 31:
        raise ValueError("timestamps overflow int32; shorten the stream")
    if kind == "uniform":
        y = rng.integers(0, height, size=(n_streams, n))
        x = rng.integers(0, width, size=(n_streams, n))
    elif kind == "edge":
        step = np.repeat(np.arange(n_steps), batch)[None, :]
        x0 = rng.integers(0, width, size=(n_streams, 1))
        col = (x0 + 2 * step) % width
        x = np.clip(col + rng.integers(-3, 4, size=(n_streams, n)), 0, width - 1)
        half = min(30, height // 2)
        y = np.clip(height // 2 + rng.integers(-half, half + 1, size=(n_streams, n)), 0, height - 1)
    else:
        raise ValueError("unknown stream kind %r" % kind)
    ev = np.stack([y, x, ts], axis=-1).astype(np.int32)
    return ev.reshape(n_streams, n_steps, batch, 3)
