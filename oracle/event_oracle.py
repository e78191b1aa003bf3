"""TEST INFRASTRUCTURE ONLY - CPU oracle for the event-driven EFCN hot path.

A numpy restatement of the reference's stateful event layers (marcocannici/async-ev-cnn), one
object per stream exactly like the reference.  Every function cites the reference file:line it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this; the product path (async-ev-cnn_b200/) never does and has no CPU fallback.

Parity pinned: tests/test_oracle_golden.py replays the committed golden vectors
(tests/golden/*.npz, minted by tests/golden/make_golden.py from the UNMODIFIED reference layers +
its compiled Cython module in this container) through this oracle and requires identical
frontiers / argmax indices / flags and float maps equal to 1e-6; when /root/reference is present
tests/test_oracle_vs_reference.py additionally runs both side by side on fresh random streams.

dtype policy: the reference leaves dtypes to NumPy's promotion rules; under the
NumPy >= 2 (NEP 50) of this image the integration surface becomes float64 after the first step
(integration.py:60,65) and the conv leak update is evaluated in float64 and rounded to float32
(conv2d.py:115).  The oracle writes those dtypes EXPLICITLY so it does not depend on the NumPy
version, and the CUDA path mirrors them.
"""
from collections import OrderedDict

import numpy as np

from . import cutils as _cutils

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------------------------
# dense helpers (reference: src/layers/functional.py:4-34, src/layers/conv2d.py:184-229,
#                src/libs/viz.py:7-24, src/models/frame_numpy.py:50-60)
# --------------------------------------------------------------------------------------------
def dense_im2col(image, kh, kw, stride=1):
    """All k x k windows as columns.  functional.py:4-34.

    image [C,H,W]   -> cols [C*kh*kw, Ho*Wo]
    image [B,C,H,W] -> cols [C*kh*kw, B*Ho*Wo]
    """
    image = np.ascontiguousarray(image)
    if image.ndim == 3:
        c, h, w = image.shape
        ho, wo = (h - kh) // stride + 1, (w - kw) // stride + 1
        shape = (c, kh, kw, ho, wo)
        steps = (h * w, w, 1, stride * w, stride)
        ncols = ho * wo
    else:
        b, c, h, w = image.shape
        ho, wo = (h - kh) // stride + 1, (w - kw) // stride + 1
        shape = (c, kh, kw, b, ho, wo)
        steps = (h * w, w, 1, c * h * w, stride * w, stride)
        ncols = b * ho * wo
    view = np.lib.stride_tricks.as_strided(image, shape=shape, strides=[s * image.itemsize for s in steps])
    cols = np.ascontiguousarray(view).reshape(c * kh * kw, ncols)
    return cols, (ho, wo)


def same_padding(size, k, stride):
    """TF-style SAME padding amounts (before, after).  conv2d.py:42-54 / :202-214."""
    total = max(k - stride, 0) if size % stride == 0 else max(k - (size % stride), 0)
    before = total // 2
    return before, total - before


def dense_conv2d(image, kernel_oihw, bias=None, padding="VALID", stride=1):
    """Dense convolution by im2col + GEMM.  conv2d.py:184-229.  image [C,H,W] -> [Cout,Ho,Wo]."""
    image = np.ascontiguousarray(image)
    cout, _, kh, kw = kernel_oihw.shape
    if padding == "SAME":
        pt, pb = same_padding(image.shape[1], kh, stride)
        pl, pr = same_padding(image.shape[2], kw, stride)
        image = np.pad(image, ((0, 0), (pt, pb), (pl, pr)), mode="constant")
    cols, (ho, wo) = dense_im2col(image, kh, kw, stride)
    out = kernel_oihw.reshape(cout, -1).dot(cols)
    if bias is not None:
        out = out + np.asarray(bias).reshape(cout, 1)
    return out.reshape(cout, ho, wo)


def dense_maxpool(x, kh, kw, stride):
    """Dense max pooling of [C,H,W] via argmax over window columns.  frame_numpy.py:50-60."""
    c, h, w = x.shape
    cols, (ho, wo) = dense_im2col(x.reshape(c, 1, h, w), kh, kw, stride)
    pick = np.argmax(cols, axis=0)
    return cols[pick, np.arange(c * ho * wo)].reshape(c, ho, wo)


def leaky(x, alpha):
    """max(x, alpha*x).  functional.py:37-47."""
    return np.maximum(x, x * alpha)


def integrate_frame(events, leak, frame_h, frame_w, prev=None):
    """Dense leaky frame with the same arithmetic / last-wins rule as the event surface.
    viz.py:7-24.  Returns (float32 frame [H,W], last_ts)."""
    ev = np.asarray(events)
    y, x, ts = ev[:, 0], ev[:, 1], ev[:, 2]
    if prev is None:
        frame, prev_ts = np.zeros((frame_h, frame_w), F32), 0
    else:
        frame, prev_ts = prev
    out = frame.astype(F64)                          # `f32 -= np.float64` evaluates in f64 (NEP 50) ...
    t_last = ts.max()
    out -= F64(int(t_last) - int(prev_ts)) * leak
    out = out.astype(F32)                            # ... and rounds back into the f32 frame (viz.py:19)
    out[out < 0] = 0
    bump = (1 - (t_last - ts) * leak)                # f64
    out[y, x] = (out[y, x].astype(F64) + bump).astype(F32)   # fancy `+=`: last duplicate wins (viz.py:21)
    out[out < 0] = 0
    return out, t_last


# --------------------------------------------------------------------------------------------
# stateful event layers (reference: src/layers/{layer,integration,conv2d,maxpool}.py)
# --------------------------------------------------------------------------------------------
class OracleLayer:
    """Contract of layer.py:2-81: reset / compute / compute_all / surface / layer_actfn /
    conv_actfn / out_shape / featuremap."""

    def featuremap(self):                            # layer.py:77-81
        return self.surface() * self.layer_actfn()

    def compute_all(self, events, delta_leak=None):  # conv2d.py:139-141, maxpool.py:163-165
        events, delta_leak = self.prev.compute_all(events, delta_leak)
        return self.compute(events, delta_leak)


class OracleIntegration(OracleLayer):
    """Leaky integration surface.  integration.py:12-95."""

    def __init__(self, leak, height, width):
        self.leak = float(leak)
        self.shape = [1, height, width]
        self.prev = None
        self.reset()

    def reset(self):                                 # integration.py:48-51
        self.prev_ts = 0
        self.S = np.zeros(self.shape, F64)
        self._mask = None

    def out_shape(self):
        return self.shape

    def surface(self):
        return self.S

    def layer_actfn(self):                           # integration.py:33-37
        if self._mask is None:
            self._mask = (self.S > 0).astype(F32)
        return self._mask

    conv_actfn = layer_actfn                         # integration.py:39-43 (same cached mask)

    def compute(self, events, _=None):               # integration.py:53-91
        ev = np.asarray(events)
        y, x, ts = ev[:, 0], ev[:, 1], ev[:, 2]
        t_last = ts.max()
        delta = F64(int(t_last) - int(self.prev_ts)) * self.leak           # :60
        alive_before = self.S > 0                                          # :63
        self.S = self.S - delta                                            # :65 (f64)
        died_leak = self.S <= 0
        self.S[died_leak] = 0                                              # :68
        self.S[:, y, x] += 1.0 - (t_last - ts) * self.leak                 # :71 last duplicate wins
        died_evt = self.S <= 0
        self.S[died_evt] = 0                                               # :74
        changed = alive_before & (died_leak | died_evt)                    # :78-79
        changed[:, y, x] = True                                            # :80
        ys, xs = np.nonzero(changed[0])                                    # :83 row-major
        self.prev_ts = t_last
        self._mask = None
        return (ys, xs), delta

    def compute_all(self, events, delta_leak=None):  # integration.py:93-95
        return self.compute(events, None)


class OracleConv(OracleLayer):
    """Event convolution: dense leak + full re-evaluation around input events.  conv2d.py:15-141."""

    def __init__(self, prev, kernel_hwio, bias, stride, alpha, padding="VALID"):
        self.prev = prev
        self.K = np.ascontiguousarray(np.asarray(kernel_hwio).transpose(3, 2, 0, 1))   # [co,ci,kh,kw] :26
        self.bias = np.asarray(bias)
        self.stride = stride
        self.alpha = float(alpha)
        cin, hin, win = prev.out_shape()
        cout, _, kh, kw = self.K.shape
        if padding == "VALID":                                             # :34-37
            ho = int(np.floor((hin - kh) / stride) + 1)
            wo = int(np.floor((win - kw) / stride) + 1)
            self.pad = (0, 0, 0, 0)
        elif padding == "SAME":                                            # :38-54
            ho, wo = int(np.ceil(hin / stride)), int(np.ceil(win / stride))
            pt, pb = same_padding(hin, kh, stride)
            pl, pr = same_padding(win, kw, stride)
            self.pad = (pt, pb, pl, pr)
        else:
            raise ValueError("'padding' must be either 'SAME' or 'VALID', but %s has been provided." % padding)
        self.shape = [cout, ho, wo]
        # init state = dense forward on the previous layer's initial map (conv2d.py:59-63; the
        # positional `stride` there lands in `padding` and is ignored because stride is 1, quirk Q7)
        self.F0 = dense_conv2d(self._padded(prev.surface() * prev.layer_actfn()), self.K, self.bias).astype(F32)
        self.A0 = np.zeros(self.shape, F32)
        self.reset()

    def _padded(self, fm):                           # conv2d.py:68-72
        pt, pb, pl, pr = self.pad
        if pt > 0 or pb > 0:
            return np.pad(fm, ((0, 0), (pt, pb), (pl, pr)), mode="constant")
        return fm

    def reset(self):                                 # conv2d.py:99-103
        self.F = self.F0.copy()
        self.A = self.A0.copy()
        self._slope = self._rate = None

    def out_shape(self):
        return self.shape

    def surface(self):
        return self.F

    def layer_actfn(self):                           # conv2d.py:83-88: 1 where F>0 else alpha
        if self._slope is None:
            pos = (self.F > 0).astype(F32)
            self._slope = pos + (1 - pos) * self.alpha
        return self._slope

    def conv_actfn(self):                            # conv2d.py:90-94
        if self._rate is None:
            self._rate = self.A * self.layer_actfn()
        return self._rate

    def _event_conv(self, image, ev, bias):          # conv2d.py:144-181
        image = np.ascontiguousarray(image).astype(F32)
        ey, ex = ev[0].astype(np.int32), ev[1].astype(np.int32)
        cout, _, kh, kw = self.K.shape
        cols, sites = _cutils.im2col_event(image, ey, ex, kh, kw, self.stride)
        out = self.K.reshape(cout, -1).dot(cols)
        if bias is not None:
            out = out + bias.reshape(cout, 1)
        return out.reshape(cout, -1), sites

    def compute(self, events, delta_leak):           # conv2d.py:105-137
        v_in = self._padded(self.prev.featuremap())
        r_in = self._padded(self.prev.conv_actfn())
        pt, pb, pl, pr = self.pad
        if pt > 0 or pb > 0:                                               # :74-78
            events = (events[0] + pt, events[1] + pl)
        sign_before = self.F >= 0                                          # :113
        # :115  `f32 -= f32 * np.float64`: product and difference in f64, result rounded to f32
        np.subtract(self.F, self.A.astype(F64) * F64(delta_leak), out=self.F, casting="same_kind")
        vals, (oy, ox) = self._event_conv(v_in, events, self.bias)         # :118-120
        self.F[:, oy, ox] = vals
        rates, (oy, ox) = self._event_conv(r_in, events, None)             # :121-123
        self.A[:, oy, ox] = rates
        flipped = np.any(sign_before != (self.F >= 0), axis=0)             # :126-128
        flipped[oy, ox] = True                                             # :130
        out_events = np.nonzero(flipped)                                   # :131 row-major
        self._slope = self._rate = None
        return out_events, delta_leak


class OraclePool(OracleLayer):
    """Event max-pool with stored argmax + sticky recompute flags.  maxpool.py:14-165."""

    def __init__(self, prev, ksize, stride):
        self.prev = prev
        self.kh, self.kw = ksize
        self.stride = stride
        c, hin, win = prev.out_shape()
        ho = int(np.floor((hin - self.kh) / stride) + 1)                   # :27-28
        wo = int(np.floor((win - self.kw) / stride) + 1)
        self.shape = [c, ho, wo]
        cols = self._window_cols(prev.surface())
        self.idx0 = cols.argmax(0).astype(np.int32)                        # :33 first maximum
        self.lin = np.arange(c * ho * wo, dtype=np.int32)                  # :34
        self.reset()

    def _window_cols(self, fm):                      # maxpool.py:45-50 (channels as batch)
        c, hin, win = self.prev.out_shape()
        cols, _ = dense_im2col(fm.reshape(c, 1, hin, win), self.kh, self.kw, self.stride)
        return cols

    def reset(self):                                 # maxpool.py:84-90
        self.idx = self.idx0.copy()
        self.flags = np.zeros(self.shape[1:], dtype=bool)
        self._cache = {}

    def out_shape(self):
        return self.shape

    def _picked(self, key, fm_fn):                   # maxpool.py:42-79
        if key not in self._cache:
            self._cache[key] = self._window_cols(fm_fn())[self.idx, self.lin].reshape(self.shape)
        return self._cache[key]

    def surface(self):
        return self._picked("s", self.prev.surface)

    def layer_actfn(self):
        return self._picked("l", self.prev.layer_actfn)

    def conv_actfn(self):
        return self._picked("c", self.prev.conv_actfn)

    def compute(self, events, delta_leak):           # maxpool.py:105-161
        fm_prev = self.prev.surface()                # PRE-activation map (:108)
        rate_prev = self.prev.conv_actfn()
        c, ho, wo = self.shape
        y, x = events
        self.flags[y // self.stride, x // self.stride] = False             # :118-120
        fy, fx = np.nonzero(self.flags)                                    # :123
        ey = np.concatenate([y, fy * self.stride]).astype(np.int32)        # :124-126
        ex = np.concatenate([x, fx * self.stride]).astype(np.int32)
        vcols, (oy, ox) = _cutils.im2col_event(fm_prev, ey, ex, self.kh, self.kw, self.stride, chan_as_cols=True)
        rcols, _ = _cutils.im2col_event(rate_prev, ey, ex, self.kh, self.kw, self.stride, chan_as_cols=True)
        amax, unstable = _cutils.min_argmax(np.asfortranarray(vcols), np.asfortranarray(rcols))   # :138
        unstable = np.any(unstable.reshape(-1, c).astype(bool), axis=1)    # :140 any channel
        self.flags[oy[unstable], ox[unstable]] = True                      # :142 sticky
        where = (oy * wo + ox).reshape(-1, 1) + np.arange(0, c * ho * wo, ho * wo)   # :145-150
        self.idx[where.reshape(-1)] = amax                                 # :151
        self._cache = {}
        return [oy, ox], delta_leak                  # ALL evaluated windows, first-touch order (:153-154)


# --------------------------------------------------------------------------------------------
# model builder (reference: src/models/event_numpy.py:53-105) and dense counterpart
# --------------------------------------------------------------------------------------------
def parse_layers(text):
    """'conv1=3,3,1,16 pool1=2,2 ...' -> OrderedDict (config.py:6-12)."""
    return OrderedDict((tok.split("=")[0], [int(v) for v in tok.split("=")[1].split(",")]) for tok in text.split(" "))


EFCN_LAYERS = ("conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,32 pool2=2,2 conv3=3,3,32,64 pool3=2,2 "
               "conv4=3,3,64,128 pool4=2,2 conv5=3,3,128,256 pool5=2,2 conv6=1,1,256,512 conv7=1,1,512,110")


class OracleEventNet:
    """One stream's chain Integration -> (Conv | Pool)*, as build_cnn_layers does
    (event_numpy.py:53-73), plus graph(events, reset) (event_numpy.py:94-103)."""

    def __init__(self, height, width, layers, weights, leak, alpha=0.1, padding="SAME"):
        if isinstance(layers, str):
            layers = parse_layers(layers)
        self.spec = layers
        self.names = ["intgr"]
        self.layers = [OracleIntegration(leak, height, width)]
        for name, size in layers.items():
            if "conv" in name:
                self.layers.append(OracleConv(self.layers[-1], weights["w_" + name], weights["b_" + name],
                                              1, alpha, padding))
            elif "pool" in name:
                self.layers.append(OraclePool(self.layers[-1], size, size[0]))
            else:
                raise NotImplementedError("non-event layer %r (EFCN has none, SURVEY 2 row 6)" % name)
            self.names.append(name)
        self.frontiers = [None] * len(self.layers)

    def reset(self):
        for layer in self.layers:
            layer.reset()

    def step(self, events, reset=False):
        """Runs every layer's compute explicitly (as test_correctness.py:33-39 does) and keeps each
        layer's output events.  Returns the last layer's feature map as [H,W,C] (event_numpy.py:79)."""
        if reset:
            self.reset()
        ev, delta = self.layers[0].compute(events, None)
        self.frontiers[0] = ev
        for i in range(1, len(self.layers)):
            ev, delta = self.layers[i].compute(ev, delta)
            self.frontiers[i] = ev
        self.delta = delta
        return self.layers[-1].featuremap().transpose(1, 2, 0)

    def frontier_mask(self, i):
        """Output-event set of layer i as a bool [H,W] mask (order-free)."""
        _, h, w = self.layers[i].out_shape()
        m = np.zeros((h, w), dtype=bool)
        ys, xs = self.frontiers[i]
        m[np.asarray(ys), np.asarray(xs)] = True
        return m


def dense_forward(frame, layers, weights, alpha=0.1, padding="SAME"):
    """Frame-mode network with ONE activation per conv and none after a pool - the semantics of the
    reference's TF model / test_correctness.py:73-80 that the event path equals.
    frame [H,W] -> list of per-layer feature maps [C,H,W] (post-activation for conv)."""
    if isinstance(layers, str):
        layers = parse_layers(layers)
    x = np.asarray(frame, F32)[None]
    outs = []
    for name, size in layers.items():
        if "conv" in name:
            k = np.ascontiguousarray(np.asarray(weights["w_" + name]).transpose(3, 2, 0, 1))
            x = leaky(dense_conv2d(x, k, np.asarray(weights["b_" + name]), padding), alpha).astype(F32)
        elif "pool" in name:
            x = dense_maxpool(x, size[0], size[1], size[0])
        outs.append(x)
    return outs
