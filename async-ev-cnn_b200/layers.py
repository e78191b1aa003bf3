"""GPU mirror of the reference's layer classes (src/layers/{layer,integration,conv2d,maxpool}.py).

Same constructors, same methods, same return conventions, so scripts written against the
reference (e.g. src/scripts/test_correctness.py:18-39) run unchanged on the B200 backend:

    intgr = IntegrationLayer(leak, frame_height, frame_width)
    conv1 = Conv2DLayer(intgr, k1, np.array([b1]), 1, alpha, "SAME")
    pool1 = MaxPoolLayer(conv1, [2, 2], 2)
    ev, d = intgr.compute(events, None); ev, d = conv1.compute(ev, d); ev, d = pool1.compute(ev, d)
    pool1.featuremap()

The layer objects hold no arithmetic: they are views onto one EventNetCuda (engine.py) that is built
when the chain is first used; every compute() launches the layer's CUDA kernels through the C ABI
and every accessor reads device state back (evaluated on the device by k_layer_view).  One chain is
one stream, as in the reference (pass n_streams/stream to IntegrationLayer for the batched form).

Differences, all documented in DESIGN.md: event lists returned by compute() are row-major sorted
for every layer (the reference's pool returns first-touch order - only the set matters downstream,
SURVEY Q8); compute(events, ...) of a conv/pool layer consumes the output events its predecessor
left on the device and, in strict mode, checks that `events` is that same set.
"""
import numpy as np

from . import _native as N
from .engine import EventNetCuda


class Layer:
    """Contract of src/layers/layer.py:2-81."""

    def reset(self):
        raise NotImplementedError('Subclasses must override reset()')

    def compute(self, events, delta_leak):
        raise NotImplementedError('Subclasses must override compute()')

    def compute_all(self, events, delta_leak=None):
        raise NotImplementedError('Subclasses must override compute_all()')

    def surface(self):
        raise NotImplementedError('Subclasses must override surface()')

    def layer_actfn(self):
        raise NotImplementedError('Subclasses must override layer_actfn()')

    def conv_actfn(self):
        raise NotImplementedError('Subclasses must override conv_actfn()')

    def out_shape(self):
        raise NotImplementedError('Subclasses must override out_shape()')

    def featuremap(self):
        return self.surface() * self.layer_actfn()


class _Chain:
    """The layers appended so far + the engine built from them on first use."""

    def __init__(self, leak, height, width, n_streams, device, stream, strict):
        self.leak, self.height, self.width = leak, height, width
        self.n_streams, self.device, self.stream, self.strict = n_streams, device, stream, strict
        self.spec = []
        self.layers = []
        self.engine = None
        self.pending_reset = set()
        self.delta = None

    def add(self, layer, item):
        if self.engine is not None:       # the chain grew after it was used: rebuild from scratch
            self.engine.close()
            self.engine = None
        self.layers.append(layer)
        if item is not None:
            self.spec.append(item)
        return len(self.layers) - 1

    def net(self):
        if self.engine is None:
            if not self.spec:
                raise RuntimeError("the B200 backend needs at least one conv layer after the IntegrationLayer")
            self.engine = EventNetCuda.from_spec(self.height, self.width, self.leak, self.spec, self.n_streams, self.device)
        return self.engine

    def flush_reset(self):
        if not self.pending_reset:
            return
        if len(self.pending_reset) != len(self.layers):
            raise NotImplementedError("the B200 backend resets whole chains: call reset() on every layer "
                                      "(as graph(events, reset=True) does, event_numpy.py:96-98)")
        mask = np.zeros(self.n_streams, np.uint8)
        mask[self.stream] = 1
        self.net().reset(mask)
        self.pending_reset.clear()

    def events_of(self, index):
        ys, xs = np.nonzero(self.net().frontier(index, self.stream))
        return ys, xs


class _GpuLayer(Layer):
    def out_shape(self):
        info = self._chain.net().infos[self._index]
        return [info.channels, info.height, info.width]

    def reset(self):
        self._chain.pending_reset.add(self._index)
        if len(self._chain.pending_reset) == len(self._chain.layers):
            self._chain.flush_reset()

    def compute_all(self, events, delta_leak=None):       # conv2d.py:139-141, maxpool.py:163-165
        events, delta_leak = self._prev_layer.compute_all(events, delta_leak)
        return self.compute(events, delta_leak)

    def compute(self, events, delta_leak):
        ch = self._chain
        ch.flush_reset()
        if ch.strict and events is not None:
            ys, xs = ch.events_of(self._index - 1)
            got = set(zip(np.asarray(events[0]).tolist(), np.asarray(events[1]).tolist()))
            if got != set(zip(ys.tolist(), xs.tolist())):
                raise ValueError("compute(): `events` is not the output of the previous layer's last compute(); "
                                 "the B200 backend consumes the events its predecessor left on the device")
        ch.net().layer_compute(self._index)
        return ch.events_of(self._index), delta_leak

    def _view(self, which):
        return self._chain.net().view(self._index, self._chain.stream, which=(which,))[which]

    def surface(self):
        return self._view("surface")

    def layer_actfn(self):
        return self._view("layer_actfn")

    def conv_actfn(self):
        return self._view("conv_actfn")

    def featuremap(self):
        return self._view("featuremap")


class IntegrationLayer(_GpuLayer):
    """Leaky integration surface (integration.py:12-95).  Extra keyword arguments select the batched
    form: `n_streams` chains share one engine and this object views stream `stream`."""

    def __init__(self, leak, h_surface, w_surface, n_streams=1, device=0, stream=0, strict=True):
        self._leak = leak
        self._chain = _Chain(float(leak), int(h_surface), int(w_surface), n_streams, device, stream, strict)
        self._index = self._chain.add(self, None)
        self._prev_layer = None

    def out_shape(self):
        return [1, self._chain.height, self._chain.width]

    def surface(self):
        s = self._chain.net().read(0, N.AEC_READ_SURFACE, self._chain.stream)
        return s.reshape(1, *s.shape)

    def layer_actfn(self):                      # integration.py:33-37
        return (self.surface() > 0).astype(np.float32)

    conv_actfn = layer_actfn                    # integration.py:39-43

    def featuremap(self):
        return self.surface() * self.layer_actfn()

    def compute(self, events, _=None):          # integration.py:53-91
        ch = self._chain
        ch.flush_reset()
        per = [None] * ch.n_streams
        per[ch.stream] = np.ascontiguousarray(events, dtype=np.int32)
        ch.net().begin_step(per)
        delta = ch.net().step_info()[0][ch.stream]
        return ch.events_of(0), np.float64(delta)

    def compute_all(self, events, delta_leak=None):     # integration.py:93-95
        return self.compute(events, None)


class Conv2DLayer(_GpuLayer):
    """Event convolution (conv2d.py:15-141): kernel float32 [k_h,k_w,c_in,c_out], bias [c_out]."""

    def __init__(self, prev_layer, kernel, bias, stride, alpha, padding='VALID'):
        if padding not in ('SAME', 'VALID'):
            raise ValueError("'padding' must be either 'SAME' or 'VALID', but %s has been provided." % padding)
        if stride != 1:
            raise NotImplementedError("the B200 backend supports stride 1 convolutions (event_numpy.py:64 always passes 1)")
        self._prev_layer = prev_layer
        self._chain = prev_layer._chain
        self._kernel = np.ascontiguousarray(np.asarray(kernel).transpose([3, 2, 0, 1]))    # conv2d.py:26, kept for introspection
        self._bias, self._stride, self._alpha, self._padding = bias, stride, alpha, padding
        name = "conv%d" % (len(self._chain.layers))
        self._index = self._chain.add(self, ("conv", name, np.asarray(kernel, np.float32), np.asarray(bias, np.float32),
                                             float(alpha), padding))

    @property
    def _featuremap(self):                      # conv2d.py:61
        return self._chain.net().state(self._index, self._chain.stream)["F"]

    @property
    def _conv_actfn(self):                      # conv2d.py:63
        return self._chain.net().state(self._index, self._chain.stream)["A"]


class MaxPoolLayer(_GpuLayer):
    """Event max-pool (maxpool.py:14-165): ksize [k_h,k_w], stride == kernel size."""

    def __init__(self, prev_layer, ksize, stride):
        k_h, k_w = ksize
        if not (stride == k_h and stride == k_w):
            raise NotImplementedError("This method only support stride equal to 1 or to the kernel's dimensions.")
        self._prev_layer = prev_layer
        self._chain = prev_layer._chain
        self._ksize, self._stride = ksize, stride
        name = "pool%d" % (len(self._chain.layers))
        self._index = self._chain.add(self, ("pool", name, int(k_h), int(k_w), int(stride)))

    @property
    def _idx_max(self):                         # maxpool.py:33-35
        idx = self._chain.net().state(self._index, self._chain.stream)["idx"].astype(np.int32).reshape(-1)
        return [idx, np.arange(idx.size, dtype=np.int32)]

    @property
    def _recompute_coords(self):                # maxpool.py:36
        return self._chain.net().state(self._index, self._chain.stream)["flags"]
