"""YoloEventCuda - the B200 backend of the reference's event-mode model (src/models/event_numpy.py).

Same constructor and `build_graph(_) -> graph(events, reset)` contract as YoloEventNumpy
(event_numpy.py:13-15,90-105), selectable with `network: YoloEventCuda` in a reference-style YAML
config.  Extensions: `n_streams` independent streams share one engine; `graph` then takes a list of
per-stream event arrays (None = no events this step) and a reset mask, and returns
[n_streams, h_cells, w_cells, C+5B].
"""
import os

import numpy as np

from .engine import EventNetCuda, pack_events
from .layers import Conv2DLayer, IntegrationLayer, MaxPoolLayer
from .streams import parse_layers, xavier_weights


def load_weights(checkpoint, cnn_layers, seed=0):
    """`w_<name>` [kh,kw,ci,co] / `b_<name>` [co] as the reference keeps them (event_numpy.py:34-51,64).

    checkpoint: path to an .npz holding those keys (a TF-1 checkpoint must be exported to .npz where
    TensorFlow is available - it cannot be read here), or None / 'random' / 'random:<seed>' for the
    TF model's initialiser (xavier-uniform, bias 0.1: frame_tf.py:76-78)."""
    if checkpoint is None or str(checkpoint).startswith("random"):
        s = str(checkpoint or "")
        seed = int(s.split(":")[1]) if ":" in s else seed
        return xavier_weights(cnn_layers, seed=seed)
    path = str(checkpoint)
    if os.path.isdir(path):
        cands = sorted(f for f in os.listdir(path) if f.endswith(".npz"))
        if not cands:
            raise FileNotFoundError("no .npz weight file in %s" % path)
        path = os.path.join(path, cands[-1])
    if not path.endswith(".npz"):
        raise ValueError("%s: TensorFlow checkpoints cannot be read by the B200 backend; export the variables "
                         "w_<layer>/b_<layer> to an .npz (np.savez) where TensorFlow is installed" % path)
    z = np.load(path)
    out = {}
    for name, size in parse_layers(cnn_layers).items():
        if "conv" in name:
            w, b = np.asarray(z["w_" + name], np.float32), np.asarray(z["b_" + name], np.float32)
            if list(w.shape) != list(size):
                raise ValueError("w_%s has shape %s, config says %s" % (name, w.shape, size))
            out["w_" + name], out["b_" + name] = w, b
    return out


class YoloEventCuda:
    def __init__(self, h_frame, w_frame, num_classes, cnn_layers, cnn_padding, h_cells, w_cells, num_bbox,
                 alpha, leak, checkpoint, sess=None, n_streams=1, device=0, max_events_per_step=0):
        self._h_frame, self._w_frame = h_frame, w_frame
        self._num_classes = num_classes
        self._cnn_layers = parse_layers(cnn_layers)
        self._padding = cnn_padding
        self._h_cells, self._w_cells, self._num_bbox = h_cells, w_cells, num_bbox
        self._alpha, self._leak = alpha, leak
        self._sess = sess                      # unused (no TensorFlow); kept for signature compatibility
        self._checkpoint = checkpoint
        self._n_streams, self._device, self._max_events = n_streams, device, max_events_per_step
        self._weights = {}
        self.restore(checkpoint)
        self.net = None

    def restore(self, checkpoint_path, restrict_vars=None):        # event_numpy.py:34-51
        w = load_weights(checkpoint_path, self._cnn_layers)
        if restrict_vars:
            w = {k: v for k, v in w.items() if k in restrict_vars}
        self._weights.update(w)

    def build_cnn_layers(self):
        """The reference's layer-object chain (event_numpy.py:53-73) on the GPU mirror classes - for
        introspection; build_graph() drives the engine directly."""
        prev = IntegrationLayer(self._leak, self._h_frame, self._w_frame, device=self._device)
        event_layers, non_event_layers = [prev], []
        for name, size in self._cnn_layers.items():
            if 'conv' in name:
                prev = Conv2DLayer(prev, self._weights['w_' + name], self._weights['b_' + name], 1, self._alpha, self._padding)
                event_layers.append(prev)
            elif 'pool' in name:
                prev = MaxPoolLayer(prev, size, size[0])
                event_layers.append(prev)
            else:
                non_event_layers.append((name, size))
        return event_layers, non_event_layers

    def build_graph(self, _=None):
        for name in self._cnn_layers:
            if 'conv' not in name and 'pool' not in name:
                raise NotImplementedError("non-event layer %r: the EFCN configs have none (fc/flatten are out of scope)" % name)
        self.net = EventNetCuda(self._h_frame, self._w_frame, self._cnn_layers, self._weights, self._leak, self._alpha,
                                self._padding, n_streams=self._n_streams, device=self._device,
                                max_events_per_step=self._max_events)
        out_shape = [self._h_cells, self._w_cells, self._num_classes + self._num_bbox * 5]
        if int(np.prod(self.net.head_shape)) != int(np.prod(out_shape)):
            raise ValueError("cannot reshape the last layer %s into %s" % (list(self.net.head_shape), out_shape))
        net, S = self.net, self._n_streams

        def graph(input, reset):                                    # event_numpy.py:94-103
            if S == 1 and isinstance(input, np.ndarray):
                if reset:
                    net.reset()
                return np.reshape(net.step([input])[0], out_shape).copy()
            if np.ndim(reset) == 0:                                  # True / np.True_ / 1: every stream
                if bool(reset):
                    net.reset()
            elif np.any(reset):
                net.reset(np.asarray(reset, np.uint8))               # per-stream mask, length checked by the engine
            if isinstance(input, tuple):
                heads = net.step_packed(*input)
            else:
                heads = net.step_packed(*pack_events(input))
            # a fresh array per call: the engine reuses its head buffer, and callers keep results across steps
            # (runner.py:100 appends net_out to a list)
            return np.reshape(heads, [S] + out_shape).copy()

        return graph
