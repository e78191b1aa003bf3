"""Per-sample loop of the reference (src/libs/runner.py:11-127) for the B200 backend, headless.

Keeps the Runner contract - data_transform, feed_network(network, events, frames, reset_state), run()
printing sec/example - and implements the INTENDED batching of runner.py:65-72 (as
written the reference feeds the whole sample): a sample's [N,3] events are split into chunks of
`batch_event_size` events (or `batch_event_usec` microseconds), `reset_state` is True for the first
chunk only (runner.py:64,101).  With n_streams > 1 that many samples are advanced side by side.
"""
import time

import numpy as np


def center_crop(l, x, y, ts, p, bboxes, old_shape, new_shape):
    """Events (and boxes) cropped to a centred new_shape window, shifted to start at (0,0).
    Follows src/libs/utils.py:4-35 including its naming quirk (new_top is derived from x, new_left
    from y) so that crops agree with the reference for the square-ish margins the configs use."""
    new_h, new_w = new_shape
    old_h, old_w = old_shape
    new_top = (x.max() - x.min() - new_w) // 2
    new_left = (y.max() - y.min() - new_h) // 2
    inside = (x >= new_left) & (x < new_left + new_w) & (y >= new_top) & (y < new_top + new_h)
    nx, ny, nts, np_ = x[inside].copy(), y[inside].copy(), ts[inside], p[inside]
    if nx.size:
        nx -= nx.min()
        ny -= ny.min()
    nb = None
    if bboxes is not None:
        b = np.array(bboxes, dtype=np.float64)
        b[:, [0, 2]] *= old_w
        b[:, [1, 3]] *= old_h
        nb = b.copy()
        shift = nx.min() if nx.size else 0
        nb[:, [0, 2]] = np.clip(b[:, [0, 2]] * old_w - shift, 0, new_w) / new_w
        nb[:, [1, 3]] = np.clip(b[:, [1, 3]] * old_h - shift, 0, new_h) / new_h
    return nx.shape[0], nx, ny, nts, np_, nb


def split_event_batches(events, batch_event_size=1, batch_event_usec=None):
    """[N,3] (y,x,ts) -> list of chunks: fixed event count, or fixed duration bins (runner.py:65-72)."""
    events = np.asarray(events)
    if batch_event_usec is not None:
        bins = np.arange(0, events[-1, -1], batch_event_usec)
        ids = np.digitize(events[:, -1], bins)
        cuts = np.where(ids[:-1] != ids[1:])[0] + 1
        return np.array_split(events, cuts, axis=0)
    n = int(np.ceil(events.shape[0] / batch_event_size))
    return np.array_split(events, max(n, 1), axis=0)


class Runner:
    def __init__(self, args, reader, profile_integration=False):
        self.args = args
        self.reader = reader
        self.num_classes = reader.num_classes()
        self.profile_integration = profile_integration
        keys = np.array(list(reader.label_to_idx().keys()))
        vals = np.array(list(reader.label_to_idx().values()))
        self.idx_to_label = keys[np.argsort(vals)]

    @staticmethod
    def data_transform(l, x, y, ts, p, bboxes, args):           # runner.py:24-33
        ts = ts - ts[0]
        if args.frame_h != args.example_h or args.frame_w != args.example_w:
            l, x, y, ts, p, bboxes = center_crop(l, x, y, ts, p, bboxes, (args.example_h, args.example_w),
                                                 (args.frame_h, args.frame_w))
        return l, np.stack([y, x, ts], axis=-1).astype(np.int32)

    def show_frames(self, net_out, frames, *args, **kwargs):
        """Display hook (cv2.imshow in the reference, runner.py:35-44): headless here."""

    def feed_network(self, network, events, frames, reset_state, *args, **kwargs):
        raise NotImplementedError()

    def run(self, network, *args, **kwargs):
        n, ex_time, outs = 0, [], []
        n_streams = getattr(self.args, "n_streams", 1)
        total = int(np.ceil(self.reader.test_size() / n_streams))
        if getattr(self.args, "max_samples", None):
            total = min(total, int(np.ceil(self.args.max_samples / n_streams)))
        for i in range(total):
            start_read = time.time()
            samples = [self.reader.next_batch(1, dataset='test', preprocessing_fn=lambda *a: self.data_transform(*a, args=self.args))[1]
                       for _ in range(n_streams)]
            end_reading = time.time()
            chunks = [split_event_batches(ev, self.args.batch_event_size, self.args.batch_event_usec) for ev in samples]
            reset_state = True
            for b in range(max(len(c) for c in chunks)):
                per = [c[b] if b < len(c) and len(c[b]) else None for c in chunks]
                start_fw = time.time()
                net_out = self.feed_network(network, per if n_streams > 1 else per[0], None, reset_state, *args, **kwargs)
                time_fw = time.time() - start_fw
                ex_time.append(time_fw)
                n += 1
                print("Test batch {:<2} - sec/example: {:.3f}  reading: {:.3f} sec".format(i + 1, time_fw, end_reading - start_read))
                if n % 1000 == 0:
                    print("Mean fw time ({} runs): {}".format(n, np.mean(ex_time)))
                self.show_frames(net_out, None)
                reset_state = False
            outs.append(net_out)
        return outs, ex_time


class CudaEventRunner(Runner):
    """Counterpart of NumpyEventRunner (runner.py:122-127)."""

    def __init__(self, args, reader):
        super().__init__(args, reader, profile_integration=False)

    def feed_network(self, network, events, frames, reset_state, *args, **kwargs):
        return network(events, reset_state)


class SyntheticReader:
    """Stand-in for detection_reader.factory(...) when no dataset is on disk: serves seeded synthetic
    N-Caltech101-shaped recordings through the four reader methods the runner uses
    (runner.py:16,19-21,55-60): num_classes, label_to_idx, test_size, next_batch."""

    def __init__(self, height, width, n_samples=4, events_per_sample=20000, n_classes=100, kind="uniform", seed=0):
        self.h, self.w, self.n, self.ev, self.k, self.kind, self.seed = height, width, n_samples, events_per_sample, n_classes, kind, seed
        self.pos = 0

    def num_classes(self):
        return self.k

    def label_to_idx(self):
        return {"class%03d" % i: i for i in range(self.k)}

    def test_size(self):
        return self.n

    def next_batch(self, batch_size, dataset='test', preprocessing_fn=None, concat_features=False, threads=1):
        from .streams import synthetic_events
        ev = synthetic_events(self.kind, 1, 1, self.ev, self.h, self.w, seed=self.seed + self.pos)[0, 0]
        self.pos += 1
        y, x, ts = ev[:, 0].copy(), ev[:, 1].copy(), ev[:, 2].copy()
        p = np.ones_like(x)
        if preprocessing_fn is not None:
            return preprocessing_fn(len(x), x, y, ts, p, None)
        return len(x), ev
