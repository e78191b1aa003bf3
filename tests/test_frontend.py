"""CPU: the oracle's front-end / back-end restatements (N-data record decoder, runner transform, YOLO decode)
against the reference's own functions where they can be imported here, and against known answers."""
import os
import sys

import numpy as np
import pytest

from oracle import frontend as F

REF = "/root/reference"
has_ref = os.path.isdir(os.path.join(REF, "src"))


def _recording(rng, n, h=172, w=232, overflow_every=0):
    x = rng.integers(0, w, n).astype(np.int32)
    y = rng.integers(0, h, n).astype(np.int32)
    y[y == 240] = 239
    ts = np.sort(rng.integers(0, 1 << 13, n)).astype(np.int32)
    p = rng.integers(0, 2, n).astype(np.int32)
    if overflow_every:
        y[overflow_every::overflow_every] = 240          # timestamp-overflow markers
    return x, y, ts, p


def test_ndata_known_answer_record():
    # x = 0x12, y = 0x34, p = 1, ts = 0x2ABCDE -> bytes 12 34 (80|2A) BC DE
    raw = np.array([0x12, 0x34, 0x80 | 0x2A, 0xBC, 0xDE], np.uint8)
    n, x, y, ts, p = F.read_ndata(raw)
    assert (n, x[0], y[0], ts[0], p[0]) == (1, 0x12, 0x34, 0x2ABCDE, 1)
    assert np.array_equal(F.encode_ndata([0x12], [0x34], [0x2ABCDE], [1]), raw)


def test_ndata_overflow_markers_shift_later_timestamps():
    #           x   y    ts
    recs = [(1, 2, 100), (0, 240, 0), (3, 4, 5), (0, 240, 0), (5, 6, 7), (7, 8, 9)]
    raw = F.encode_ndata([r[0] for r in recs], [r[1] for r in recs], [r[2] for r in recs], [0] * len(recs))
    n, x, y, ts, p = F.read_ndata(raw)
    assert n == 4 and x.tolist() == [1, 3, 5, 7] and y.tolist() == [2, 4, 6, 8]
    assert ts.tolist() == [100, 5 + 8192, 7 + 2 * 8192, 9 + 2 * 8192]


def test_ndata_round_trip_random():
    rng = np.random.default_rng(3)
    x, y, ts, p = _recording(rng, 4000)
    n, dx, dy, dts, dp = F.read_ndata(F.encode_ndata(x, y, ts, p))
    assert n == 4000 and np.array_equal(dx, x) and np.array_equal(dy, y) and np.array_equal(dts, ts) and np.array_equal(dp, p)
    assert F.read_ndata(np.zeros(0, np.uint8))[0] == 0


def test_data_transform_zero_base_and_crop():
    rng = np.random.default_rng(4)
    x, y, ts, p = _recording(rng, 3000)
    ts = ts + 77
    ev, pol = F.data_transform(x, y, ts, p, (172, 232), (160, 224))
    assert ev.shape[1] == 3 and ev[:, 0].min() == 0 and ev[:, 1].min() == 0
    assert ev[:, 0].max() < 160 and ev[:, 1].max() < 224      # (y, x) inside the frame
    same, _ = F.data_transform(x, y, ts, p, (172, 232), (172, 232))
    assert np.array_equal(same[:, 2], ts - ts[0]) and np.array_equal(same[:, 0], y)


@pytest.mark.skipif(not has_ref, reason="reference tree not present")
def test_center_crop_equals_reference():
    sys.path.insert(0, REF)
    from src.libs import utils as RU
    rng = np.random.default_rng(5)
    for shape, new in [((172, 232), (160, 224)), ((180, 240), (160, 224)), ((124, 124), (96, 100))]:
        x, y, ts, p = _recording(rng, 2500, shape[0], shape[1])
        bb = np.array([[0.1, 0.2, 0.5, 0.6]])
        l, rx, ry, rts, rp, _ = RU.center_crop(len(x), x.copy(), y.copy(), ts.copy(), p.copy(), bb, shape, new)
        ox, oy, ots, op = F.center_crop_events(x, y, ts, p, new)
        assert l == len(ox) and np.array_equal(rx, ox) and np.array_equal(ry, oy) and np.array_equal(rts, ots) and np.array_equal(rp, op)


@pytest.mark.skipif(not has_ref, reason="reference tree not present")
def test_convert_bboxes_equals_reference():
    sys.path.insert(0, REF)
    try:
        from src.libs import viz as RV
    except Exception as e:      # cv2 missing
        pytest.skip("reference viz not importable: %s" % e)
    rng = np.random.default_rng(6)
    b = (rng.random((4, 5, 7, 2, 4)) * 1.5 - 0.2).astype(np.float32)
    for sqrt in (True, False):
        assert np.array_equal(RV.convert_bboxes(b, 5, 7, 160, 224, sqrt), F.convert_bboxes(b, 5, 7, 160, 224, sqrt))


def test_decode_head_known_answer():
    C, B, gh, gw = 3, 2, 2, 2
    head = np.zeros((1, gh, gw, C + 5 * B), np.float32)
    head[0, 1, 0, :C] = [0.2, 0.7, 0.1]
    head[0, 1, 0, C:C + 5] = [0.5, 0.25, 0.5, 0.5, 0.9]          # box 0 of cell (row 1, col 0)
    boxes, conf, valid, label = F.decode_head(head, gh, gw, C, 100, 200, 0.1)
    i = (1 * gw + 0) * B + 0
    assert np.allclose(boxes[0, i], [(0.5 + 0) / 2 * 200, (0.25 + 1) / 2 * 100, 0.25 * 200, 0.25 * 100])
    assert conf[0, i] == np.float32(0.9) and valid[0, i] and label[0, i] == 1
    assert valid.sum() == 1


# ---------------------------------------------------------------------------------------------------------------
# f1 on the device: batching of a sample into steps (runner.py:65-72) and recordings -> detections without a host
# round trip
# ---------------------------------------------------------------------------------------------------------------
def _samples(rng):
    """Ragged samples: empty, one event, fewer than a batch, exact multiples, long; timestamps with repeats and gaps."""
    out = []
    for n in [0, 1, 3, 39, 40, 41, 80, 130, 257, 5000]:
        ts = np.cumsum(rng.integers(0, 60, n)) if n else np.zeros(0, np.int64)
        if n > 50:
            ts[n // 2:] += 7000                                  # a long silence: empty duration bins in between
        out.append(np.stack([rng.integers(0, 16, n), rng.integers(0, 24, n), ts], axis=-1).astype(np.int32))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("size,usec", [(40, None), (1, None), (7, None), (100000, None), (40, 500), (40, 1), (40, 100000), (40, 64)])
def test_device_batching_equals_the_runner_split(size, usec):
    from async_ev_cnn_b200.frontend import split_batches
    from async_ev_cnn_b200.runner import split_event_batches
    samples = _samples(np.random.default_rng(5))
    got = split_batches(samples, batch_event_size=size, batch_event_usec=usec)
    for ev, b in zip(samples, got):
        if len(ev) == 0 and usec is not None:
            assert b.tolist() == [0, 0]                          # the reference indexes events[-1] here and raises; one empty batch
            continue
        want = split_event_batches(ev, size, usec)
        bounds = np.concatenate([[0], np.cumsum([len(c) for c in want])]).astype(np.int32)
        assert np.array_equal(b, bounds), "N=%d size=%s usec=%s" % (len(ev), size, usec)


@pytest.mark.gpu
@pytest.mark.parametrize("size,usec", [(60, None), (60, 4000)])
def test_recordings_to_detections_on_the_device_equal_the_python_loop(size, usec):
    """aec_net_run_ndata (decode + transform + batching + steps, all on the device) == decoding, splitting and feeding the
    chunks from Python (the runner's loop), bit for bit; streams with fewer batches idle."""
    import async_ev_cnn_b200 as P
    from oracle import frontend as F
    from async_ev_cnn_b200.engine import EventNetCuda
    from async_ev_cnn_b200.frontend import decode_ndata, run_recordings
    from async_ev_cnn_b200.runner import split_event_batches
    layers = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
    h, w, S = 32, 48, 4
    rng = np.random.default_rng(3)
    recs = []
    for n in [900, 0, 333, 1500]:
        x = rng.integers(0, 60, n).astype(np.int32)
        y = rng.integers(0, 44, n).astype(np.int32)
        ts = np.sort(rng.integers(0, 1 << 15, n)).astype(np.int32)
        recs.append(F.encode_ndata(x, y, ts, rng.integers(0, 2, n).astype(np.int32)))
    wts = P.xavier_weights(layers, seed=2)
    a = EventNetCuda(h, w, layers, wts, 1e-4, 0.1, "SAME", n_streams=S, max_events_per_step=4096)
    b = EventNetCuda(h, w, layers, wts, 1e-4, 0.1, "SAME", n_streams=S, max_events_per_step=4096)
    heads, steps, kept = run_recordings(a, recs, crop_to=(h, w), batch_event_size=size, batch_event_usec=usec)
    evs = decode_ndata(recs, zero_base_ts=True, crop_to=(h, w))
    assert [len(e) for e in evs] == kept.tolist()
    chunks = [split_event_batches(e, size, usec) if len(e) else [e] for e in evs]
    assert steps == max(len(c) for c in chunks)
    want = None
    b.reset()
    for t in range(steps):
        per = [c[t] if t < len(c) and len(c[t]) else None for c in chunks]
        want = b.step(per).copy()
    assert np.array_equal(heads, want)
    for li in range(len(a.names)):
        for s in range(S):
            sa, sb = a.state(li, s), b.state(li, s)
            for k in sa:
                assert np.array_equal(sa[k], sb[k]), "layer %s stream %d %s" % (a.names[li], s, k)
    # a second call without reset continues the streams; with reset it reproduces the first
    heads2, _, _ = run_recordings(a, recs, crop_to=(h, w), batch_event_size=size, batch_event_usec=usec, reset=True)
    assert np.array_equal(heads2, heads)
    with pytest.raises(ValueError):
        run_recordings(a, recs[:2])
    a.close()
    b.close()
