"""GPU: the reference-facing Python surface (layers.py, models.py, runner.py) - written the way the reference's own
src/scripts/test_correctness.py:18-39,96-140 drives its layer classes: build the chain from layer objects, call
compute() layer by layer, compare the accessors.  The oracle's layer classes (same constructors and methods,
restating src/layers/*.py) stand in for the reference's NumPy layers."""
import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.layers import Conv2DLayer, IntegrationLayer, MaxPoolLayer
from async_ev_cnn_b200.models import YoloEventCuda
from async_ev_cnn_b200.runner import CudaEventRunner, SyntheticReader
from oracle.event_oracle import OracleConv, OracleEventNet, OracleIntegration, OraclePool

pytestmark = pytest.mark.gpu


def _as_set(ev):
    return set(zip(np.asarray(ev[0]).tolist(), np.asarray(ev[1]).tolist()))


def _chain(intgr_cls, conv_cls, pool_cls, h, w, leak, alpha, k1, b1, k2, b2):
    intgr = intgr_cls(leak, h, w)                                      # test_correctness.py:20-27
    conv1 = conv_cls(intgr, k1, b1, 1, alpha, "SAME")
    pool1 = pool_cls(conv1, [2, 2], 2)
    conv2 = conv_cls(pool1, k2, b2, 1, alpha, "SAME")
    pool2 = pool_cls(conv2, [2, 2], 2)
    return [intgr, conv1, pool1, conv2, pool2]


def test_correctness_protocol_on_the_layer_classes():
    """8x8 frame, Intgr -> Conv3x3 -> Pool2 -> Conv3x3 -> Pool2, integer kernel [[-2,-1,1]]*3, bias 10, leak 0.1,
    alpha 0.1, one event at ts 0 then 5 events per step with ts = sort(randint(1,10,5)) + prev
    (test_correctness.py:96-105,124-126,165-168; seeded here)."""
    h = w = 8
    k = np.array([[-2, -1, 1]] * 3, np.float32).reshape(3, 3, 1, 1)     # HWIO, as the layers take it
    b = np.array([10.0], np.float32)
    gpu = _chain(IntegrationLayer, Conv2DLayer, MaxPoolLayer, h, w, 0.1, 0.1, k, b, k, b)
    ref = _chain(OracleIntegration, OracleConv, OraclePool, h, w, 0.1, 0.1, k, b, k, b)
    assert [l.out_shape() for l in gpu] == [l.out_shape() for l in ref]
    rng = np.random.default_rng(0)
    ts_prev = 0
    events = np.array([[3, 4, 0]], np.int32)
    for step in range(300):
        ev_g, d_g = gpu[0].compute(events, None)
        ev_r, d_r = ref[0].compute(events, None)
        assert d_g == d_r
        for lg, lr in zip(gpu[1:], ref[1:]):                            # test_correctness.py:33-39
            ev_g, d_g = lg.compute(ev_g, d_g)
            ev_r, d_r = lr.compute(ev_r, d_r)
            assert _as_set(ev_g) == _as_set(ev_r), "step %d" % step
        for lg, lr in zip(gpu, ref):                                    # test_correctness.py:137-140 (allclose there)
            assert np.allclose(lg.featuremap(), lr.featuremap(), rtol=1e-5, atol=1e-5), "step %d featuremap" % step
        if step % 50 == 0:
            for lg, lr in zip(gpu[1:], ref[1:]):
                assert np.allclose(lg.surface(), lr.surface(), rtol=1e-5, atol=1e-5)
                assert np.allclose(lg.conv_actfn(), lr.conv_actfn(), rtol=1e-5, atol=1e-5)
                assert np.array_equal(lg.layer_actfn(), lr.layer_actfn())
        ts = np.sort(rng.integers(1, 10, 5)) + ts_prev
        ts_prev = int(ts[-1])
        events = np.stack([rng.integers(0, h, 5), rng.integers(0, w, 5), ts], axis=-1).astype(np.int32)
    for l in gpu:                                                       # event_numpy.py:96-98
        l.reset()
    for l in ref:
        l.reset()
    ev_g, d_g = gpu[-1].compute_all(events)
    ev_r, d_r = ref[-1].compute_all(events)
    assert np.allclose(gpu[-1].featuremap(), ref[-1].featuremap(), rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):                                     # events that are not the predecessor's output
        gpu[1].compute((np.array([0]), np.array([0])), d_g)


def test_model_class_graph_contract():
    """YoloEventCuda(...) / build_graph(_) -> graph(events, reset), event_numpy.py:13-15,90-105."""
    layers = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,7"
    h, w = 16, 24                                                       # head 4 x 6 x 7 = cells 4 x 6, 2 classes + 1 box
    model = YoloEventCuda(h, w, 2, layers, "SAME", 4, 6, 1, 0.1, 0.001, "random:5")
    graph = model.build_graph(None)
    oracle = OracleEventNet(h, w, layers, P.xavier_weights(layers, seed=5), 0.001, 0.1, "SAME")
    evs = P.synthetic_events("uniform", 1, 12, 15, h, w, seed=2, dt_int=(1, 20))[0]
    for t in range(12):
        out = graph(evs[t], t == 0)
        want = oracle.step(evs[t], reset=(t == 0))
        assert out.shape == (4, 6, 7) and out.dtype == np.float32
        assert np.allclose(out, want, rtol=1e-4, atol=1e-5)
    out = graph(evs[0], True)                                           # reset restarts the stream
    assert np.allclose(out, OracleEventNet(h, w, layers, P.xavier_weights(layers, seed=5), 0.001, 0.1, "SAME").step(evs[0]), rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        YoloEventCuda(h, w, 2, layers, "SAME", 5, 6, 1, 0.1, 0.001, "random:5").build_graph(None)   # head does not reshape
    with pytest.raises(NotImplementedError):
        YoloEventCuda(h, w, 2, layers + " fc1=168,10", "SAME", 4, 6, 1, 0.1, 0.001, "random:5").build_graph(None)


def test_multi_stream_graph_returns_fresh_arrays_and_validates_reset():
    """graph() with n_streams > 1: a caller may keep results across steps (runner.py:100 appends them), so two
    consecutive outputs must not alias; a scalar truthy reset (np.True_, 1) is a full reset, a mask of the wrong
    length is rejected instead of being read out of bounds."""
    layers = "conv1=3,3,1,4 pool1=2,2 conv2=1,1,4,7"
    h, w, S = 16, 24, 3
    model = YoloEventCuda(h, w, 2, layers, "SAME", 8, 12, 1, 0.1, 0.001, "random:1", n_streams=S)
    graph = model.build_graph(None)
    evs = P.synthetic_events("uniform", S, 4, 20, h, w, seed=3, dt_int=(1, 20))
    a = graph([evs[s, 0] for s in range(S)], True)
    keep = a.copy()
    b = graph([evs[s, 1] for s in range(S)], False)
    assert not np.shares_memory(a, b) and np.array_equal(a, keep) and not np.array_equal(a, b)
    first = graph([evs[s, 0] for s in range(S)], np.True_)             # numpy bool scalar: full reset
    assert np.array_equal(first, keep)
    graph([evs[s, 1] for s in range(S)], False)
    again = graph([evs[s, 0] for s in range(S)], 1)                    # truthy int: full reset
    assert np.array_equal(again, keep)
    one = graph([evs[s, 1] if s != 1 else evs[1, 0] for s in range(S)], [0, 1, 0])   # stream 1 restarts, the others go on
    assert np.array_equal(one[1], keep[1]) and np.array_equal(one[0], b[0])
    with pytest.raises(ValueError):
        graph([evs[s, 2] for s in range(S)], [1, 0])                   # mask shorter than n_streams
    with pytest.raises(ValueError):
        model.net.reset(np.uint8(1))                                   # 0-d mask
    with pytest.raises(ValueError):
        model.net.step_packed(np.zeros((2, 3), np.int32), np.array([0, 1, 2, 2], np.int32), out=np.empty((S, 8, 12, 6), np.float32))
    with pytest.raises(Exception):                                     # offsets that do not end at total / decrease
        model.net.step_packed(np.zeros((2, 3), np.int32), np.array([0, 2, 1, 2], np.int32))
    with pytest.raises(Exception):
        model.net.step_packed(np.zeros((2, 3), np.int32), np.array([1, 1, 2, 2], np.int32))


def test_runner_feeds_chunks_and_resets_per_sample(capsys):
    """CudaEventRunner.run: a sample is split into batch_event_size chunks, reset_state only on the first chunk
    (runner.py:64-72,101 as intended); the result equals feeding the oracle the same chunks."""
    import argparse
    layers = "conv1=3,3,1,4 pool1=2,2 conv2=1,1,4,7"
    h, w = 16, 24
    args = argparse.Namespace(frame_h=h, frame_w=w, example_h=h, example_w=w, batch_event_size=40, batch_event_usec=None,
                              n_streams=1, max_samples=None)
    reader = SyntheticReader(h, w, n_samples=3, events_per_sample=130, n_classes=2, kind="uniform", seed=9)
    model = YoloEventCuda(h, w, 2, layers, "SAME", 8, 12, 1, 0.1, 0.001, "random:1")
    calls = []
    graph = model.build_graph(None)

    def network(events, reset):
        calls.append((len(events), bool(reset)))
        return graph(events, reset)

    outs, times = CudaEventRunner(args, reader).run(network)
    assert len(outs) == 3 and len(times) == len(calls) == 3 * 4        # 130 events -> 4 chunks of <= 40
    assert [c[1] for c in calls] == [True, False, False, False] * 3
    assert sum(c[0] for c in calls[:4]) == 130
    ref_reader = SyntheticReader(h, w, n_samples=3, events_per_sample=130, n_classes=2, kind="uniform", seed=9)
    wts = P.xavier_weights(layers, seed=1)
    from async_ev_cnn_b200.runner import Runner, split_event_batches
    for i in range(3):
        _, ev = ref_reader.next_batch(1, preprocessing_fn=lambda *a: Runner.data_transform(*a, args=args))
        oracle = OracleEventNet(h, w, layers, wts, 0.001, 0.1, "SAME")
        for chunk in split_event_batches(ev, 40):
            want = oracle.step(chunk)
        assert np.allclose(outs[i], want.reshape(outs[i].shape), rtol=1e-4, atol=1e-5)
    assert "sec/example" in capsys.readouterr().out


def test_run_networks_cli_with_the_efcn_config(capsys):
    """python -m async_ev_cnn_b200.run_networks -c configs/efcn_event_cuda.yml (run_networks.py:15-59): the reference's
    YAML with `network: YoloEventCuda`; no dataset on disk -> seeded synthetic recordings, centre-cropped 172x232 -> 160x224."""
    import os
    from async_ev_cnn_b200.run_networks import main
    yml = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "efcn_event_cuda.yml")
    outs = main(["-c", yml, "--max_samples", "1"])
    assert len(outs) == 1 and outs[0].shape == (5, 7, 110) and np.isfinite(outs[0]).all()
    text = capsys.readouterr().out
    assert "sec/example" in text and "Mean fw time" in text
    outs2 = main(["-c", yml, "--max_samples", "2", "--n_streams", "2", "--batch_event_size", "500"])
    assert len(outs2) == 1 and outs2[0].shape == (2, 5, 7, 110) and np.isfinite(outs2[0]).all()
    with pytest.raises(SystemExit):
        main(["-c", yml, "--network", "YoloFrameTf"])
