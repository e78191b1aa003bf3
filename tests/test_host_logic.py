"""CPU: host-side mirror of the reference's runner / config / weight handling - no GPU.

split_event_batches is checked against a direct restatement of runner.py:65-72 (the INTENDED batching),
the config layer against configs/efcn_event_cuda.yml, load_weights against the dict event_numpy.py:34-51 builds."""
import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200 import config as cfg
from async_ev_cnn_b200.models import load_weights
from async_ev_cnn_b200.runner import split_event_batches


def _events(n, seed=0, max_gap=40):
    rng = np.random.default_rng(seed)
    ts = np.cumsum(rng.integers(0, max_gap, size=n)).astype(np.int32)
    return np.stack([rng.integers(0, 160, n), rng.integers(0, 224, n), ts], axis=-1).astype(np.int32)


@pytest.mark.parametrize("n,size", [(1, 200), (199, 200), (200, 200), (201, 200), (1000, 7), (5, 1)])
def test_split_by_event_count_follows_runner(n, size):
    ev = _events(n, seed=n)
    chunks = split_event_batches(ev, batch_event_size=size)
    want = np.array_split(ev, int(np.ceil(n / size)), axis=0)          # runner.py:71-72
    assert len(chunks) == len(want) == int(np.ceil(n / size))
    for a, b in zip(chunks, want):
        assert np.array_equal(a, b)
    assert np.array_equal(np.concatenate(chunks), ev)
    sizes = [len(c) for c in chunks]
    assert max(sizes) <= size and max(sizes) - min(sizes) <= 1


def test_split_by_duration_follows_runner():
    ev = _events(3000, seed=5)
    usec = 1000
    chunks = split_event_batches(ev, batch_event_usec=usec)
    bins = np.arange(0, ev[-1, -1], usec)                              # runner.py:66-70
    ids = np.digitize(ev[:, -1], bins)
    want = np.array_split(ev, np.where(ids[:-1] != ids[1:])[0] + 1, axis=0)
    assert len(chunks) == len(want)
    for a, b in zip(chunks, want):
        assert np.array_equal(a, b)
    assert np.array_equal(np.concatenate(chunks), ev)
    for c in chunks:                                                   # one duration bin per chunk, bins in order
        assert c[-1, -1] // usec == c[0, -1] // usec
    firsts = [c[0, -1] // usec for c in chunks]
    assert firsts == sorted(firsts) and len(set(firsts)) == len(firsts)


def test_yaml_config_and_flag_override(tmp_path):
    import os
    yml = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "efcn_event_cuda.yml")
    a = cfg.config(["-c", yml])
    assert a.network == "YoloEventCuda"
    assert list(a.yolo_cnn_layers.keys())[0] == "conv1" and a.yolo_cnn_layers["conv1"] == [3, 3, 1, 16]
    assert a.yolo_cnn_layers["pool1"] == [2, 2] and a.yolo_cnn_layers["conv7"][-1] == 110
    assert (a.frame_h, a.frame_w) == (160, 224) and a.batch_event_size == 200 and a.yolo_cnn_padding == "SAME"
    b = cfg.config(["-c", yml, "--batch_event_size", "50", "--leak", "0.001", "--some_unknown_flag", "1"])   # parse_known_args
    assert b.batch_event_size == 50 and b.leak == 0.001 and b.frame_w == 224
    with pytest.raises(SystemExit):
        cfg.build_parser().parse_args(["-c", yml, "--yolo_cnn_layers", "conv1=3;3"])


def test_load_weights_npz_random_and_errors(tmp_path):
    layers = "conv1=3,3,1,4 pool1=2,2 conv2=1,1,4,5"
    w = P.xavier_weights(layers, seed=3)
    path = tmp_path / "weights.npz"
    np.savez(path, **w)
    got = load_weights(str(path), layers)
    assert set(got) == {"w_conv1", "b_conv1", "w_conv2", "b_conv2"}
    for k in got:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k])
    assert np.array_equal(load_weights(str(tmp_path), layers)["w_conv2"], w["w_conv2"])     # a directory: newest .npz
    r1, r2 = load_weights("random:3", layers), load_weights(None, layers, seed=3)
    assert np.array_equal(r1["w_conv1"], w["w_conv1"]) and np.array_equal(r2["b_conv2"], w["b_conv2"])
    lim = np.sqrt(6.0 / (9 * (1 + 4)))                                  # xavier-uniform, frame_tf.py:76-78
    assert np.abs(w["w_conv1"]).max() <= lim and np.allclose(w["b_conv1"], 0.1)
    with pytest.raises(ValueError):
        load_weights(str(tmp_path / "model.ckpt"), layers)              # TF checkpoints cannot be read here
    np.savez(tmp_path / "bad.npz", w_conv1=np.zeros((3, 3, 1, 8), np.float32), b_conv1=np.zeros(8, np.float32),
             w_conv2=w["w_conv2"], b_conv2=w["b_conv2"])
    with pytest.raises(ValueError):
        load_weights(str(tmp_path / "bad.npz"), layers)
    with pytest.raises(FileNotFoundError):
        load_weights(str(tmp_path / "nothing_here"), layers) if (tmp_path / "nothing_here").mkdir() is None else None
