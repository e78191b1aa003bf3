"""GPU: BASELINE-size checks through size-independent properties (the oracle is too slow for
thousands of streams): replicated streams agree bit-for-bit, runs are deterministic, a stream's
result does not depend on which other streams share the batch, and the event net equals a dense
torch fp32 frame network."""
import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda

pytestmark = pytest.mark.gpu
H, W = 160, 224


def _run(net, evs, steps):
    heads = []
    for t in range(steps):
        heads.append(net.step([evs[s % evs.shape[0], t] for s in range(net.n_streams)]).copy())
    return np.stack(heads)


def test_replicas_determinism_and_batch_independence():
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    evs = P.synthetic_events("edge", 4, 12, 200, H, W, seed=3)
    net = EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=64)
    a = _run(net, evs, 12)
    for s in range(4, 64):                              # stream s replays stream s % 4
        assert np.array_equal(a[:, s], a[:, s % 4])
    net.reset()
    b = _run(net, evs, 12)
    assert np.array_equal(a, b)                         # deterministic, reset restores the initial state
    net.close()
    small = EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=4)
    c = _run(small, evs, 12)
    assert np.array_equal(c, a[:, :4])                  # result independent of batch composition
    sites, steps = small.counters()
    assert steps == 12 and sites[1] > 0
    small.close()


def test_efcn_event_equals_dense_torch_frame_network():
    """configs/efcn_frame_np.yml shape: dense conv -> leaky -> max_pool on the integrated frame."""
    import torch
    import torch.nn.functional as F
    from oracle.event_oracle import integrate_frame
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    evs = P.synthetic_events("uniform", 1, 16, 200, H, W, seed=5)[0]
    net = EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=1)
    state = None
    for t in range(16):
        head = net.step(evs[t])[0]
        frame, ts = integrate_frame(evs[t], 5e-5, H, W, state)
        state = (frame, ts)
    x = torch.from_numpy(frame)[None, None].double().cuda()
    for name, size in P.parse_layers(P.EFCN_LAYERS).items():
        if "conv" in name:
            k = torch.from_numpy(wts["w_" + name]).permute(3, 2, 0, 1).double().cuda()
            x = F.conv2d(x, k, torch.from_numpy(wts["b_" + name]).double().cuda(), padding=(size[0] - 1) // 2)
            x = torch.maximum(x, 0.1 * x)
        else:
            x = F.max_pool2d(x, size[0], size[0])
    dense = x[0].permute(1, 2, 0).cpu().numpy()
    scale = np.abs(dense).max()
    assert np.abs(head - dense).max() <= 1e-4 * scale, "event vs dense frame: %.3e (scale %.3e)" % (np.abs(head - dense).max(), scale)
    net.close()


@pytest.mark.parametrize("kind,pairs", [("edge", False), ("uniform", False), ("edge", True)])
def test_efcn_benchmark_regime_against_live_oracle(kind, pairs, monkeypatch):
    """The regime bench.py measures (BASELINE config 2): EFCN 160x224, B = 200, sweep skipping on, compared with the
    live oracle step by step from reset THROUGH the 160-step settling phase and 64 steps of the steady state behind it
    (sticky pool flags saturated, the leak sweep skipping what the step re-evaluates).  Every step: surface and
    surface events exact, the integer decisions bit-exactly the reference's rules on the CUDA path's own maps, float
    maps within 1e-4 of the scale, and any integer difference from the oracle explained by a near tie in the
    oracle's own values.  The counters of both phases go to the parity log (profiles/parity_r2.json).
    `pairs`: with 32 streams or more (the benchmark has 1024) conv5 and conv6 run on CTA pairs (tcgen05 cta_group::2: aec_tc.cuh kPair); AEC_TC_PAIR=1 selects those kernels for the two streams of this test."""
    if pairs:
        monkeypatch.setenv("AEC_TC_PAIR", "1")
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from parity import OracleAdapter, compare_live, record_parity
    from async_ev_cnn_b200.engine import CudaAdapter
    from oracle.event_oracle import OracleEventNet
    settle, steady = 160, 64
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    evs = P.synthetic_events(kind, 1, settle + steady, 200, H, W, seed=100)[0]
    net = EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=2)
    ora = OracleEventNet(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME")
    ad, oa = CudaAdapter(net, stream=1), OracleAdapter(ora)
    a = compare_live(ad, oa, list(evs[:settle]), exact=False)
    tag = kind + ("_cta_pairs" if pairs else "")
    record_parity("live/efcn160x224_%s_steps0-%d_settling" % (tag, settle - 1), a)
    b = compare_live(ad, oa, list(evs[settle:]), exact=False, carry=a.carry)
    record_parity("live/efcn160x224_%s_steps%d-%d_steady_state" % (tag, settle, settle + steady - 1), b)
    if pairs:
        kernels = [g["kernel"] for g in (net.tc_geometry(i) for i in range(1, len(net.names))) if g is not None]
        assert sum("CTA pairs" in k for k in kernels) == 2, kernels
    st = net.sweep_stats()
    assert st["swept_conv_elems"] < st["live_conv_elems"], "the steady state is where the sweep skips work"
    net.close()


STRESS_LAYERS = ("conv1=3,3,1,16 conv1b=3,3,16,16 pool1=2,2 conv2=3,3,16,32 conv2b=3,3,32,32 pool2=2,2 conv3=3,3,32,64 pool3=2,2 "
                 "conv4=3,3,64,128 pool4=2,2 conv5=3,3,128,256 pool5=2,2 conv6=1,1,256,512 conv7=1,1,512,110")


def test_stress_config_davis346_deeper_variant_against_live_oracle():
    """BASELINE config 5: DAVIS346-sized stream cropped to 256x320 (pooled dims must be even), high
    event rate (1000 events per step), a deeper EFCN variant (two 3x3 convs in the first two stages).  The
    reference has no such config; its layers are generic, so the oracle runs it."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from parity import OracleAdapter, compare_live, record_parity
    from async_ev_cnn_b200.engine import CudaAdapter
    from oracle.event_oracle import OracleEventNet
    h, w, steps = 256, 320, 10
    wts = P.xavier_weights(STRESS_LAYERS, seed=3)
    evs = P.synthetic_events("edge", 1, steps, 1000, h, w, seed=7)[0]
    net = EventNetCuda(h, w, STRESS_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=2, max_events_per_step=4096)
    ora = OracleEventNet(h, w, STRESS_LAYERS, wts, 5e-5, 0.1, "SAME")
    mm = compare_live(CudaAdapter(net, stream=1), OracleAdapter(ora), list(evs), exact=False)
    record_parity("live/stress256x320_deeper_edge_10steps", mm)
    assert net.head_shape == (8, 10, 110)
    net.close()


@pytest.mark.parametrize("kind", ["edge", "uniform"])
def test_efcn_sweep_skipping_is_bit_neutral_at_steady_state(kind, monkeypatch):
    """EFCN 160x224 run into its steady state (the regime where the sweep skips most of its work): every head,
    map, index, flag and frontier must be bit-identical with the skip bitmaps switched off (AEC_SWEEP_SKIP=0)."""
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    S, steps = 6, 180
    evs = P.synthetic_events(kind, S, steps, 200, H, W, seed=17)
    nets = []
    for flag in ("1", "0"):
        monkeypatch.setenv("AEC_SWEEP_SKIP", flag)
        nets.append(EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=S))
    for t in range(steps):
        per = [evs[s, t] if (s + t) % 13 else None for s in range(S)]      # an idle stream now and then
        ha, hb = nets[0].step(per), nets[1].step(per)
        assert np.array_equal(ha, hb), "step %d heads differ" % t
        if t == 90:
            for n in nets:
                n.reset(stream_mask=[0, 1, 0, 0, 0, 0])                    # one stream restarts mid-run
    for s in (0, 1, S - 1):
        for i in range(len(nets[0].names)):
            sa, sb = nets[0].state(i, s), nets[1].state(i, s)
            for key in sa:
                assert np.array_equal(sa[key], sb[key]), "stream %d layer %s %s" % (s, nets[0].names[i], key)
            assert np.array_equal(nets[0].frontier(i, s), nets[1].frontier(i, s))
    st = nets[0].sweep_stats()
    assert st["swept_conv_elems"] < st["live_conv_elems"], "at steady state the skip bitmaps must remove work"
    for n in nets:
        n.close()
