"""GPU: the CUDA path (through the C ABI) against the golden vectors minted from the reference and
against the live CPU oracle.  Bars: frontier sets / argmax / flags bit-exact; float maps within
1e-4 relative (parity.FLOAT_RTOL); nets with exactly-representable arithmetic bit-equal."""
import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import CudaAdapter, EventNetCuda
from oracle.event_oracle import OracleEventNet, dense_forward, integrate_frame
from parity import (FLOAT_RTOL, Golden, OracleAdapter, assert_close_map, compare_live, golden_path, record_parity, replay_golden)

pytestmark = pytest.mark.gpu

SMALL = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
DEEP = "conv1=3,3,1,8 conv1b=3,3,8,8 pool1=2,2 conv2=3,3,8,16 conv2b=1,1,16,12 pool2=2,2 conv3=3,3,12,20"


@pytest.mark.parametrize("case,exact,n_streams", [
    ("small32_exact", True, 3),
    ("ragged16", True, 2),
    ("proto8x8", False, 1),
    ("small32_float", False, 2),
    ("efcn_edge", False, 2),
    ("efcn_uniform", False, 1),
])
def test_cuda_matches_reference_golden(case, exact, n_streams):
    g = Golden(golden_path(case))
    net = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=n_streams)
    assert net.shapes() == g.shapes()
    for i, nm in enumerate(g.names):           # state after construction
        st = net.init_state(i)
        if "F" in st:
            assert_close_map(st["F"], g.z["init_F_%s" % nm], exact, "init F %s" % nm)
        if "idx" in st:
            assert np.array_equal(st["idx"], g.z["init_idx_%s" % nm])
    ad = CudaAdapter(net, stream=n_streams - 1)
    if exact:
        mm = replay_golden(ad, g, exact=True, steps=150)
    else:
        # float net: the live oracle runs beside the CUDA path so that every integer disagreement can be explained by
        # the oracle's own values (near ties); the oracle itself is held to the fixture minted from the reference
        ora = OracleEventNet(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME")
        n = min(150, g.n_steps)
        mm = compare_live(ad, OracleAdapter(ora), [g.events(s) for s in range(n)], exact=False, golden=g)
    record_parity("golden/" + case, mm)
    net.close()


@pytest.mark.parametrize("kind,layers,h,w,batch", [
    ("uniform", SMALL, 32, 48, 24),
    ("edge", DEEP, 40, 64, 30),
])
def test_many_streams_against_live_oracle_exact(kind, layers, h, w, batch):
    """Different events per stream; arithmetic exactly representable -> everything bit-equal."""
    S, steps = 6, 60
    wts = P.xavier_weights(layers, seed=9, exact=True)
    evs = P.synthetic_events(kind, S, steps, batch, h, w, seed=5, dt_int=(1, 5))
    net = EventNetCuda(h, w, layers, wts, 1.0 / 64, 0.5, "SAME", n_streams=S)
    oracles = [OracleEventNet(h, w, layers, wts, 1.0 / 64, 0.5, "SAME") for _ in range(S)]
    for t in range(steps):
        per = [evs[s, t] if (s + t) % 5 else None for s in range(S)]     # some streams idle in some steps
        heads = net.step(per)
        delta, active = net.step_info()
        for s in range(S):
            if per[s] is None:
                assert not active[s]
                continue
            ho = oracles[s].step(per[s])
            assert active[s] and delta[s] == oracles[s].delta
            assert np.array_equal(heads[s], ho), "step %d stream %d head" % (t, s)
        if t % 10 == 9:
            for s in range(S):
                oa = OracleAdapter(oracles[s])
                for i in range(len(net.names)):
                    so, sc = oa.state(i), net.state(i, s)
                    for key in so:
                        assert np.array_equal(sc[key], so[key]), "step %d stream %d layer %s %s" % (t, s, net.names[i], key)
                    if per[s] is not None:
                        assert np.array_equal(net.frontier(i, s), oa.frontier(i))
    net.close()


def test_float_net_against_live_oracle_with_explained_mismatches():
    h, w, steps = 48, 64, 120
    wts = P.xavier_weights(DEEP, seed=2)
    evs = P.synthetic_events("uniform", 1, steps, 40, h, w, seed=8, dt_int=(1, 12))[0]
    net = EventNetCuda(h, w, DEEP, wts, 0.004, 0.1, "SAME", n_streams=1)
    ora = OracleEventNet(h, w, DEEP, wts, 0.004, 0.1, "SAME")
    mm = compare_live(CudaAdapter(net), OracleAdapter(ora), list(evs), exact=False)
    record_parity("live/deep48x64_uniform", mm)
    net.close()


def test_layer_at_a_time_equals_fused_step():
    g = Golden(golden_path("small32_float"))
    a = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=2)
    b = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=2)
    for s in range(40):
        ev = g.events(s)
        ha = a.step([ev, ev])
        b.begin_step([ev, ev])
        for li in range(1, len(b.names)):
            b.layer_compute(li)
        b.compute_head()
        for li in range(len(a.names)):
            sa, sb = a.state(li, 1), b.state(li, 1)
            for k in sa:
                assert np.array_equal(sa[k], sb[k])
            assert np.array_equal(a.frontier(li, 1), b.frontier(li, 1))
    a.close()
    b.close()


def test_reset_mask_and_idle_streams():
    g = Golden(golden_path("small32_exact"))
    net = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=3)
    for s in range(10):
        net.step([g.events(s), g.events(s), None])
    fresh = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=1)
    for li in range(len(net.names)):          # stream 2 never received events: still the initial state
        a, b = net.state(li, 2), fresh.state(li, 0)
        for k in a:
            assert np.array_equal(a[k], b[k])
    net.reset(stream_mask=[0, 1, 0])
    for li in range(len(net.names)):
        a, b = net.state(li, 1), fresh.state(li, 0)
        for k in a:
            assert np.array_equal(a[k], b[k]), "reset stream differs at %s" % net.names[li]
    # stream 0 kept its state; replaying the fixture from step 10 on it must still match the golden
    ad = CudaAdapter(net, stream=0, mirror=False)
    for s in range(10, 30):
        head = ad.step(g.events(s))
        assert np.array_equal(head, g.z["heads"][s])
    # the reset stream restarted from scratch
    ad1 = CudaAdapter(net, stream=1, mirror=False)
    for s in range(0, 10):
        assert np.array_equal(ad1.step(g.events(s)), g.z["heads"][s])
    net.close()
    fresh.close()


def test_bad_events_raise_like_the_reference():
    g = Golden(golden_path("small32_exact"))
    net = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=1, max_events_per_step=16)
    with pytest.raises(IndexError):
        net.step(np.array([[g.height, 0, 5]], np.int32))          # y out of range (reference: IndexError)
    with pytest.raises(IndexError):
        net.step(np.zeros((17, 3), np.int32))                      # more than max_events_per_step
    with pytest.raises(ValueError):
        EventNetCuda(8, 8, "conv1=3,3,1,1", {"w_conv1": np.zeros((3, 3, 1, 1)), "b_conv1": np.zeros(1)}, 0.1, padding="FULL")
    with pytest.raises(Exception):
        EventNetCuda(9, 9, "conv1=3,3,1,1 pool1=2,2", {"w_conv1": np.zeros((3, 3, 1, 1)), "b_conv1": np.zeros(1)}, 0.1)
    net.close()


def test_valid_padding_and_odd_kernels_exact():
    layers = "conv1=5,5,1,4 conv2=3,3,4,4 pool1=2,2 conv3=1,1,4,4"
    h, w = 30, 34                                      # VALID: 30x34 -> 26x30 -> 24x28 -> 12x14
    wts = P.xavier_weights(layers, seed=6, exact=True)
    evs = P.synthetic_events("uniform", 1, 50, 15, h, w, seed=4, dt_int=(1, 5))[0]
    net = EventNetCuda(h, w, layers, wts, 1.0 / 64, 0.5, "VALID", n_streams=1)
    ora = OracleEventNet(h, w, layers, wts, 1.0 / 64, 0.5, "VALID")
    compare_live(CudaAdapter(net), OracleAdapter(ora), list(evs), exact=True)
    net.close()


def test_event_vs_frame_equivalence_on_gpu():
    """test_correctness.py protocol with the CUDA event net: event-driven == dense frame network."""
    g = Golden(golden_path("proto8x8"))
    w = g.weights()
    net = EventNetCuda(g.height, g.width, g.layers, w, g.leak, g.alpha, "SAME", n_streams=1)
    state = None
    for s in range(300):
        ev = g.events(s)
        frame, ts = integrate_frame(ev, g.leak, g.height, g.width, state)
        state = (frame, ts)
        net.step(ev)
        dense = dense_forward(frame, g.layers, w, g.alpha, "SAME")
        for i, d in enumerate(dense, start=1):
            fm = net.view(i, 0, which=("featuremap",))["featuremap"]
            assert np.allclose(fm, d, rtol=1e-5, atol=1e-5), "step %d layer %s" % (s, net.names[i])
    net.close()


def test_pipelined_host_steps_equal_blocking_steps():
    """aec_net_step_host_async (two steps in flight, separate copy streams) == aec_net_step_host, bit for bit."""
    g = Golden(golden_path("small32_float"))
    S, steps = 3, 24
    a = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=S)
    b = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=S)
    from async_ev_cnn_b200.engine import pack_events
    packed = [pack_events([g.events(t), g.events(t) if t % 3 else None, g.events((t + 5) % g.n_steps)]) for t in range(steps)]
    want = [a.step_packed(ev, off).copy() for ev, off in packed]
    outs = [np.empty_like(want[0]) for _ in range(steps)]
    for t, (ev, off) in enumerate(packed):
        b.step_packed_async(ev, off, outs[t])
        if t % 5 == 4:
            b.host_sync()
    b.host_sync()
    for t in range(steps):
        assert np.array_equal(outs[t], want[t]), "step %d" % t
    for li in range(len(a.names)):
        sa, sb = a.state(li, 2), b.state(li, 2)
        for k in sa:
            assert np.array_equal(sa[k], sb[k])
    with pytest.raises(IndexError):
        b.step_packed_async(np.array([[g.height, 0, 5]], np.int32), np.array([0, 1, 1, 1], np.int32), outs[0])
        b.host_sync()
    a.close()
    b.close()


def test_decode_head_kernel_equals_oracle():
    from oracle import frontend as F
    g = Golden(golden_path("efcn_edge"))
    net = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=3)
    for s in range(6):
        heads = net.step([g.events(s), g.events(s + 1), None]).copy()
    C, B, gh, gw = 100, 2, 5, 7
    boxes, conf, valid, label = net.decode_head(C, B, gh, gw, conf_threshold=0.1)
    ob, oc, ov, ol = F.decode_head(heads, gh, gw, C, g.height, g.width, 0.1)
    assert np.array_equal(boxes, ob) and np.array_equal(conf, oc) and np.array_equal(valid, ov) and np.array_equal(label, ol)
    with pytest.raises(Exception):
        net.decode_head(C, B, gh, gw + 1)
    net.close()


def test_ndata_decode_kernel_equals_oracle():
    from oracle import frontend as F
    from async_ev_cnn_b200.frontend import decode_ndata
    rng = np.random.default_rng(11)
    recs, want = [], {}
    for r, n in enumerate([0, 1, 7, 255, 256, 257, 5000, 40000]):
        x = rng.integers(0, 232, n).astype(np.int32)
        y = rng.integers(0, 172, n).astype(np.int32)
        ts = np.sort(rng.integers(0, 1 << 13, n)).astype(np.int32)
        p = rng.integers(0, 2, n).astype(np.int32)
        if n > 10:
            y[rng.integers(0, n, max(1, n // 300))] = 240          # overflow markers, a few in a row sometimes
            y[3:5] = 240
        recs.append(F.encode_ndata(x, y, ts, p))
    for crop in (None, (160, 224)):
        for zero in (True, False):
            got, pols = decode_ndata(recs, zero_base_ts=zero, crop_to=crop, with_polarity=True)
            for r, raw in enumerate(recs):
                n, x, y, ts, p = F.read_ndata(raw)
                if zero and n:
                    ts = ts - ts[0]
                if crop is not None and n:
                    x, y, ts, p = F.center_crop_events(x, y, ts, p, crop)
                ev = np.stack([y, x, ts], axis=-1).astype(np.int32) if len(x) else np.zeros((0, 3), np.int32)
                assert np.array_equal(got[r], ev), "recording %d crop %s zero %s" % (r, crop, zero)
                assert np.array_equal(pols[r], np.asarray(p, np.int32).reshape(-1))
    with pytest.raises(ValueError):
        decode_ndata([np.zeros(7, np.uint8)])


@pytest.mark.parametrize("layers,h,w", [
    ("conv1=3,3,1,4 conv2=7,7,4,8 pool1=2,2 conv3=1,1,8,4", 28, 36),        # 49 taps: the SIMT GEMM fallback (tap bitmask is 32 wide)
    ("conv1=3,3,1,8 pool1=3,3 conv2=3,3,8,12 pool2=3,3 conv3=3,3,12,4", 36, 54),   # 3x3 / stride-3 pools: generic window path
    ("conv1=3,3,1,6 conv2=3,3,6,10 pool1=2,2 conv3=3,3,10,7", 24, 40),       # channel counts that are not multiples of 4
])
def test_unusual_shapes_against_live_oracle_exact(layers, h, w):
    S, steps = 3, 40
    wts = P.xavier_weights(layers, seed=13, exact=True)
    evs = P.synthetic_events("uniform", S, steps, 20, h, w, seed=17, dt_int=(1, 5))
    net = EventNetCuda(h, w, layers, wts, 1.0 / 64, 0.5, "SAME", n_streams=S)
    oracles = [OracleEventNet(h, w, layers, wts, 1.0 / 64, 0.5, "SAME") for _ in range(S)]
    for t in range(steps):
        heads = net.step([evs[s, t] for s in range(S)])
        for s in range(S):
            assert np.array_equal(heads[s], oracles[s].step(evs[s, t])), "step %d stream %d head" % (t, s)
        if t % 8 == 7:
            for s in range(S):
                oa = OracleAdapter(oracles[s])
                for i in range(len(net.names)):
                    so, sc = oa.state(i), net.state(i, s)
                    for key in so:
                        assert np.array_equal(sc[key], so[key]), "step %d stream %d layer %s %s" % (t, s, net.names[i], key)
                    assert np.array_equal(net.frontier(i, s), oa.frontier(i))
    net.close()


@pytest.mark.parametrize("layers,h,w,padding", [
    # conv2: 16 -> 32 channels on a 24-wide map: 4 output rows per unit (slot 32); conv3: 32 -> 8 on 12 columns: 8 rows per unit
    ("conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,32 pool2=2,2 conv3=3,3,32,8", 32, 48, "SAME"),
    # rows wider than one tile: x segments of 126 sites (272 columns = 3 segments), 64-byte and 128-byte tile rows
    ("conv1=3,3,1,16 conv2=3,3,16,16 conv3=3,3,16,32 conv4=3,3,32,12", 10, 272, "SAME"),
    # VALID padding and a 5x5 window (tile rows 0..131, taps read the tile from 0..4 rows further on)
    ("conv1=3,3,1,16 conv2=5,5,16,32 pool1=2,2 conv3=3,3,32,64", 30, 70, "VALID"),
    # two channel blocks per pixel (Cin = 64)
    ("conv1=3,3,1,16 conv2=3,3,16,64 conv3=3,3,64,16", 20, 40, "SAME"),
])
def test_row_tile_conv_layers_exact(layers, h, w, padding):
    """Layers that take the row-tile kernel (aec_rt.cuh: Cin = 16 or a multiple of 32, Cout <= 64) on exactly
    representable nets: every map, frontier, argmax and flag bit-equal to the oracle, several streams with idle steps."""
    S, steps = 3, 40
    wts = P.xavier_weights(layers, seed=21, exact=True)
    evs = P.synthetic_events("edge", S, steps, 30, h, w, seed=23, dt_int=(1, 5))
    net = EventNetCuda(h, w, layers, wts, 1.0 / 64, 0.5, padding, n_streams=S)
    assert any(g is not None and "k_conv_rows" in g["kernel"] for g in (net.tc_geometry(i) for i in range(1, len(net.names)))), \
        "no layer of this net took the row-tile kernel"
    oracles = [OracleEventNet(h, w, layers, wts, 1.0 / 64, 0.5, padding) for _ in range(S)]
    for t in range(steps):
        per = [evs[s, t] if (s + t) % 6 else None for s in range(S)]
        heads = net.step(per)
        for s in range(S):
            if per[s] is not None:
                assert np.array_equal(heads[s], oracles[s].step(per[s])), "step %d stream %d head" % (t, s)
        if t % 8 == 7 or t == steps - 1:
            for s in range(S):
                oa = OracleAdapter(oracles[s])
                for i in range(len(net.names)):
                    so, sc = oa.state(i), net.state(i, s)
                    for key in so:
                        assert np.array_equal(sc[key], so[key]), "step %d stream %d layer %s %s" % (t, s, net.names[i], key)
                    if per[s] is not None:
                        assert np.array_equal(net.frontier(i, s), oa.frontier(i))
    units = net.unit_counters()
    assert units.sum() > 0
    net.close()


def test_long_run_through_pixel_deaths_exact():
    """300 steps with a leak that kills a pixel within a few steps: the surface kernel walks only the pixels alive
    after the previous step (its alive bitmap), so births, deaths and re-births must keep that bitmap exact."""
    layers, h, w, S, steps, batch = SMALL, 32, 64, 3, 300, 12
    wts = P.xavier_weights(layers, seed=2, exact=True)
    evs = P.synthetic_events("edge", S, steps, batch, h, w, seed=31, dt_int=(1, 9))
    net = EventNetCuda(h, w, layers, wts, 1.0 / 128, 0.25, "SAME", n_streams=S)
    oracles = [OracleEventNet(h, w, layers, wts, 1.0 / 128, 0.25, "SAME") for _ in range(S)]
    deaths, alive_before = 0, [None] * S
    for t in range(steps):
        per = [evs[s, t] if (s * 7 + t) % 11 else None for s in range(S)]
        heads = net.step(per)
        for s in range(S):
            if per[s] is None:
                continue
            ho = oracles[s].step(per[s])
            assert np.array_equal(heads[s], ho), "step %d stream %d head" % (t, s)
        if t % 25 == 24:
            for s in range(S):
                surf_o = OracleAdapter(oracles[s]).state(0)["S"]
                assert np.array_equal(net.state(0, s)["S"], surf_o), "step %d stream %d surface" % (t, s)
                alive = np.asarray(surf_o) > 0
                if alive_before[s] is not None:
                    deaths += int((alive_before[s] & ~alive).sum())
                alive_before[s] = alive
    assert deaths > 0, "the test is meant to run through pixel deaths"
    for s in range(S):
        oa = OracleAdapter(oracles[s])
        for i in range(len(net.names)):
            so, sc = oa.state(i), net.state(i, s)
            for key in so:
                assert np.array_equal(sc[key], so[key]), "final state stream %d layer %s %s" % (s, net.names[i], key)
    net.close()


def test_duplicate_pixels_and_event_count_boundary_exact():
    """All events of a step on one pixel (last duplicate wins, integration.py:71-80), a step of exactly
    max_events_per_step events, and one more than that (rejected, state untouched)."""
    layers, h, w = SMALL, 32, 48
    wts = P.xavier_weights(layers, seed=4, exact=True)
    cap = 64
    net = EventNetCuda(h, w, layers, wts, 1.0 / 64, 0.5, "SAME", n_streams=2, max_events_per_step=cap)
    orc = OracleEventNet(h, w, layers, wts, 1.0 / 64, 0.5, "SAME")
    rng = np.random.default_rng(0)
    ts = 0

    def batch(n, same_pixel):
        nonlocal ts
        t = ts + np.cumsum(rng.integers(1, 4, size=n))
        ts = int(t[-1])
        if same_pixel:
            y = np.full(n, 7); x = np.full(n, 9)
        else:
            y = rng.integers(0, h, size=n); x = rng.integers(0, w, size=n)
        return np.stack([y, x, t], axis=1).astype(np.int32)

    for n, same in [(cap, True), (cap, False), (1, False), (cap, True), (5, False)]:
        ev = batch(n, same)
        heads = net.step([ev, ev])
        ho = orc.step(ev)
        assert np.array_equal(heads[0], ho) and np.array_equal(heads[1], ho)
    oa = OracleAdapter(orc)
    for i in range(len(net.names)):
        so, sc = oa.state(i), net.state(i, 1)
        for key in so:
            assert np.array_equal(sc[key], so[key]), "layer %s %s" % (net.names[i], key)
    with pytest.raises(IndexError):
        net.step([batch(cap + 1, False), None])
    ev = batch(3, False)                                   # the rejected call left no trace
    heads = net.step([ev, ev])
    assert np.array_equal(heads[1], orc.step(ev))
    net.close()


def test_diagnostics_api_is_consistent_with_the_maps():
    """aec_net_sweep_stats / aec_net_tc_timing: the live-site bitmap must cover every element with a non-zero rate
    (it decides what the leak sweep touches), and the role timers must tick on the tensor-core layers."""
    g = Golden(golden_path("efcn_edge"))
    net = EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=2)
    ad = CudaAdapter(net, stream=1)
    for s in range(6):
        ad.step(g.events(s))
    st = net.sweep_stats()
    nz = tot = 0
    for i, nm in enumerate(net.names):
        if "conv" in nm:
            for s in range(2):
                a = net.state(i, s)["A"]
                nz += int(np.count_nonzero(a)); tot += a.size
    assert st["conv_elems"] == tot
    assert nz <= st["live_conv_elems"] <= tot
    assert 0 < st["nz_groups"] <= st["groups"]
    assert len(net.tc_layers()) >= 5
    net.tc_timing(True)
    ad.step(g.events(6))
    tm = net.read_tc_timing()
    net.tc_timing(False)
    assert tm, "no tensor-core layer reported timing"
    for nm, d in tm.items():
        assert d["ctas"] > 0 and d["mma_total"] > 0 and d["prod_total"] > 0 and d["epi_total"] > 0, (nm, d)
    assert np.array_equal(ad.step(g.events(7)).shape, g.z["heads"][7].shape)
    net.close()


def test_sweep_skipping_changes_no_bit(monkeypatch):
    """The leak sweep leaves alone the sites that the same step re-evaluates (k_frontier_skip computes a subset of
    every layer's work set before the sweep).  With the skip switched off (AEC_SWEEP_SKIP=0) every map, index and
    frontier must come out bit-identical, on a float net with sign flips and sticky pool flags."""
    g = Golden(golden_path("small32_float"))
    nets = []
    for flag in ("1", "0"):
        monkeypatch.setenv("AEC_SWEEP_SKIP", flag)
        nets.append(EventNetCuda(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME", n_streams=2))
    n_steps = min(120, g.z["heads"].shape[0])
    for t in range(n_steps):
        ev = g.events(t)
        per = [ev, ev if t % 3 else None]
        ha, hb = nets[0].step(per), nets[1].step(per)
        assert np.array_equal(ha[0], hb[0]), "step %d head" % t
        if t % 10 == 9 or t == n_steps - 1:
            for s in range(2):
                for i in range(len(nets[0].names)):
                    sa, sb = nets[0].state(i, s), nets[1].state(i, s)
                    for key in sa:
                        assert np.array_equal(sa[key], sb[key]), "step %d stream %d layer %s %s" % (t, s, nets[0].names[i], key)
                    assert np.array_equal(nets[0].frontier(i, s), nets[1].frontier(i, s))
    st = nets[0].sweep_stats()
    assert st["swept_conv_elems"] <= st["live_conv_elems"]
    assert nets[1].sweep_stats()["swept_conv_elems"] == nets[1].sweep_stats()["live_conv_elems"]
    for n in nets:
        n.close()


POOL_IN_CONV = "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,96 pool2=2,2 conv3=3,3,96,160 pool3=2,2 conv4=1,1,160,12"


def test_pool_in_conv_epilogue_exact():
    """Conv layers with more than 64 channels run on the gathered weights-as-M kernel, whose epilogue evaluates the 2x2 pool
    behind them for the windows all four sites of which are re-evaluated (work list ordered by window, aec_tc.cuh kPool;
    the rest of the pool's work set stays with k_pool_eval).  One weight tile (96 channels) and two (160); exactly
    representable arithmetic -> maps, argmax rows, copies, flags and frontiers bit-equal to the oracle
    (maxpool.py:118-151, cutils.pyx:161-177)."""
    S, steps, h, w = 5, 50, 32, 48
    wts = P.xavier_weights(POOL_IN_CONV, seed=21, exact=True)
    evs = P.synthetic_events("edge", S, steps, 24, h, w, seed=23, dt_int=(1, 5))
    net = EventNetCuda(h, w, POOL_IN_CONV, wts, 1.0 / 64, 0.5, "SAME", n_streams=S)
    oracles = [OracleEventNet(h, w, POOL_IN_CONV, wts, 1.0 / 64, 0.5, "SAME") for _ in range(S)]
    for t in range(steps):
        per = [evs[s, t] if (s + 2 * t) % 7 else None for s in range(S)]
        heads = net.step(per)
        for s in range(S):
            if per[s] is None:
                continue
            assert np.array_equal(heads[s], oracles[s].step(per[s])), "step %d stream %d head" % (t, s)
        if t % 5 == 4:
            for s in range(S):
                oa = OracleAdapter(oracles[s])
                for i in range(len(net.names)):
                    so, sc = oa.state(i), net.state(i, s)
                    for key in so:
                        assert np.array_equal(sc[key], so[key]), "step %d stream %d layer %s %s" % (t, s, net.names[i], key)
                    if per[s] is not None:
                        assert np.array_equal(net.frontier(i, s), oa.frontier(i))
    units = net.unit_counters()
    names = list(net.names)
    assert units[names.index("pool2")] > 0 and units[names.index("pool3")] > 0, "no window was evaluated in a conv epilogue"
    net.close()


def test_pool_in_conv_epilogue_changes_no_bit(monkeypatch):
    """The same float net with the fusion on and off (AEC_POOL_FUSE=0: every pool window goes through k_pool_eval): heads,
    maps, argmax rows, (Fp, Ap) copies, flags and frontiers must be bit-identical, and the per-layer work counters equal."""
    S, steps, h, w = 3, 80, 32, 48
    wts = P.xavier_weights(POOL_IN_CONV, seed=4)
    evs = P.synthetic_events("edge", S, steps, 30, h, w, seed=8)
    nets = []
    for flag in ("1", "0"):
        monkeypatch.setenv("AEC_POOL_FUSE", flag)
        nets.append(EventNetCuda(h, w, POOL_IN_CONV, wts, 5e-5, 0.1, "SAME", n_streams=S))
    for t in range(steps):
        per = [evs[s, t] if (s + t) % 4 else None for s in range(S)]
        ha, hb = nets[0].step(per), nets[1].step(per)
        assert np.array_equal(ha, hb), "step %d head" % t
        if t % 8 == 7 or t == steps - 1:
            for s in range(S):
                for i in range(len(nets[0].names)):
                    sa, sb = nets[0].state(i, s), nets[1].state(i, s)
                    for key in sa:
                        assert np.array_equal(sa[key], sb[key]), "step %d stream %d layer %s %s" % (t, s, nets[0].names[i], key)
                    assert np.array_equal(nets[0].frontier(i, s), nets[1].frontier(i, s))
    ca, cb = nets[0].counters()[0], nets[1].counters()[0]
    assert np.array_equal(ca, cb), (ca, cb)
    names = list(nets[0].names)
    for nm in ("pool2", "pool3"):
        assert nets[0].unit_counters()[names.index(nm)] > 0 and nets[1].unit_counters()[names.index(nm)] == 0
    for n in nets:
        n.close()


PAIR_NET = "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,96 pool2=2,2 conv3=3,3,96,160 pool3=2,2 conv4=3,3,160,392 conv5=1,1,392,12"


def test_pair_units_change_no_bit(monkeypatch):
    """Layers with an even number of weight tiles run on CTA pairs (tcgen05 cta_group::2, M = 256: each CTA of a cluster
    holds one weight tile and converts half of the unit's sites, aec_tc.cuh kPair) when there are many streams.  An
    accumulator column sees the same instruction sequence as on a single CTA: AEC_TC_PAIR=1 (forced for this small job)
    and AEC_TC_PAIR=0 must give bit-identical heads, maps, pool state and frontiers.  conv3 (two tiles, pool in its
    epilogue) and conv4 (four tiles, plain epilogue) take both variants of the pair kernel."""
    S, steps, h, w = 6, 40, 32, 48
    wts = P.xavier_weights(PAIR_NET, seed=16)
    evs = P.synthetic_events("uniform", S, steps, 30, h, w, seed=12)
    nets = []
    for flag in ("1", "0"):
        monkeypatch.setenv("AEC_TC_PAIR", flag)
        nets.append(EventNetCuda(h, w, PAIR_NET, wts, 5e-5, 0.1, "SAME", n_streams=S))
    for t in range(steps):
        per = [evs[s, t] for s in range(S)]
        ha, hb = nets[0].step(per), nets[1].step(per)
        assert np.array_equal(ha, hb), "step %d head" % t
        if t % 8 == 7 or t == steps - 1:
            for s in range(S):
                for i in range(len(nets[0].names)):
                    sa, sb = nets[0].state(i, s), nets[1].state(i, s)
                    for key in sa:
                        assert np.array_equal(sa[key], sb[key]), "step %d stream %d layer %s %s" % (t, s, nets[0].names[i], key)
                    assert np.array_equal(nets[0].frontier(i, s), nets[1].frontier(i, s))
    for n in nets:
        n.close()


def test_pair_units_exact_against_oracle(monkeypatch):
    """The CTA-pair kernel against the oracle on exactly representable arithmetic (bit-equal maps, argmax rows, flags and
    frontiers; conv3 of the net has two weight tiles and the pool in its epilogue): conv2d.py:118-123,144-181, maxpool.py:118-151."""
    monkeypatch.setenv("AEC_TC_PAIR", "1")
    S, steps, h, w = 3, 30, 32, 48
    wts = P.xavier_weights(POOL_IN_CONV, seed=31, exact=True)
    evs = P.synthetic_events("edge", S, steps, 24, h, w, seed=33, dt_int=(1, 5))
    net = EventNetCuda(h, w, POOL_IN_CONV, wts, 1.0 / 64, 0.5, "SAME", n_streams=S)
    oracles = [OracleEventNet(h, w, POOL_IN_CONV, wts, 1.0 / 64, 0.5, "SAME") for _ in range(S)]
    for t in range(steps):
        per = [evs[s, t] if (s + 2 * t) % 7 else None for s in range(S)]
        heads = net.step(per)
        for s in range(S):
            if per[s] is None:
                continue
            assert np.array_equal(heads[s], oracles[s].step(per[s])), "step %d stream %d head" % (t, s)
        if t % 5 == 4:
            for s in range(S):
                oa = OracleAdapter(oracles[s])
                for i in range(len(net.names)):
                    so, sc = oa.state(i), net.state(i, s)
                    for key in so:
                        assert np.array_equal(sc[key], so[key]), "step %d stream %d layer %s %s" % (t, s, net.names[i], key)
                    if per[s] is not None:
                        assert np.array_equal(net.frontier(i, s), oa.frontier(i))
    net.close()


def test_half_units_change_no_bit(monkeypatch):
    """With few sites on a layer's work list the gathered kernel cuts units of 64 sites (one MMA of N = 128 per product)
    instead of 128, so that more CTAs work on a small job.  Same accumulation order per site: AEC_TC_HALF=0 (always 128-site
    units) must give bit-identical heads, maps, pool state and frontiers - which also runs this small net through the
    full-unit path that otherwise only many-stream workloads take."""
    S, steps, h, w = 3, 40, 32, 48
    wts = P.xavier_weights(POOL_IN_CONV, seed=6)
    evs = P.synthetic_events("uniform", S, steps, 30, h, w, seed=10)
    nets = []
    for flag in ("1", "0"):
        monkeypatch.setenv("AEC_TC_HALF", flag)
        nets.append(EventNetCuda(h, w, POOL_IN_CONV, wts, 5e-5, 0.1, "SAME", n_streams=S))
    for t in range(steps):
        per = [evs[s, t] for s in range(S)]
        ha, hb = nets[0].step(per), nets[1].step(per)
        assert np.array_equal(ha, hb), "step %d head" % t
        if t % 8 == 7 or t == steps - 1:
            for s in range(S):
                for i in range(len(nets[0].names)):
                    sa, sb = nets[0].state(i, s), nets[1].state(i, s)
                    for key in sa:
                        assert np.array_equal(sa[key], sb[key]), "step %d stream %d layer %s %s" % (t, s, nets[0].names[i], key)
                    assert np.array_equal(nets[0].frontier(i, s), nets[1].frontier(i, s))
    for n in nets:
        n.close()
