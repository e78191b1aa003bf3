"""CPU: the claim behind the leak sweep's skip bitmaps (k_frontier_skip), checked on the oracle.

The frontier chain run with the sign flips left out - event pixels and pixel deaths, dilation through the conv
kernels, pool windows, sticky flags as they stand BEFORE the step - must give, for every layer, a subset of the
sites / windows the reference re-evaluates in that step (conv2d.py:118-123, maxpool.py:118-151).  Only then may the
sweep leave those sites alone.  The test also measures how large the subset is, so that a change that makes it
trivially empty does not pass unnoticed."""
import numpy as np

import async_ev_cnn_b200 as P
from oracle.event_oracle import OracleConv, OracleEventNet, OraclePool

LAYERS = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
H, W = 32, 48


def _conv_reach(prev, layer):
    """Sites of a SAME/VALID conv whose receptive field contains a set bit of `prev` (cutils.pyx:78-105)."""
    _, ho, wo = layer.shape
    kh, kw = layer.K.shape[2], layer.K.shape[3]
    pt, pl = layer.pad[0], layer.pad[2]
    out = np.zeros((ho, wo), bool)
    hin, win = prev.shape
    for ky in range(kh):
        for kx in range(kw):
            # out (y, x) reads in (y + ky - pt, x + kx - pl)
            ys = np.arange(ho) + ky - pt
            xs = np.arange(wo) + kx - pl
            vy, vx = (ys >= 0) & (ys < hin), (xs >= 0) & (xs < win)
            sub = np.zeros((ho, wo), bool)
            sub[np.ix_(vy, vx)] = prev[np.ix_(ys[vy], xs[vx])]
            out |= sub
    return out


def _pool_hit(prev, layer):
    _, ho, wo = layer.shape
    s = layer.stride
    return prev[: ho * s, : wo * s].reshape(ho, s, wo, s).any(axis=(1, 3))


def test_flip_free_frontier_is_a_subset_of_every_work_set():
    wts = P.xavier_weights(LAYERS, seed=7)
    net = OracleEventNet(H, W, LAYERS, wts, 0.002, 0.1, "SAME")
    evs = P.synthetic_events("edge", 1, 80, 20, H, W, seed=3, dt_int=(1, 30))[0]
    evaluated = {}
    for i, layer in enumerate(net.layers):
        if isinstance(layer, OracleConv):
            orig = layer._event_conv

            def wrapped(img, events, bias, _orig=orig, _i=i):
                vals, (oy, ox) = _orig(img, events, bias)
                m = np.zeros(net.layers[_i].shape[1:], bool)
                m[oy, ox] = True
                evaluated[_i] = m
                return vals, (oy, ox)
            layer._event_conv = wrapped
    covered = total = flips_seen = 0
    for t in range(80):
        flags_before = {i: l.flags.copy() for i, l in enumerate(net.layers) if isinstance(l, OraclePool)}
        net.step(evs[t])
        g = net.frontier_mask(0)                       # known before the leak: event pixels and pixel deaths
        for i in range(1, len(net.layers)):
            layer = net.layers[i]
            if isinstance(layer, OracleConv):
                g = _conv_reach(g, layer)
                work = evaluated[i]
                flips_seen += int((net.frontier_mask(i) & ~work).sum())
            else:
                g = _pool_hit(g, layer) | flags_before[i]
                work = net.frontier_mask(i)            # a pool layer reports every evaluated window (maxpool.py:153-154)
            assert not (g & ~work).any(), "step %d layer %s: skip set is not a subset of the work set" % (t, net.names[i])
            covered += int(g.sum())
            total += int(work.sum())
    assert flips_seen > 0, "the stream never produced a sign flip outside the evaluated sites: the test would be vacuous"
    assert covered > 0.5 * total, "the flip-free chain should cover most of the work sets (got %d of %d)" % (covered, total)
