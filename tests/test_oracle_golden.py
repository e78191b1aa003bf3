"""CPU: the oracle port replayed against the golden vectors minted from the reference itself."""
import numpy as np
import pytest

from oracle.event_oracle import OracleEventNet, dense_forward, integrate_frame
from parity import Golden, OracleAdapter, golden_path, replay_golden

CASES = ["proto8x8", "small32_float", "small32_exact", "ragged16", "efcn_uniform", "efcn_edge"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    g = Golden(golden_path(case))
    net = OracleEventNet(g.height, g.width, g.layers, g.weights(), g.leak, g.alpha, "SAME")
    # same numpy/BLAS calls as the reference -> bit-equal, so hold the port to exactness everywhere
    replay_golden(OracleAdapter(net), g, exact=True, steps=None if case.startswith("efcn") else 200)


def test_event_vs_frame_equivalence_protocol():
    """src/scripts/test_correctness.py:92-171 with the TF half restated densely: event net ==
    dense net (conv -> leaky -> pool, one activation) on the leaky frame, np.allclose per layer."""
    g = Golden(golden_path("proto8x8"))
    w = g.weights()
    net = OracleEventNet(g.height, g.width, g.layers, w, g.leak, g.alpha, "SAME")
    state = None
    for s in range(g.n_steps):
        ev = g.events(s)
        frame, ts = integrate_frame(ev, g.leak, g.height, g.width, state)
        state = (frame, ts)
        net.step(ev)
        dense = dense_forward(frame, g.layers, w, g.alpha, "SAME")
        for i, d in enumerate(dense, start=1):
            assert np.allclose(net.layers[i].featuremap(), d), "step %d layer %s" % (s, net.names[i])


def test_event_vs_frame_equivalence_multichannel():
    g = Golden(golden_path("small32_float"))
    w = g.weights()
    net = OracleEventNet(g.height, g.width, g.layers, w, g.leak, g.alpha, "SAME")
    state = None
    for s in range(g.n_steps):
        ev = g.events(s)
        frame, ts = integrate_frame(ev, g.leak, g.height, g.width, state)
        state = (frame, ts)
        net.step(ev)
        dense = dense_forward(frame, g.layers, w, g.alpha, "SAME")
        for i, d in enumerate(dense, start=1):
            fm = net.layers[i].featuremap()
            assert np.abs(fm - d).max() <= 1e-4 * max(1.0, np.abs(d).max()), "step %d layer %s" % (s, net.names[i])
