"""CPU, in-container only: oracle port side by side with the UNMODIFIED reference layers on fresh
random streams (skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle.event_oracle import OracleEventNet
from oracle.ref_loader import reference_available
import async_ev_cnn_b200 as P

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present")

SMALL = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 conv2b=3,3,8,8 pool2=2,2 conv3=1,1,8,6"


def _ref_chain(h, w, layers, weights, leak, alpha):
    from oracle.ref_loader import load_reference_layers
    I, C, Pl, _, _ = load_reference_layers()
    chain = [I(leak, h, w)]
    for name, size in P.parse_layers(layers).items():
        chain.append(C(chain[-1], weights["w_" + name], weights["b_" + name], 1, alpha, "SAME") if "conv" in name
                     else Pl(chain[-1], size, size[0]))
    return chain


@pytest.mark.parametrize("kind,h,w,layers,steps,batch,leak,seed", [
    ("uniform", 24, 40, SMALL, 120, 12, 0.02, 1),
    ("edge", 24, 40, SMALL, 120, 12, 0.02, 2),
    ("uniform", 160, 224, P.EFCN_LAYERS, 6, 200, 5e-5, 3),
])
def test_port_is_bit_equal_to_reference(kind, h, w, layers, steps, batch, leak, seed):
    wts = P.xavier_weights(layers, seed=seed)
    ref = _ref_chain(h, w, layers, wts, leak, 0.1)
    ora = OracleEventNet(h, w, layers, wts, leak, 0.1, "SAME")
    evs = P.synthetic_events(kind, 1, steps, batch, h, w, seed=seed + 10, dt_int=(1, 12) if h < 100 else None)[0]
    for s in range(steps):
        e, d = ref[0].compute(evs[s], None)
        fr = [e]
        for layer in ref[1:]:
            e, d = layer.compute(e, d)
            fr.append(e)
        head = ora.step(evs[s])
        assert np.array_equal(ref[-1].featuremap().transpose(1, 2, 0), head)
        assert d == ora.delta
        for i, (lr, lo) in enumerate(zip(ref, ora.layers)):
            assert np.array_equal(np.asarray(fr[i][0]), np.asarray(ora.frontiers[i][0]))
            assert np.array_equal(np.asarray(fr[i][1]), np.asarray(ora.frontiers[i][1]))
            assert lr.surface().dtype == lo.surface().dtype
            assert np.array_equal(lr.surface(), lo.surface())
            assert np.array_equal(lr.conv_actfn(), lo.conv_actfn())
            if hasattr(lr, "_idx_max"):
                assert np.array_equal(lr._idx_max[0], lo.idx)
                assert np.array_equal(lr._recompute_coords, lo.flags)
