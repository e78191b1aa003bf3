"""CPU: the parity harness itself (tests/parity.py:compare_live).  The GPU parity tests lean on its rule that a
float net may disagree with the oracle on integer state only where the oracle's own values explain it, so the
rule is exercised here without a GPU: identical nets give all-zero counters, a net whose weights differ in the
last bits passes with every disagreement explained, and injected faults (a frontier bit, an argmax flip away from
a tie, a recompute flag) are rejected."""
import numpy as np
import pytest

import async_ev_cnn_b200 as P
from oracle.event_oracle import OracleEventNet
from parity import Mismatch, OracleAdapter, compare_live

LAYERS = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
H, W = 32, 48


def _nets(perturb=0.0, seed=3):
    wts = P.xavier_weights(LAYERS, seed=seed)
    a = OracleEventNet(H, W, LAYERS, wts, 0.004, 0.1, "SAME")
    if perturb:
        rng = np.random.default_rng(1)
        wts = {k: (v * (1 + perturb * rng.standard_normal(v.shape))).astype(np.float32) if k.startswith("w_") else v for k, v in wts.items()}
    b = OracleEventNet(H, W, LAYERS, wts, 0.004, 0.1, "SAME")
    return a, b


def _events(steps, seed=8):
    return list(P.synthetic_events("uniform", 1, steps, 30, H, W, seed=seed, dt_int=(1, 12))[0])


class Faulty(OracleAdapter):
    """Oracle adapter with one injected fault at step `at`."""

    def __init__(self, net, fault, at=5):
        super().__init__(net)
        self.fault, self.at, self.t = fault, at, -1

    def step(self, events):
        self.t += 1
        return super().step(events)

    def frontier(self, i):
        f = super().frontier(i).copy()
        if self.fault == "frontier" and self.t == self.at and self.names[i] == "conv2":
            ys, xs = np.nonzero(~f)
            f[ys[len(ys) // 2], xs[len(xs) // 2]] = True
        return f

    def state(self, i):
        st = {k: np.array(v, copy=True) for k, v in super().state(i).items()}
        if self.t == self.at and self.names[i] == "pool1":
            if self.fault == "argmax":
                F = self.net.layers[i - 1].F                      # flip towards a clearly smaller candidate
                wins = F[0].reshape(F.shape[1] // 2, 2, F.shape[2] // 2, 2).transpose(0, 2, 1, 3).reshape(F.shape[1] // 2, F.shape[2] // 2, 4)
                spread = wins.max(axis=2) - wins.min(axis=2)
                y, x = np.unravel_index(np.argmax(spread), spread.shape)
                st["idx"][0, y, x] = int(np.argmin(wins[y, x]))
                assert spread[y, x] > 1e-3 * np.abs(F).max()
            if self.fault == "flag":
                conv = self.net.layers[i - 1]                     # the window whose argmax rate is farthest from the smallest rate
                R = (conv.A * np.where(conv.F > 0, 1.0, conv.alpha)).astype(np.float32)
                c, h, w = R.shape
                Rw = R.reshape(c, h // 2, 2, w // 2, 2).transpose(0, 1, 3, 2, 4).reshape(c, h // 2, w // 2, 4)
                r_arg = np.take_along_axis(Rw, st["idx"][..., None].astype(np.int64), axis=3)[..., 0]
                gap = (r_arg - Rw.min(axis=3)).max(axis=0)
                y, x = np.unravel_index(np.argmax(gap), gap.shape)
                assert gap[y, x] > 1e-3 * np.abs(R).max()
                st["flags"][y, x] = not st["flags"][y, x]
        return st


def test_identical_nets_have_nothing_to_explain():
    a, b = _nets()
    mm = compare_live(OracleAdapter(b), OracleAdapter(a), _events(40), exact=False)
    d = mm.as_dict()
    assert d["steps"] == 40 and d["front_total"] > 0 and d["idx_total"] > 0
    for k in ("front_explained", "front_unexplained", "idx_roots", "idx_downstream", "flag_roots", "flag_explained",
              "flag_unexplained", "slope_flips", "bad_sites"):
        assert d[k] == 0, (k, d)
    assert d["head_max_rel_err"] == 0.0


def test_last_bit_differences_are_explained():
    """Weights differing by ~1e-7 relative: float maps agree within tolerance; whatever integer state differs is a near tie
    or downstream of one."""
    a, b = _nets(perturb=1e-7)
    mm = compare_live(OracleAdapter(b), OracleAdapter(a), _events(120), exact=False)
    assert mm.front_unexplained == 0 and mm.flag_unexplained == 0
    assert mm.head_max_rel_err < 1e-4


@pytest.mark.parametrize("fault,msg,msg_rules", [
    ("frontier", "frontier bits differ with no explanation", "conv frontier is not dilate"),
    ("argmax", "not a near tie", "argmax entries are not the reference rule"),
    ("flag", "recompute flags differ with no explanation", "recompute flags are not the reference rule")])
def test_injected_faults_are_rejected(fault, msg, msg_rules):
    """Both layers of the check catch a fault on their own: the comparison with the oracle (explained-only) and the
    bit-exact rules applied to the implementation's own maps."""
    a, b = _nets()
    with pytest.raises(AssertionError, match=msg):
        compare_live(Faulty(b, fault), OracleAdapter(a), _events(12), exact=False, rules=False)
    a, b = _nets()
    with pytest.raises(AssertionError, match=msg_rules):
        compare_live(Faulty(b, fault), OracleAdapter(a), _events(12), exact=False, rules=True)


def test_root_cause_budget():
    mm = Mismatch()
    mm.idx_total, mm.idx_roots = 10 ** 6, 9
    with pytest.raises(AssertionError):
        mm.check()
    mm.idx_roots = 5
    mm.check()
    mm.front_unexplained = 1
    with pytest.raises(AssertionError):
        mm.check()
