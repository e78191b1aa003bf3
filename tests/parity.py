"""Shared parity harness: replays a golden fixture (or a live oracle) against any implementation that
offers the small adapter protocol below.  Used for the CPU oracle (tests/test_oracle_golden.py)
and for the CUDA path (tests/test_cuda_parity.py) so both are held to the same checks.

Adapter protocol
    names                 list of layer names, "intgr" first
    step(events)          events int32 [B,3] (y,x,ts) -> head float32 [H,W,C]
    frontier(i)           bool [H_i,W_i] mask of layer i's output events of the last step
    state(i)              dict: {"S": f64 [H,W]} | {"F": f32 [C,H,W], "A": f32 [C,H,W]}
                                | {"idx": int [C,Ho,Wo], "flags": bool [Ho,Wo]}
    delta()               float64 leak amount of the last step

Bars (north_star): frontier sets, argmax indices and recompute flags bit-exact; float maps within
FLOAT_RTOL = 1e-4 relative to the map's scale (exact nets: bit-equal).
"""
import os

import numpy as np

FLOAT_RTOL = 1e-4


class Golden:
    def __init__(self, path):
        self.z = np.load(path, allow_pickle=False)
        self.names = [str(n) for n in self.z["names"]]
        self.height = int(self.z["height"])
        self.width = int(self.z["width"])
        self.layers = str(self.z["layers"])
        self.leak = float(self.z["leak"])
        self.alpha = float(self.z["alpha"])
        self.n_steps = int(self.z["n_steps"])
        self.offsets = self.z["ev_offsets"]
        self.full_steps = [int(s) for s in self.z["full_steps"]]

    def events(self, s):
        return self.z["events"][self.offsets[s]:self.offsets[s + 1]]

    def shapes(self):
        """[C,H,W] of every layer, derived like conv2d.py:38-58 / maxpool.py:25-30 (SAME, stride 1)."""
        from async_ev_cnn_b200.streams import parse_layers
        out = [[1, self.height, self.width]]
        for name, size in parse_layers(self.layers).items():
            c, h, w = out[-1]
            out.append([size[3], h, w] if "conv" in name else [c, (h - size[0]) // size[0] + 1, (w - size[1]) // size[0] + 1])
        return out

    def front(self, i, s):
        _, h, w = self.shapes()[i]
        return np.unpackbits(self.z["front_%s" % self.names[i]][s])[:h * w].reshape(h, w).astype(bool)

    def weights(self):
        from async_ev_cnn_b200.streams import xavier_weights
        tag = str(self.z["weight_tag"])
        if tag.startswith("test_correctness"):
            k = np.array([[-2, -1, 1]] * 3).reshape(3, 3, 1, 1)
            w = {"w_conv1": k, "b_conv1": np.array([10]), "w_conv2": k, "b_conv2": np.array([10])}
        else:
            from async_ev_cnn_b200.streams import EFCN_LAYERS  # noqa: F401 (used by eval)
            SMALL_NET = self.layers  # noqa: F841 (used by eval)
            w = eval(tag, {"xavier_weights": xavier_weights, "SMALL_NET": self.layers, "EFCN_LAYERS": EFCN_LAYERS})
        cat = np.concatenate([np.asarray(v, np.float64).ravel() for _, v in sorted(w.items())])
        dg = np.array([cat.sum(), np.abs(cat).sum(), (cat * cat).sum()])
        assert np.array_equal(dg, self.z["weight_digest"]), "weight generator drifted from the golden fixture"
        return w


def digest(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def assert_close_map(got, want, exact, what, bad_sites=False):
    """exact: bit-equal.  Otherwise max |err| <= FLOAT_RTOL * max|want|.  With bad_sites=True (float nets,
    [C,H,W] state maps) a bounded number of SITES may exceed the tolerance: a pre-activation within
    rounding of 0 (slope 1 vs alpha) or a pool tie decided by GEMM rounding legitimately changes the
    rate map of the sites it feeds until they are next re-evaluated (parity policy)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    if exact:
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)
            raise AssertionError("%s: %d elements differ (exact net), first %s got %r want %r" % (
                what, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])]))
        return
    scale = max(float(np.abs(want).max()), 1e-30)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    if bad_sites and got.ndim == 3:
        nbad = int((err > FLOAT_RTOL * scale).any(axis=0).sum())
        allowed = max(2, int(np.ceil(2e-3 * got.shape[1] * got.shape[2])))
        assert nbad <= allowed, "%s: %d sites exceed %.0e * scale (allowed %d), max err %.3e scale %.3e" % (
            what, nbad, FLOAT_RTOL, allowed, float(err.max()), scale)
        return
    assert float(err.max()) <= FLOAT_RTOL * scale, "%s: max abs err %.3e > %.0e * scale %.3e" % (what, float(err.max()), FLOAT_RTOL, scale)


def assert_close_digest(got, want, exact, what, rtol=FLOAT_RTOL):
    if exact:
        assert np.array_equal(got, want), "%s: digest %s vs %s" % (what, got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * max(1.0, float(np.abs(want).max())), err_msg=what)


NEAR_TIE = 1e-5          # a decision counts as a near tie when the oracle's own margin is <= NEAR_TIE * map scale


class Mismatch:
    """What a random-float net may disagree on with the oracle, and only when the ORACLE'S OWN VALUES explain it
    (GEMM rounding decides an exact or near tie differently from the oracle's BLAS, whose summation order is
    unspecified too: SURVEY 'frontier bit-exactness vs non-reproducible GEMM').  Root causes are counted
    (`idx_roots`, `slope_flips`, `flag_roots`); every other disagreement must lie in the reach of one
    (`*_explained`), and anything else is `*_unexplained` and fails the test.  Exact nets never touch this:
    they are held to bit-equality."""

    FIELDS = ("steps", "front_total", "front_explained", "front_unexplained", "idx_total", "idx_roots", "idx_downstream",
              "flag_total", "flag_roots", "flag_explained", "flag_unexplained", "slope_total", "slope_flips", "bad_sites",
              "head_max_rel_err")

    def __init__(self):
        self.steps = 0
        self.front_total = 0          # frontier sites of the oracle, summed over layers and steps
        self.front_explained = 0      # differing frontier bits inside the reach of a root cause
        self.front_unexplained = 0    # differing frontier bits with no explanation (must stay 0)
        self.idx_total = 0            # argmax decisions compared
        self.idx_roots = 0            # argmax differs between two candidates the oracle holds within NEAR_TIE (root cause)
        self.idx_downstream = 0       # argmax differs in a window whose inputs already differ
        self.flag_total = 0
        self.flag_roots = 0           # recompute flag differs where the oracle's rate at the argmax is within NEAR_TIE of the minimum
        self.flag_explained = 0       # ... or the window's inputs / argmax / earlier sticky flag already differ
        self.flag_unexplained = 0     # (must stay 0)
        self.slope_total = 0
        self.slope_flips = 0          # activation slope differs where |F| <= NEAR_TIE * scale (root cause)
        self.bad_sites = 0            # conv sites beyond FLOAT_RTOL, all inside the reach of a root cause
        self.head_max_rel_err = 0.0   # over the steps whose head is not downstream of a root cause

    def as_dict(self):
        return {k: (float(getattr(self, k)) if k == "head_max_rel_err" else int(getattr(self, k))) for k in self.FIELDS}

    def check(self, root_rate=5e-6):
        """Nothing unexplained; root causes rare (a broken tie rule would show up as thousands of them)."""
        assert self.front_unexplained == 0, "%d frontier bits differ without an explanation" % self.front_unexplained
        assert self.flag_unexplained == 0, "%d recompute flags differ without an explanation" % self.flag_unexplained
        assert self.idx_roots <= max(8, root_rate * self.idx_total), "near-tie argmax flips %d of %d" % (self.idx_roots, self.idx_total)
        assert self.slope_flips <= max(8, root_rate * self.slope_total), "near-zero slope flips %d of %d" % (self.slope_flips, self.slope_total)
        assert self.flag_roots <= max(8, root_rate * self.flag_total * 16), "near-equal-rate flag flips %d of %d" % (self.flag_roots, self.flag_total)

    def __repr__(self):
        return "Mismatch(%s)" % ", ".join("%s=%s" % (k, ("%.2e" % getattr(self, k)) if k == "head_max_rel_err" else getattr(self, k)) for k in self.FIELDS)


def record_parity(case, mm, extra=None):
    """Prints the counters and, when AEC_PARITY_LOG names a file, appends them as one JSON line (the GPU run's log is
    committed as profiles/parity_r2.json)."""
    import json
    row = {"case": case}
    row.update(mm.as_dict())
    if extra:
        row.update(extra)
    print("\n[parity] %s" % json.dumps(row))
    path = os.environ.get("AEC_PARITY_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(row) + "\n")
    return row


def replay_golden(adapter, g, exact, steps=None, check_init=True):
    """Feeds the fixture's events to `adapter` step by step and checks everything the fixture holds, bit for bit
    (exact=True: the oracle port on any net, the CUDA path on exactly representable nets).  A float net on the CUDA
    path goes through compare_live(..., golden=g) instead: there every integer disagreement has to be explained by
    the oracle's own values, which a fixture of digests cannot do."""
    assert exact, "float nets are compared with compare_live(golden=...): every integer mismatch must be explained"
    names = g.names
    mm = Mismatch()
    assert list(adapter.names) == names
    if check_init:
        for i, nm in enumerate(names):
            st = adapter.state(i)
            if "F" in st:
                assert_close_map(st["F"], g.z["init_F_%s" % nm], exact, "init F %s" % nm)
                assert not np.any(st["A"]), "init A %s must be zero (conv2d.py:62)" % nm
            if "idx" in st:
                assert np.array_equal(st["idx"], g.z["init_idx_%s" % nm]), "init idx %s" % nm
                assert not np.any(st["flags"])
    n = g.n_steps if steps is None else min(steps, g.n_steps)
    for s in range(n):
        head = adapter.step(g.events(s))
        assert adapter.delta() == g.z["delta"][s], "step %d: delta_leak %r vs %r" % (s, adapter.delta(), g.z["delta"][s])
        for i, nm in enumerate(names):
            got, want = adapter.frontier(i), g.front(i, s)
            mm.front_total += int(want.sum())
            if not np.array_equal(got, want):
                raise AssertionError("step %d layer %s: frontier differs: %d extra, %d missing (want %d sites)" % (
                    s, nm, int((got & ~want).sum()), int((~got & want).sum()), int(want.sum())))
        assert_close_map(head, g.z["heads"][s], exact, "step %d head" % s)
        full = s in g.full_steps
        fi = g.full_steps.index(s) if full else -1
        for i, nm in enumerate(names):
            st = adapter.state(i)
            if "S" in st:
                assert_close_digest(digest(st["S"]), g.z["dgS_%s" % nm][s], True, "step %d S digest" % s)
                if full and "S_%s" % nm in g.z:
                    assert np.array_equal(st["S"], g.z["S_%s" % nm][fi]), "step %d surface" % s
            elif "F" in st:
                assert_close_digest(digest(st["F"]), g.z["dgF_%s" % nm][s], exact, "step %d F digest %s" % (s, nm))
                # rate maps are discontinuous at sign decisions / pool ties: looser digest bound for float nets
                assert_close_digest(digest(st["A"]), g.z["dgA_%s" % nm][s], exact, "step %d A digest %s" % (s, nm), rtol=3e-2)
                if full and "F_%s" % nm in g.z:
                    assert_close_map(st["F"], g.z["F_%s" % nm][fi], exact, "step %d F %s" % (s, nm), bad_sites=True)
                    assert_close_map(st["A"], g.z["A_%s" % nm][fi], exact, "step %d A %s" % (s, nm), bad_sites=True)
            else:
                nflag, wflag = int(st["flags"].sum()), int(g.z["flagcnt_%s" % nm][s])
                dg, wdg = digest(st["idx"]), g.z["dgI_%s" % nm][s]
                mm.flag_total += st["flags"].size
                mm.idx_total += st["idx"].size
                assert nflag == wflag, "step %d flag count %s" % (s, nm)
                assert np.array_equal(dg, wdg), "step %d argmax digest %s" % (s, nm)
                if full and "idx_%s" % nm in g.z:
                    wi, wf = g.z["idx_%s" % nm][fi], g.z["flags_%s" % nm][fi].astype(bool)
                    assert np.array_equal(st["idx"], wi), "step %d argmax %s" % (s, nm)
                    assert np.array_equal(st["flags"], wf), "step %d flags %s" % (s, nm)
    mm.steps = n
    return mm


class OracleAdapter:
    """Adapter over oracle.event_oracle.OracleEventNet."""

    def __init__(self, net):
        self.net = net
        self.names = net.names
        self.alpha = next((l.alpha for l in net.layers if hasattr(l, "alpha")), 0.1)

    def step(self, events):
        return self.net.step(events)

    def delta(self):
        return np.float64(self.net.delta)

    def frontier(self, i):
        return self.net.frontier_mask(i)

    def state(self, i):
        layer = self.net.layers[i]
        if hasattr(layer, "S"):
            return {"S": layer.S[0]}
        if hasattr(layer, "A"):
            return {"F": layer.F, "A": layer.A}
        return {"idx": layer.idx.reshape(layer.shape), "flags": layer.flags}


def _dilate(mask, r):
    """Binary dilation of a [H,W] mask by a (2r+1)^2 square."""
    out = np.zeros_like(mask)
    h, w = mask.shape
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            ys, xs = slice(max(0, dy), min(h, h + dy)), slice(max(0, dx), min(w, w + dx))
            yd, xd = slice(max(0, -dy), min(h, h - dy)), slice(max(0, -dx), min(w, w - dx))
            out[ys, xs] |= mask[yd, xd]
    return out


class IntegerRules:
    """Bit-exact check of an implementation's INTEGER decisions against the reference's rules applied to the
    implementation's OWN float maps.  On a float net the conv maps differ from the oracle's in the last bits, so
    frontier sets / argmax / flags can only be compared with the oracle up to near ties (compare_live); but given
    its own maps every decision is an exact function of them:
      conv frontier = {sites whose receptive field holds an input event} U {other sites where sign(F >= 0) of some
                      channel changed in the leak}                          (conv2d.py:113-131, cutils.pyx:73-112)
      pool          : hit = windows of the conv's output events; flags[hit] = False; W = hit U flags;
                      evaluated windows: argmax by (F, then smaller rate R = A*slope, then smaller row), unstable when
                      R[argmax] != min R, flags |= any-channel unstable; others keep idx and flag
                                                                            (maxpool.py:116-154, cutils.pyx:161-177)
    Any deviation is a bug, whatever the float noise.  Built from the oracle's layer objects (kernel sizes, pads)."""

    def __init__(self, oracle_net):
        self.spec = []
        for l in oracle_net.layers:
            if hasattr(l, "K"):
                self.spec.append(("conv", l.K.shape[2], l.K.shape[3], l.pad[0], l.pad[2], np.float32(l.alpha)))
            elif hasattr(l, "kh"):
                assert l.kh == l.kw == l.stride
                self.spec.append(("pool", l.kh))
            else:
                self.spec.append(("intgr",))
        self.prev = None

    def prime(self, impl):
        """State before the first compared step."""
        self.prev = [{k: np.array(v, copy=True) for k, v in impl.state(i).items()} for i in range(len(self.spec))]

    @staticmethod
    def min_argmax(Fw, Rw):
        """cutils.pyx:161-177 on [..., rows] arrays: rows scanned ascending; strict > updates; on equality switch only
        when the rate is smaller.  Returns (argmax rows, unstable)."""
        bf, br = Fw[..., 0].copy(), Rw[..., 0].copy()
        row = np.zeros(Fw.shape[:-1], np.int64)
        for r in range(1, Fw.shape[-1]):
            f, q = Fw[..., r], Rw[..., r]
            take = (f > bf) | ((f == bf) & (q < br))
            bf, br = np.where(take, f, bf), np.where(take, q, br)
            row[take] = r
        return row, br != Rw.min(axis=-1)

    def check(self, impl, step, states=None):
        assert self.prev is not None, "IntegerRules.prime() was not called"
        cur = states if states is not None else [impl.state(i) for i in range(len(self.spec))]
        fronts = [impl.frontier(i) for i in range(len(self.spec))]
        for i, sp in enumerate(self.spec):
            if sp[0] == "conv":
                _, kh, kw, pt, pl, alpha = sp
                pf = fronts[i - 1]
                ho, wo = fronts[i].shape
                N = np.zeros((ho, wo), bool)
                for dy in range(kh):
                    for dx in range(kw):
                        # output o sees input p = o - pad + d
                        y0, x0 = dy - pt, dx - pl
                        oy0, oy1 = max(0, -y0), min(ho, pf.shape[0] - y0)
                        ox0, ox1 = max(0, -x0), min(wo, pf.shape[1] - x0)
                        if oy1 > oy0 and ox1 > ox0:
                            N[oy0:oy1, ox0:ox1] |= pf[oy0 + y0:oy1 + y0, ox0 + x0:ox1 + x0]
                flips = ((self.prev[i]["F"] >= 0) != (cur[i]["F"] >= 0)).any(axis=0) & ~N
                want = N | flips
                if not np.array_equal(fronts[i], want):
                    raise AssertionError("step %d layer %d: conv frontier is not dilate(input events) U sign flips of its own map: %d extra, %d missing" % (
                        step, i, int((fronts[i] & ~want).sum()), int((~fronts[i] & want).sum())))
                # sites outside N are not re-evaluated: their rate map must be untouched
                keep = ~N
                assert np.array_equal(cur[i]["A"][:, keep], self.prev[i]["A"][:, keep]), "step %d layer %d: rate map changed at a site that was not re-evaluated" % (step, i)
            elif sp[0] == "pool":
                k = sp[1]
                cf = fronts[i - 1]
                ho, wo = fronts[i].shape
                hit = cf[:ho * k, :wo * k].reshape(ho, k, wo, k).any(axis=(1, 3))
                Wset = hit | self.prev[i]["flags"]
                if not np.array_equal(fronts[i], Wset):
                    raise AssertionError("step %d layer %d: evaluated windows are not hit U sticky flags: %d extra, %d missing" % (
                        step, i, int((fronts[i] & ~Wset).sum()), int((~fronts[i] & Wset).sum())))
                F, A = cur[i - 1]["F"], cur[i - 1]["A"]
                alpha = self.spec[i - 1][5]
                R = (A * np.where(F > 0, np.float32(1), alpha)).astype(np.float32)
                row, unstable = self.min_argmax(_pool_windows(F, k), _pool_windows(R, k))
                want_idx = np.where(Wset[None], row, self.prev[i]["idx"])
                want_flags = (self.prev[i]["flags"] & ~hit) | (Wset & unstable.any(axis=0))
                nbad = int((cur[i]["idx"] != want_idx).sum())
                assert nbad == 0, "step %d layer %d: %d argmax entries are not the reference rule applied to the layer's own input maps" % (step, i, nbad)
                nbad = int((cur[i]["flags"] != want_flags).sum())
                assert nbad == 0, "step %d layer %d: %d recompute flags are not the reference rule applied to the layer's own input maps" % (step, i, nbad)
        self.prev = [{k: np.array(v, copy=True) for k, v in st.items()} for st in cur]     # an adapter may hand out live arrays


def _pool_windows(a, k):
    """[C,H,W] -> [C,Ho,Wo,k*k] with the window rows in the reference's order (ky*k + kx)."""
    c, h, w = a.shape
    ho, wo = h // k, w // k
    return a[:, :ho * k, :wo * k].reshape(c, ho, k, wo, k).transpose(0, 1, 3, 2, 4).reshape(c, ho, wo, k * k)


def compare_live(impl, oracle, event_batches, exact, reach=1, golden=None, alpha=None, rules=True, carry=None):
    """Steps `impl` and the live `oracle` adapter together over the same batches.

    exact=True: everything bit-equal.  exact=False: float maps within FLOAT_RTOL and integer state (frontier sets,
    pool argmax, recompute flags) identical, except where a difference is EXPLAINED by the oracle's own values:
      ROOT CAUSES (counted, must stay rare - Mismatch.check):
      * a pool argmax differs between candidates whose pre-activations the oracle holds within NEAR_TIE of the
        map scale (a tie decided by GEMM rounding; the oracle's BLAS order is unspecified too);
      * an activation slope differs (F > 0 on one side only) where |F| <= NEAR_TIE * scale;
      * a recompute flag differs where the oracle's rate at the argmax is within NEAR_TIE * scale of the window's
        smallest rate (cutils.pyx:177 compares them by value);
      CONSEQUENCES (counted, each must lie in the reach of a root cause):
      * a conv site beyond the tolerance inside the receptive-field reach (`reach` = (k-1)/2 of the largest
        float-net kernel) of a site whose visible output already differs, or that already differed after the
        previous step (it stays different until it is next re-evaluated);
      * a frontier bit of a conv layer inside the dilated frontier difference of its input, or at a site that
        differs / holds a value within NEAR_TIE of zero (the leak's sign test, conv2d.py:113,126-128);
      * a frontier bit / flag of a pool layer in a window whose inputs, input frontier, argmax or earlier sticky
        flag differ.
    On top of that (`rules`, float nets): the implementation's integer decisions must be bit-exactly the reference's
    rules applied to its OWN float maps at every step (IntegerRules) - that part of the parity bar has no tolerance.
    Anything else fails.  `golden`: a Golden fixture recorded on the same batches - the live oracle is then held
    to it (heads to 1e-6, frontiers exactly), which keeps the committed reference vectors in the loop.
    `carry`: the `.carry` of the Mismatch returned by an earlier call on the same adapters - the comparison goes on
    where that one stopped (separate counters for, say, the settling phase and the steady state of one run).
    Returns the Mismatch."""
    mm = Mismatch()
    if alpha is None:
        alpha = getattr(oracle, "alpha", 0.1)
    if carry is not None:
        still_bad, flag_bad, last_F, checker, s0 = carry
    else:
        still_bad = {}            # conv layer -> sites that differed after the previous step
        flag_bad = {}             # pool layer -> windows whose sticky flag differed after the previous step
        last_F = {}               # conv layer -> the oracle's F after the previous step (the leak's sign test compares with it)
        checker, s0 = None, 0
        if rules and not exact and hasattr(oracle, "net"):
            checker = IntegerRules(oracle.net)
            checker.prime(impl)
    for s, ev in enumerate(event_batches, start=s0):
        h0 = oracle.step(ev)
        h1 = impl.step(ev)
        mm.steps += 1
        impl_states = [impl.state(i) for i in range(len(oracle.names))]
        if checker is not None:
            checker.check(impl, s, impl_states)
        assert impl.delta() == oracle.delta(), "step %d delta" % s
        if golden is not None:
            assert_close_map(h0, golden.z["heads"][s], False, "step %d: live oracle head vs golden fixture" % s)
        prev_F = prev_A = None
        out_bad = None            # [H,W] sites of the previous layer whose visible output (V or R) differs
        front_bad = None          # [H,W] frontier bits of the previous layer that differ
        for i, nm in enumerate(oracle.names):
            got, want = impl.frontier(i), oracle.frontier(i)
            if golden is not None:
                assert np.array_equal(want, golden.front(i, s)), "step %d %s: live oracle frontier vs golden fixture" % (s, nm)
            mm.front_total += int(want.sum())
            fd = got ^ want
            if exact and fd.any():
                raise AssertionError("step %d layer %s frontier: %d extra %d missing" % (s, nm, int((got & ~want).sum()), int((~got & want).sum())))
            so, si = oracle.state(i), impl_states[i]
            if "S" in so:
                assert np.array_equal(si["S"], so["S"]), "step %d surface" % s
                assert not fd.any(), "step %d: the surface layer's output events differ (integer / float64 work: must be exact)" % s
                out_bad = np.zeros(so["S"].shape[-2:], bool)
                front_bad = fd
            elif "F" in so:
                if exact:
                    assert_close_map(si["F"], so["F"], True, "step %d %s F" % (s, nm))
                    assert_close_map(si["A"], so["A"], True, "step %d %s A" % (s, nm))
                else:
                    sF = max(float(np.abs(so["F"]).max()), 1e-30)
                    sA = max(float(np.abs(so["A"]).max()), 1e-30)
                    bad = (np.abs(si["F"].astype(np.float64) - so["F"]) > FLOAT_RTOL * sF).any(axis=0) | \
                          (np.abs(si["A"].astype(np.float64) - so["A"]) > FLOAT_RTOL * sA).any(axis=0)
                    allowed = _dilate(out_bad, reach)
                    if i in still_bad:
                        allowed |= still_bad[i]
                    assert not (bad & ~allowed).any(), "step %d %s: %d sites differ beyond %.0e * scale outside the reach of any explained upstream difference" % (
                        s, nm, int((bad & ~allowed).sum()), FLOAT_RTOL)
                    flip = (si["F"] > 0) != (so["F"] > 0)
                    unexplained = flip & (np.abs(so["F"]) > NEAR_TIE * sF) & ~bad[None]
                    assert not unexplained.any(), "step %d %s: %d activation-slope flips away from zero" % (s, nm, int(unexplained.sum()))
                    mm.slope_flips += int((flip & ~bad[None]).sum())
                    mm.slope_total += flip.size
                    mm.bad_sites += int(bad.sum())
                    if fd.any():
                        near0 = (np.abs(so["F"]) <= NEAR_TIE * sF).any(axis=0)
                        if i in last_F:
                            near0 |= (np.abs(last_F[i]) <= NEAR_TIE * sF).any(axis=0)
                        ok = _dilate(front_bad, reach) | near0 | bad | still_bad.get(i, False)
                        nun = int((fd & ~ok).sum())
                        mm.front_unexplained += nun
                        mm.front_explained += int(fd.sum()) - nun
                        assert nun == 0, "step %d %s: %d frontier bits differ with no explanation (%d extra, %d missing)" % (
                            s, nm, nun, int((got & ~want & ~ok).sum()), int((~got & want & ~ok).sum()))
                    still_bad[i] = bad
                    last_F[i] = so["F"].copy()
                    out_bad = bad | flip.any(axis=0)
                front_bad = fd
                prev_F, prev_A = so["F"], so["A"]
            else:
                mm.idx_total += so["idx"].size
                mm.flag_total += so["flags"].size
                if exact:
                    assert np.array_equal(si["idx"], so["idx"]), "step %d %s argmax" % (s, nm)
                    assert np.array_equal(si["flags"], so["flags"]), "step %d %s flags" % (s, nm)
                else:
                    k = int(round((prev_F.shape[1] / so["idx"].shape[1])))
                    ho, wo = so["idx"].shape[1:]
                    pooled_bad = out_bad[:ho * k, :wo * k].reshape(ho, k, wo, k).any(axis=(1, 3))
                    pooled_front = front_bad[:ho * k, :wo * k].reshape(ho, k, wo, k).any(axis=(1, 3))
                    idx_diff = si["idx"] != so["idx"]
                    scale = max(float(np.abs(prev_F).max()), 1e-30)
                    for c, y, x in np.argwhere(idx_diff):
                        if pooled_bad[y, x]:
                            mm.idx_downstream += 1            # inputs of this window already differ: not a root cause
                            continue
                        mm.idx_roots += 1
                        a, b = int(si["idx"][c, y, x]), int(so["idx"][c, y, x])
                        fa = prev_F[c, y * k + a // k, x * k + a % k]
                        fb = prev_F[c, y * k + b // k, x * k + b % k]
                        assert abs(float(fa) - float(fb)) <= NEAR_TIE * scale, (
                            "step %d %s argmax flip at %s is not a near tie: %r vs %r" % (s, nm, (c, y, x), fa, fb))
                    idx_win = idx_diff.any(axis=0)
                    fdiff = si["flags"] != so["flags"]
                    was_bad = flag_bad.get(i, np.zeros_like(fdiff))
                    if fdiff.any():
                        slope = np.where(prev_F > 0, np.float32(1), np.float32(alpha))
                        Rw = _pool_windows((prev_A * slope).astype(np.float32), k)           # [C,Ho,Wo,k*k]
                        r_arg = np.take_along_axis(Rw, so["idx"][..., None].astype(np.int64), axis=3)[..., 0]
                        sR = max(float(np.abs(Rw).max()), 1e-30)
                        gap = np.abs(r_arg - Rw.min(axis=3)).max(axis=0)                     # [Ho,Wo] largest over the channels
                        # oracle says unstable, implementation stable: every unstable channel must be a near tie of rates;
                        # oracle says stable (all gaps exactly 0), implementation unstable: rounding can only separate equal
                        # rates that are not exact zeros (a rate is exactly 0 on both sides when its whole patch is)
                        near_r = np.where(so["flags"], gap <= NEAR_TIE * sR, (Rw != 0).any(axis=(0, 3)))
                        downstream = pooled_bad | idx_win | was_bad
                        nun = int((fdiff & ~downstream & ~near_r).sum())
                        mm.flag_roots += int((fdiff & ~downstream & near_r).sum())
                        mm.flag_explained += int((fdiff & downstream).sum())
                        mm.flag_unexplained += nun
                        assert nun == 0, "step %d %s: %d recompute flags differ with no explanation" % (s, nm, nun)
                    if fd.any():
                        ok = pooled_front | was_bad | pooled_bad
                        nun = int((fd & ~ok).sum())
                        mm.front_unexplained += nun
                        mm.front_explained += int(fd.sum()) - nun
                        assert nun == 0, "step %d %s: %d evaluated-window bits differ with no explanation" % (s, nm, nun)
                    flag_bad[i] = fdiff
                    out_bad = pooled_bad | idx_win
                front_bad = fd
        if exact or not out_bad.any():
            assert_close_map(h1, h0, exact, "step %d head" % s)
            scale = max(float(np.abs(h0).max()), 1e-30)
            mm.head_max_rel_err = max(mm.head_max_rel_err, float(np.abs(np.asarray(h1, np.float64) - h0).max()) / scale)
    mm.carry = (still_bad, flag_bad, last_F, checker, s0 + mm.steps)
    if not exact:
        mm.check()
    return mm


def golden_path(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
