"""Mints the committed golden vectors from the UNMODIFIED reference (in-container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

It imports the reference's own layers from /root/reference (oracle/ref_loader.py) together with
its Cython module compiled by oracle/build_ref.sh, chains them exactly like
src/models/event_numpy.py:53-73 (that module itself imports TensorFlow and cannot be imported
here), drives them with the seeded streams / weights of async_ev_cnn_b200.streams and records
inputs and outputs.  The reference publishes no golden vectors of its own; these are
outputs of the reference itself run here, which is what pins the oracle and the CUDA path.

Cases
  proto8x8        the test_correctness.py protocol (8x8, integer kernel, bias 10, leak .1), seeded
  small32_float   32x32, 1->4->8->6 channels, random-float weights, uniform events
  small32_exact   same net with {-1,0,1} weights, alpha=.5, leak=1/64 (all arithmetic exact in f32), edge events
  ragged16        16x16 with 0..7 events per step incl. duplicate pixels, backwards-in-batch timestamps
  efcn_uniform / efcn_edge   full EFCN 160x224, B=200, 24 steps: head outputs, all frontiers, state digests
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference_layers  # noqa: E402
import async_ev_cnn_b200 as P  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SMALL_NET = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
PROTO_NET = "conv1=3,3,1,1 pool1=2,2 conv2=3,3,1,1 pool2=2,2"


def build_reference_chain(h, w, layers, weights, leak, alpha, padding):
    IntegrationLayer, Conv2DLayer, MaxPoolLayer, _, _ = load_reference_layers()
    chain = [IntegrationLayer(leak, h, w)]
    for name, size in P.parse_layers(layers).items():
        if "conv" in name:
            chain.append(Conv2DLayer(chain[-1], weights["w_" + name], weights["b_" + name], 1, alpha, padding))
        else:
            chain.append(MaxPoolLayer(chain[-1], size, size[0]))
    return chain


def mask_of(shape_hw, ev):
    m = np.zeros(shape_hw, dtype=bool)
    m[np.asarray(ev[0]), np.asarray(ev[1])] = True
    return m


def digest(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def record(name, h, w, layers, weights, leak, alpha, event_batches, full_every, weight_tag, keep_full_layers=None):
    """Runs the reference chain over `event_batches` and writes tests/golden/<name>.npz.

    Layout (arrays are stacked over steps to keep the archive small):
      events [total,3] int32 + ev_offsets [steps+1]      inputs
      delta [steps] f64, heads [steps,H,W,C] f32           per-step outputs
      front_<layer> [steps, ceil(H*W/8)] u8                packed output-event masks (row-major bits)
      dg{S,F,A,I}_<layer> [steps,3] f64                    (sum, abs-sum, square-sum) state digests
      flagcnt_<layer> [steps]                              number of sticky recompute flags
      full_steps [n] + {S,F,A,idx,flags}_<layer> [n,...]   full state at the listed steps
      init_F_<layer>, init_idx_<layer>                     state after construction / reset
    """
    chain = build_reference_chain(h, w, layers, weights, leak, alpha, "SAME")
    names = ["intgr"] + list(P.parse_layers(layers).keys())
    wcat = np.concatenate([np.asarray(v, np.float64).ravel() for _, v in sorted(weights.items())])
    rec = {"height": h, "width": w, "layers": layers, "leak": leak, "alpha": alpha, "names": np.array(names),
           "n_steps": len(event_batches), "weight_tag": np.array(weight_tag), "weight_digest": digest(wcat)}
    for nm, layer in zip(names, chain):
        if hasattr(layer, "_conv_actfn"):
            rec["init_F_%s" % nm] = layer._featuremap.copy()
        if hasattr(layer, "_idx_max"):
            rec["init_idx_%s" % nm] = layer._idx_max[0].reshape(layer.out_shape()).astype(np.uint8)
    rec["events"] = np.concatenate([np.asarray(e, np.int32) for e in event_batches])
    rec["ev_offsets"] = np.cumsum([0] + [len(e) for e in event_batches]).astype(np.int64)
    per_step = {}
    full = {}
    full_steps = []

    def put(store, key, val):
        store.setdefault(key, []).append(val)

    for s, ev in enumerate(event_batches):
        e, d = chain[0].compute(np.asarray(ev, np.int32), None)
        fr = [e]
        for layer in chain[1:]:
            e, d = layer.compute(e, d)
            fr.append(e)
        put(per_step, "delta", np.float64(d))
        put(per_step, "heads", chain[-1].featuremap().transpose(1, 2, 0).astype(np.float32))
        cnt = []
        for i, (nm, layer) in enumerate(zip(names, chain)):
            _, hh, ww = layer.out_shape()
            m = mask_of((hh, ww), fr[i])
            cnt.append(int(m.sum()))
            put(per_step, "front_%s" % nm, np.packbits(m))
            if hasattr(layer, "_conv_actfn"):
                put(per_step, "dgF_%s" % nm, digest(layer._featuremap))
                put(per_step, "dgA_%s" % nm, digest(layer._conv_actfn))
            elif hasattr(layer, "_idx_max"):
                put(per_step, "dgI_%s" % nm, digest(layer._idx_max[0]))
                put(per_step, "flagcnt_%s" % nm, int(layer._recompute_coords.sum()))
            else:
                put(per_step, "dgS_%s" % nm, digest(layer.surface()))
        put(per_step, "front_counts", np.array(cnt, np.int32))
        if full_every and (s % full_every == full_every - 1 or s == len(event_batches) - 1):
            full_steps.append(s)
            for nm, layer in zip(names, chain):
                if keep_full_layers is not None and nm not in keep_full_layers:
                    continue
                if hasattr(layer, "_conv_actfn"):
                    put(full, "F_%s" % nm, layer._featuremap.copy())
                    put(full, "A_%s" % nm, layer._conv_actfn.copy())
                elif hasattr(layer, "_idx_max"):
                    put(full, "idx_%s" % nm, layer._idx_max[0].reshape(layer.out_shape()).astype(np.uint8))
                    put(full, "flags_%s" % nm, layer._recompute_coords.copy())
                else:
                    put(full, "S_%s" % nm, layer.surface()[0].copy())
    for store in (per_step, full):
        for key, vals in store.items():
            rec[key] = np.stack(vals)
    rec["full_steps"] = np.array(full_steps, np.int32)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-16s %3d steps  frontier means %s  -> %s (%.0f KB)" % (
        name, len(event_batches), np.round(np.mean(rec["front_counts"], axis=0), 1).tolist(), os.path.basename(path),
        os.path.getsize(path) / 1024))


def proto_events(steps, seed):
    """test_correctness.py:124-126,165-168 with a seed: 1 event at ts 0, then 5 events per step with
    ts = sort(randint(1,10,5)) + prev_ts, uniform y,x on 8x8."""
    rng = np.random.RandomState(seed)
    out = [np.array([[rng.randint(0, 8), rng.randint(0, 8), 0]], np.int32)]
    prev = 0
    for _ in range(steps - 1):
        ts = np.sort(rng.randint(1, 10, size=5)) + prev
        y = rng.randint(0, 8, size=5)
        x = rng.randint(0, 8, size=5)
        out.append(np.stack([y, x, ts], -1).astype(np.int32))
        prev = ts.max()
    return out


def ragged_events(steps, h, w, seed):
    """0..7 events per step (0 -> the step is skipped by the caller), deliberate duplicate pixels,
    timestamps unsorted inside a batch, occasional long gaps that kill the whole surface."""
    rng = np.random.default_rng(seed)
    out, t = [], 0
    for s in range(steps):
        n = int(rng.integers(1, 8))
        t += int(rng.integers(1, 6)) if s % 17 else 400
        ts = t + rng.integers(0, 6, size=n)
        y = rng.integers(0, h, size=n)
        x = rng.integers(0, w, size=n)
        if n >= 3:                       # duplicate pixel inside the batch: last one must win
            y[-1], x[-1] = y[0], x[0]
        out.append(np.stack([y, x, ts], -1).astype(np.int32))
        t = int(ts.max())
    return out


def main():
    # 1. reference test protocol, integer kernel [[-2,-1,1]]*3, bias 10 (test_correctness.py:96-105)
    k = np.array([[-2, -1, 1]] * 3).reshape(3, 3, 1, 1)
    w_proto = {"w_conv1": k, "b_conv1": np.array([10]), "w_conv2": k, "b_conv2": np.array([10])}
    record("proto8x8", 8, 8, PROTO_NET, w_proto, 0.1, 0.1, proto_events(400, 1234), 1, "test_correctness integer kernel")

    # 2/3. small multi-channel nets
    ev = P.synthetic_events("uniform", 1, 80, 20, 32, 32, seed=11, dt_int=(1, 10))[0]
    record("small32_float", 32, 32, SMALL_NET, P.xavier_weights(SMALL_NET, seed=3), 0.01, 0.1, list(ev), 4,
           "xavier_weights(SMALL_NET, seed=3)")
    ev = P.synthetic_events("edge", 1, 80, 20, 32, 32, seed=12, dt_int=(1, 4))[0]
    record("small32_exact", 32, 32, SMALL_NET, P.xavier_weights(SMALL_NET, seed=4, exact=True), 1.0 / 64, 0.5, list(ev), 1,
           "xavier_weights(SMALL_NET, seed=4, exact=True)")

    # 4. ragged / duplicate / dying-surface edge cases
    record("ragged16", 16, 16, SMALL_NET, P.xavier_weights(SMALL_NET, seed=5, exact=True), 1.0 / 64, 0.5,
           ragged_events(120, 16, 16, 21), 1, "xavier_weights(SMALL_NET, seed=5, exact=True)")

    # 5. full EFCN, configs/efcn_event.yml shape
    for kind in ("uniform", "edge"):
        ev = P.synthetic_events(kind, 1, 24, 200, 160, 224, seed=7)[0]
        record("efcn_" + kind, 160, 224, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, list(ev), 24,
               "xavier_weights(EFCN_LAYERS, seed=0)", keep_full_layers={"pool3", "conv4", "pool4", "conv5", "pool5", "conv6", "conv7"})


if __name__ == "__main__":
    main()
