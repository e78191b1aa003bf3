"""Diagnostic (GPU box): EFCN float net, CUDA vs live oracle, per-layer error statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
from oracle.event_oracle import OracleEventNet

kind = sys.argv[1] if len(sys.argv) > 1 else "edge"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
H, W = 160, 224
wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
evs = P.synthetic_events(kind, 1, steps, 200, H, W, seed=7)[0]
net = EventNetCuda(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME", n_streams=1)
ora = OracleEventNet(H, W, P.EFCN_LAYERS, wts, 5e-5, 0.1, "SAME")
for t in range(steps):
    ho = ora.step(evs[t]); hc = net.step(evs[t])[0]
    line = []
    prevF = None
    for i, nm in enumerate(net.names):
        st = net.state(i); lo = ora.layers[i]
        fx = int((net.frontier(i) ^ ora.frontier_mask(i)).sum())
        if "F" in st:
            sF = np.abs(lo.F).max(); sA = max(np.abs(lo.A).max(), 1e-30)
            eF = np.abs(st["F"] - lo.F); eA = np.abs(st["A"] - lo.A)
            msg = "%s F %.1e(%d) A %.1e(%d)" % (nm, eF.max() / sF, int((eF > 1e-4 * sF).sum()), eA.max() / sA, int((eA > 1e-4 * sA).sum()))
            prevF = lo.F
        elif "idx" in st:
            oi = lo.idx.reshape(lo.shape)
            d = np.argwhere(st["idx"] != oi)
            gaps = []
            for c, y, x in d:
                a, b = int(st["idx"][c, y, x]), int(oi[c, y, x])
                gaps.append(abs(float(prevF[c, 2*y + a//2, 2*x + a%2]) - float(prevF[c, 2*y + b//2, 2*x + b%2])) / np.abs(prevF).max())
            msg = "%s idx %d maxgap %.1e flags %d" % (nm, len(d), max(gaps) if gaps else 0, int((st["flags"] != lo.flags).sum()))
        else:
            msg = "S eq %s" % np.array_equal(st["S"], lo.S[0])
        if fx: msg += " FRONT^%d" % fx
        line.append(msg)
    eh = np.abs(hc - ho).max() / np.abs(ho).max()
    print("step %2d head %.1e | " % (t, eh) + " | ".join(line[1:]))
