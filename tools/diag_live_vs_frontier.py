"""Developer tool: how much of the leak sweep's work lies on sites that the same step re-evaluates anyway?
Per conv layer at steady state (edge streams): live sites (some channel has a non-zero rate before the step),
frontier sites of the step, and their intersection."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda

H, W, B, S, STEPS = 160, 224, 200, 8, 200
kind = sys.argv[1] if len(sys.argv) > 1 else "edge"
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, "SAME", n_streams=S, max_events_per_step=2048)
ev = P.synthetic_events(kind, S, STEPS, B, H, W, seed=100)
for t in range(STEPS - 1):
    net.step([ev[s, t] for s in range(S)])
live = {}
for i, nm in enumerate(net.names):
    if "conv" in nm:
        live[i] = [np.any(net.state(i, s)["A"] != 0, axis=0) for s in range(S)]
net.step([ev[s, STEPS - 1] for s in range(S)])
print("%-6s %8s %8s %8s %8s   (sites per stream, mean of %d streams)" % ("layer", "sites", "live", "frontier", "both", S))
tl = tb = 0
for i, nm in enumerate(net.names):
    if "conv" not in nm:
        continue
    c, h, w = net.shapes()[i]
    l = np.mean([live[i][s].sum() for s in range(S)])
    f = np.mean([net.frontier(i, s).sum() for s in range(S)])
    b = np.mean([(live[i][s] & net.frontier(i, s)).sum() for s in range(S)])
    tl += l * c; tb += b * c
    print("%-6s %8d %8.0f %8.0f %8.0f   live & re-evaluated = %.0f %% of live" % (nm, h * w, l, f, b, 100 * b / max(l, 1)))
print("element-weighted: %.0f %% of the swept conv elements are overwritten by the same step's re-evaluation" % (100 * tb / tl))
net.close()

# ---- second question: how many evaluated pool windows have all four conv sites re-evaluated in the same step?
# (those could be reduced in the conv epilogue instead of going through HBM)
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, "SAME", n_streams=S, max_events_per_step=2048)
for t in range(STEPS):
    net.step([ev[s, t] for s in range(S)])
print("%-6s %10s %10s %10s" % ("pool", "windows", "evaluated", "complete"))
for i, nm in enumerate(net.names):
    if "pool" not in nm:
        continue
    c, h, w = net.shapes()[i]
    ev_w = comp = 0
    for s in range(S):
        fw = net.frontier(i, s)
        fc = net.frontier(i - 1, s)[: 2 * h, : 2 * w].reshape(h, 2, w, 2).all(axis=(1, 3))
        ev_w += fw.sum(); comp += (fw & fc).sum()
    print("%-6s %10d %10.0f %10.0f   %.0f %% of the evaluated windows are complete" % (nm, h * w, ev_w / S, comp / S, 100.0 * comp / max(ev_w, 1)))
net.close()
