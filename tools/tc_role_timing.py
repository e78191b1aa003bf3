"""Developer tool: per-role cycle accounting of the tensor-core conv kernel on the bench workload
(aec_net_tc_timing).  Prints, per conv layer, what fraction of its life each warp role spent waiting."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--kind", default="edge")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--preroll", type=int, default=48)
a = ap.parse_args()
H, W, B = 160, 224, 200
S = a.streams
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, "SAME", n_streams=S, max_events_per_step=2048)
n = a.preroll + a.steps
ev = P.synthetic_events(a.kind, S, n, B, H, W, seed=100)
ev = np.ascontiguousarray(ev.transpose(1, 0, 2, 3)).reshape(n, S * B, 3)
evd = torch.from_numpy(ev).cuda()
off = torch.from_numpy((np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)).cuda()
for t in range(a.preroll):
    net.step_device(evd[t].data_ptr(), off.data_ptr(), S * B, None)
torch.cuda.synchronize()
net.tc_timing(True)
for t in range(a.preroll, n):
    net.step_device(evd[t].data_ptr(), off.data_ptr(), S * B, None)
torch.cuda.synchronize()
for nm, d in net.read_tc_timing().items():
    c = max(1, d["ctas"])
    f = lambda k, tot: 100.0 * d[k] / max(1, d[tot])
    print("%-6s ctas %4d units/cta %.1f | mma %7.0f kcyc/cta: gate-blocked %4.1f%% (gate waits: acc %4.1f%% sites %4.1f%% weights %4.1f%%) | prod %7.0f: wait info %4.1f%% stage %4.1f%% | "
          "epi %7.0f: wait acc %4.1f%% info %4.1f%% | load wait %4.1f%% | mma section (slot 15) %4.1f%%" % (
              nm, d["ctas"], d["units"] / c, d["mma_total"] / c / 1e3, f("mma_wait_sites", "mma_total"), f("mma_wait_acc", "mma_total"), f("gate_wait_sites", "mma_total"),
              f("mma_wait_weights", "mma_total"), d["prod_total"] / c / 1e3, f("prod_wait_siteinfo", "prod_total"), f("prod_wait_stage", "prod_total"),
              d["epi_total"] / c / 1e3, f("epi_wait_acc", "epi_total"), f("epi_wait_siteinfo", "epi_total"), f("load_wait", "load_total"),
              f("mma_section", "mma_total")))
net.close()
