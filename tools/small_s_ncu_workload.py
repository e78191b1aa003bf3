import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench as BN
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
S, H, W, B = 1, 160, 224, 200
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), BN.LEAK, BN.ALPHA, "SAME", n_streams=S, device=0, max_events_per_step=2048)
ev = BN.gen_events(P, "edge", S, 164, B, H, W, 4242)
evd = torch.from_numpy(ev).cuda()
off = torch.from_numpy((np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)).cuda()
for t in range(164):
    net.step_device(evd[t].data_ptr(), off.data_ptr(), S * B, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
net.close()
