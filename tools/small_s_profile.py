"""Developer measurement (not part of the product path): where a step's time goes for FEW streams.

    python tools/small_s_profile.py [S ...]          default S = 1 16 64

For each S: EFCN 160x224, B = 200 edge events per stream and step, 160-step pre-roll, then (a) the graph-replayed step
timed between CUDA events with device-resident events, (b) the blocking host call (events in, head out), (c) the per-launch
table (CUDA events after every launch of the un-graphed step, aec_net_profile).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench as BN
    import async_ev_cnn_b200 as P
    from async_ev_cnn_b200.engine import EventNetCuda
    sizes = [int(a) for a in sys.argv[1:]] or [1, 16, 64]
    H, W, B, pre, K = 160, 224, 200, 160, 50
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    for S in sizes:
        ev = BN.gen_events(P, "edge", S, pre + 3 * K + 1, B, H, W, 4242)
        net = EventNetCuda(H, W, P.EFCN_LAYERS, wts, BN.LEAK, BN.ALPHA, "SAME", n_streams=S, device=0, max_events_per_step=2048)
        ev_dev = torch.from_numpy(ev).cuda()
        off = (np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)
        off_dev = torch.from_numpy(off).cuda()
        stream = torch.cuda.current_stream()
        sh = stream.cuda_stream
        t = 0
        for _ in range(pre):
            net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
            t += 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
            t += 1
        e1.record(stream)
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / K
        evh = torch.from_numpy(np.ascontiguousarray(ev[t:t + K])).pin_memory().numpy()
        offh = torch.from_numpy(off).pin_memory().numpy()
        out = torch.empty((S,) + net.head_shape, dtype=torch.float32).pin_memory().numpy()
        w0 = time.perf_counter()
        for i in range(K):
            net.step_packed(evh[i], offh, out=out, cuda_stream=sh)
        host_ms = 1e3 * (time.perf_counter() - w0) / K
        t += K
        net.profile(True)
        for _ in range(K):
            net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
            t += 1
        torch.cuda.synchronize()
        prof, steps = net.read_profile()
        net.profile(False)
        tot = sum(prof.values())
        print("S = %d: graph step %.4f ms (device-resident events), blocking host call %.4f ms, sum of per-launch times %.4f ms over %d steps"
              % (S, dev_ms, host_ms, tot, steps))
        print("   " + "  ".join("%s %.1f" % (k, 1e3 * v) for k, v in prof.items()) + "   (us)")
        net.close()
        del ev_dev


if __name__ == "__main__":
    main()
