"""Developer tool: is the slow 'sustained' rate power or workload?  Runs a long NON-recycled edge sequence
(timestamps keep increasing) back to back, prints ms/step per block of 32 with the live-site fraction, then idles
2 s and times one more block from the same state (a cool GPU on the same workload)."""
import os, sys, time
import numpy as np, torch, pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda

H, W, B, S = 160, 224, 200, 1024
NSTEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 48 + 640
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, "SAME", n_streams=S, max_events_per_step=2048)
ev = P.synthetic_events("edge", S, NSTEPS, B, H, W, seed=100)
ev = np.ascontiguousarray(ev.transpose(1, 0, 2, 3)).reshape(NSTEPS, S * B, 3)
off = (np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)
evd = torch.from_numpy(ev).cuda(); offd = torch.from_numpy(off).cuda()

def block(t0, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(t0, t0 + n):
        net.step_device(evd[t].data_ptr(), offd.data_ptr(), S * B, None)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def live():
    st = net.sweep_stats()
    return st["live_conv_elems"] / max(1, st["conv_elems"])

for t in range(0, 48, 16):
    block(t, 16)
print("after pre-roll: live fraction %.3f" % live())
t, w0 = 48, time.perf_counter()
while t + 32 <= NSTEPS - 32:
    ms = block(t, 32); t += 32
    print("t=%5.2fs step %4d  %.3f ms/step  live %.3f  %4.0f W  sm %d MHz" % (time.perf_counter() - w0, t, ms, live(),
          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
time.sleep(2.0)
print("after 2 s idle: %.3f ms/step (same state, cool GPU)  %4.0f W" % (block(t, 32), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
