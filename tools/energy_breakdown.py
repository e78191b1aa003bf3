"""Developer tool: time, energy (NVML total-energy counter) and mean power of each layer's update kernels,
by re-running one layer's (sweep + frontier + evaluation) many times on the bench workload's state.
The 1 kW power cap decides the sustained step rate, so joules per step matter as much as milliseconds."""
import os, sys, time
import numpy as np, torch, pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
H, W, B, S = 160, 224, 200, 1024
net = EventNetCuda(H, W, P.EFCN_LAYERS, P.xavier_weights(P.EFCN_LAYERS, seed=0), 5e-5, 0.1, "SAME", n_streams=S, max_events_per_step=2048)
n = 50
ev = P.synthetic_events("edge", S, n, B, H, W, seed=100)
for t in range(n - 1):
    net.step([ev[s, t] for s in range(S)]) if False else None
evp = np.ascontiguousarray(ev.transpose(1, 0, 2, 3)).reshape(n, S * B, 3)
evd = torch.from_numpy(evp).cuda()
off = torch.from_numpy((np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)).cuda()
for t in range(n - 1):
    net.step_device(evd[t].data_ptr(), off.data_ptr(), S * B, None)
torch.cuda.synchronize()
def measure(name, fn, reps):
    fn(); torch.cuda.synchronize()
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    t1 = time.perf_counter(); e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    dt = (t1 - t0) / reps; de = (e1 - e0) * 1e-3 / reps
    print("%-10s %8.3f ms  %7.3f J  %6.0f W" % (name, dt * 1e3, de, de / dt))
    return dt, de
time.sleep(1.0)
e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); time.sleep(1.0); e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
print("idle power %.0f W" % ((e1 - e0) * 1e-3))
tot_t = tot_e = 0.0
for li in range(1, len(net.names)):
    reps = int(max(50, min(2000, 1.5 / 0.0005)))
    dt, de = measure(net.names[li], lambda: net.layer_compute(li), 600)
    tot_t += dt; tot_e += de
print("sum of layers %.3f ms %.3f J" % (tot_t * 1e3, tot_e))
measure("full step", lambda: net.step_device(evd[n - 1].data_ptr(), off.data_ptr(), S * B, None), 400)
