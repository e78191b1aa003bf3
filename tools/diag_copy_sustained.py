"""How much of the measured HBM copy peak survives the 1 kW power cap?  Runs torch's device copy back to back
for a few seconds and prints the bandwidth of the first and last iterations with the board power (NVML)."""
import sys
import time

import torch
import pynvml

def main(seconds=4.0, gib=2):
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n = gib * (1 << 30) // 2
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    evs, pw = [], []
    t0 = time.time()
    while time.time() - t0 < seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            b.copy_(a)
        e1.record()
        evs.append((e0, e1))
        if len(evs) % 16 == 0:
            torch.cuda.synchronize()
            pw.append((time.time() - t0, pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                       pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)))
    torch.cuda.synchronize()
    gbs = [8 * 2 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9 for e0, e1 in evs]
    k = max(1, len(gbs) // 10)
    print(f"iterations {len(gbs)} x 8 copies of {gib} GiB; first {k}: {sum(gbs[:k]) / k:.0f} GB/s, last {k}: {sum(gbs[-k:]) / k:.0f} GB/s, best {max(gbs):.0f}")
    for t, w, sm, mem in pw[:: max(1, len(pw) // 12)]:
        print(f"  t={t:5.2f}s  {w:6.0f} W  sm {sm} MHz  mem {mem} MHz")

if __name__ == "__main__":
    main(*(float(x) for x in sys.argv[1:2]))
