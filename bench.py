#!/usr/bin/env python
"""bench.py - EFCN event-mode throughput (events/s) on B200, next to the host-core CPU baseline.

    python bench.py --gpus 1 --steps K --warmup W                 (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus 1 --steps K --warmup W

Workload (BASELINE.json configs[1]): EFCN of configs/efcn_event.yml
(5 x [3x3 conv + 2x2 pool] + two 1x1 convs -> 5x7x110) on 160x224 frames (the 240x180-sensor stream
centre-cropped as the config does), leak 5e-5/us, alpha 0.1, B = 200 events per stream per step,
random-init weights (xavier, seed 0), seeded synthetic streams.  A "step" advances EVERY stream by
one batch of B events; streams are independent, `--streams` of them per GPU (weak scaling: per-GPU
work fixed, no data-path collective).  Before timing, every stream is pre-rolled to the steady
state of the workload: the edge of the synthetic stream needs 112 steps to cross the frame and a pixel
that collected a few events stays alive for 40-100 steps (1/leak = 20 ms = 40 steps per unit of
surface value), so the live-site fraction and the frontier sizes settle only after ~120 steps
(tools/diag_sustained2.py); the default pre-roll is 160 steps, on the GPU and in the CPU arms alike.

What the one JSON line holds beyond the contract keys:
  parity_check   the CPU baseline's streams ARE GPU streams 0..P-1 (same events); after the pre-roll the oracle's
                 head / frontier sets / pool argmax / flags of those streams are compared with the GPU's
  roofline       dominant kernel (gathered GEMM on the tensor cores) + a per-layer table of every launch of the step
  roofline_hbm   the leak sweep and the whole step in algorithmic bytes
  legs           BASELINE configs 4 and 5 measured in the same run (K = 5 each): uniform streams, 256 and 4096
                 streams per GPU, and the DAVIS346-sized stress net (256x320, deeper EFCN, 1000 / 2000 events/step)
  e2e            through ShardedEventNet (the multi-GPU product path): pinned host events in, every rank's head
                 copied straight into one shared page-locked host array -> [n_gpus*S, 5, 7, 110] assembled on rank 0

Nothing here reads /root/reference.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, LEAK, ALPHA = 160, 224, 5e-5, 0.1
METRIC = "efcn_event_inference_throughput"
UNIT = "events/s"
# BASELINE config 5 ("deeper EFCN variant", builder-defined as SURVEY 8(d) allows: two 3x3 convs in the first two
# stages), on a DAVIS346 frame cropped to 256x320 (pooled dimensions must be even)
STRESS_LAYERS = ("conv1=3,3,1,16 conv1b=3,3,16,16 pool1=2,2 conv2=3,3,16,32 conv2b=3,3,32,32 pool2=2,2 conv3=3,3,32,64 pool3=2,2 "
                 "conv4=3,3,64,128 pool4=2,2 conv5=3,3,128,256 pool5=2,2 conv6=1,1,256,512 conv7=1,1,512,110")
STRESS_H, STRESS_W, STRESS_RATE = 256, 320, 10.0        # 10 events/us = 10 Mev/s per stream
CPU_MAX_STEPS = 600                                      # timed CPU steps available to a baseline worker


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="concurrent event streams PER GPU")
    ap.add_argument("--batch", type=int, default=200, help="events per stream per step (batch_event_size)")
    ap.add_argument("--kind", default="edge", choices=["edge", "uniform"], help="synthetic stream kind (SURVEY 8d)")
    ap.add_argument("--preroll", type=int, default=160, help="untimed steps to reach the workload's steady state (live-site fraction settles after ~120)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the BASELINE config 4 / 5 legs (they run at --gpus 1 only)")
    ap.add_argument("--leg-steps", type=int, default=5)
    ap.add_argument("--latency-steps", type=int, default=100, help="steps of the single-stream latency measurement (0 = skip)")
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="extra back-to-back steps after the timed region (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="timed CPU work per core for the baseline")
    ap.add_argument("--cpu-worker", type=int, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-steps", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-events", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-dump", default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port (one reference-style network object per stream, one process per core)
# ------------------------------------------------------------------------------------------------
def cpu_worker(args):
    """One process = one stream on one core, OMP_NUM_THREADS=1.  Times only the
    compute chain + final featuremap, as runner.py:84-89 does for the event runner.  The stream is row
    `--cpu-worker` of the event file the parent wrote (the same events GPU stream `--cpu-worker` gets); with
    --cpu-dump the integer state and the head after the untimed steps are saved for the parity check."""
    import async_ev_cnn_b200 as P
    from oracle.event_oracle import OracleEventNet
    idx = args.cpu_worker
    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    net = OracleEventNet(H, W, P.EFCN_LAYERS, wts, LEAK, ALPHA, "SAME")
    evs = np.load(args.cpu_events, mmap_mode="r")[idx]
    untimed = args.preroll + args.warmup
    max_steps = args.cpu_steps if args.cpu_steps > 0 else evs.shape[0] - untimed
    t = 0
    head = None
    for _ in range(untimed):
        head = net.step(np.asarray(evs[t]))
        t += 1
    if args.cpu_dump:
        rec = {"head": np.asarray(head, np.float32)}
        for i, nm in enumerate(net.names):
            rec["front_" + nm] = np.packbits(net.frontier_mask(i))
            layer = net.layers[i]
            if hasattr(layer, "idx"):
                rec["idx_" + nm] = layer.idx.reshape(layer.shape).astype(np.uint8)
                rec["flags_" + nm] = np.packbits(layer.flags)
                # the oracle's own pre-activations under the windows: what decides whether an argmax disagreement is a near tie
                rec["below_" + nm] = np.asarray(net.layers[i - 1].F, np.float32)
        np.savez(args.cpu_dump + "%d.npz" % idx, **rec)
    done = 0
    t0 = time.perf_counter()
    while done < max_steps:
        net.step(np.asarray(evs[t]))
        t += 1
        done += 1
        if args.cpu_steps <= 0 and time.perf_counter() - t0 >= args.cpu_seconds:
            break
    dt = time.perf_counter() - t0
    print(json.dumps({"steps": done, "seconds": dt}))


def check_streams(args, P, n_check, untimed, timed):
    """Events of the streams that both the CPU workers and GPU streams 0..n_check-1 consume: [n_check, steps, B, 3]."""
    return P.synthetic_events(args.kind, n_check, untimed + timed, args.batch, H, W, seed=7000)


def run_cpu_processes(args, fixed_steps, events_path, warmup, cores, dump_prefix=None):
    """Runs `cores` workers (one per host core) concurrently; returns (aggregate events/s, cores, per-core steps, seconds)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--kind", args.kind, "--batch", str(args.batch),
           "--preroll", str(args.preroll), "--warmup", str(warmup), "--cpu-seconds", str(args.cpu_seconds),
           "--cpu-steps", str(fixed_steps), "--cpu-events", events_path]
    if dump_prefix:
        cmd += ["--cpu-dump", dump_prefix]
    procs = [subprocess.Popen(cmd + ["--cpu-worker", str(i)], stdout=subprocess.PIPE, env=env, text=True) for i in range(cores)]
    rate, steps, secs = 0.0, [], []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("CPU baseline worker failed")
        r = json.loads(out.strip().splitlines()[-1])
        rate += r["steps"] * args.batch / r["seconds"]
        steps.append(r["steps"])
        secs.append(r["seconds"])
    return rate, cores, steps, secs


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores.
    The reference is Python + one Cython module and cannot travel to the GPU box, so this times the
    oracle port (bit-identical to the reference and equally fast: tests/test_oracle_vs_reference.py),
    one stream per core on all cores; a step = every core advances its stream by one batch."""
    import async_ev_cnn_b200 as P
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = max(args.warmup, 2)
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="aec_bench_")
    try:
        path = os.path.join(tmp, "events.npy")
        np.save(path, check_streams(args, P, cores, args.preroll + warm, steps))
        rate, cores, st, secs = run_cpu_processes(args, steps, path, warm, cores)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ms = 1e3 * float(np.mean(secs)) / steps
    sample = "%d cores x 1 stream x %d steps x %d events (%s stream, %d pre-roll steps)" % (cores, steps, args.batch, args.kind, args.preroll)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.streams),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, streams):
    return {"workload": "configs/efcn_event.yml EFCN event-mode, 160x224 (240x180 sensor centre-cropped), B=%d events/stream/step, "
                        "synthetic %s streams, random-init weights" % (args.batch, args.kind),
            "streams_per_gpu": streams, "batch_event_size": args.batch, "stream_kind": args.kind, "frame": [H, W],
            "leak": LEAK, "alpha": ALPHA, "preroll_steps": args.preroll,
            "l2": "inputs larger than L2: %.1f GB of stream state per GPU, of which every step reads or writes several GB "
                  "(re-evaluated sites, pool windows, the leak sweep; roofline_hbm.whole_step); L2 = 126 MB" % (streams * 12.1e-3)}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (a query takes well under a millisecond,
    so even a 70 ms region gets samples), `nvidia-smi -lms` as the fallback (its first line can take longer than the region:
    start() then waits for it)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]          # NVML clocks event reasons: HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag, self.source = [], None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._sample_nvml()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._loop_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
            self.rows = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t_end = time.perf_counter() + 3.0
            while not self.rows and time.perf_counter() < t_end:      # the first line can take a second on a busy host
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        sm = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
        mask = int(self._reasons(self.handle))
        self.rows.append((time.perf_counter(), [str(sm), str(self.max_mhz), ""] + ["Active" if mask & b else "Not Active" for b in self.BITS]))

    def _loop_nvml(self):
        while not self.stop_flag:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"], "samples": 0}
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.15)
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for t, r in self.rows if t0 - 0.2 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for r in rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows), "source": self.source}


# ------------------------------------------------------------------------------------------------
# GPU side: accounting
# ------------------------------------------------------------------------------------------------
def layer_bytes(net, i, n, streams):
    """Algorithmic bytes of layer i's evaluation for n re-evaluated sites / windows: a conv writes F and A
    at its sites (8 B x Cout) and reads every distinct input value + rate once (8 B x Cin, at most the whole input
    map); a pool reads 4 values + 4 rates and writes the index per (window, channel) (36 B)."""
    shapes = net.shapes()
    c, h, w = shapes[i]
    nm = net.names[i]
    if "conv" in nm:
        cin, hin, win = shapes[i - 1]
        k = net.infos[i].k_h * net.infos[i].k_w
        in_bytes = 8 * cin if i > 1 else 8            # first conv reads the float64 surface
        return n * 8 * c + min(k * n, streams * hin * win) * in_bytes
    if "pool" in nm:
        return n * c * 36.0
    return 0.0


def algorithmic_bytes(net, sites_per_step, sw, streams, batch, h, w):
    """SURVEY 8(d) / BASELINE.md section 4 per step over all streams, with the leak term stated for what the
    implemented algorithm needs: A and F are read and F written (12 B/elem) at the live sites (non-zero-rate bit
    set) that the step does not re-evaluate anyway - a re-evaluated site gets F and A overwritten, so its leak is
    not needed (`swept_conv_elems`) - plus the bitmaps (non-zero-rate, skip, sign-change).
    The pool layers' (Fp, Ap) copies are an implementation choice and are NOT counted as algorithmic."""
    shapes = net.shapes()
    bitmaps = sum(hh * ((ww + 31) // 32) * 4 for nm, (c, hh, ww) in zip(net.names, shapes) if "conv" in nm) * streams
    leak = 12.0 * sw.get("swept_conv_elems", sw["live_conv_elems"]) + 3 * bitmaps
    surface = streams * 2 * 8 * h * w
    ev = streams * 12 * batch
    conv = sum(layer_bytes(net, i, float(sites_per_step[i]), streams) for i, nm in enumerate(net.names) if "conv" in nm)
    pool = sum(layer_bytes(net, i, float(sites_per_step[i]), streams) for i, nm in enumerate(net.names) if "pool" in nm)
    return {"leak_sweep": leak, "surface": surface, "conv": conv, "pool": pool, "events": ev,
            "total": leak + surface + conv + pool + ev}


def layer_flops(net, i, n):
    """Useful FLOPs of conv layer i for n re-evaluated sites: sites x {value, rate} x 2 x K x Cout."""
    info = net.infos[i]
    return float(n) * 2 * 2 * info.k_h * info.k_w * info.in_channels * net.shapes()[i][0]


def tf32_peak(peaks, sm_mhz):
    """Measured tensor peak for kind::tf32: profiles/r2_umma_peak.json holds the cycles one tcgen05.mma
    (M128 x N256 x K8) takes when issued back to back (tools/bench_umma.cu, run on the box); x 148 SMs x the SM clock
    sampled during the timed region.  Fallback: MEASURED_PEAKS bf16 sustained / 2."""
    try:
        um = json.load(open(os.path.join(ROOT, "profiles", "r2_umma_peak.json")))
        cyc = float(um["tf32_m128_n256_k8_cycles_per_mma"])
        mhz = float(sm_mhz or um.get("sm_mhz_assumed", 1800.0))
        sms = int(um.get("sms", 148))
        return 2.0 * 128 * 256 * 8 / cyc * sms * mhz * 1e6 / 1e12, "profiles/r2_umma_peak.json: %.1f cycles per M128xN256xK8 kind::tf32 MMA (tools/bench_umma.cu on the box) x %d SMs x %.0f MHz (SM clock sampled in the timed region)" % (cyc, sms, mhz)
    except Exception:
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        return bf16 / 2.0, ("MEASURED_PEAKS.json bf16_tflops_sustained / 2" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s bf16 / 2")


def layer_table(net, prof, sites_per_step, streams, tf32_pk, hbm_pk, units_per_step=None):
    """Per-launch table of one step: ms, work, useful TFLOP/s (conv) and algorithmic GB/s against the measured peaks."""
    rows = {}
    for i, nm in enumerate(net.names):
        if i == 0:
            continue
        ms = prof.get(nm + ".eval", 0.0)
        n = float(sites_per_step[i])
        row = {"ms": round(ms, 4), "sites_per_stream": round(n / streams, 1)}
        b = layer_bytes(net, i, n, streams)
        if ms > 0:
            row["alg_GB"] = round(b / 1e9, 4)
            row["alg_GBps"] = round(b / (ms * 1e-3) / 1e9, 1)
            row["hbm_frac"] = round(b / (ms * 1e-3) / 1e9 / hbm_pk, 4)
            if "conv" in nm:
                fl = layer_flops(net, i, n)
                tc = net.tc_geometry(i)
                row["useful_TFLOPs"] = round(fl / (ms * 1e-3) / 1e12, 2)
                row["useful_frac_of_tf32_peak"] = round(fl / (ms * 1e-3) / 1e12 / tf32_pk, 4)
                if tc is not None:
                    n_units = float(units_per_step[i]) if (tc.get("units_counted") and units_per_step is not None) else np.ceil(n / tc["unit_sites"])
                    row["units_of_128_sites"] = round(n_units, 1)
                    issued = tc["mma_flops_per_unit"] * n_units * tc["m_groups"]
                    row["issued_TFLOPs"] = round(issued / (ms * 1e-3) / 1e12, 2)
                    row["issued_frac_of_tf32_peak"] = round(issued / (ms * 1e-3) / 1e12 / tf32_pk, 4)
                    row["kernel"] = tc["kernel"]
                else:
                    row["kernel"] = "k_conv_stencil / k_conv_eval (SIMT)"
            else:
                row["kernel"] = "k_pool_eval"
                in_sweep = float(units_per_step[i]) if units_per_step is not None else 0.0
                if in_sweep > 0:
                    # windows evaluated inside k_sweep_windows (pool behind the first conv layer) or in the epilogue of the gathered conv
                    # kernel (complete windows of a pool behind it) cost no time here: rate this launch on the windows it evaluated
                    b = layer_bytes(net, i, n - in_sweep, streams)
                    row.update({"windows_evaluated_elsewhere_per_stream": round(in_sweep / streams, 1), "alg_GB": round(b / 1e9, 4),
                                "alg_GBps": round(b / (ms * 1e-3) / 1e9, 1), "hbm_frac": round(b / (ms * 1e-3) / 1e9 / hbm_pk, 4),
                                "kernel": "k_pool_eval (the remaining windows; the others in k_sweep_windows / the conv kernel's epilogue)"})
        rows[nm] = row
    return rows


def gen_events(P, kind, S, n_steps, B, h, w, seed, rate=0.4, block=256):
    """[n_steps, S*B, 3] int32, streams packed per step; generated in blocks of streams to bound the temporaries."""
    out = np.empty((n_steps, S * B, 3), np.int32)
    for b0 in range(0, S, block):
        nb = min(block, S - b0)
        ev = P.synthetic_events(kind, nb, n_steps, B, h, w, seed=seed + 7919 * (b0 // block), rate=rate)
        out[:, b0 * B:(b0 + nb) * B] = ev.transpose(1, 0, 2, 3).reshape(n_steps, nb * B, 3)
    return out


def run_leg(torch, P, EventNetCuda, name, layers, h, w, S, B, kind, preroll, K, local, peaks, rate=0.4, seed=300):
    """One extra configuration measured like the headline: pre-roll on device-resident events, K timed steps between
    CUDA events, then K profiled steps for the per-launch table."""
    wts = P.xavier_weights(layers, seed=0 if layers == P.EFCN_LAYERS else 3)
    n_steps = preroll + 3 + 2 * K
    t_gen = time.perf_counter()
    ev_np = gen_events(P, kind, S, n_steps, B, h, w, seed, rate=rate)
    t_gen = time.perf_counter() - t_gen
    net = EventNetCuda(h, w, layers, wts, LEAK, ALPHA, "SAME", n_streams=S, device=local, max_events_per_step=max(2048, B))
    ev_dev = torch.from_numpy(ev_np).cuda()
    off_dev = torch.from_numpy((np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)).cuda()
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    t = 0
    for _ in range(preroll + 3):
        net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
        t += 1
    net.counters(reset=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    time.sleep(0.2)
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
        t += 1
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop(w0, time.perf_counter())
    ms = e0.elapsed_time(e1) / K
    units_per_step = net.unit_counters().astype(np.float64) / K
    sites, _ = net.counters(reset=True)
    sites_per_step = sites.astype(np.float64) / K
    sw = net.sweep_stats()
    net.profile(True)
    for _ in range(K):
        net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)
        t += 1
    prof, _ = net.read_profile()
    net.profile(False)
    hbm_pk = float(peaks.get("hbm_gbs", 6650.0))
    tf_pk, _ = tf32_peak(peaks, clocks.get("sm_mhz"))
    ab = algorithmic_bytes(net, sites_per_step, sw, S, B, h, w)
    table = layer_table(net, prof, sites_per_step, S, tf_pk, hbm_pk, units_per_step)
    top = max(prof.items(), key=lambda kv: kv[1])
    out = {"value": S * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": K, "preroll_steps": preroll,
           "streams_per_gpu": S, "batch_event_size": B, "stream_kind": kind, "frame": [h, w], "event_rate_per_us": rate,
           "state_GB": round(net.device_bytes() / 1e9, 2),
           "whole_step": {"algorithmic_bytes": ab["total"], "achieved_gbs": ab["total"] / (ms * 1e-3) / 1e9,
                          "frac_of_hbm_peak": ab["total"] / (ms * 1e-3) / 1e9 / hbm_pk},
           "dominant_launch": {"name": top[0], "ms": round(top[1], 4), "share": round(top[1] / max(1e-9, sum(prof.values())), 3)},
           "ms_by_launch": {k: round(v, 4) for k, v in prof.items()}, "layers": table,
           "live_site_fraction": sw["live_conv_elems"] / max(1, sw["conv_elems"]),
           "clocks": clocks, "host_seconds_generating_events": round(t_gen, 1)}
    net.close()
    del ev_dev
    torch.cuda.empty_cache()
    return out


def parity_check(net, dump_prefix, n_check, steps_done):
    """GPU streams 0..n_check-1 against the oracle workers that consumed the same events, after `steps_done` steps."""
    heads = net.read_head(0, n_check)
    out = {"streams": n_check, "steps": steps_done, "head_max_rel_err": 0.0, "frontier_sites": 0, "frontier_mismatch": 0,
           "argmax_entries": 0, "argmax_mismatch": 0, "argmax_near_ties": 0, "argmax_not_near_ties": 0, "flag_windows": 0, "flag_mismatch": 0, "streams_with_any_mismatch": 0}
    for s in range(n_check):
        z = np.load(dump_prefix + "%d.npz" % s)
        want = z["head"]
        err = float(np.abs(heads[s].astype(np.float64) - want).max()) / max(float(np.abs(want).max()), 1e-30)
        out["head_max_rel_err"] = max(out["head_max_rel_err"], err)
        any_bad = False
        for i, nm in enumerate(net.names):
            hh, ww = net.infos[i].height, net.infos[i].width
            wf = np.unpackbits(z["front_" + nm])[:hh * ww].reshape(hh, ww).astype(bool)
            gf = net.frontier(i, s)
            out["frontier_sites"] += int(wf.sum())
            d = int((gf ^ wf).sum())
            out["frontier_mismatch"] += d
            any_bad |= d > 0
            if "idx_" + nm in z:
                st = net.state(i, s)
                wi = z["idx_" + nm]
                wfl = np.unpackbits(z["flags_" + nm])[:hh * ww].reshape(hh, ww).astype(bool)
                out["argmax_entries"] += wi.size
                diff = st["idx"] != wi
                d = int(diff.sum())
                out["argmax_mismatch"] += d
                any_bad |= d > 0
                if d:
                    # the two candidates by the ORACLE'S values: within NEAR_TIE (1e-5, tests/parity.py) of the map scale = a tie that
                    # rounding decides (the reference's BLAS order is unspecified too); anything else is counted separately
                    below = z["below_" + nm]
                    k = below.shape[1] // wi.shape[1]
                    scale = max(float(np.abs(below).max()), 1e-30)
                    for c, y, x in zip(*np.nonzero(diff)):
                        a, b = int(st["idx"][c, y, x]), int(wi[c, y, x])
                        fa, fb = float(below[c, y * k + a // k, x * k + a % k]), float(below[c, y * k + b // k, x * k + b % k])
                        out["argmax_near_ties" if abs(fa - fb) <= 1e-5 * scale else "argmax_not_near_ties"] += 1
                out["flag_windows"] += wfl.size
                d = int((st["flags"] != wfl).sum())
                out["flag_mismatch"] += d
                any_bad |= d > 0
        out["streams_with_any_mismatch"] += int(any_bad)
    out["note"] = ("the CPU baseline's oracle workers and GPU streams 0..%d consume the same events; compared after the "
                   "pre-roll + warm-up (the state the timed region starts from).  A non-zero integer mismatch on a float net "
                   "must be a near tie of the oracle's own values (argmax_near_ties: the two candidates' pre-activations in the "
                   "oracle differ by <= 1e-5 of the map scale; argmax_not_near_ties counts the rest, e.g. windows downstream of "
                   "an earlier flip): tests/test_cuda_fullsize.py::test_efcn_benchmark_regime_against_live_oracle asserts the "
                   "full explanation chain over 224 steps" % (n_check - 1))
    return out


def native_arm(args):
    import torch
    import torch.distributed as dist
    import async_ev_cnn_b200 as P
    from async_ev_cnn_b200.engine import EventNetCuda
    from async_ev_cnn_b200.sharding import ShardedEventNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    S, B, K, Wm = args.streams, args.batch, args.steps, max(args.warmup, 3)
    n_e2e = 2 * K + 4
    n_steps = args.preroll + Wm + 2 * K + 2 + n_e2e

    # ---- CPU baseline first (before CUDA is touched in this process); its streams become GPU streams 0..P-1
    cpu = None
    chk = None
    tmp = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, S)
        tmp = tempfile.mkdtemp(prefix="aec_bench_")
        chk = check_streams(args, P, cores, args.preroll + Wm, max(CPU_MAX_STEPS, n_steps))
        path = os.path.join(tmp, "events.npy")
        np.save(path, chk)
        rate, cores, st, secs = run_cpu_processes(args, 0, path, Wm, cores, dump_prefix=os.path.join(tmp, "state"))
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d cores x 1 stream each, %.0f s timed per core after %d pre-roll steps (%d..%d steps of %d events, %s stream)" % (
                   cores, args.cpu_seconds, args.preroll, min(st), max(st), args.batch, args.kind)}

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wts = P.xavier_weights(P.EFCN_LAYERS, seed=0)
    # the product's multi-GPU object: world*S global streams, this rank owns a contiguous block of S (sharding.py)
    shard = ShardedEventNet(H, W, P.EFCN_LAYERS, wts, LEAK, ALPHA, "SAME", n_streams=world * S, device=local,
                            max_events_per_step=max(2048, B))
    net = shard.net
    assert net.n_streams == S
    ev_np = P.synthetic_events(args.kind, S, n_steps, B, H, W, seed=100 + rank)        # [S, steps, B, 3]
    n_check = 0
    if chk is not None:
        n_check = chk.shape[0]
        ev_np[:n_check] = chk[:, :n_steps]
    ev_np = np.ascontiguousarray(ev_np.transpose(1, 0, 2, 3)).reshape(n_steps, S * B, 3)   # per step: streams packed
    off_np = (np.arange(S + 1, dtype=np.int64) * B).astype(np.int32)
    ev_dev = torch.from_numpy(ev_np).cuda()
    off_dev = torch.from_numpy(off_np).cuda()
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def gpu_step(t):
        net.step_device(ev_dev[t].data_ptr(), off_dev.data_ptr(), S * B, sh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = 0
    for _ in range(args.preroll + Wm):
        gpu_step(t)
        t += 1
    pc = None
    if n_check:
        torch.cuda.synchronize()
        pc = parity_check(net, os.path.join(tmp, "state"), n_check, t)
    if tmp:
        shutil.rmtree(tmp, ignore_errors=True)
    # ---- timed region: device-resident inputs, CUDA events on the launching stream
    launches0 = net.launch_count()
    net.counters(reset=True)
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        gpu_step(t)
        t += 1
    e1.record(stream)
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1)
    ms_total = e0.elapsed_time(e1)
    launches = net.launch_count() - launches0
    units_per_step = net.unit_counters().astype(np.float64) / K
    sites, _ = net.counters(reset=True)
    sites_per_step = sites.astype(np.float64) / K
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    value = world * S * B * K / (ms_total * 1e-3)

    # ---- per-kernel durations inside the real step (events after every launch), same K steps again
    sw = net.sweep_stats()
    net.profile(True)
    for _ in range(K):
        gpu_step(t)
        t += 1
    prof, psteps = net.read_profile()
    net.profile(False)
    step_ms_prof = sum(prof.values())
    tc_layers = net.tc_layers()
    tc_names = {net.names[i] + ".eval" for i in tc_layers}
    by_kernel = {}
    for name, ms in prof.items():
        key = name.split(".")[-1] if "." in name else name
        key = {"frontier": "frontier_bitmaps", "eval": "site_eval"}.get(key, key)
        if key == "site_eval":
            key = ("conv_eval_tc" if name in tc_names else "conv_eval_simt") if "conv" in name else "pool_eval"
        by_kernel[key] = by_kernel.get(key, 0.0) + ms
    ab = algorithmic_bytes(net, sites_per_step, sw, S, B, H, W)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel of the step: the tcgen05 conv re-evaluation - k_conv_rows (row tiles) or k_conv_eval_tc (gathered),
    # whichever holds more of the step; the per-layer table carries every launch of both
    tf_pk, tf_src = tf32_peak(peaks, clocks.get("sm_mhz"))
    table = layer_table(net, prof, sites_per_step, S, tf_pk, peak, units_per_step)
    fam = {}
    for i in tc_layers:
        row = table[net.names[i]]
        k = "k_conv_rows" if row.get("kernel", "").startswith("k_conv_rows") else "k_conv_eval_tc"
        f = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "issued": 0.0, "n": 0})
        f["ms"] += row["ms"]
        f["flops"] += layer_flops(net, i, sites_per_step[i])
        f["issued"] += row.get("issued_TFLOPs", 0.0) * row["ms"]
        f["bytes"] = f.get("bytes", 0.0) + row.get("alg_GB", 0.0) * 1e9
        f["n"] += 1
    dom = max(fam, key=lambda k: fam[k]["ms"]) if fam else "k_conv_eval_tc"
    fd = fam.get(dom, {"ms": 0.0, "flops": 0.0, "issued": 0.0, "n": 1})
    tc_ms = by_kernel.get("conv_eval_tc", 0.0)
    dom_tflops = fd["flops"] / (fd["ms"] * 1e-3) / 1e12 if fd["ms"] > 0 else 0.0
    # Which roof bounds it: arithmetic intensity (useful FLOPs per algorithmic byte) against the ridge of the two measured peaks.
    # Row tiles (Cout <= 64: 33-72 FLOP/B) sit left of it - the HBM roof is the attainable one - the gathered layers
    # (Cout >= 128) on it or right of it.  `frac` = attainable time / measured time of the binding roof; both are reported.
    dom_bytes = fd.get("bytes", 0.0)
    dom_gbs = dom_bytes / (fd["ms"] * 1e-3) / 1e9 if fd["ms"] > 0 else 0.0
    intensity = fd["flops"] / dom_bytes if dom_bytes > 0 else float("inf")
    ridge = tf_pk * 1e12 / (peak * 1e9)
    hbm_bound = intensity < ridge
    roofline = {"bound": "hbm" if hbm_bound else "tensor", "kernel": dom,
                "achieved": dom_gbs if hbm_bound else dom_tflops, "peak": peak if hbm_bound else tf_pk, "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": (dom_gbs / peak) if hbm_bound else (dom_tflops / tf_pk), "traffic": None,
                "peak_source": (peak_src if hbm_bound else tf_src),
                "intensity_flop_per_byte": round(intensity, 1), "ridge_flop_per_byte": round(ridge, 1),
                "bound_rule": "useful FLOPs / algorithmic bytes of the kernel's launches against the ridge tensor peak / HBM peak: left of it the HBM roof "
                              "is the attainable one, right of it the tensor roof; both fractions follow",
                "hbm": {"achieved_GBps": round(dom_gbs, 1), "peak_GBps": peak, "frac": round(dom_gbs / peak, 4), "algorithmic_bytes_per_launch": dom_bytes / max(1, fd["n"]),
                        "peak_source": peak_src},
                "tensor": {"useful_TFLOPs": round(dom_tflops, 2), "peak_TFLOPs": round(tf_pk, 1), "frac": round(dom_tflops / tf_pk, 4), "peak_source": tf_src},
                "precision": "3xTF32 (two or three tcgen05.mma.kind::tf32 per product for fp32-grade results): `achieved` counts "
                             "USEFUL FLOPs (2 x sites x {value, rate} x K x Cout); `issued` what the tensor pipe was asked to do, "
                             "padding rows, gap sites of a row tile and split products included",
                "issued_TFLOPs": fd["issued"] / fd["ms"] if fd["ms"] > 0 else 0.0,
                "issued_frac_of_peak": (fd["issued"] / fd["ms"] / tf_pk) if fd["ms"] > 0 else 0.0,
                "algorithmic_flops_per_launch": fd["flops"] / max(1, fd["n"]), "launches_per_step": fd["n"], "kernel_ms": fd["ms"] / max(1, fd["n"]),
                "kernel_share_of_step": fd["ms"] / step_ms_prof if step_ms_prof else None,
                "all_tcgen05_conv_kernels": {k: {"ms": round(v["ms"], 4), "launches": v["n"], "useful_TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                                                 "issued_TFLOPs": round(v["issued"] / v["ms"], 2) if v["ms"] > 0 else 0.0} for k, v in fam.items()},
                "ms_by_kernel": {k: round(v, 4) for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1])},
                "ms_by_launch": {k: round(v, 4) for k, v in prof.items()},
                "layers": table,
                "sites_per_step_per_stream": {nm: round(float(sites_per_step[i]) / S, 1) for i, nm in enumerate(net.names) if i}}
    # DRAM traffic per launch from the committed ncu --set full capture of the same workload (profiles/traffic.json)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        same = (tj["config"]["streams_per_gpu"] == S and tj["config"]["stream_kind"] == args.kind and tj["config"]["batch_event_size"] == B)
    except Exception:
        tj, same = None, False
    if same:
        roofline["traffic"] = tj["dram_bytes_per_launch"].get(dom)
        roofline["traffic_source"] = tj["source"]
    # the HBM-bound group of the step: leak sweep (plain + window form, which also evaluates the first pool layer's sticky
    # windows) and the pool evaluations - their algorithmic bytes against their summed in-step durations
    sweep_ms = prof.get("leak_sweep", 0.0) + prof.get("window_sweep", 0.0)
    pool_ms = sum(v for k, v in prof.items() if k.endswith(".eval") and "pool" in k)
    grp_ms = sweep_ms + pool_ms
    # pool windows evaluated in a conv kernel's epilogue (unit counter of a pool layer other than the one behind the first conv,
    # whose share is evaluated by k_sweep_windows, i.e. inside this group) cost no HBM reads and no time in this group
    pool_in_conv_bytes = sum(layer_bytes(net, i, float(units_per_step[i]), S) for i, nm in enumerate(net.names) if "pool" in nm and i != 2)
    grp_bytes = ab["leak_sweep"] + ab["pool"] - pool_in_conv_bytes
    achieved = grp_bytes / (grp_ms * 1e-3) / 1e9 if grp_ms > 0 else 0.0
    roofline_hbm = {"bound": "hbm", "kernel": "k_sweep_windows + k_leak_sweep + k_pool_eval (leak of the conv maps and pool re-evaluation)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_step": grp_bytes, "leak_bytes": ab["leak_sweep"], "pool_bytes": ab["pool"] - pool_in_conv_bytes,
                    "pool_bytes_saved_in_conv_epilogues": pool_in_conv_bytes,
                    "kernel_ms": grp_ms, "sweep_ms": sweep_ms, "pool_ms": pool_ms,
                    "kernel_share_of_step": grp_ms / step_ms_prof if step_ms_prof else None,
                    "live_site_fraction": sw["live_conv_elems"] / max(1, sw["conv_elems"]),
                    "swept_fraction": sw.get("swept_conv_elems", sw["live_conv_elems"]) / max(1, sw["conv_elems"]),
                    "nonzero_rate_group_fraction": sw["nz_groups"] / max(1, sw["groups"]),
                    "unaccounted": "the pool layers' (Fp, Ap) copies swept by the same launches (%.3g of %.3g elements x 12 B) are not in the algorithmic bytes" % (
                        sw.get("swept_pool_elems", sw["live_pool_elems"]), sw["pool_elems"]),
                    "whole_step": {"algorithmic_bytes": ab["total"], "achieved_gbs": ab["total"] / (ms_total / K * 1e-3) / 1e9,
                                   "frac": ab["total"] / (ms_total / K * 1e-3) / 1e9 / peak,
                                   "note": "bytes the implemented algorithm needs (sparse leak sweep + measured frontier terms, SURVEY 8(d) "
                                           "with the leak term restricted to the sites that are swept)"}}
    if same:
        d = tj["dram_bytes_per_launch"]
        lp = tj["launches_per_step"]
        roofline_hbm["traffic"] = sum(d.get(k, 0.0) * lp.get(k, 1) for k in ("k_leak_sweep", "k_sweep_windows", "k_pool_eval"))

    # ---- end to end through the multi-GPU product path (ShardedEventNet): pinned host events in; every rank's head is
    # copied by its own GPU straight into its rows of ONE shared page-locked host array, so after sync() rank 0 holds
    # the assembled [world*S, 5, 7, 110] detections ("gathered on the host", north_star) - no collective, no staging.
    # (a) the pipelined form (two steps in flight, copies overlap the kernels) is the throughput a user gets,
    # (b) the blocking form waits for the assembled detections of every step before the next one starts.
    ev_host = torch.from_numpy(ev_np[t:t + n_e2e]).pin_memory()
    off_host = torch.from_numpy(off_np).pin_memory()
    evh, offh = ev_host.numpy(), off_host.numpy()
    shard.open_host_gather(slots=2)
    shard.step_packed_async(evh[0], offh, cuda_stream=sh)          # warm the path (staging buffers, copy streams)
    shard.step_packed_async(evh[1], offh, cuda_stream=sh)
    shard.sync(sh)
    half = max(3, n_e2e // 2)
    barrier()
    w0 = time.perf_counter()
    last_slot = 0
    for i in range(2, half):
        last_slot = shard.step_packed_async(evh[i], offh, cuda_stream=sh)
    shard.sync(sh)                                                  # every rank's rows of every enqueued step are in host memory
    e2e_s = time.perf_counter() - w0
    e2e_steps = half - 2
    assembled = None
    if rank == 0:
        a = shard.assembled(last_slot)
        blocks = [float(np.abs(a[r * S:(r + 1) * S]).sum()) for r in range(world)]
        assembled = {"shape": list(a.shape), "abs_sum": float(sum(blocks)), "rank_blocks_nonzero": int(sum(b > 0 for b in blocks)),
                     "finite": bool(np.isfinite(a).all()),
                     "where": "one POSIX shared-memory array on the host, page-locked by every rank (cudaHostRegister); each rank's "
                              "device-to-host copy writes its own rows"}
    barrier()
    w0 = time.perf_counter()
    for i in range(half, n_e2e):
        shard.step_packed_async(evh[i], offh, cuda_stream=sh)
        shard.sync(sh)
    sync_s = time.perf_counter() - w0
    sync_steps = n_e2e - half
    if world > 1:
        tt = torch.tensor([e2e_s, sync_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s, sync_s = float(tt[0].item()), float(tt[1].item())
    e2e_value = world * S * B * e2e_steps / e2e_s
    e2e_sync_value = world * S * B * sync_steps / sync_s
    head_bytes = int(np.prod((S,) + net.head_shape)) * 4

    # ---- sustained rate: the same step back to back for a few seconds (the board reaches its 1 kW power cap after
    # about a second of this).  The event batches are reused with their timestamps shifted forward by the length of
    # the sequence at every wrap, so stream time keeps increasing.  Reported, not the headline.
    sustained = None
    if args.sustained_seconds > 0:
        n_avail = ev_dev.shape[0]
        span = int(ev_np[:, :, 2].max()) - int(ev_np[:, :, 2].min()) + 500
        t_end = time.perf_counter() + args.sustained_seconds
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = 0
        last_ms, last_n = 0.0, 0
        max_wraps = (2 ** 31 - 1 - int(ev_np[:, :, 2].max())) // span
        while time.perf_counter() < t_end and done // n_avail < max_wraps:
            s0.record(stream)
            for i in range(32):
                if (done + i) % n_avail == 0:
                    ev_dev[:, :, 2] += span
                net.step_device(ev_dev[(done + i) % n_avail].data_ptr(), off_dev.data_ptr(), S * B, sh)
            s1.record(stream)
            torch.cuda.synchronize()
            done += 32
            last_ms, last_n = s0.elapsed_time(s1), 32
        if last_n:
            sustained = {"value": world * S * B * last_n / (last_ms * 1e-3), "unit": UNIT, "ms_per_step": last_ms / last_n,
                         "after_seconds": args.sustained_seconds, "steps_run": done,
                         "note": "last 32 of the back-to-back steps (event batches reused, timestamps shifted forward at every wrap), "
                                 "rank 0's GPU, at the 1 kW power cap"}
    state_bytes, device_bytes = net.state_bytes_per_stream(), net.device_bytes()
    shard.close()
    del ev_dev
    torch.cuda.empty_cache()

    # ---- single-stream latency (the reference's own operating point: batch_size 1, one network object = one stream)
    latency = None
    if rank == 0 and args.latency_steps > 0:
        one = EventNetCuda(H, W, P.EFCN_LAYERS, wts, LEAK, ALPHA, "SAME", n_streams=1, device=local, max_events_per_step=max(2048, B))
        ev1 = P.synthetic_events(args.kind, 1, args.preroll + args.latency_steps + 1, B, H, W, seed=4242)[0]
        ev1h = torch.from_numpy(np.ascontiguousarray(ev1)).pin_memory().numpy()
        off1 = torch.from_numpy(np.array([0, B], np.int32)).pin_memory().numpy()
        out1 = torch.empty((1,) + one.head_shape, dtype=torch.float32).pin_memory().numpy()
        for i in range(args.preroll):
            one.step_packed(ev1h[i], off1, out=out1, cuda_stream=sh)
        w0 = time.perf_counter()
        for i in range(args.preroll, args.preroll + args.latency_steps):
            one.step_packed(ev1h[i], off1, out=out1, cuda_stream=sh)      # blocking: events in, detections out
        lat = (time.perf_counter() - w0) / args.latency_steps
        latency = {"ms_per_step": 1e3 * lat, "events_per_s": B / lat, "streams": 1, "steps": args.latency_steps,
                   "api": "aec_net_step_host, one stream, host events in / host head out per call"}
        one.close()

    # ---- BASELINE configs 4 and 5 in the same run (one GPU): stream-count sweep, saturated streams, the stress net
    legs = None
    if rank == 0 and world == 1 and not args.no_legs:
        KL = max(1, args.leg_steps)
        legs = {}
        specs = [("uniform_S%d" % S, P.EFCN_LAYERS, H, W, S, B, "uniform", args.preroll, 0.4),
                 ("edge_S256", P.EFCN_LAYERS, H, W, 256, B, "edge", args.preroll, 0.4),
                 ("edge_S4096", P.EFCN_LAYERS, H, W, 4096, B, "edge", args.preroll, 0.4),
                 ("stress_256x320_deeper_B1000", STRESS_LAYERS, STRESS_H, STRESS_W, 256, 1000, "edge", max(args.preroll, 200), STRESS_RATE),
                 ("stress_256x320_deeper_B2000", STRESS_LAYERS, STRESS_H, STRESS_W, 256, 2000, "edge", max(args.preroll, 200), STRESS_RATE)]
        for name, layers, h, w, s_, b_, kind, pre, rate in specs:
            try:
                legs[name] = run_leg(torch, P, EventNetCuda, name, layers, h, w, s_, b_, kind, pre, KL, local, peaks, rate=rate)
            except Exception as e:                       # a leg must not take the headline down with it
                legs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
        legs["note"] = ("BASELINE configs 4 (256-4096 concurrent streams; uniform = every site on the frontier) and 5 (DAVIS346-sized frame "
                        "cropped to 256x320, ~10 Mev/s per stream, deeper EFCN variant: two 3x3 convs in the first two stages), each measured "
                        "like the headline: pre-roll, K timed steps between CUDA events on device-resident events, K profiled steps")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, S),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * (evh[0].nbytes + offh.nbytes)),
                    "d2h_bytes_per_step": int(world * head_bytes), "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "api": "ShardedEventNet.step_packed_async + sync (aec_net_step_host_async + aec_net_host_sync per rank: pipelined, "
                           "2 steps in flight, pinned host buffers; bytes are the whole job's)",
                    "assembled_on_rank0": assembled,
                    "blocking_api": {"value": e2e_sync_value, "ms_per_step": 1e3 * sync_s / sync_steps, "steps": sync_steps,
                                     "api": "step_packed_async + sync after every step: the assembled detections of step t are on rank 0's "
                                            "host before step t+1 starts"}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "state_bytes_per_stream": state_bytes, "device_bytes": device_bytes,
        }
        if pc is not None:
            line["parity_check"] = pc
        if sustained is not None:
            line["sustained"] = sustained
        if latency is not None:
            line["single_stream"] = latency
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if legs is not None:
            line["legs"] = legs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE line, the JSON: anything a library prints there (NCCL's version banner under torchrun,
    # nvcc / loader chatter) goes to stderr instead, and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    args = parse_args()
    if args.cpu_worker is not None:
        cpu_worker(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
