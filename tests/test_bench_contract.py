"""CPU: bench.py's reference arm (the oracle port timed on the host cores) prints ONE JSON line with the keys the
driver reads, and the GPU arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--preroll", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "efcn_event_inference_throughput" and d["unit"] == "events/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("configs/efcn_event.yml") and d["config"]["preroll_steps"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the native arm would run the full benchmark")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=120)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)


def test_committed_traffic_file_matches_the_default_workload():
    """bench.py fills roofline.traffic from profiles/traffic.json only when that capture was taken on the default
    workload; keep the two in step so the driver's bench line carries the number."""
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert tj["config"] == {"streams_per_gpu": 1024, "stream_kind": "edge", "batch_event_size": 200}
    for k in ("k_conv_rows", "k_conv_eval_tc", "k_leak_sweep", "k_pool_eval"):
        assert tj["dram_bytes_per_launch"][k] > 0 and tj["launches_per_step"][k] > 0
    assert "ncu --set full" in tj["source"]
