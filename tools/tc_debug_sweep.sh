#!/bin/bash
# Developer experiment: per-layer conv_eval time with parts of the tcgen05 kernel disabled
# (AEC_TC_DEBUG bits; results are invalid, only the timings mean something).
#   1 no gather loads (every tap reads the zero line)   2 no operand stores   4 no MMA   8 no epilogue stores   16 no weight copies
mkdir -p gpurun_out
for m in ${MODES:-0 1 2 3 4 8 16 7 15 31}; do
  AEC_TC_DEBUG=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-legs --sustained-seconds 0 --latency-steps 0 ${BENCH_ARGS} > gpurun_out/dbg_$m.json 2> gpurun_out/dbg_$m.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dbg_$m.json"))
    r=d["roofline"]; l=r["ms_by_launch"]
    print("mode %2d step %.3f ms | "%($m, d["ms_per_step"]) + " ".join("%s %.3f"%(k.split('.')[0],v) for k,v in l.items() if k.endswith(".eval") and "conv" in k) + " | sites " + " ".join("%.0f"%v for k,v in r["sites_per_step_per_stream"].items() if "conv" in k))
except Exception as e:
    print("mode $m failed", e)
PY
done
