#!/bin/bash
# Developer script: the evidence run committed under profiles/ - GPU test suite with parity counters, the full bench line, the ncu
# launch list of a short bench run and one `--set full` capture of the HBM / tensor kernels of one steady-state step.
tag=${1:-final}
mkdir -p gpurun_out
AEC_PARITY_LOG=gpurun_out/${tag}_parity.jsonl timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
Q="--steps 2 --warmup 3 --no-cpu-baseline --no-legs --sustained-seconds 0 --latency-steps 0"
python bench.py $Q > gpurun_out/${tag}_quick.json 2> gpurun_out/${tag}_quick.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_full.csv python bench.py $Q > gpurun_out/${tag}_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_conv_rows|k_conv_eval_tc|k_pool_eval|k_leak_sweep|k_sweep_windows' -s 2130 -c 26 -o gpurun_out/${tag}_full -f python bench.py $Q > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2> gpurun_out/${tag}_ncu3.log
ls -la gpurun_out/${tag}_*
