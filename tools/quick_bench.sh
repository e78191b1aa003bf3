#!/bin/bash
# usage: quick.sh tag [ENV=val ...]
tag=$1; shift
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs --sustained-seconds 0 --latency-steps 0 > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
