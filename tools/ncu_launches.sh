#!/bin/bash
# ncu launch list (per-launch gpu__time_duration) of the default bench command; run under gpurun on one GPU
TAG=${1:-launches}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${TAG}.log 2>&1
tail -1 gpurun_out/${TAG}.log | cut -c1-200
