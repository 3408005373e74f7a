#!/bin/bash
# `ncu --set full` capture of the laserscan kernel on the BASELINE configs[3] per-GPU share (1024 envs x 20 agents)
PREC=${1:-f64}; TAG=${2:-laser}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_laserscan" \
    --launch-skip 10 --launch-count 2 -f -o gpurun_out/${TAG}_${PREC} \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision ${PREC} --envs 1024 --agents 20 --laser --history 0 > gpurun_out/${TAG}_${PREC}.log 2>&1
tail -1 gpurun_out/${TAG}_${PREC}.log | cut -c1-200
