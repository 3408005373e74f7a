#!/bin/bash
# One `ncu --set full` capture of the two step kernels (run under gpurun, one GPU).  usage: tools/ncu_capture.sh f64|f32 TAG
PREC=${1:-f64}; TAG=${2:-cap}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pairwise_obs|k_policy_move" \
    --launch-skip 16 --launch-count 2 -f -o gpurun_out/${TAG}_${PREC} \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision ${PREC} > gpurun_out/${TAG}_${PREC}.log 2>&1
tail -2 gpurun_out/${TAG}_${PREC}.log | cut -c1-300
