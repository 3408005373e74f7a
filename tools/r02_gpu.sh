#!/bin/bash
# round-2 GPU check (run under gpurun, one GPU): parity suite, bench lines, ncu launch list + full capture of the step kernels
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 100 2>gpurun_out/${TAG}_bench_f64.err | tail -1 > gpurun_out/${TAG}_bench_f64.json
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_f64.json')); r=d['roofline']
print('f64', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pairwise_obs|k_policy_move" \
    --launch-skip 16 --launch-count 2 -f -o gpurun_out/${TAG}_f64 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision f64 > gpurun_out/${TAG}_f64.log 2>&1
tail -2 gpurun_out/${TAG}_f64.log | cut -c1-200
