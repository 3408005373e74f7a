#!/bin/bash
# full GPU parity suite + the two headline bench lines (run under gpurun)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for pr in f64 f32; do python bench.py --no-cpu-baseline --precision $pr --steps 100 2>/dev/null | tail -1 > gpurun_out/bench_$pr.json; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$pr.json'))
r=d['roofline']
print('$pr', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'])"; done
