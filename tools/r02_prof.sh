#!/bin/bash
# round-2 profiles of the current build (run under gpurun, one GPU)
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_f64.err | tail -1 > gpurun_out/${TAG}_bench_f64.json
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_f64.json')); r=d['roofline']
print('f64', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'], 'live', round(d['live_fraction'],3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_pairwise_obs|k_policy_move" \
    --launch-skip 16 --launch-count 2 -f -o gpurun_out/${TAG}_f64 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision f64 > gpurun_out/${TAG}_f64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_laserscan" \
    --launch-skip 6 --launch-count 3 -f -o gpurun_out/${TAG}_laser \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --envs 1024 --agents 20 --laser > gpurun_out/${TAG}_laser.log 2>&1
tail -1 gpurun_out/${TAG}_laser.log | cut -c1-200
CAV_POLICY_GROUP=1 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench
quick_bench.run(1024, 20, 'f64', 0, 10, laser=True)" 2>&1 | grep "^E=" | cut -c1-170
CAV_TPA_MAX_N=0 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench
quick_bench.run(1024, 20, 'f64', 0, 10, laser=True)" 2>&1 | grep "^E=" | cut -c1-170
