#!/bin/bash
# round-2 GPU check #2: parity suite, bench line, ncu of the three step kernels at 16384x10 and of the laserscan at 1024x20
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_f64.err | tail -1 > gpurun_out/${TAG}_bench_f64.json
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_f64.json')); r=d['roofline']
print('f64', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'], 'live', round(d['live_fraction'],3))
PY
python tools/quick_bench.py --steps 20 > gpurun_out/${TAG}_quick.log 2>&1; grep -E "^E=" gpurun_out/${TAG}_quick.log | cut -c1-170
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pairwise_obs|k_policy_move" \
    --launch-skip 16 --launch-count 2 -f -o gpurun_out/${TAG}_f64 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision f64 > gpurun_out/${TAG}_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_laserscan|k_policy_move" \
    --launch-skip 6 --launch-count 4 -f -o gpurun_out/${TAG}_laser \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --envs 1024 --agents 20 --laser > gpurun_out/${TAG}_laser.log 2>&1
tail -1 gpurun_out/${TAG}_laser.log | cut -c1-200
