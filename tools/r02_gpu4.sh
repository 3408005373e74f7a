#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vec_env.py tests/test_gpu_scale.py tests/test_gpu_abi_raw.py -q --tb=short -p no:cacheprovider --timeout 150 -k "vec_env or full_size or laserscan or abi or stats" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for sp in 1 0; do
  echo "== CAV_STEP_SPLIT=$sp"
  CAV_STEP_SPLIT=$sp timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'], 'launches', d['gpu_launches'])"
done
timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench
quick_bench.run(1024, 20, 'f64', 0, 10, laser=True)" 2>&1 | grep "^E=" | cut -c1-170
