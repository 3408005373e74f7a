#!/bin/bash
# round-2 GPU check #3: parity suite (per-test timeout), bench line, A/B libs, laser timing
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 150 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for lib in gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so gpurun_libs/*.so; do
  cp gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so gpurun_out/orig_lib.so 2>/dev/null
  [ "$lib" != "gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so" ] && cp $lib gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
  echo "== $lib"
  timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'frac', round(r['frac'],3), 'e2e %.2e'%d['e2e']['value'])"
  cp gpurun_out/orig_lib.so gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
done
timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench
quick_bench.run(1024, 20, 'f64', 0, 10, laser=True)
quick_bench.run(16384, 10, 'f64', 0, 20)" 2>&1 | grep "^E=" | cut -c1-170
