#!/bin/bash
# A/B of the schedule key (gpurun_libs/sched_key*.so) on the bench workload, then the GA3C GPU tests on the default build
TAG=${1:-r02n}
mkdir -p gpurun_out
orig=gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
cp $orig gpurun_out/orig_lib.so
: > gpurun_out/${TAG}_sched_key_ab.txt
for rep in 1 2; do
for lib in gpurun_libs/sched_key*.so; do
  cp $lib $orig
  for extra in "--steps 20 --warmup 5" "--steps 200 --warmup 20"; do
  python bench.py --no-cpu-baseline $extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib $extra: value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f' % (d['value'], d['ms_per_step'], r['policy_move_kernel_ms']*1e3, r['kernel_ms']*1e3))" | tee -a gpurun_out/${TAG}_sched_key_ab.txt
  done
done
done
cp gpurun_out/orig_lib.so $orig
timeout 600 python -m pytest tests/test_gpu_ga3c.py -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_ga3c.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_ga3c.log
