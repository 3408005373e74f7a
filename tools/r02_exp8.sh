#!/bin/bash
# A/B of library variants in gpurun_libs/*.so on the bench workload (driver settings), twice each
TAG=${1:-r02q}
mkdir -p gpurun_out
orig=gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
cp $orig gpurun_out/orig_lib.so
: > gpurun_out/${TAG}_ab.txt
for rep in 1 2; do
for lib in gpurun_libs/*.so; do
  cp $lib $orig
  python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib: value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f frac %.3f e2e %.3e' % (d['value'], d['ms_per_step'], r['policy_move_kernel_ms']*1e3, r['kernel_ms']*1e3, r['frac'], d['e2e']['value']))" | tee -a gpurun_out/${TAG}_ab.txt
done
done
cp gpurun_out/orig_lib.so $orig
