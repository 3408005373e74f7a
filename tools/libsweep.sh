#!/bin/bash
# A/B experiment helper: run bench.py with alternative builds of the library (gpurun_libs/*.so); $PREC = f64|f32
orig=gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
mkdir -p gpurun_out; cp $orig gpurun_out/orig_lib.so
for lib in "$@"; do
  [ "$lib" != "$orig" ] && cp $lib $orig
  echo "== $lib"
  python bench.py --no-cpu-baseline --steps 50 --warmup 10 --precision ${PREC:-f64} 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%.3e'%d['value'], round(d['ms_per_step']*1e3,1), round(d['roofline']['kernel_ms']*1e3,1), round(d['roofline']['policy_move_kernel_ms']*1e3,1), round(d['roofline']['frac'],3))"
done
cp gpurun_out/orig_lib.so $orig
