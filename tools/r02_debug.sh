#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest "tests/test_gpu_scale.py::test_full_size_properties" -x -q -k "RVO" -p no:cacheprovider > gpurun_out/dbg_pytest2.log 2>&1
echo "coop+syncwarp rc=$?"; tail -3 gpurun_out/dbg_pytest2.log | cut -c1-220
timeout 200 python -m pytest tests/test_gpu_scale.py -x -q -k "full_size or anti_coll or across_agent" -p no:cacheprovider > gpurun_out/dbg_pytest3.log 2>&1
echo "scale rc=$?"; tail -3 gpurun_out/dbg_pytest3.log | cut -c1-220
timeout 60 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('f64', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'e2e %.2e'%d['e2e']['value'])"
cp gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so gpurun_out/orig_lib.so; cp gpurun_libs/nocoop.so gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
timeout 60 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('nocoop f64', '%.3e'%d['value'], 'step_us', round(d['ms_per_step']*1e3,1), 'pairwise_us', round(r['kernel_ms']*1e3,1), 'policy_us', round(r['policy_move_kernel_ms']*1e3,1), 'e2e %.2e'%d['e2e']['value'])"
