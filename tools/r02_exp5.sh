#!/bin/bash
# GA3C kernel variants (gpurun_libs/ga_*.so): parity of the forward tests + kernel time; then the GA3C GPU tests on the default build
TAG=${1:-r02m}
mkdir -p gpurun_out
orig=gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so
cp $orig gpurun_out/orig_lib.so
: > gpurun_out/${TAG}_ga3c_variants.txt
for lib in gpurun_libs/ga_*.so; do
  cp $lib $orig
  echo "== $lib" | tee -a gpurun_out/${TAG}_ga3c_variants.txt
  timeout 300 python -m pytest tests/test_gpu_ga3c.py -m gpu -q -x --tb=line -p no:cacheprovider --timeout 200 -k "forward or predict" 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ga3c_variants.txt
  timeout 200 python tools/bench_ga3c.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ga3c_variants.txt
done
cp gpurun_out/orig_lib.so $orig
timeout 600 python -m pytest tests/test_gpu_ga3c.py -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_ga3c.log 2>&1; tail -25 gpurun_out/${TAG}_pytest_ga3c.log
