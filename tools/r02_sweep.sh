#!/bin/bash
# round-2 tuning sweep (run under gpurun, one GPU): parity suite on the default build, then A/B bench lines over gpurun_libs/*.so
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python tools/quick_bench.py --steps 30 > gpurun_out/${TAG}_quick.log 2>&1; grep -E "^E=" gpurun_out/${TAG}_quick.log | cut -c1-200
bash tools/libsweep.sh gym_collision_avoidance_b200/csrc/libcavoid_sm100a.so gpurun_libs/*.so 2>&1 | tee gpurun_out/${TAG}_libsweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_laserscan" \
    --launch-skip 4 --launch-count 1 -f -o gpurun_out/${TAG}_laser \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --envs 1024 --agents 20 --laser > gpurun_out/${TAG}_laser.log 2>&1
tail -2 gpurun_out/${TAG}_laser.log | cut -c1-200
