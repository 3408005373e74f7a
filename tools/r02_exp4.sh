#!/bin/bash
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ga3c.py -m gpu -q -x --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_ga3c.log 2>&1; tail -25 gpurun_out/${TAG}_pytest_ga3c.log
timeout 200 python tools/bench_ga3c.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_ga3c_bench.jsonl
timeout 200 python tools/bench_ga3c.py --mix 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ga3c_bench.jsonl
timeout 200 python tools/bench_ga3c.py --envs 4096 --agents 4 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ga3c_bench.jsonl
