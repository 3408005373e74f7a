#!/bin/bash
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ga3c.py tests/test_gpu_scale.py -m gpu -q -x --tb=short -p no:cacheprovider --timeout 200 -k "ga3c or schedule" > gpurun_out/${TAG}_pytest_ga3c.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_ga3c.log
: > gpurun_out/${TAG}_ga3c_bench.jsonl
for args in "" "--mix" "--envs 4096 --agents 4" "--steps 200"; do timeout 200 python tools/bench_ga3c.py $args 2>/dev/null | tail -1 | tee -a gpurun_out/${TAG}_ga3c_bench.jsonl; done
