#!/bin/bash
# full GPU suite + smoke + driver bench on the current build
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_f64_driver.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_f64_driver.json")); r=d["roofline"]
print("value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f frac %.3f e2e %.3e launches %s cpu %s" % (d["value"], d["ms_per_step"], r["policy_move_kernel_ms"]*1e3, r["kernel_ms"]*1e3, r["frac"], d["e2e"]["value"], d.get("gpu_launches"), d.get("cpu_baseline",{}).get("value")))
PY
