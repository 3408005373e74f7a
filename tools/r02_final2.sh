#!/bin/bash
# last verification of the final tree: full GPU suite, smoke, both bench arms with the driver's settings
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_reference_arm.json
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_f64_driver.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_f64_driver.json")); r=d["roofline"]
print("value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f frac %.3f e2e %.3e windows %s launches %s cpu %s clocks %s" % (d["value"], d["ms_per_step"], r["policy_move_kernel_ms"]*1e3, r["kernel_ms"]*1e3, r["frac"], d["e2e"]["value"], [round(x,3) for x in d["e2e"]["windows_ms_per_step"]], d.get("gpu_launches"), d.get("cpu_baseline",{}).get("value"), d.get("clocks")))
d=json.load(open("gpurun_out/${TAG}_bench_reference_arm.json")); print("reference arm", d.get("value"), d.get("impl"), d.get("cpu_baseline",{}).get("cores"))
PY
