#!/bin/bash
# Round-end evidence on one GPU: tests, smoke, bench (both arms), per-config lines, ncu launch list + full captures.
TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_f64.json
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_reference_arm.json
python bench.py --precision f32 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_f32.json
tools/baseline_configs.sh > gpurun_out/${TAG}_configs.txt 2>&1; cp gpurun_out/configs.jsonl gpurun_out/${TAG}_baseline_configs.jsonl
tools/ncu_launches.sh ${TAG}_launches > /dev/null
tools/ncu_capture.sh f64 ${TAG} > /dev/null
tools/ncu_capture.sh f32 ${TAG} > /dev/null
python - <<PY
import json
for n in ("f64","f32","reference_arm"):
    d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%n))
    print(n, "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("clocks"))
PY
cat gpurun_out/${TAG}_configs.txt | tail -6
