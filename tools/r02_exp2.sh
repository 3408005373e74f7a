#!/bin/bash
# experiment (one GPU): environment schedule on / off
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_vec_env.py -m gpu -q -x --tb=short -p no:cacheprovider --timeout 200 \
   -k "schedule or full_size or vec or lookahead or graph" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
for O in 1 0; do
  CAV_ENV_ORDER=$O timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/${TAG}_order$O.json
  CAV_ENV_ORDER=$O timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${TAG}_order${O}_200.json
  python - <<PY
import json
for f in ("gpurun_out/${TAG}_order$O.json", "gpurun_out/${TAG}_order${O}_200.json"):
    d=json.load(open(f)); r=d["roofline"]
    print("order $O: value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f e2e %.3e live %.2f launches %s" % (d["value"], d["ms_per_step"], r["policy_move_kernel_ms"]*1e3, r["kernel_ms"]*1e3, d["e2e"]["value"], d.get("live_fraction", 0), d.get("gpu_launches")))
PY
done
