#!/bin/bash
# experiment (one GPU): environment grouping / ordering potential of the policy kernel, and the number of step ranges
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 200 python tools/exp_grouping.py > gpurun_out/${TAG}_grouping.txt 2>&1; cat gpurun_out/${TAG}_grouping.txt | tail -8
for P in 2 3 4 0; do
  CAV_STEP_SPLIT=$P timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/${TAG}_split$P.json
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_split$P.json")); r=d["roofline"]
print("split $P: value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f e2e %.3e" % (d["value"], d["ms_per_step"], r["policy_move_kernel_ms"]*1e3, r["kernel_ms"]*1e3, d["e2e"]["value"]))
PY
done
