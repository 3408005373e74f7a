#!/bin/bash
# BASELINE configs[4]: agent-count x env-count sweep on one GPU (run under gpurun)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python tools/sweep.py --steps 8 2>gpurun_out/${TAG}_sweep_1gpu.err | grep '^{' > gpurun_out/${TAG}_sweep_1gpu.jsonl
python - <<PY
import json
print("%3s %7s %10s %12s %9s %9s %7s" % ("N","E","ms/step","agent-st/s","policy_us","pairw_us","frac"))
for l in open("gpurun_out/${TAG}_sweep_1gpu.jsonl"):
    d=json.loads(l)
    print("%3d %7d %10.4f %12.3e %9.1f %9.1f %7.3f" % (d["N"], d["E"], d["ms_per_step"], d["agent_steps_per_s"], d["policy_move_kernel_ms"]*1e3, d["pairwise_obs_kernel_ms"]*1e3, d["pairwise_obs_roofline_frac"]))
PY
