#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider --timeout 100 \
  -k "agent_counts or ragged or anti or compile_time or schedule or circle or teacher" > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 100 python tools/sweep.py --agents 12,16 --envs 4096,16384 --steps 6 2>/dev/null | grep "^{" > gpurun_out/${TAG}_sweep.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_sweep.jsonl"):
    d=json.loads(l); print("N %d E %d ms/step %.4f policy_us %.1f pairwise_us %.1f" % (d["N"], d["E"], d["ms_per_step"], d["policy_move_kernel_ms"]*1e3, d["pairwise_obs_kernel_ms"]*1e3))
PY
