#!/bin/bash
# multi-GPU evidence (run under gpurun --gpus N): headline weak scaling, BASELINE configs[3] on all GPUs, a sweep subset
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1 > gpurun_out/${TAG}_bench_${N}gpu.json
if [ -n "$CAV_MULTI_FULL" ]; then
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-numa 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_${N}gpu_nonuma.json
fi
if [ "$N" = "8" ]; then
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --envs 1024 --agents 20 --laser --history 0 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c4_${N}gpu.json
  [ -n "$CAV_MULTI_FULL" ] && timeout 400 $TR tools/sweep.py --agents 2,8,32 --envs 4096,32768 --steps 8 2>/dev/null | grep '^{' > gpurun_out/${TAG}_sweep_${N}gpu.jsonl
fi
python - <<PY
import json
for n in ("bench_${N}gpu","bench_${N}gpu_nonuma","bench_c4_${N}gpu"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%n))
        print("%-22s n_gpus %d value %.3e ms/step %.4f e2e %.3e (%.3f ms) per-rank %s e2e-per-rank %s numa %s" % (n, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], [round(x,4) for x in d["per_rank_ms_per_step"]], [round(x,3) for x in d["e2e"].get("per_rank_ms_per_step", [])], d.get("numa_binding_per_rank")))
    except Exception as e:
        print(n, "missing", e)
PY
cat gpurun_out/${TAG}_sweep_${N}gpu.jsonl 2>/dev/null | cut -c1-230
head -14 gpurun_out/${TAG}_topo_${N}gpu.txt
