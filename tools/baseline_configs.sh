#!/bin/bash
# One bench line per BASELINE.json config that fits one GPU (configs[1..3]) + the headline; writes gpurun_out/configs.jsonl
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
run() { python bench.py --no-cpu-baseline --steps 100 --warmup 10 "$@" 2>/dev/null | tail -1 >> gpurun_out/configs.jsonl; }
run                                            # headline: 16384 x 10 RVO f64
run --envs 4096                                # configs[1]: 4096 x 10 RVO
run --envs 32768 --policy noncoop              # configs[2]: 32768 x 10 NonCooperative + sorted-neighbour obs
run --envs 1024 --agents 20 --laser --history 0   # configs[3] per-GPU share: 1024 x 20 RVO + LaserScan 512
run --precision f32                            # float32 layout
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    d=json.loads(l); c=d['config']
    print("%-70s %s value=%.3e ms/step=%.4f e2e=%.3e roofline=%.3f" % (c['workload'][:70], d['dtype'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
PY
