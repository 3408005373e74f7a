#!/bin/bash
# kernel time vs env count (tail-effect probe); prints: envs pairwise_obs_us policy_move_us step_us roofline_frac
for e in "$@"; do
  python bench.py --no-cpu-baseline --envs $e --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print($e, round(d['roofline']['kernel_ms']*1e3,1), round(d['roofline']['policy_move_kernel_ms']*1e3,1), round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))"
done
