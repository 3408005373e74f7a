#!/bin/bash
TAG=${1:-r02r}
bash tools/r02_exp8.sh $TAG
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu.log
for extra in "--precision f32" "--obs-f64" "--envs 32768 --policy noncoop" "--envs 4096"; do
python bench.py --no-cpu-baseline --steps 20 --warmup 5 $extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$extra: value %.3e ms/step %.4f policy_us %.1f pairwise_us %.1f frac %.3f' % (d['value'], d['ms_per_step'], r['policy_move_kernel_ms']*1e3, r['kernel_ms']*1e3, r['frac']))" | tee -a gpurun_out/${TAG}_ab.txt
done
