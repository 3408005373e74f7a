#!/bin/bash
# Round-end evidence on one GPU: parity suite, smoke, bench lines (both arms + variants), per-config lines, ncu launch list
# and full captures of the step kernels and the laserscan.   usage (under gpurun): bash tools/final_check.sh r02h
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
b() { out=$1; shift; timeout 300 python bench.py "$@" 2>gpurun_out/${TAG}_$out.err | tail -1 > gpurun_out/${TAG}_$out.json; }
b bench_f64_driver --steps 20 --warmup 5                 # what the driver runs
b bench_reference_arm --impl reference --steps 20 --warmup 5
b bench_f64 --no-cpu-baseline                            # default K / W (200 / 20)
b bench_f64_steady --no-cpu-baseline --preroll 150 --steps 50 --warmup 5
b bench_f64_obsf64 --no-cpu-baseline --obs-f64 --steps 20 --warmup 5
b bench_f64_nolookahead --no-cpu-baseline --no-lookahead --steps 20 --warmup 5
b bench_f32 --no-cpu-baseline --precision f32 --steps 20 --warmup 5
: > gpurun_out/${TAG}_baseline_configs.jsonl
c() { timeout 300 python bench.py --no-cpu-baseline --steps 50 --warmup 10 "$@" 2>/dev/null | tail -1 >> gpurun_out/${TAG}_baseline_configs.jsonl; }
c --envs 4096                                # configs[1]: 4096 x 10 RVO
c --envs 32768 --policy noncoop              # configs[2]: 32768 x 10 NonCooperative + sorted-neighbour obs
c --envs 1024 --agents 20 --laser --history 0   # configs[3] per-GPU share: 1024 x 20 RVO + LaserScan 512
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_pairwise_obs|k_policy_move" \
    --launch-skip 32 --launch-count 4 -f -o gpurun_out/${TAG}_f64 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --precision f64 > gpurun_out/${TAG}_f64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_laserscan" \
    --launch-skip 6 --launch-count 3 -f -o gpurun_out/${TAG}_laser \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --envs 1024 --agents 20 --laser --history 0 > gpurun_out/${TAG}_laser.log 2>&1
# GA3C-CADRL agents: device-timed step + the network kernel alone (fp32 FLOP rate), and its ncu capture
: > gpurun_out/${TAG}_ga3c_bench.jsonl
for args in "" "--mix" "--envs 4096 --agents 4"; do timeout 200 python tools/bench_ga3c.py $args 2>/dev/null | tail -1 >> gpurun_out/${TAG}_ga3c_bench.jsonl; done
cat gpurun_out/${TAG}_ga3c_bench.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_ga3c$" --launch-skip 8 --launch-count 1 -f -o gpurun_out/${TAG}_ga3c \
    python tools/bench_ga3c.py --steps 4 > gpurun_out/${TAG}_ga3c.log 2>&1
python - <<PY
import json
for n in ("bench_f64_driver","bench_reference_arm","bench_f64","bench_f64_steady","bench_f64_obsf64","bench_f64_nolookahead","bench_f32"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%n))
        r=d.get("roofline",{})
        print("%-24s value %.3e ms/step %.4f e2e %.3e frac %s pw_us %s pm_us %s live %s cpu %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac") and round(r["frac"],3), r.get("kernel_ms") and round(r["kernel_ms"]*1e3,1), r.get("policy_move_kernel_ms") and round(r["policy_move_kernel_ms"]*1e3,1), d.get("live_fraction") and round(d["live_fraction"],3), d.get("cpu_baseline",{}).get("value")))
    except Exception as e:
        print(n, "FAILED", e)
for l in open("gpurun_out/${TAG}_baseline_configs.jsonl"):
    d=json.loads(l); c=d['config']
    print("%-60s value=%.3e ms/step=%.4f e2e=%.3e frac=%.3f" % (c['workload'][:60], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
PY
