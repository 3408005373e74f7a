#!/bin/bash
# tools/ncu_report.sh REP.ncu-rep  -> prints the tracked metrics, stall mix, and per-source-line attribution of each kernel
REP=$1; B=${REP%.ncu-rep}
ncu -i $REP --page raw --csv > $B.csv 2>/dev/null
ncu -i $REP --page source --csv --print-source sass > ${B}_sass.csv 2>/dev/null
python - "$B" <<'PY'
import csv, sys
B = sys.argv[1]
rows = list(csv.reader(open(B + ".csv")))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:70], d.get("Grid Size"), d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print("   %-62s %s" % (k, d[k]))
    st = [(float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and d[h] not in ("", "n/a")]
    st.sort(reverse=True)
    print("   stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
PY
