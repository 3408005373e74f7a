#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output into the handful of metrics we track in profiles/."""
import csv
import json
import sys

KEYS = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks', 'launch__waves_per_multiprocessor',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max']


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        o = {}
        for k in KEYS:
            if k in d:
                o[k] = d[k] + (" " + units[hdr.index(k)] if units[hdr.index(k)] else "")
        for h in hdr:
            if h.startswith("smsp__warp_issue_stalled") and h.endswith("per_warp_active.pct"):
                try:
                    v = float(d[h])
                except ValueError:
                    continue
                if v >= 3.0:
                    o[h.replace("smsp__warp_issue_stalled_", "stall_").replace("_per_warp_active.pct", "")] = v
        res.append(o)
    txt = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
