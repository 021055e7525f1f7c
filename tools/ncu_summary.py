#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): key roofline/occupancy/stall metrics of the first
kernel in the report, optionally the hottest SASS lines.  Usage: ncu_summary.py rep [--source N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_read.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
for r in rows[2:3]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print("%-75s %-12s %s" % (k, u[k], d[k]))
    st = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k]]
    for v, k in sorted(st, reverse=True)[:8]:
        print("  stall %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    cs, ci, csrc = col("# Samples") or col("Samples"), col("Instructions Executed"), col("Source")
    print(h[:12])
    body = [r for r in rows[1:] if len(r) == len(h)]
    tot = sum(float(r[cs] or 0) for r in body)
    top = sorted(body, key=lambda r: -float(r[cs] or 0))[:n]
    for r in top:
        print("%6.2f%%  %s" % (100 * float(r[cs] or 0) / max(tot, 1), r[csrc][:120]))
