#!/usr/bin/env python
"""Print the headline counters and stall breakdown of an .ncu-rep (first profiled launch)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {name: (v[i], u[i]) for i, name in enumerate(h)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "smsp__inst_executed.sum"]
for k in keys:
    for name in d:
        if name.endswith(k):
            print(f"{k:90s} {d[name][0]:>18s} {d[name][1]}")
            break
st = []
for name, (val, unit) in d.items():
    if "issue_stalled" in name and name.endswith("per_issue_active.ratio"):
        try:
            st.append((float(val.replace(",", "")), name.split("issue_stalled_")[1].split("_per_issue")[0]))
        except ValueError:
            pass
print("stalls per issue:", ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
