#!/usr/bin/env python
"""tools/ncu_trim.py REPORT.ncu-rep OUT.csv -- the raw page of an ncu report cut down to the metrics the design discussion uses
(one row per profiled launch): what gets committed under profiles/ instead of the multi-megabyte report."""
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_atom.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
stalls = [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
cols = [n for n in KEEP + stalls if n in h]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([rows[1][h.index(c)] for c in cols])
    for r in rows[2:]:
        w.writerow([r[h.index(c)] for c in cols])
print("%d launches, %d metrics -> %s" % (len(rows) - 2, len(cols), sys.argv[2]))
