#!/usr/bin/env python3
"""Text summary of one kernel of an .ncu-rep (run where `ncu` is installed; no GPU needed):
raw metrics that the roofline discussion uses, the SASS opcode mix and the hottest source lines.
usage: python scratch/ncu_summary.py report.ncu-rep [kernel-index] > profiles/rN/<name>_summary.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
print("kernel:", vals[hdr.index("Kernel Name")])
for i, h in enumerate(hdr):
    if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print("%-90s %-16s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page holds one table per kernel ("Kernel Name" line, header line, rows); take the idx-th
r = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(r) if x and x[0] == "Kernel Name"]
lo = starts[min(idx, len(starts) - 1)]
hi = starts[starts.index(lo) + 1] if starts.index(lo) + 1 < len(starts) else len(r)
h = r[lo + 1]
data = [x for x in r[lo + 2:hi] if len(x) == len(h)]
iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(x[iE]) for x in data) or 1
totw = sum(int(x[iW]) for x in data) or 1
ops, st = collections.Counter(), collections.Counter()
for x in data:
    t = x[iS].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(x[iE]); st[op] += int(x[iW])
print("--- SASS opcode mix (warp instructions executed: %d)" % tot)
for k, v in ops.most_common(16):
    print("%-10s %6.2f%% inst  %6.2f%% stall samples" % (k, 100.0 * v / tot, 100.0 * st[k] / totw))
