#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` capture for profiles/: headline metrics, stall
reasons per issued instruction, stall samples and executed instructions by opcode, hottest SASS instructions.
Usage: ncu_summary.py <file.ncu-rep> [header text]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
for h, u, v in zip(hdr, units, vals):
    if h == "Kernel Name":
        print("kernel: " + v)
    if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print("%-92s %-16s %s" % (h, u, v))
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]


def opcode(text):
    t = text.strip().split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    return op.split(".")[0]


print("\n# stall samples by reason -> opcode of the stalled instruction")
for key in ("stall_wait", "stall_long_sb", "stall_no_inst", "stall_not_selected", "stall_short_sb", "stall_branch_resolving", "stall_math"):
    c, tot = collections.Counter(), 0
    for r in body:
        v = int(float(r[col[key]] or 0))
        tot += v
        c[opcode(r[1])] += v
    print("%-24s %7d  %s" % (key, tot, " ".join("%s:%d" % kv for kv in c.most_common(10))))
c, tot = collections.Counter(), 0
for r in body:
    v = int(float(r[col["Instructions Executed"]] or 0))
    tot += v
    c[opcode(r[1])] += v
print("\n# executed warp instructions by opcode (total %d, %d SASS instructions in the kernel)" % (tot, len(body)))
print(" ".join("%s:%.1f%%" % (k, 100.0 * v / tot) for k, v in c.most_common(28)))
print("\n# hottest SASS instructions by stall samples (samples, executed, instruction)")
for r in sorted(body, key=lambda r: -int(float(r[col["# Samples"]] or 0)))[:25]:
    print("%6d %10d  %s" % (int(float(r[col["# Samples"]] or 0)), int(float(r[col["Instructions Executed"]] or 0)), r[1].strip()[:90]))
