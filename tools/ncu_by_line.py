#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv` SASS dump per CUDA source line, using the line table
nvdisasm prints for the same kernel.  Usage:
  ncu_by_line.py <src.csv> <cubin> <kernel-substring> [top]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.check_output(["nvdisasm", "--print-line-info", cubin]).decode(errors="replace").splitlines()
# find the function section
start = None
for i, l in enumerate(dis):
    if l.startswith(".text.") and kname in l:
        start = i
        break
    if re.match(r"\s*\.section\s+\.text\..*" + re.escape(kname), l):
        start = i
        break
assert start is not None, "kernel not found"
line_of = {}
cur = ("?", 0)
for l in dis[start + 1:]:
    if re.match(r"\s*\.section\s+\.text\.", l) and kname not in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: [0, 0, 0])
ops = defaultdict(int)
tot_inst = tot_samp = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    off = a - base
    inst = int(float(r[col["Instructions Executed"]] or 0))
    samp = int(float(r[col["# Samples"]] or 0))
    (ln, sass) = line_of.get(off, (("?", 0), r[1]))
    agg[ln][0] += inst
    agg[ln][1] += samp
    agg[ln][2] += 1
    ops[r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]] += inst
    tot_inst += inst
    tot_samp += samp
print("total warp-instructions %d, samples %d, SASS instructions %d" % (tot_inst, tot_samp, sum(v[2] for v in agg.values())))
print("--- by source line (sorted by stall samples) ---")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s line %-5d inst %10d (%5.1f%%)  samples %6d (%5.1f%%)  sass %d" % (
        ln[0], ln[1], v[0], 100.0 * v[0] / max(1, tot_inst), v[1], 100.0 * v[1] / max(1, tot_samp), v[2]))
print("--- by opcode ---")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print("%-16s %10d (%5.1f%%)" % (k, v, 100.0 * v / max(1, tot_inst)))
