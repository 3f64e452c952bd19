#!/usr/bin/env python
"""Inclusive cost per source STATEMENT of one function: attributes every SASS instruction of an
`ncu --page source --csv` dump to the call-site line inside [lo, hi] of <file> found in its inline
chain (nvdisasm --print-line-info-inline).  Usage:
  ncu_by_stmt.py <src.csv> <cubin> <kernel-substring> <file> <lo> <hi> [top]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
top = int(sys.argv[7]) if len(sys.argv) > 7 else 50
dis = subprocess.check_output(["nvdisasm", "--print-line-info-inline", cubin]).decode(errors="replace").splitlines()
start = [i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\..*" + re.escape(kname), l)][0]
chain, frames_of = [], {}
for l in dis[start + 1:]:
    if re.match(r"\s*\.section\s+\.text\.", l) and kname not in l:
        break
    if "//## File" in l:
        m = re.search(r'File "([^"]+)", line (\d+)', l)
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        if "inlined at" not in l and chain and False:
            pass
        chain.append(fr)
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        if chain:
            last = chain
            frames_of[int(m.group(1), 16)] = list(chain)
            chain = []
        else:
            frames_of[int(m.group(1), 16)] = last
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = base or a
    inst = int(float(r[col["Instructions Executed"]] or 0))
    samp = int(float(r[col["# Samples"]] or 0))
    frames = frames_of.get(a - base, [("?", 0)])
    key = None
    for f, n in frames:
        if f == fname and lo <= n <= hi:
            key = n
    if key is None:
        key = "outside:%s:%d" % frames[-1] if frames else "?"
    agg[key][0] += inst
    agg[key][1] += samp
    tot_i += inst
    tot_s += samp
lines = open([p for p in [fname, "raytracing_rb_b200/csrc/" + fname] if __import__("os").path.exists(p)][0]).read().splitlines()
print("total inst %d samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = lines[k - 1].strip()[:70] if isinstance(k, int) else ""
    print("%-34s inst %9d (%5.1f%%) samples %5d (%5.1f%%)  %s" % (k, v[0], 100.0 * v[0] / tot_i, v[1], 100.0 * v[1] / max(1, tot_s), text))
