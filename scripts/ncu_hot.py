#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples, and the opcode histogram, from an .ncu-rep
(needs --import-source on / -lineinfo).  Usage: python scripts/ncu_hot.py REP [kernel-regex] [N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["-k", "regex:" + sys.argv[2]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
si, ei, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
body = []
for r in rows[h + 1:]:
    if len(r) <= max(si, ei, wi):
        break
    try:
        body.append((int(r[wi].replace(",", "") or 0), int(r[ei].replace(",", "") or 0), r[si]))
    except ValueError:
        break
tot_s = sum(b[0] for b in body) or 1
tot_i = sum(b[1] for b in body) or 1
print(f"{len(body)} SASS lines, {tot_i} warp instructions, {tot_s} stall samples")
print("--- top by stall samples (share, executed, SASS)")
for s_, e_, src in sorted(body, reverse=True)[:n]:
    print(f"{100 * s_ / tot_s:6.2f}%  {e_:12d}  {src[:90]}")
agg = collections.Counter()
for s_, e_, src in body:
    op = src.split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    agg[o.split(".")[0]] += e_
print("--- opcode histogram")
for op, c in agg.most_common(18):
    print(f"{100 * c / tot_i:6.2f}%  {op}")
