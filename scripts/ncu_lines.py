#!/usr/bin/env python3
"""Attribute an ncu source page (per-SASS-instruction samples and executed counts) to CUDA source
lines, using nvdisasm --print-line-info on the object the kernel was built from (compile with
-lineinfo).  Usage: python scripts/ncu_lines.py REP.ncu-rep OBJECT.o MANGLED_KERNEL [top_n]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = dis.split(f".text.{kern}:")[1]
line_of = {}
cur = None
inl = None
for l in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
    if l.startswith("//-----") or ".section" in l:
        if line_of:
            break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg = defaultdict(lambda: [0, 0])
for n, r in enumerate(data):
    key = line_of.get(16 * n, ("?", 0))
    agg[key][0] += int(r[ix["# Samples"]] or 0)
    agg[key][1] += int(r[ix["Instructions Executed"]] or 0)
tot_s = sum(v[0] for v in agg.values()); tot_n = sum(v[1] for v in agg.values())
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    text = ""
    for cand in (os.path.join(os.path.dirname(obj), "..", "csrc", f), f):
        if os.path.exists(cand):
            if cand not in src:
                src[cand] = open(cand).read().splitlines()
            if 0 < ln <= len(src[cand]):
                text = src[cand][ln - 1].strip()
            break
    print(f"{100.0 * v[0] / tot_s:5.1f}% samples {100.0 * v[1] / tot_n:5.1f}% instr  {f}:{ln}  {text[:100]}")
