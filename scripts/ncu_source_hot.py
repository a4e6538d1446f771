#!/usr/bin/env python3
"""Per-region view of an ncu source page (ncu -i REP --page source --csv): splits the SASS at
backward branches (loop bodies) and prints, per region, executed warp instructions, stall samples
and the opcode mix.  Usage: python scripts/ncu_source_hot.py REP.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
ins = []
for r in data:
    try:
        addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    except ValueError:
        continue
    ins.append((addr, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
base = ins[0][0]
pos = {a: i for i, (a, _, _, _) in enumerate(ins)}
loops = []
for i, (a, t, n, s) in enumerate(ins):
    m = re.search(r'BRA\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        for cand in (tgt, base + tgt):
            if cand in pos and pos[cand] <= i and i - pos[cand] < 400:
                loops.append((pos[cand], i))
                break
tot_n = sum(x[2] for x in ins)
tot_s = sum(x[3] for x in ins)
covered = [False] * len(ins)
print(f"total warp instructions {tot_n}, samples {tot_s}")
for lo, hi in loops:
    n = sum(x[2] for x in ins[lo:hi + 1]); s = sum(x[3] for x in ins[lo:hi + 1])
    if n < 0.002 * tot_n:
        continue
    for j in range(lo, hi + 1):
        covered[j] = True
    c = Counter(re.sub(r'^@!?U?P\d+\s+', '', x[1]).split()[0].split('.')[0] for x in ins[lo:hi + 1])
    trips = ins[hi][2]
    print(f"loop [{lo}:{hi}] {hi - lo + 1} instr, trips {trips}, {100.0 * n / tot_n:5.1f}% of instr, {100.0 * s / tot_s:5.1f}% of samples  "
          + " ".join(f"{k}:{v}" for k, v in c.most_common(8)))
n = sum(x[2] for x, cv in zip(ins, covered) if not cv); s = sum(x[3] for x, cv in zip(ins, covered) if not cv)
print(f"outside those loops: {100.0 * n / tot_n:5.1f}% of instr, {100.0 * s / tot_s:5.1f}% of samples")
if len(sys.argv) > 2 and sys.argv[2] == "--top":
    for x in sorted(ins, key=lambda x: -x[3])[:40]:
        print(f"  {x[3]:6d} samples {x[2]:10d} exec  {x[1][:90]}")
