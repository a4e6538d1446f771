#!/usr/bin/env python3
"""List the loops (backward branches) of a cuobjdump -sass dump with an opcode histogram per body.
Usage: cuobjdump -sass -fun <mangled> file.o | python scripts/sass_loops.py [max_body]"""
import re
import sys
from collections import Counter

max_body = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ins = []
for l in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr:
        body = [x[1] for x in ins[addr[tgt]:i + 1]]
        if len(body) <= max_body:
            c = Counter(re.sub(r'^@!?U?P\d+\s+', '', b).split()[0].split('.')[0] for b in body)
            print(f"loop 0x{tgt:x}..0x{a:x}: {len(body)} instr  " + " ".join(f"{k}:{v}" for k, v in c.most_common()))
