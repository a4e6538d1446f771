#!/usr/bin/env python3
"""Print the measured issue rate of the opcodes the MC kernels are made of (lanes per clock per SM)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
from optpricer_b200 import _native as nv  # noqa: E402
names = ["FFMA", "MUFU.EX2", "DFMA", "LDS.64 bytes", "IMAD.WIDE+LOP3", "IMAD.WIDE", "IMAD.HI", "IMAD", "LOP3", "I2FP",
         "sin.approx (FMUL+MUFU.SIN)", "MUFU.LG2", "MUFU.SQRT", "FMNMX", "FADD",
         "group: IMAD.WIDE + 3 FFMA", "group: IMAD.WIDE + 3 FMNMX", "group: IMAD.WIDE + MUFU.EX2",
         "CRR node: DFMA,DFMA,DSETP,2 FSEL", "CRR node: DFMA"]
nv.init(0)
for i, n in enumerate(names):
    v, _ = nv.microbench(i)
    print(f"{i:2d} {n:28s} {v:.4e} /s  = {v / (148 * 1.965e9):7.2f} per clk per SM (at 1965 MHz)")
