#!/usr/bin/env python3
"""One launch of every production kernel at a size ncu can replay quickly (the `ncu --set full`
target; see profiles/README.md).  Usage: python scripts/profile_target.py [which ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
import numpy as np  # noqa: E402
import optpricer_b200 as op  # noqa: E402
from optpricer_b200 import _native as nv, exotics as ex, processes as pr  # noqa: E402
from bench import cfg2_contracts  # noqa: E402

which = set(sys.argv[1:]) or {"gbm", "heston", "lvm", "terminal", "crr"}
nv.init(0)
reps = 3  # first launches warm the instruction cache; ncu captures with -s
for _ in range(reps):
    if "gbm" in which:
        print("gbm", op.barrier_price(pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, 1 << 23, seed=1), 100.0, 0.05, 1.0,
                                      "call", 130.0, "up-and-out"))
    if "heston" in which:
        print("heston", op.asian_price(pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, 1 << 22,
                                                       seed=1), 100.0, 0.05, 1.0, "call"))
    if "lvm" in which:
        z = np.load(os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz"))
        tab = pr.SigmaTable(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
        print("lvm", op.barrier_price(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 1 << 22, tab, seed=1),
                                      100.0, 0.05, 1.0, "call", 130.0, "up-and-out"))
    if "terminal" in which:
        print("terminal", op.euro_price_mc(op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2), "call", n_paths=1 << 28, seed=1))
    if "crr" in which:   # banded kernel (default) and the full-width kernel on the same contracts
        cols = cfg2_contracts(1 << 16)
        print("crr band", op.crr_batch(*cols, 1000, american=True)[:3])
        nv.set_option("crr_band", 0)
        print("crr full", op.crr_batch(*cols, 1000, american=True)[:3])
        nv.set_option("crr_band", 1)
print("launches", nv.launch_count())
