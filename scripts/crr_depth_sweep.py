#!/usr/bin/env python3
"""Every tree depth N = 1 ... 140 (segment boundaries, odd / even last levels, trees narrower than the window) and a
few deeper ones: the two-zone kernel (all layouts) against the full-width kernel on random contracts -- puts, calls,
dividends, deep in / out of the money.  Prints the largest relative difference per block of depths and any status
mismatch.  Usage: python scripts/crr_depth_sweep.py [contracts_per_depth]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]
import numpy as np  # noqa: E402
from optpricer_b200 import _native as nv, binomial as bn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
nv.init(0)
g = np.random.default_rng(404)
S0 = g.uniform(20.0, 200.0, n)
K = S0 * np.exp(g.uniform(-0.8, 0.8, n))
T = np.exp(g.uniform(np.log(0.02), np.log(5.0), n))
sig = np.exp(g.uniform(np.log(0.05), np.log(1.0), n))
r = np.exp(g.uniform(np.log(0.002), np.log(0.2), n))
q = np.where(g.random(n) < 0.4, 0.0, g.uniform(0.0, 0.12, n))
is_call = (g.random(n) < 0.3).astype(np.uint8)
depths = list(range(1, 141)) + [191, 192, 193, 255, 256, 511, 777, 1279, 1280, 1281, 2047, 3001]
worst, bad, block = 0.0, 0, []
for N in depths:
    nv.set_option("crr_band", 0)
    full, st_f = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    nv.set_option("crr_band", 1)
    for G in (0, 1, 2, 4):
        nv.set_option("crr_zone_group", G)
        zone, st = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
        ok = (st == 0) & (st_f == 0)
        bad += int(np.count_nonzero(st != st_f))
        err = float(np.max(np.abs(zone[ok] - full[ok]) / np.maximum(np.abs(full[ok]), 1e-8 * K[ok]))) if ok.any() else 0.0
        worst = max(worst, err)
    nv.set_option("crr_zone_group", 0)
    block.append(N)
    if len(block) == 20 or N == depths[-1]:
        print(json.dumps({"depths": f"{block[0]}..{block[-1]}", "contracts": n, "max_rel_diff_vs_full_width": worst,
                          "status_mismatches": bad}), flush=True)
        worst, bad, block = 0.0, 0, []
