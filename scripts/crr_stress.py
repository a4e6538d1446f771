#!/usr/bin/env python3
"""Random stress of the two-zone CRR kernel's segment schedule (csrc/crr_batch.cu: kZoneSeg0 / kZoneSeg): wide
parameter ranges, puts and calls, several tree depths; every price against the full-width kernel
(optp_set_option("crr_band", 0)), plus how many contracts the two-zone passes declined (those are priced by the
full-width kernel anyway: a decline costs time, never accuracy).  Usage: python scripts/crr_stress.py [n]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]
import numpy as np  # noqa: E402
from optpricer_b200 import _native as nv, binomial as bn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
nv.init(0)
g = np.random.default_rng(2026)
for N in (64, 250, 1000, 1023, 2500):
    S0 = g.uniform(20.0, 200.0, n)
    K = S0 * np.exp(g.uniform(-0.6, 0.6, n))
    T = np.exp(g.uniform(np.log(0.01), np.log(10.0), n))
    sig = np.exp(g.uniform(np.log(0.03), np.log(1.5), n))
    r = np.exp(g.uniform(np.log(0.001), np.log(0.3), n))
    q = np.where(g.random(n) < 0.5, 0.0, g.uniform(0.0, 0.12, n))
    is_call = (g.random(n) < 0.3).astype(np.uint8)
    band, st = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    t0 = time.perf_counter()
    bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True)
    secs = time.perf_counter() - t0
    nv.set_option("crr_band_only", 1)
    _, st_only = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    nv.set_option("crr_band_only", 0)
    nv.set_option("crr_band", 0)
    full, st_full = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    nv.set_option("crr_band", 1)
    ok = (st == 0) & (st_full == 0)
    err = np.abs(band[ok] - full[ok]) / np.maximum(np.abs(full[ok]), 1e-8 * K[ok])
    puts = (is_call == 0) & (r > q) & ok
    print(json.dumps({"N": N, "contracts": n, "priced": int(ok.sum()), "bad_p": int((st == 1).sum()),
                      "status_equal": bool(np.array_equal(st, st_full)),
                      "max_rel_err_vs_full_width": float(err.max()), "declined_by_two_zone_passes": int((st_only == 2).sum()),
                      "declined_puts_r_gt_q": int(((st_only == 2) & puts).sum()), "puts_r_gt_q": int(puts.sum()),
                      "never_exercised": int((st_only == 3).sum()), "seconds_host_api": round(secs, 4)}), flush=True)
