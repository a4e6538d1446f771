"""American puts on deep trees (N > 1023): two-zone kernel + one-option-per-CTA fallback vs the CTA kernel alone."""
import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from optpricer_b200 import binomial as bn, _native as nv
from bench import cfg2_contracts
nv.init(0)
for N, n in ((2000, 50_000), (5000, 20_000), (9000, 10_000)):
    S0, K, T, r, q, sig, c = cfg2_contracts(n)
    nv.set_option("crr_band", 1)
    bn.crr_batch(S0, K, T, r, q, sig, c, N, american=True)
    t0 = time.perf_counter(); a, st = bn.crr_batch(S0, K, T, r, q, sig, c, N, american=True, return_status=True); dt = time.perf_counter() - t0
    nv.set_option("crr_band_only", 1)
    _, st = bn.crr_batch(S0, K, T, r, q, sig, c, N, american=True, return_status=True)
    nv.set_option("crr_band_only", 0)
    nv.set_option("crr_band", 0)
    m = min(n, 2000)
    t0 = time.perf_counter(); b = bn.crr_batch(S0[:m], K[:m], T[:m], r[:m], q[:m], sig[:m], c[:m], N, american=True); dt2 = time.perf_counter() - t0
    nv.set_option("crr_band", 1)
    print("N=%d: banded %.3e options/s (declined %.2f%%), CTA kernel %.3e options/s; max rel diff %.2e"
          % (N, n / dt, 100.0 * np.mean(st == 2), m / dt2, np.max(np.abs(a[:m] - b) / np.maximum(np.abs(b), 1e-8))))
