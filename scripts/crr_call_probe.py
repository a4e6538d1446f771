import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from optpricer_b200 import binomial as bn, _native as nv
from bench import cfg2_contracts
nv.init(0)
S0, K, T, r, q, sig, _ = cfg2_contracts(200_000)
call = np.ones(200_000, dtype=np.uint8)
bn.crr_batch(S0[:1000], K[:1000], T[:1000], r[:1000], q[:1000], sig[:1000], call[:1000], 1000, american=True)
t0 = time.perf_counter(); a = bn.crr_batch(S0, K, T, r, q, sig, call, 1000, american=True); dt = time.perf_counter() - t0
nv.set_option("crr_band", 0)
t0 = time.perf_counter(); b = bn.crr_batch(S0, K, T, r, q, sig, call, 1000, american=True); dt2 = time.perf_counter() - t0
e = bn.crr_batch(S0, K, T, r, q, sig, call, 1000, american=False)
print("american calls q=0: routed %.3e options/s, full-width american %.3e options/s; max rel diff %.2e; vs european %.2e"
      % (200_000 / dt, 200_000 / dt2, np.max(np.abs(a - b) / np.maximum(b, 1e-10)), np.max(np.abs(a - e))))
