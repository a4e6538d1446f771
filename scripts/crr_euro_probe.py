"""European CRR trees: O(N) expectation kernel against the backward induction (throughput and agreement)."""
import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from optpricer_b200 import binomial as bn, _native as nv
from bench import cfg2_contracts
nv.init(0)
for N, n in ((1000, 1_000_000), (500, 1_000_000), (5000, 200_000)):
    cols = cfg2_contracts(n)
    bn.crr_batch(*cols, N)
    t0 = time.perf_counter(); a = bn.crr_batch(*cols, N); dt = time.perf_counter() - t0
    m = min(n, 200_000 if N <= 1000 else 2000)
    sub = [c[:m] for c in cols]
    nv.set_option("crr_euro_sum", 0)
    bn.crr_batch(*[c[:1000] for c in cols], N)
    t0 = time.perf_counter(); b = bn.crr_batch(*sub, N); dt2 = time.perf_counter() - t0
    nv.set_option("crr_euro_sum", 1)
    print("N=%d european: expectation kernel %.3e options/s (host arrays), backward induction %.3e; max rel diff %.2e"
          % (N, n / dt, m / dt2, np.max(np.abs(a[:m] - b) / np.maximum(1.0, np.abs(b)))))
