import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from optpricer_b200 import binomial as bn, _native as nv
from bench import cfg2_contracts
nv.init(0)
n = 200_000
S0, K, T, r, q, sig, _ = cfg2_contracts(n)
g = np.random.default_rng(1)
for name, call, rr, qq in (("calls q~U(.5%,2.5%), r=3%", 1, np.full(n, 0.03), g.uniform(0.005, 0.025, n)),
                           ("puts  q=0, r=3% (cfg2)", 0, np.full(n, 0.03), np.zeros(n))):
    c = np.full(n, call, dtype=np.uint8)
    for G in (0, 4, 2, 1):
        nv.set_option("crr_zone_group", G)
        bn.crr_batch(S0[:1000], K[:1000], T[:1000], rr[:1000], qq[:1000], sig[:1000], c[:1000], 1000, american=True)
        t0 = time.perf_counter(); a = bn.crr_batch(S0, K, T, rr, qq, sig, c, 1000, american=True); dt = time.perf_counter() - t0
        print("%s G=%d: %.3e options/s" % (name, G, n / dt))
nv.set_option("crr_zone_group", 0)
