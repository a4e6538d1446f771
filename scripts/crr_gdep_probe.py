import sys
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from optpricer_b200 import binomial as bn, _native as nv
nv.init(0)
g = np.random.default_rng(99)
n = 30_000
S0 = g.uniform(20, 300, n); K = S0 * np.exp(g.uniform(-1.2, 1.2, n)); T = np.exp(g.uniform(np.log(0.01), np.log(5.0), n))
sig = g.uniform(0.05, 0.9, n); r = g.uniform(0.0005, 0.12, n); q = np.where(g.random(n) < 0.5, 0.0, g.uniform(0.0, 0.1, n))
is_call = (g.random(n) < 0.2).astype(np.uint8)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
res = {}; st = {}
for G in (1, 2, 4):
    nv.set_option("crr_zone_group", G)
    res[G] = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True)
    nv.set_option("crr_band_only", 1)
    _, st[G] = bn.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    nv.set_option("crr_band_only", 0)
for G in (2, 4):
    bad = np.flatnonzero(res[1] != res[G])
    print("G", G, "mismatch", bad.size, "status differs", np.count_nonzero(st[1] != st[G]))
    b = bad[:12]
    for i in b:
        print("  i=%d call=%d r=%.4f q=%.4f sig=%.3f T=%.3f S0/K=%.3f st1=%d stG=%d p1=%.15g pG=%.15g" % (i, is_call[i], r[i], q[i], sig[i], T[i], S0[i]/K[i], st[1][i], st[G][i], res[1][i], res[G][i]))
    if bad.size:
        print("  calls among bad:", np.mean(is_call[bad]), " q>r among bad:", np.mean(q[bad] > r[bad]), " status(G=1) of bad:", np.bincount(st[1][bad], minlength=4), " status(G) of bad:", np.bincount(st[G][bad], minlength=4))
