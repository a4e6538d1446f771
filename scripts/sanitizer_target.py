"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitizer_target.py [crr|mc ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
import numpy as np
import optpricer_b200 as op
from optpricer_b200 import _native as nv, processes as pr
which = set(sys.argv[1:]) or {"crr", "mc"}
nv.init(0)
if "crr" in which:
    g = np.random.default_rng(3)
    n = 96
    S0 = g.uniform(50, 150, n); K = S0 * g.uniform(0.8, 1.2, n); T = g.uniform(0.1, 2.0, n); sig = g.uniform(0.1, 0.6, n)
    r = g.uniform(0.0, 0.08, n); q = g.uniform(0.0, 0.08, n); call = (np.arange(n) % 2).astype(np.uint8)
    for G in (1, 2, 4):
        nv.set_option("crr_zone_group", G)
        for N in (200, 1000, 1500):
            p = op.crr_batch(S0, K, T, r, q, sig, call, N, american=True)
            print("crr zone G=%d N=%d" % (G, N), float(p.sum()))
    nv.set_option("crr_zone_group", 0)
    print("crr european", float(op.crr_batch(S0, K, T, r, q, sig, call, 300).sum()))
    nv.set_option("crr_band", 0)
    print("crr full-width", float(op.crr_batch(S0, K, T, r, q, sig, call, 300, american=True).sum()))
    nv.set_option("crr_band", 1)
    print("crr deep", float(op.crr_batch(S0[:4], K[:4], T[:4], r[:4], q[:4], sig[:4], call[:4], 10_000, american=True).sum()))
if "mc" in which:
    print("euro", op.euro_price_mc(op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2), "call", n_paths=50_000, seed=1))
    print("gbm", op.barrier_price(pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 50, 20_000, seed=1), 100.0, 0.05, 1.0, "call", 130.0, "up-and-out"))
    print("heston", op.asian_price(pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 51, 20_000, seed=1), 100.0, 0.05, 1.0, "call"))
    print("heston qe", op.asian_price(pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 16, 20_000, seed=1, scheme="qe"), 100.0, 0.05, 1.0, "call"))
    z = np.load(os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz"))
    tab = pr.SigmaTable(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
    print("lvm", op.barrier_price(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 10_000, tab, seed=1), 100.0, 0.05, 1.0, "call", 130.0, "up-and-out"))
    print("merton", op.lookback_price(pr.merton_jump_paths(100.0, 0.05, 0.0, 0.2, 1.0, 30, 10_000, lam=0.5, mJ=-0.1, sJ=0.15, seed=1), 0.05, 1.0, "call"))
    print("materialised", np.asarray(pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 10, 100, seed=1)).shape)
    from optpricer_b200 import book
    print("book", sum(r["price"] for r in book.price_book(book.synthetic_book(40, n_paths=2000), mc_paths=4000)))
print("launches", nv.launch_count())
