"""Latency of single public-API calls (the shape the reference's own tests and scripts use)."""
import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import optpricer_b200 as op
from optpricer_b200 import processes as pr, _native as nv
nv.init(0)
opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
def t(name, f, n=200):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    print("%-60s %8.1f us/call" % (name, (time.perf_counter() - t0) / n * 1e6))
t("crr(N=500, american put)", lambda: op.crr(opt, "put", 500, american=True))
t("crr(N=500, european call)", lambda: op.crr(opt, "call", 500))
t("crr(N=1000, american put)", lambda: op.crr(opt, "put", 1000, american=True))
t("crr(N=1000, american call, q=0)", lambda: op.crr(opt, "call", 1000, american=True))
t("crr_vec(21 strikes, N=500, american put)", lambda: op.crr_vec(100.0, [90 + i for i in range(21)], 1.0, 0.03, 0.0, 0.2, "put", 500, american=True))
t("euro_price_mc(n_paths=200k)", lambda: op.euro_price_mc(opt, "call", n_paths=200_000, seed=1))
t("barrier_price(gbm_paths(500 steps, 100k paths))", lambda: op.barrier_price(pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 500, 100_000, seed=42), 100.0, 0.05, 1.0, "call", 130.0, "up-and-out"))
t("asian_price(heston_paths(365 steps, 50k paths))", lambda: op.asian_price(pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, 50_000, seed=42), 100.0, 0.05, 1.0, "call"))
