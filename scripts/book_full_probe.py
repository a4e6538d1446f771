import sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
from optpricer_b200 import book as bk, _native as nv
nv.init(0)
rows = bk.synthetic_book(10_000, n_paths=5_000_000)
bk.price_book(rows[:50], mc_paths=5_000_000)
t0 = time.perf_counter()
out = bk.price_book(rows, mc_paths=5_000_000)
dt = time.perf_counter() - t0
steps = sum(2.0 * 5_000_000 * (1 if r["method"] == "mc" else r["n_steps"]) for r in rows)
print("full-size cfg5 book: %.2f s, %.3e path-steps/s, %d priced" % (dt, steps / dt, sum(1 for r in out if r.get("price") is not None)))
