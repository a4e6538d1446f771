"""Host-side cost of price_book: a 10k-row config-5 book with almost no paths (the kernels vanish, the Python
planning / struct filling / launch / estimator loop remains)."""
import cProfile, pstats, sys, time
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
from optpricer_b200 import book, _native as nv
nv.init(0)
rows = book.synthetic_book(10_000, n_paths=1024)
book.price_book(rows, mc_paths=2048)
t0 = time.perf_counter(); book.price_book(rows, mc_paths=2048); print("price_book host-bound: %.3f s" % (time.perf_counter() - t0))
cProfile.run("book.price_book(rows, mc_paths=2048)", "/tmp/prof")
pstats.Stats("/tmp/prof").sort_stats("tottime").print_stats(12)
