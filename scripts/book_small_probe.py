#!/usr/bin/env python3
"""Time price_book on a 10k-row book at the REFERENCE's default size (100k base paths per row, scripts/price_book.py):
items that do not fill the machine, where optp_mc_book's four launch streams matter (0.43 s on one stream, 0.21 s on
four).  Usage: [OPTP_LIB=variant.so] python scripts/book_small_probe.py"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'options-pricing-engine-black-scholes-binomial-monte-carlo_b200')]
from optpricer_b200 import book as bk, _native as nv
nv.init(0)
rows = bk.synthetic_book(10000, n_paths=100_000, exotic_steps=252)
bk.price_book(rows[:300], mc_paths=100_000)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); out = bk.price_book(rows, mc_paths=100_000); best = min(best, time.perf_counter() - t0)
print(os.path.basename(nv.LIB_PATH), 'book 10k rows x 2e5 effective paths: %.3f s' % best, 'checksum', sum(r['price'] for r in out[:100]))
