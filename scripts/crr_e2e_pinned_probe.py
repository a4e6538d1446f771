#!/usr/bin/env python3
"""End-to-end optp_crr_batch (host pointers in, host pointers out) with pageable against pinned caller arrays:
how much of the gap to the device-resident rate is the driver's staging of pageable copies."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]
import numpy as np, torch
from bench import cfg2_contracts
from optpricer_b200 import _native as nv, binomial as bn
nv.init(0)
n = 1_000_000
cols = cfg2_contracts(n)
def best(fn, reps=7):
    fn(); t = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t = min(t, time.perf_counter() - t0)
    return t
t_page = best(lambda: bn.crr_batch(*cols, 1000, american=True))
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in cols]
t_pin = best(lambda: bn.crr_batch(*pinned, 1000, american=True))
print(json.dumps({"pageable_options_per_s": n / t_page, "pinned_inputs_options_per_s": n / t_pin}))
