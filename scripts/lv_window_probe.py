import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/options-pricing-engine-black-scholes-binomial-monte-carlo_b200")
import numpy as np
import optpricer_b200 as op
from optpricer_b200 import _native as nv, processes as pr, exotics as ex
z = np.load("/root/repo/tests/golden/dupire_table_252x256.npz")
tab = pr.SigmaTable(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
nv.init(0)
specs = [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.asian_spec(100.0, "put"), ex.lookback_spec("call")]
for win in (2, 1, 0):
    nv.set_option("lv_window", 1 if win else 0)
    nv.set_option("lv_direct", 1 if win == 2 else 0)
    lp = pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 1 << 23, tab, seed=1)
    ex.price_payoffs(lp, specs[:1], 0.05, 1.0)
    t0 = time.perf_counter(); r = ex.price_payoffs(lp, specs[:1], 0.05, 1.0); dt = time.perf_counter() - t0
    small = ex.price_payoffs(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 200000, tab, seed=1), specs, 0.05, 1.0)
    print("window", win, r, f"{2 * (1 << 23) * 252 / dt:.4e} path-steps/s", small)
nv.set_option("lv_window", 1)
nv.set_option("lv_direct", 0)
