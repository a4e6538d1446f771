import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/options-pricing-engine-black-scholes-binomial-monte-carlo_b200")
import numpy as np
from optpricer_b200 import _native as nv, processes as pr, exotics as ex
z = np.load("/root/repo/tests/golden/dupire_table_252x256.npz")
tab = pr.SigmaTable(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
nv.init(0)
nv.set_option("lv_direct", 1)
for _ in range(3):
    print(ex.price_payoffs(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 1 << 22, tab, seed=1), [ex.barrier_spec(100.0, "call", 130.0, "up-and-out")], 0.05, 1.0))
