#!/usr/bin/env python3
"""A/B of the banded CRR kernels on config-2 contracts (American put, N=1000): the single-zone
band kernel, the two-zone kernel with 1, 2 and 4 options per warp, and the
full-width kernel as the parity anchor.  Usage: python scripts/crr_zone_probe.py [n_contracts]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from optpricer_b200 import _native as nv  # noqa: E402
from bench import cfg2_contracts  # noqa: E402

n = min(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18, 1_000_000)
groups = [1, 2, 4, 0]
if len(sys.argv) > 2:
    groups = [int(x) for x in sys.argv[2:]]
lib = nv.init(0)
cols = [a[:n] for a in cfg2_contracts(1_000_000)]
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cols]
price = torch.empty(n, dtype=torch.float64, device=dev)
status = torch.empty(n, dtype=torch.uint8, device=dev)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    nv.check(lib.optp_crr_batch_dev(*[C.c_void_p(t.data_ptr()) for t in d], n, 1000, 1,
                                    C.c_void_p(price.data_ptr()), C.c_void_p(status.data_ptr()), s), "crr")


def timed(reps=3):
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nv.set_option("crr_band", 0)
ms = timed(1)
full = price.cpu().numpy().copy()
print(f"full-width: {ms:.2f} ms {n / ms * 1e3:.4e} options/s")
nv.set_option("crr_band", 1)
nv.set_option("crr_band_kernel", 1)
ms = timed()
err = np.max(np.abs(price.cpu().numpy() - full) / full)
print(f"band (single zone): {ms:.2f} ms {n / ms * 1e3:.4e} options/s  max rel err vs full {err:.2e}")
nv.set_option("crr_band_kernel", 2)
for G in groups:
    nv.set_option("crr_zone_group", G)
    ms = timed()
    err = np.max(np.abs(price.cpu().numpy() - full) / full)
    zone = price.cpu().numpy().copy()
    nv.set_option("crr_band_only", 1)
    status.zero_()
    run()
    torch.cuda.synchronize()
    nv.set_option("crr_band_only", 0)
    declined = int((status == 2).sum().item())
    same = "" if G == groups[0] else f"  bit-identical to G={groups[0]}: {bool(np.array_equal(zone, first))}"
    if G == groups[0]:
        first = zone
    print(f"zone G={G}: {ms:.2f} ms {n / ms * 1e3:.4e} options/s  max rel err vs full {err:.2e}  declined {declined}{same}")
