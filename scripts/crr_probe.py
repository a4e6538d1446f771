#!/usr/bin/env python3
"""Time the batched CRR kernel (config 2 contracts, American put, N=1000) and check the first 64
prices against the reference fixture.  Usage: python scripts/crr_probe.py [n_contracts]"""
import ctypes as C
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from optpricer_b200 import _native as nv  # noqa: E402
from bench import cfg2_contracts  # noqa: E402

n = min(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18, 1_000_000)
lib = nv.init(0)
cols = [a[:n] for a in cfg2_contracts(1_000_000)]  # the fixture's first 64 come from the 1M-draw stream
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cols]
price = torch.empty(n, dtype=torch.float64, device=dev)
status = torch.empty(n, dtype=torch.uint8, device=dev)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for american in (1, 0):
    def run():
        nv.check(lib.optp_crr_batch_dev(*[C.c_void_p(t.data_ptr()) for t in d], n, 1000, american,
                                        C.c_void_p(price.data_ptr()), C.c_void_p(status.data_ptr()), s), "crr")
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    msg = f"american={american} n={n}: {ms:.3f} ms  {n / ms * 1e3:.4e} options/s"
    if american:
        ref = np.array(json.load(open(os.path.join(ROOT, "tests/golden/crr.json")))["cfg2_first64_american_put_N1000"])
        got = price[:64].cpu().numpy()
        msg += f"  max rel err vs reference (64 contracts) {np.max(np.abs(got - ref) / ref):.2e}"
        msg += f"  status counts {np.bincount(status.cpu().numpy(), minlength=3).tolist()}"
    print(msg)
