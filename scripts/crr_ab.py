#!/usr/bin/env python3
"""A/B runs of CRR kernel variants on the GPU box: every library given (scripts/build_variant.sh) prices
BASELINE config 2 (1M American puts, N = 1000) device-resident in its own process; prints options/s, how many
contracts the two-zone passes declined (left for the full-width kernel), and the largest price difference to
the first library.  Usage: python scripts/crr_ab.py lib1.so lib2.so ...   (worker: --worker out.npy)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]


def worker(out):
    import ctypes as C
    import numpy as np
    import torch
    from bench import cfg2_contracts
    from optpricer_b200 import _native as nv
    lib = nv.init(0)
    n = int(os.environ.get("AB_N", "1000000"))
    cols = cfg2_contracts(n)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cols]
    price = torch.empty(n, dtype=torch.float64, device=dev)
    status = torch.empty(n, dtype=torch.uint8, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        nv.check(lib.optp_crr_batch_dev(*[C.c_void_p(t.data_ptr()) for t in d], n, 1000, 1, C.c_void_p(price.data_ptr()),
                                        C.c_void_p(status.data_ptr()), sp), "crr")
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(7):
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    p = price.cpu().numpy()
    nv.set_option("crr_band_only", 1)
    step(); torch.cuda.synchronize()
    declined = int((status != 0).sum().item())
    nv.set_option("crr_band_only", 0)
    np.save(out, p)
    print(json.dumps({"lib": os.path.basename(nv.LIB_PATH), "ms": best, "options_per_s": n / best * 1e3, "declined": declined,
                      "checksum8": float(p[:8].sum())}))


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2])
        sys.exit(0)
    import numpy as np
    ref = None
    for i, lib in enumerate(sys.argv[1:]):
        out = f"/tmp/ab_{i}.npy"
        r = subprocess.run([sys.executable, __file__, "--worker", out], env=dict(os.environ, OPTP_LIB=os.path.abspath(lib)),
                           capture_output=True, text=True)
        line = (r.stdout.strip().splitlines() or ["{}"])[-1]
        if r.returncode != 0:
            print(lib, "FAILED", r.stderr[-500:])
            continue
        p = np.load(out)
        if ref is None:
            ref = p
        rec = json.loads(line)
        rec["max_rel_diff_vs_first"] = float(np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1e-8)))
        print(json.dumps(rec), flush=True)
