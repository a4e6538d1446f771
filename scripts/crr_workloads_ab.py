#!/usr/bin/env python3
"""A/B of CRR library variants (scripts/build_variant.sh) on the workloads next to BASELINE config 2: American calls
with dividends and puts with q > r (dual put / boundary jump), other tree depths, the risk batch; host API, best of 5.
Usage: python scripts/crr_workloads_ab.py lib1.so lib2.so ...   (worker: --worker)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]


def worker():
    import numpy as np
    from bench import cfg2_contracts
    from optpricer_b200 import _native as nv, binomial as bn
    nv.init(0)
    out = {"lib": os.path.basename(nv.LIB_PATH)}

    def best(fn, reps=5):
        fn()
        t = 1e9
        for _ in range(reps):
            t0 = time.perf_counter(); r = fn(); t = min(t, time.perf_counter() - t0)
        return t, r
    n = 200_000
    S0, K, T, r, q, sig, is_call = cfg2_contracts(n)
    g = np.random.default_rng(7)
    qd = g.uniform(0.005, 0.025, n)
    cases = {
        "puts_N1000": (S0, K, T, r, q, sig, is_call, 1000),
        "calls_div_N1000": (S0, K, T, r, qd, sig, np.ones(n, np.uint8), 1000),
        "puts_q_gt_r_N1000": (S0, K, T, r, r + qd, sig, is_call, 1000),
        "puts_N500": (S0, K, T, r, q, sig, is_call, 500),
        "puts_N250": (S0, K, T, r, q, sig, is_call, 250),
        "puts_N2000": tuple(a[:50_000] for a in (S0, K, T, r, q, sig, is_call)) + (2000,),
        "puts_N5000": tuple(a[:10_000] for a in (S0, K, T, r, q, sig, is_call)) + (5000,),
        "puts_N1001": (S0, K, T, r, q, sig, is_call, 1001),
    }
    if os.environ.get("AB_1M"):
        big = cfg2_contracts(1_000_000)
        cases = {"puts_N1000_1M_host_api": tuple(big) + (1000,), "puts_N1000_600k_host_api": tuple(a[:600_000] for a in big) + (1000,)}
    for name, a in cases.items():
        t, p = best(lambda: bn.crr_batch(*a[:7], a[7], american=True))
        out[name] = {"options_per_s": round(len(a[0]) / t), "checksum": float(np.sum(p[:1000]))}
    print(json.dumps(out))


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker()
        sys.exit(0)
    for lib in sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, "--worker"], env=dict(os.environ, OPTP_LIB=os.path.abspath(lib)),
                           capture_output=True, text=True)
        print((r.stdout.strip().splitlines() or [r.stderr[-400:]])[-1], flush=True)
