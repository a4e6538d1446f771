#!/usr/bin/env python3
"""A/B runs of Monte Carlo kernel variants on the GPU box: every library given (scripts/build_variant.sh) times the
BASELINE config-4 kernel (Heston FTE, 365 steps, arithmetic Asian call), the config-3 kernel (LV Milstein on the
Dupire table, barrier call) and the headline GBM + barrier kernel, device-resident, best of 5; prints path-steps/s
and the prices.  Usage: python scripts/mc_ab.py lib1.so lib2.so ...   (worker: --worker)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")]


def worker():
    import ctypes as C
    import numpy as np
    import torch
    from optpricer_b200 import _native as nv, exotics as ex
    lib = nv.init(0)
    dev = torch.device("cuda", 0)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.zeros(nv.NSTATS, dtype=torch.float64, device=dev)

    def timed(p, spec, n_steps, nb, r, T):
        pay = ex._native_payoffs([spec], r, T)
        run = lambda: nv.check(lib.optp_mc_path_dev(C.byref(p), pay, 1, n_steps, nb, 1, nv.MODE_FP32_PHILOX, 42, 0, None, None,  # noqa: E731
                                                    None, C.c_void_p(out.data_ptr()), sp), "optp_mc_path_dev")
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        s = out.cpu().numpy()
        return 2.0 * nb * n_steps / (best * 1e-3), ex.estimate(s)

    res = {"lib": os.path.basename(nv.LIB_PATH)}
    hp = nv.OptpProcess()
    hp.kind = nv.PROC_HESTON_FTE
    hp.S0, hp.r, hp.q, hp.T = 100.0, 0.05, 0.0, 1.0
    hp.v0, hp.kappa, hp.theta, hp.xi, hp.rho = 0.04, 2.0, 0.04, 0.3, -0.7
    res["heston_asian"] = timed(hp, ex.asian_spec(100.0, "call"), 365, 1 << 25, 0.05, 1.0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz"))
    table = np.ascontiguousarray(z["table"], dtype=np.float64)
    nv.check(lib.optp_sigma_table_upload(nv.ptr(table), table.shape[0], table.shape[1]), "table upload")
    lp = nv.OptpProcess()
    lp.kind = nv.PROC_LV_MILSTEIN
    lp.S0, lp.r, lp.q, lp.T, lp.dS_bump = 100.0, 0.05, 0.02, 1.0, 0.01
    lp.n_nodes, lp.x0, lp.h = table.shape[1], float(z["x0"]), float(z["h"])
    res["lvm_barrier"] = timed(lp, ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), 252, 1 << 25, 0.05, 1.0)
    gp = nv.OptpProcess()
    gp.kind = nv.PROC_GBM
    gp.S0, gp.r, gp.q, gp.T, gp.sigma = 100.0, 0.05, 0.0, 1.0, 0.2
    res["gbm_barrier"] = timed(gp, ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), 252, 1 << 26, 0.05, 1.0)
    res["gbm_asian"] = timed(gp, ex.asian_spec(100.0, "call"), 252, 1 << 25, 0.05, 1.0)
    print(json.dumps(res))


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker()
        sys.exit(0)
    for lib in sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, "--worker"], env=dict(os.environ, OPTP_LIB=os.path.abspath(lib)),
                           capture_output=True, text=True)
        print((r.stdout.strip().splitlines() or [lib + " FAILED " + r.stderr[-400:]])[-1], flush=True)
