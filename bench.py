#!/usr/bin/env python3
"""Benchmark of the B200 pricing hot paths (BASELINE.json metric: MC path-steps/s and CRR
options/s at N=1000 American).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on rank 0.  The headline `metric` is Monte Carlo path-steps/s on the fused GBM
path kernel (252-step up-and-out barrier call, antithetic, Philox in registers); the second
headline, CRR options/s on BASELINE config 2 (1M American puts, N=1000), is the `crr` object of
the same line with the same sub-keys (value / e2e / roofline / cpu_baseline).  `workloads` lists
the other BASELINE configurations at bounded size.

  value  device-resident: kernels launched through the *_dev C-ABI entry points on torch's
         current stream, timed with CUDA events on that stream, L2 flushed between steps.
  e2e    the same work through the public Python API (host buffers in, host results out:
         optp_mc_path / optp_crr_batch stage, launch, copy back and synchronise), wall clock
         between device synchronisations.
  --impl reference: the numpy oracle restatement of the reference (oracle/, kind "port" -- the
         reference itself is pure Python and is not present on the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# ---- workload definitions ------------------------------------------------------------
MC_STEPS = 252
MC_BASE_PER_GPU = 1 << 26           # base paths per GPU per step (x2 antithetic mirrors)
MC_CONTRACT = dict(S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0, K=100.0, barrier=130.0)
CRR_N = 1000
CRR_CONTRACTS_PER_GPU = 1_000_000
# algorithmic work per unit (restated in DESIGN.md)
OPS_PER_PATH_STEP = {"gbm_barrier": 13.0, "heston_asian": 32.0, "lvm_barrier": 50.0, "terminal": 18.0}
CRR_SMEM_BYTES_PER_OPTION = 12.012e6   # 500 500 node updates x 24 B
CRR_FP64_OPS_PER_OPTION = 2.5e6        # 500 500 node updates x 5


def cfg2_contracts(n, seed=0):
    """BASELINE config 2 synthetic batch: draw order S0, K-ratio, T, sigma."""
    g = np.random.default_rng(seed)
    S0 = g.uniform(50.0, 150.0, n)
    K = S0 * g.uniform(0.8, 1.2, n)
    T = g.uniform(0.1, 2.0, n)
    sigma = g.uniform(0.1, 0.6, n)
    return S0, K, T, np.full(n, 0.03), np.zeros(n), sigma, np.zeros(n, dtype=np.uint8)


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = []
        with open(self.tmp.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) == 7:
                    rows.append(parts)
        os.unlink(self.tmp.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        def num(x):
            try:
                return float(x)
            except ValueError:
                return float("nan")
        power = [num(r[2]) for r in rows]
        hot = [r for r, p in zip(rows, power) if p >= 0.5 * max(power)] or rows
        sm = sorted(num(r[0]) for r in hot)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": num(rows[0][1]), "power_w_max": max(power),
                "samples": len(rows), "samples_under_load": len(hot), "reasons": reasons}


# ---- CPU arm (oracle port of the reference) ----------------------------------------------------
def _cpu_mc_task(args):
    seed, n_steps, n_base = args
    from oracle import path_schemes, payoffs
    c = MC_CONTRACT
    t0 = time.perf_counter()
    (z,) = path_schemes.draw_normals(seed, n_steps, n_base)
    paths = path_schemes.gbm_exact(z, S0=c["S0"], r=c["r"], q=c["q"], sigma=c["sigma"], T=c["T"])
    price = payoffs.barrier(paths, c["K"], c["r"], c["T"], "call", c["barrier"], "up-and-out")
    return 2 * n_base * n_steps, time.perf_counter() - t0, price[0]


def _cpu_crr_task(args):
    lo, hi = args
    from oracle import crr_tree
    S0, K, T, r, q, sig, is_call = cfg2_contracts(hi)
    t0 = time.perf_counter()
    p, _ = crr_tree.crr_loop(S0[lo:hi], K[lo:hi], T[lo:hi], r[lo:hi], q[lo:hi], sig[lo:hi], is_call[lo:hi], CRR_N, True)
    return hi - lo, time.perf_counter() - t0, float(p[0])


def _cpu_crr_c_task(args):
    """The table-based C restatement (oracle/crr_tree.c) on a slice of config-2 contracts: what a
    hand-optimised scalar CPU implementation of the same lattice does (not the reference's numpy path)."""
    lo, hi = args
    from oracle import crr_c
    S0, K, T, r, q, sig, is_call = cfg2_contracts(hi)
    t0 = time.perf_counter()
    p, _ = crr_c.crr_loop_fast(S0[lo:hi], K[lo:hi], T[lo:hi], r[lo:hi], q[lo:hi], sig[lo:hi], is_call[lo:hi], CRR_N, True)
    return hi - lo, time.perf_counter() - t0, float(p[0])


def _cpu_extra_task(args):
    """The other BASELINE configurations on the oracle port (SURVEY 8d CPU baseline (i) and (iii)):
    returns (path-steps, seconds) of one bounded call."""
    what, seed = args
    from oracle import path_schemes, payoffs, terminal_mc, sigma_grid
    t0 = time.perf_counter()
    if what == "terminal":     # cfg 1: euro_price_mc, chunks of 100k as the reference streams them
        n_paths = 2_000_000
        terminal_mc.euro_mc(100.0, 110.0, 1.0, 0.03, 0.2, 0.0, "call", n_paths=n_paths, seed=seed)
        units = 2 * n_paths
    elif what == "heston_asian":   # cfg 4: heston_paths + asian_price, 365 steps
        n_steps, n_base = 365, 10_000
        z2, zp = path_schemes.draw_normals(seed, n_steps, n_base, 2)
        paths = path_schemes.heston_fte(z2, zp, S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3,
                                        rho=-0.7, T=1.0)
        payoffs.asian(paths, 100.0, 0.05, 1.0, "call")
        units = 2 * n_base * n_steps
    else:                          # cfg 3: milstein_local_vol_paths on the tabulated Dupire surface + barrier_price
        n_steps, n_base = 252, 10_000
        z = np.load(os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz"))
        grid = sigma_grid.SigmaGrid(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
        (zz,) = path_schemes.draw_normals(seed, n_steps, n_base)
        paths = path_schemes.lv_milstein(zz, grid, S0=100.0, r=0.05, q=0.02, T=1.0)
        payoffs.barrier(paths, 100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
        units = 2 * n_base * n_steps
    return units, time.perf_counter() - t0


def cpu_step(pool, workers, mc_base, crr_per_worker, step, mc_calls=1):
    """One bounded sample of each workload on `workers` host processes (`mc_calls` calls of mc_base base paths
    per process, the way the reference would chunk a larger run); returns units and seconds."""
    t0 = time.perf_counter()
    mc = list(pool.map(_cpu_mc_task, [(1000 * step + w, MC_STEPS, mc_base) for w in range(workers * mc_calls)]))
    t_mc = time.perf_counter() - t0
    t0 = time.perf_counter()
    cr = list(pool.map(_cpu_crr_task, [(w * crr_per_worker, (w + 1) * crr_per_worker) for w in range(workers)]))
    t_crr = time.perf_counter() - t0
    return sum(m[0] for m in mc), t_mc, sum(c[0] for c in cr), t_crr


def cpu_baseline_single_core():
    """Scalar (1 core) oracle port, ~10-20 s: the `cpu_baseline` objects of the b200 line."""
    n, dt, _ = _cpu_mc_task((1, MC_STEPS, 20_000))
    reps = max(1, int(6.0 / max(dt, 1e-3)))
    tot_n, tot_t = n, dt
    for i in range(reps - 1):
        n, dt, _ = _cpu_mc_task((2 + i, MC_STEPS, 20_000))
        tot_n += n
        tot_t += dt
    mc = {"value": tot_n / tot_t, "unit": "path-steps/s", "cores": 1, "kind": "port",
          "sample": f"oracle gbm_exact+barrier_payoff (numpy), {reps} calls of 252 steps x 20000 base paths (+mirrors)"}
    n, dt, _ = _cpu_crr_task((0, 160))
    crr = {"value": n / dt, "unit": "options/s", "cores": 1, "kind": "port",
           "sample": "oracle crr_loop (numpy restatement of binomial.py:7-38) on the first 160 config-2 contracts"}
    try:
        nc, dtc, _ = _cpu_crr_c_task((0, 4000))
        crr["optimised_c"] = {"value": nc / dtc, "unit": "options/s", "cores": 1, "kind": "port",
                              "sample": "oracle/crr_tree.c table-based loop (no pow() per node, gcc -O2) on the first "
                                        "4000 config-2 contracts; NOT the reference's numpy path"}
    except Exception as e:  # the C oracle is optional test infrastructure
        crr["optimised_c"] = {"unavailable": str(e)[:200]}
    return mc, crr


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import concurrent.futures as cf
    workers = os.cpu_count() or 1
    mc_base, mc_calls, crr_per = 20_000, 3, 24   # ~1.2 s of work on every core per step
    with cf.ProcessPoolExecutor(max_workers=workers) as pool:
        for s in range(args.warmup):
            cpu_step(pool, workers, mc_base // 4, 2, s)
        u_mc = t_mc = u_crr = t_crr = 0.0
        for s in range(args.steps):
            a, b, c, d = cpu_step(pool, workers, mc_base, crr_per, 100 + s, mc_calls)
            u_mc += a; t_mc += b; u_crr += c; t_crr += d
        try:
            t0 = time.perf_counter()
            cc = list(pool.map(_cpu_crr_c_task, [(w * 2000, (w + 1) * 2000) for w in range(workers)]))
            crr_c_v = sum(c[0] for c in cc) / (time.perf_counter() - t0)
        except Exception:
            crr_c_v = None
        extra = []
        for what, label in (("terminal", "cfg1 euro_price_mc (terminal GBM, antithetic + CV), 2M paths per process"),
                            ("heston_asian", "cfg4 heston_paths + asian_price, 365 steps x 10k base paths per process"),
                            ("lvm_barrier", "cfg3 milstein_local_vol_paths (tabulated Dupire sigma) + barrier_price, "
                                            "252 steps x 10k base paths per process")):
            try:
                t0 = time.perf_counter()
                res = list(pool.map(_cpu_extra_task, [(what, 7 + w) for w in range(workers)]))
                extra.append({"workload": label, "value": sum(x[0] for x in res) / (time.perf_counter() - t0),
                              "unit": "path-steps/s", "cores": workers, "kind": "port"})
            except Exception as e:
                extra.append({"workload": label, "unavailable": str(e)[:200]})
    mc_v, crr_v = u_mc / t_mc, u_crr / t_crr
    sample_mc = (f"{workers} processes x {mc_calls} calls of (252 steps x {mc_base} base paths + mirrors) per step, "
                 "oracle numpy port")
    sample_crr = f"{workers} processes x {crr_per} config-2 contracts per step, oracle numpy port"
    line = {
        "impl": "reference", "metric": "mc_path_steps_per_sec", "value": mc_v, "unit": "path-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_mc / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gbm_paths 252 steps + barrier_price up-and-out call (S0=100,K=100,B=130), antithetic",
                   "sample": sample_mc},
        "cpu_baseline": {"value": mc_v, "unit": "path-steps/s", "cores": workers, "kind": "port", "sample": sample_mc},
        "e2e": {"value": mc_v, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "crr": {"metric": "crr_american_N1000_options_per_sec", "value": crr_v, "unit": "options/s",
                "ms_per_step": 1e3 * t_crr / args.steps,
                "cpu_baseline": {"value": crr_v, "unit": "options/s", "cores": workers, "kind": "port",
                                 "sample": sample_crr},
                "optimised_c": {"value": crr_c_v, "unit": "options/s", "cores": workers, "kind": "port",
                                "sample": f"{workers} processes x 2000 config-2 contracts, oracle/crr_tree.c table-based "
                                          "loop (not the reference's numpy path)"},
                "e2e": {"value": crr_v, "unit": "options/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "workloads": extra,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- GPU arm ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mc-base", type=int, default=MC_BASE_PER_GPU)
    ap.add_argument("--crr-contracts", type=int, default=CRR_CONTRACTS_PER_GPU)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-extra", action="store_true", help="skip the `workloads` list")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from optpricer_b200 import _native as nv, exotics as ex, processes as pr, binomial as bn
    from optpricer_b200 import distributed as dd, monte_carlo as mc
    lib = nv.init(local)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # roofline denominators measured in this run (MEASURED_PEAKS.json carries only HBM and bf16)
    peaks = {}
    for which, name in ((0, "fp32_ffma_lane_ops"), (1, "mufu_ex2_ops"), (2, "fp64_dfma_lane_ops"),
                        (3, "smem_lds64_bytes"), (4, "imad_wide_lane_ops")):
        v, mhz = nv.microbench(which)
        peaks[name] = v
        peaks.setdefault("sm_mhz_in_microbench", mhz)

    sampler = ClockSampler(local) if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    launches0 = nv.launch_count()

    # ================= Monte Carlo: fused GBM path kernel, up-and-out barrier call =================
    c = MC_CONTRACT
    proc = nv.OptpProcess()
    proc.kind = nv.PROC_GBM
    proc.S0, proc.r, proc.q, proc.T, proc.sigma = c["S0"], c["r"], c["q"], c["T"], c["sigma"]
    spec = ex.barrier_spec(c["K"], "call", c["barrier"], "up-and-out")
    pay = ex._native_payoffs([spec], c["r"], c["T"])
    stats_t = torch.zeros(nv.NSTATS, dtype=torch.float64, device=dev)
    n_base = args.mc_base

    def mc_step(i):
        off = (i * world + rank) * n_base
        nv.check(lib.optp_mc_path_dev(C.byref(proc), pay, 1, MC_STEPS, n_base, 1, nv.MODE_FP32_PHILOX, 20261018,
                                      off, None, None, None, C.c_void_p(stats_t.data_ptr()), sptr), "optp_mc_path_dev")

    for i in range(W):
        flush.zero_()
        mc_step(i)
        if world > 1:
            dist.all_reduce(stats_t)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    wall0 = time.perf_counter()
    mc_launch0 = nv.launch_count()
    price_sum = 0.0
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        mc_step(W + i)
        ev[i][1].record()
        if world > 1:
            dist.all_reduce(stats_t)
        ev[i][2].record()
    barrier()
    mc_wall = time.perf_counter() - wall0
    mc_launches = nv.launch_count() - mc_launch0
    mc_ms = max_over_ranks(sum(e[0].elapsed_time(e[2]) for e in ev) / K)
    mc_kernel_ms = max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in ev) / K)
    st = stats_t.cpu().numpy()
    mc_price = ex.estimate(st)
    path_steps_per_step = 2.0 * n_base * MC_STEPS * world
    mc_value = path_steps_per_step / (mc_ms * 1e-3)
    mc_kernel_rate = 2.0 * n_base * MC_STEPS / (mc_kernel_ms * 1e-3)   # per GPU

    # e2e: public API, host in / host out, synchronous
    def mc_e2e_step(i):
        lp = pr.gbm_paths(c["S0"], c["r"], c["q"], c["sigma"], c["T"], MC_STEPS, n_base * world, seed=777 + i)
        return dd.price_payoffs(lp, [spec], c["r"], c["T"])[0]
    for i in range(2):
        mc_e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_price = mc_e2e_step(10 + i)
    barrier()
    mc_e2e_s = max_over_ranks((time.perf_counter() - t0) / K)
    mc_e2e = {"value": path_steps_per_step / mc_e2e_s, "unit": "path-steps/s",
              "h2d_bytes_per_step": C.sizeof(nv.OptpProcess) + C.sizeof(nv.OptpPayoff),
              "d2h_bytes_per_step": 8 * nv.NSTATS, "ms_per_step": 1e3 * mc_e2e_s,
              "api": "barrier_price(gbm_paths(...)) -> optp_mc_path (host stats out, synchronous)",
              "price": e2e_price[0], "stderr": e2e_price[1]}

    # ================= CRR: config 2, American put, N = 1000 ==========================================
    n_c = args.crr_contracts
    cols = cfg2_contracts(n_c, seed=rank)
    dcols = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cols]
    price_t = torch.empty(n_c, dtype=torch.float64, device=dev)
    status_t = torch.empty(n_c, dtype=torch.uint8, device=dev)

    def crr_step():
        nv.check(lib.optp_crr_batch_dev(*[C.c_void_p(t.data_ptr()) for t in dcols], n_c, CRR_N, 1,
                                        C.c_void_p(price_t.data_ptr()), C.c_void_p(status_t.data_ptr()), sptr),
                 "optp_crr_batch_dev")
    crr_w = max(1, min(W, 3))
    for _ in range(crr_w):
        flush.zero_()
        crr_step()
    barrier()
    cev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    crr_launch0 = nv.launch_count()
    for i in range(K):
        flush.zero_()
        cev[i][0].record()
        crr_step()
        cev[i][1].record()
    barrier()
    crr_launches = nv.launch_count() - crr_launch0
    crr_ms = max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in cev) / K)
    crr_value = n_c * world / (crr_ms * 1e-3)
    crr_rate_gpu = n_c / (crr_ms * 1e-3)
    crr_check = float(price_t[:8].sum().item())
    # the same batch through the full-width kernel only (every node of every level is updated):
    # the honest pipe-utilisation figure, since the banded kernel skips provably irrelevant nodes
    nv.set_option("crr_band", 0)
    crr_step()
    torch.cuda.synchronize()
    fev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush.zero_()
    fev[0].record()
    crr_step()
    fev[1].record()
    torch.cuda.synchronize()
    nv.set_option("crr_band", 1)
    crr_full_ms = max_over_ranks(fev[0].elapsed_time(fev[1]))
    crr_full_rate = n_c / (crr_full_ms * 1e-3)
    crr_full_check = float(price_t[:8].sum().item())

    # e2e: pinned host columns in, host prices out
    def pinned(a):
        p = C.c_void_p()
        nv.check(lib.optp_host_alloc(C.byref(p), a.nbytes), "optp_host_alloc")
        buf = (C.c_char * a.nbytes).from_address(p.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size)
        out[:] = a
        return out, p
    hcols = [pinned(np.ascontiguousarray(a)) for a in cols]
    harrs = [h[0] for h in hcols]
    for _ in range(2):
        bn.crr_batch(*harrs, CRR_N, american=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_price = bn.crr_batch(*harrs, CRR_N, american=True)
    barrier()
    crr_e2e_s = max_over_ranks((time.perf_counter() - t0) / K)
    crr_e2e = {"value": n_c * world / crr_e2e_s, "unit": "options/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in harrs)), "d2h_bytes_per_step": 9 * n_c,
               "ms_per_step": 1e3 * crr_e2e_s, "api": "crr_batch(host arrays) -> optp_crr_batch",
               "checksum_first8": float(host_price[:8].sum())}
    for _, p in hcols:
        lib.optp_host_free(p)

    # ================= other BASELINE configurations at bounded size ===================================
    workloads = []
    if not args.skip_extra:
        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(reps):
                flush.zero_()
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best

        def path_rate(p, specs, n_steps, nb, r_, T_):
            pa = ex._native_payoffs(specs, r_, T_)
            out = torch.zeros(len(specs) * nv.NSTATS, dtype=torch.float64, device=dev)
            ms = timed(lambda: nv.check(lib.optp_mc_path_dev(C.byref(p), pa, len(specs), n_steps, nb, 1,
                                                            nv.MODE_FP32_PHILOX, 99, rank * nb, None, None, None,
                                                            C.c_void_p(out.data_ptr()), sptr), "optp_mc_path_dev"))
            return 2.0 * nb * n_steps / (ms * 1e-3), ms, out.cpu().numpy().reshape(len(specs), nv.NSTATS)

        # cfg 4: Heston FTE, 365 steps, arithmetic Asian call, antithetic + control variate
        hp = nv.OptpProcess()
        hp.kind = nv.PROC_HESTON_FTE
        hp.S0, hp.r, hp.q, hp.T = 100.0, 0.05, 0.0, 1.0
        hp.v0, hp.kappa, hp.theta, hp.xi, hp.rho = 0.04, 2.0, 0.04, 0.3, -0.7
        nb = 1 << 23
        rate, ms, s = path_rate(hp, [ex.asian_spec(100.0, "call")], 365, nb, 0.05, 1.0)
        pcv = ex.estimate(s[0], control_variate=True, expected_y=100.0)
        workloads.append({"workload": "cfg4 Heston FTE 365 steps, arithmetic Asian call, antithetic+CV",
                          "n_paths_effective": 2 * nb, "value": rate, "unit": "path-steps/s/GPU", "ms": ms,
                          "fp32_pipe_frac": rate * OPS_PER_PATH_STEP["heston_asian"] / peaks["fp32_ffma_lane_ops"],
                          "price": pcv[0], "stderr": pcv[1]})
        # cfg 4, QE variant (no reference code -- parity unpinned, checked against oracle/heston_qe):
        # the scheme is meant for coarse grids, 32 steps here
        hq = nv.OptpProcess()
        C.memmove(C.byref(hq), C.byref(hp), C.sizeof(nv.OptpProcess))
        hq.kind = nv.PROC_HESTON_QE
        rate, ms, s = path_rate(hq, [ex.asian_spec(100.0, "call")], 32, nb, 0.05, 1.0)
        pq = ex.estimate(s[0], control_variate=True, expected_y=100.0)
        workloads.append({"workload": "cfg4 variant: Heston QE (Andersen 2008) 32 steps, arithmetic Asian call, antithetic+CV",
                          "n_paths_effective": 2 * nb, "value": rate, "unit": "path-steps/s/GPU", "ms": ms,
                          "price": pq[0], "stderr": pq[1], "parity": "unpinned (no reference implementation)"})
        # cfg 3: LV Milstein on the Dupire table, 252 steps, up-and-out barrier call
        tab_file = os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz")
        if os.path.exists(tab_file):
            z = np.load(tab_file)
            table = np.ascontiguousarray(z["table"], dtype=np.float64)
            nv.check(lib.optp_sigma_table_upload(nv.ptr(table), table.shape[0], table.shape[1]), "table upload")
            lp_ = nv.OptpProcess()
            lp_.kind = nv.PROC_LV_MILSTEIN
            lp_.S0, lp_.r, lp_.q, lp_.T, lp_.dS_bump = 100.0, 0.05, 0.02, 1.0, 0.01
            lp_.n_nodes, lp_.x0, lp_.h = table.shape[1], float(z["x0"]), float(z["h"])
            nb = 1 << 23
            rate, ms, s = path_rate(lp_, [ex.barrier_spec(100.0, "call", 130.0, "up-and-out")], 252, nb, 0.05, 1.0)
            pb = ex.estimate(s[0])
            workloads.append({"workload": "cfg3 LV Milstein on Dupire table 252x256, up-and-out call B=130",
                              "n_paths_effective": 2 * nb, "value": rate, "unit": "path-steps/s/GPU", "ms": ms,
                              "fp32_pipe_frac": rate * OPS_PER_PATH_STEP["lvm_barrier"] / peaks["fp32_ffma_lane_ops"],
                              "price": pb[0], "stderr": pb[1], "clamped_lookups": float(s[0][6])})
        # cfg 1 throughput variant: terminal-only European call, 2^30 base paths
        tstats = torch.zeros(nv.NSTATS, dtype=torch.float64, device=dev)
        nb = 1 << 30
        ms = timed(lambda: nv.check(lib.optp_mc_terminal_dev(100.0, 110.0, 1.0, 0.03, 0.0, 0.2, 1, nb, 1,
                                                             nv.MODE_FP32_PHILOX, 0, 5, rank * nb, None,
                                                             C.c_void_p(tstats.data_ptr()), sptr), "terminal"))
        ts = tstats.cpu().numpy()
        pt = mc.estimator((int(ts[0]), *ts[1:6]), S0=100.0, q=0.0, T=1.0)
        rate = 2.0 * nb / (ms * 1e-3)
        workloads.append({"workload": "cfg1 contract, terminal-only euro_price_mc kernel, 2^31 effective paths",
                          "value": rate, "unit": "path-steps/s/GPU", "ms": ms,
                          "fp32_pipe_frac": rate * OPS_PER_PATH_STEP["terminal"] / peaks["fp32_ffma_lane_ops"],
                          "price": pt[0], "stderr": pt[1]})
        # cfg 5 at BASELINE size: 10k-contract book (1/3 European, 1/3 barrier, 1/3 arithmetic Asian)
        # through price_book, 5e6 base paths (1e7 effective) per contract; items are dealt over the ranks
        from optpricer_b200 import book as bk
        n_rows, book_paths = 10_000, 5_000_000
        rows = bk.synthetic_book(n_rows, n_paths=book_paths)
        bk.price_book(rows[:200], mc_paths=book_paths)
        barrier()
        t0 = time.perf_counter()
        priced = bk.price_book(rows, mc_paths=book_paths)
        barrier()
        book_s = max_over_ranks(time.perf_counter() - t0)
        steps_of = lambda r: 1 if r["method"] == "mc" else r["n_steps"]  # noqa: E731
        book_path_steps = sum(2.0 * book_paths * steps_of(r) for r in rows)
        workloads.append({"workload": "cfg5 book: 10k contracts (European / barrier / arithmetic Asian), price_book(), "
                                      f"{2 * book_paths} effective paths each, 252 steps, items dealt over the ranks (book.shard_items)",
                          "value": book_path_steps / book_s, "unit": "path-steps/s (all GPUs, end to end)",
                          "contracts_per_s": n_rows / book_s, "seconds": book_s,
                          "priced": sum(1 for r in priced if r.get("price") is not None),
                          "checksum": float(sum(r["price"] for r in priced[:100]))})
        # the bump-and-reprice caller (risk.py:146-203) over a 20k-instrument book,
        # 8 reprices each, N=500 American trees, as ONE device batch
        from optpricer_b200 import risk as rk
        rS0, rK, rT, rr, rq, rsig, _ = cfg2_contracts(20_000, seed=3)
        book_r = [dict(S=float(rS0[i]), K=float(rK[i]), T=float(rT[i]), r=float(rr[i]), q=0.0, sigma=float(rsig[i]),
                       kind="put", position=1.0) for i in range(20_000)]
        tp = rk.TreePricer(500, american=True)
        rk.portfolio_risk(book_r[:100], tp)
        t0 = time.perf_counter()
        pr_ = rk.portfolio_risk(book_r, tp)
        risk_s = time.perf_counter() - t0
        workloads.append({"workload": "portfolio_risk (risk.py:146-203) on 20k American puts, N=500 trees, 8 reprices each, one batch",
                          "value": 8 * 20_000 / risk_s, "unit": "tree reprices/s (end to end, incl. Python)",
                          "seconds": risk_s, "total_delta": pr_["total_delta"], "total_value": pr_["total_value"]})
        # American calls on a dividend-paying asset (q < r): priced as dual puts by the two-zone kernel
        from optpricer_b200 import binomial as bn
        dS0, dK, dT, dr, _, dsig, _ = cfg2_contracts(400_000, seed=7)
        dq = np.random.default_rng(8).uniform(0.005, 0.025, dS0.size)
        dcall = np.ones(dS0.size, dtype=np.uint8)
        bn.crr_batch(dS0, dK, dT, dr, dq, dsig, dcall, CRR_N, american=True)
        t0 = time.perf_counter()
        dprice = bn.crr_batch(dS0, dK, dT, dr, dq, dsig, dcall, CRR_N, american=True)
        dual_s = time.perf_counter() - t0
        workloads.append({"workload": "CRR American CALLS with dividends (config-2 contracts as calls, q~U(0.5%,2.5%), r=3%), N=1000, "
                                      "crr_batch(host arrays): dual-put route of the two-zone kernel",
                          "value": dS0.size / dual_s, "unit": "options/s (end to end)", "seconds": dual_s,
                          "checksum_first8": float(np.sum(dprice[:8]))})
        # cfg 1 as is: latency of one euro_price_mc call (200k paths) through the public API
        opt = __import__("optpricer_b200").OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
        mc.euro_price_mc(opt, "call", n_paths=200_000, seed=1)
        t0 = time.perf_counter()
        for _ in range(20):
            lat_p = mc.euro_price_mc(opt, "call", n_paths=200_000, seed=1)
        workloads.append({"workload": "cfg1 euro_price_mc(n_paths=200k, seed=1) public API latency",
                          "value": (time.perf_counter() - t0) / 20 * 1e6, "unit": "us/call", "price": lat_p[0],
                          "stderr": lat_p[1]})

    total_launches = nv.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    cpu_mc = cpu_crr = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_mc, cpu_crr = cpu_baseline_single_core()

    if rank == 0:
        measured = {}
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            with open(mp) as f:
                measured = json.load(f)
        p32 = peaks["fp32_ffma_lane_ops"]
        line = {
            "metric": "mc_path_steps_per_sec", "value": mc_value, "unit": "path-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": mc_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 path state, f64 accumulators", "data": "synthetic",
            "config": {"workload": "fused GBM exact log-Euler paths, 252 steps, up-and-out barrier call "
                                   "(S0=100,K=100,B=130,r=0.05,sigma=0.2,T=1), antithetic, Philox4x32-10 in registers",
                       "base_paths_per_gpu_per_step": n_base, "effective_paths_per_step": 2 * n_base * world,
                       "l2": "flushed between steps (256 MiB write); the kernel has no HBM inputs",
                       "collective": "one NCCL all-reduce of the 8-double statistics block per step" if world > 1 else "none"},
            "price": mc_price[0], "stderr": mc_price[1], "wall_ms_per_step": 1e3 * mc_wall / K,
            "e2e": mc_e2e,
            "roofline": {"bound": "fp32_pipe", "achieved": mc_kernel_rate * OPS_PER_PATH_STEP["gbm_barrier"] / 1e12,
                         "peak": p32 / 1e12, "unit": "T lane-ops/s",
                         "frac": mc_kernel_rate * OPS_PER_PATH_STEP["gbm_barrier"] / p32,
                         "traffic": 105984,  # ncu --set full of THIS launch (profiles/r1_ncu_full_bench_gbm.md): dram read 106 KB, write 0
                         # the same figure against the nominal issue rate: 148 SMs x 128 FP32 lanes x the SM clock
                         # sampled during the timed region (the measured FFMA rate above is ~91 % of it)
                         "peak_nominal": 148 * 128 * (clocks or {}).get("sm_mhz", 1965.0) * 1e6 / 1e12
                         if (clocks or {}).get("sm_mhz") else 148 * 128 * 1965.0e6 / 1e12,
                         "frac_nominal": mc_kernel_rate * OPS_PER_PATH_STEP["gbm_barrier"]
                         / (148 * 128 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6),
                         # secondary bound: one MUFU per effective path-step against the measured MUFU.EX2 rate
                         "mufu_frac": mc_kernel_rate * 1.0 / peaks["mufu_ex2_ops"],
                         "ops_per_path_step": OPS_PER_PATH_STEP["gbm_barrier"], "kernel_ms": mc_kernel_ms,
                         "path_steps_per_s_per_gpu": mc_kernel_rate,
                         "peak_source": "optp_microbench FFMA issue rate measured in this run (MEASURED_PEAKS.json "
                                        "has no FP32-pipe figure); compute-bound kernel, HBM traffic ~0"},
            "cpu_baseline": cpu_mc,
            "crr": {
                "metric": "crr_american_N1000_options_per_sec", "value": crr_value, "unit": "options/s",
                "ms_per_step": crr_ms, "dtype": "f64",
                "config": {"workload": "BASELINE config 2: American put, N=1000, synthetic contracts "
                                       "(default_rng: S0~U(50,150), K=S0*U(.8,1.2), T~U(.1,2), sigma~U(.1,.6), r=.03)",
                           "contracts_per_gpu_per_step": n_c, "l2": "flushed between steps"},
                "e2e": crr_e2e,
                "roofline": {"bound": "smem", "achieved": crr_rate_gpu * CRR_SMEM_BYTES_PER_OPTION / 1e9,
                             "peak": peaks["smem_lds64_bytes"] / 1e9, "unit": "GB/s",
                             "frac": crr_rate_gpu * CRR_SMEM_BYTES_PER_OPTION / peaks["smem_lds64_bytes"],
                             # ncu --set full of this launch at 1M contracts (profiles/r1_ncu_full_bench_zone.md):
                             # dram read 49.47 MB + write 0.48 MB (the 57 MB of contract columns and results)
                             "traffic": 49_949_440 if n_c == 1_000_000 else None,
                             "ncu_pipe_pct": {"fp64": 48.9, "issue_slots": 58.3, "smem_wavefronts": 29.9,
                                              "source": "profiles/r1_ncu_full_bench_zone.md"},
                             "bytes_per_option": CRR_SMEM_BYTES_PER_OPTION,
                             "fp64_frac": crr_rate_gpu * CRR_FP64_OPS_PER_OPTION / peaks["fp64_dfma_lane_ops"],
                             "note": "ALGORITHMIC figures (500 500 node updates x 24 B / x 5 FP64 ops per option) over "
                                     "measured LDS.64 bandwidth / DFMA rate. They exceed 1 because (a) tree levels live in "
                                     "registers, not shared memory, and (b) the two-zone banded kernel updates only the nodes "
                                     "between the early-exercise boundary and the 2^-80 K zero region (about 140k of the 500k "
                                     "node updates; only the 32 nodes next to the boundary take the exercise test, the rest "
                                     "cost one DFMA), four options per warp, with prices equal to the full-width kernel to "
                                     "1e-10 (tests/test_gpu_parity.py::test_crr_banded_kernel_equals_full_width_kernel). The "
                                     "physical pipe fraction is `full_width.fp64_frac`.",
                             "peak_source": "optp_microbench LDS.64 bandwidth / DFMA rate measured in this run"},
                "full_width": {"value": crr_full_rate * world, "unit": "options/s", "ms_per_step": crr_full_ms,
                               "fp64_frac": crr_full_rate * CRR_FP64_OPS_PER_OPTION / peaks["fp64_dfma_lane_ops"],
                               "smem_algorithmic_frac": crr_full_rate * CRR_SMEM_BYTES_PER_OPTION / peaks["smem_lds64_bytes"],
                               "checksum_first8": crr_full_check,
                               "what": "optp_set_option('crr_band', 0): every node of every level updated in registers"},
                "cpu_baseline": cpu_crr, "gpu_launches": crr_launches, "checksum_first8": crr_check,
            },
            "workloads": workloads,
            "peaks_measured_in_run": peaks,
            "measured_peaks_file": {k: measured.get(k) for k in ("hbm_gbs", "bf16_tflops", "sm_max_mhz")},
            "clocks": clocks,
            "gpu_launches": mc_launches + crr_launches,
            "gpu_launches_total_in_run": total_launches,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
