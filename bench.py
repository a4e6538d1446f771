#!/usr/bin/env python3
"""Benchmark of the B200 pricing hot paths (BASELINE.json metric: MC path-steps/s and CRR
options/s at N=1000 American).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on rank 0.  The headline `metric` is Monte Carlo path-steps/s on the fused GBM
path kernel (252-step up-and-out barrier call, antithetic, Philox in registers).  The second half of
BASELINE's metric, CRR options/s on config 2 (1M American puts, N=1000), is the compact `crr` object
that ENDS the line (so that it survives a truncated tail), preceded by `strong` (fixed-total-work
time-to-price of config 4 and of the config-5 book on the N GPUs); `crr_detail` and `workloads` carry
the long form and the other BASELINE configurations.

  value  device-resident: kernels launched through the *_dev C-ABI entry points on torch's
         current stream, timed with CUDA events on that stream, L2 flushed between steps.
  e2e    the same work through the public Python API (host buffers in, host results out:
         optp_mc_path / optp_crr_batch stage, launch, copy back and synchronise), wall clock
         between device synchronisations.
  --impl reference: the UNMODIFIED reference (`optpricer` imported from baseline/_ref, else from
         /root/reference/src; `kind: "reference"`) on all host cores; only when neither is present the
         numpy oracle restatement (oracle/, `kind: "port"`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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
CRR_NODES_PER_OPTION = 500_500
N_SM, FP32_LANES, FP64_LANES = 148, 128, 64
REF_DIRS = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src")


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


# ---- CPU arm: the unmodified reference when it is present, else the oracle port --------------------
_REF = None


def reference_module():
    """(`optpricer` module, where it came from) or (None, None): baseline/_ref (the pip --target install
    that travels to the GPU box), then /root/reference/src (build container only)."""
    global _REF
    if _REF is None:
        _REF = (None, None)
        for d in REF_DIRS:
            if os.path.isdir(os.path.join(d, "optpricer")):
                sys.path.insert(0, d)
                try:
                    import optpricer
                    import optpricer.processes  # noqa: F401  (gbm_paths / heston_paths are not re-exported)
                    _REF = (optpricer, os.path.relpath(d, ROOT) if d.startswith(ROOT) else d)
                    break
                except Exception:
                    sys.path.remove(d)
    return _REF


def cpu_kind():
    return "reference" if reference_module()[0] is not None else "port"


def _cpu_mc_task(args):
    """gbm_paths + barrier_price, one call: (path-steps, seconds, price)."""
    seed, n_steps, n_base = args
    c = MC_CONTRACT
    ref, _ = reference_module()
    t0 = time.perf_counter()
    if ref is not None:    # processes.py:30-56 + exotics.py:33-96, unmodified
        paths = ref.processes.gbm_paths(c["S0"], c["r"], c["q"], c["sigma"], c["T"], n_steps, n_base, antithetic=True, seed=seed)
        price = ref.barrier_price(paths, c["K"], c["r"], c["T"], "call", c["barrier"], "up-and-out")
    else:
        from oracle import path_schemes, payoffs
        (z,) = path_schemes.draw_normals(seed, n_steps, n_base)
        paths = path_schemes.gbm_exact(z, S0=c["S0"], r=c["r"], q=c["q"], sigma=c["sigma"], T=c["T"])
        price = payoffs.barrier(paths, c["K"], c["r"], c["T"], "call", c["barrier"], "up-and-out")
    return 2 * n_base * n_steps, time.perf_counter() - t0, price[0]


def _cpu_crr_task(args):
    """crr(N=1000, american=True) looped over config-2 contracts [lo, hi): (options, seconds, first price)."""
    lo, hi = args
    S0, K, T, r, q, sig, is_call = cfg2_contracts(hi)
    ref, _ = reference_module()
    t0 = time.perf_counter()
    if ref is not None:    # binomial.py:7-38, unmodified (the reference has no batched heterogeneous API)
        p = [ref.crr(ref.OptionSpec(S0=float(S0[i]), K=float(K[i]), T=float(T[i]), r=float(r[i]), sigma=float(sig[i]),
                                    q=float(q[i])), "put", CRR_N, american=True) for i in range(lo, hi)]
    else:
        from oracle import crr_tree
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
    """The other BASELINE configurations (SURVEY 8d CPU baseline (i) and (iii)) on the reference (or the
    port): returns (path-steps, seconds) of one bounded call."""
    what, seed = args
    ref, _ = reference_module()
    from oracle import sigma_grid   # the table callable handed to the reference's Milstein generator
    t0 = time.perf_counter()
    if what == "terminal":     # cfg 1: euro_price_mc, chunks of 100k as the reference streams them
        n_paths = 2_000_000
        if ref is not None:
            ref.euro_price_mc(ref.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2), "call", n_paths=n_paths, seed=seed)
        else:
            from oracle import terminal_mc
            terminal_mc.euro_mc(100.0, 110.0, 1.0, 0.03, 0.2, 0.0, "call", n_paths=n_paths, seed=seed)
        units = 2 * n_paths
    elif what == "heston_asian":   # cfg 4: heston_paths + asian_price, 365 steps
        n_steps, n_base = 365, 10_000
        if ref is not None:
            paths = ref.processes.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, n_steps, n_base, seed=seed)
            ref.asian_price(paths, 100.0, 0.05, 1.0, "call")
        else:
            from oracle import path_schemes, payoffs
            z2, zp = path_schemes.draw_normals(seed, n_steps, n_base, 2)
            paths = path_schemes.heston_fte(z2, zp, S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3,
                                            rho=-0.7, T=1.0)
            payoffs.asian(paths, 100.0, 0.05, 1.0, "call")
        units = 2 * n_base * n_steps
    else:                          # cfg 3: milstein_local_vol_paths on the tabulated Dupire surface + barrier_price
        n_steps, n_base = 252, 10_000
        z = np.load(os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz"))
        grid = sigma_grid.SigmaGrid(z["table"], float(z["x0"]), float(z["h"]), float(z["dt"]))
        if ref is not None:
            paths = ref.processes.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, n_steps, n_base, grid, seed=seed)
            ref.barrier_price(paths, 100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
        else:
            from oracle import path_schemes, payoffs
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


def _cpu_label():
    mod, where = reference_module()
    if mod is not None:
        return f"unmodified reference optpricer {getattr(mod, '__version__', '?')} from {where}"
    return "oracle numpy port (the reference is not present on this box)"


def cpu_baseline_single_core():
    """Scalar (1 core) CPU run, ~10-20 s: the `cpu_baseline` objects of the b200 line."""
    kind, label = cpu_kind(), _cpu_label()
    n, dt, _ = _cpu_mc_task((1, MC_STEPS, 20_000))
    reps = max(1, int(6.0 / max(dt, 1e-3)))
    tot_n, tot_t = n, dt
    for i in range(reps - 1):
        n, dt, _ = _cpu_mc_task((2 + i, MC_STEPS, 20_000))
        tot_n += n
        tot_t += dt
    mc = {"value": tot_n / tot_t, "unit": "path-steps/s", "cores": 1, "kind": kind,
          "sample": f"{label}: gbm_paths + barrier_price, {reps} calls of 252 steps x 20000 base paths (+mirrors)"}
    n, dt, _ = _cpu_crr_task((0, 160))
    crr = {"value": n / dt, "unit": "options/s", "cores": 1, "kind": kind,
           "sample": f"{label}: crr(N=1000, american=True) looped over the first 160 config-2 contracts"}
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
    kind, label = cpu_kind(), _cpu_label()
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
        for what, label_w in (("terminal", "cfg1 euro_price_mc (terminal GBM, antithetic + CV), 2M paths per process"),
                              ("heston_asian", "cfg4 heston_paths + asian_price, 365 steps x 10k base paths per process"),
                              ("lvm_barrier", "cfg3 milstein_local_vol_paths (tabulated Dupire sigma) + barrier_price, "
                                              "252 steps x 10k base paths per process")):
            try:
                t0 = time.perf_counter()
                res = list(pool.map(_cpu_extra_task, [(what, 7 + w) for w in range(workers)]))
                extra.append({"workload": label_w, "value": sum(x[0] for x in res) / (time.perf_counter() - t0),
                              "unit": "path-steps/s", "cores": workers, "kind": kind})
            except Exception as e:
                extra.append({"workload": label_w, "unavailable": str(e)[:200]})
        # crr_vec, many strikes on one shared tree (binomial.py:44-109): the reference's own batched tree API
        vec = None
        ref, _ = reference_module()
        if ref is not None:
            t0 = time.perf_counter()
            pv = ref.crr_vec(100.0, np.linspace(60.0, 140.0, 1000), 1.0, 0.03, 0.0, 0.2, "put", N=1000, american=True)
            dtv = time.perf_counter() - t0
            vec = {"workload": "crr_vec(S0=100, 1000 strikes in [60,140], T=1, r=.03, sigma=.2, put, N=1000, american=True)",
                   "value": 1000 / dtv, "unit": "options/s", "seconds": dtv, "cores": 1, "kind": "reference",
                   "checksum": float(np.sum(pv))}
            extra.append(vec)
    mc_v, crr_v = u_mc / t_mc, u_crr / t_crr
    sample_mc = (f"{workers} processes x {mc_calls} calls of (252 steps x {mc_base} base paths + mirrors) per step; {label}; "
                 "the process pool is harness-level (the reference parallelises euro_price_mc only)")
    sample_crr = f"{workers} processes x {crr_per} config-2 contracts per step; {label}"
    line = {
        "impl": "reference", "metric": "mc_path_steps_per_sec", "value": mc_v, "unit": "path-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_mc / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gbm_paths 252 steps + barrier_price up-and-out call (S0=100,K=100,B=130), antithetic",
                   "sample": sample_mc},
        "cpu_baseline": {"value": mc_v, "unit": "path-steps/s", "cores": workers, "kind": kind, "sample": sample_mc},
        "e2e": {"value": mc_v, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "workloads": extra,
        "crr": {"metric": "crr_american_N1000_options_per_sec", "value": crr_v, "unit": "options/s",
                "ms_per_step": 1e3 * t_crr / args.steps,
                "cpu_baseline": {"value": crr_v, "unit": "options/s", "cores": workers, "kind": kind,
                                 "sample": sample_crr},
                "optimised_c": {"value": crr_c_v, "unit": "options/s", "cores": workers, "kind": "port",
                                "sample": f"{workers} processes x 2000 config-2 contracts, oracle/crr_tree.c table-based "
                                          "loop (not the reference's numpy path)"},
                "e2e": {"value": crr_v, "unit": "options/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
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
    ap.add_argument("--skip-extra", action="store_true", help="skip the `workloads` list and the strong-scaling jobs")
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
    if world > 1:
        dd.init_comm()   # the library's own NCCL communicator: statistics are reduced by optp_allreduce_stats[_dev]
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

    def allreduce_dev(t):   # in place, on the stream the kernels run on, no host round trip
        if world > 1:
            nv.check(lib.optp_allreduce_stats_dev(C.c_void_p(t.data_ptr()), t.numel(), sptr), "optp_allreduce_stats_dev")

    # pipe rates measured in this run (MEASURED_PEAKS.json carries only HBM and bf16)
    peaks = {}
    for which, name in ((0, "fp32_ffma_lane_ops"), (1, "mufu_ex2_ops"), (2, "fp64_dfma_lane_ops"),
                        (3, "smem_lds64_bytes"), (4, "imad_wide_lane_ops")):
        peaks[name] = nv.microbench(which)[0]

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

    def mc_step(i, mode=nv.MODE_FP32_PHILOX, nb=None):
        nb = n_base if nb is None else nb
        off = (i * world + rank) * nb
        nv.check(lib.optp_mc_path_dev(C.byref(proc), pay, 1, MC_STEPS, nb, 1, mode, 20261018,
                                      off, None, None, None, C.c_void_p(stats_t.data_ptr()), sptr), "optp_mc_path_dev")

    for i in range(W):
        flush.zero_()
        mc_step(i)
        allreduce_dev(stats_t)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    wall0 = time.perf_counter()
    mc_launch0 = nv.launch_count()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        mc_step(W + i)
        ev[i][1].record()
        allreduce_dev(stats_t)
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

    # the same kernel contract in the reference's precision: fp64 path arithmetic (reference operation order,
    # -fmad=false) on the same Philox draws -- the like-for-like figure next to the fp32 headline
    nb64 = max(1 << 16, n_base >> 4)
    ms64 = timed(lambda: mc_step(1, nv.MODE_FP64_PHILOX, nb64), reps=2)
    fp64_rate = 2.0 * nb64 * MC_STEPS / (ms64 * 1e-3)
    st64 = stats_t.cpu().numpy()
    mc_step(1, nv.MODE_FP32_PHILOX, nb64)
    torch.cuda.synchronize()
    st32 = stats_t.cpu().numpy()
    fp64_mode = {"value": fp64_rate, "unit": "path-steps/s/GPU", "ms": ms64, "n_paths_effective": 2 * nb64,
                 "price_fp64": ex.estimate(st64)[0], "price_fp32_same_draws": ex.estimate(st32)[0],
                 "stderr": ex.estimate(st64)[1],
                 "what": "OPTP_MODE_FP64_PHILOX: fp64 path state in the reference's operation order, same Philox/Box-Muller "
                         "normals as the fp32 kernel (the normals themselves are built in fp32)"}

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
              "api": "distributed.price_payoffs(gbm_paths(...), [barrier_spec], r, T) -> exotics.path_stats -> optp_mc_path "
                     "(descriptors in, host statistics out, synchronous)"
                     + (" + optp_allreduce_stats over the ranks" if world > 1 else "")
                     + "; barrier_price(gbm_paths(...)) reaches the same optp_mc_path call",
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
    # executed work of the same batch: the diagnostic build of the same kernel counts the FP64-pipe
    # lane-instructions and node updates of its level loops (not timed)
    nv.set_option("crr_count_ops", 1)
    crr_step()
    torch.cuda.synchronize()
    crr_fp64_exec = nv.get_counter("crr_fp64_lane_ops") / n_c
    crr_nodes_exec = nv.get_counter("crr_node_updates") / n_c
    crr_count_check = float(price_t[:8].sum().item())
    nv.set_option("crr_count_ops", 0)
    # the same batch through the full-width kernel only (every node of every level is updated)
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

    # ================= other BASELINE configurations, at BASELINE size ===================================
    workloads = []
    strong = None
    if not args.skip_extra:
        def path_job(p, specs, n_steps, nb_total, r_, T_, seed=99, reps=3):
            """A FIXED-total-size job: nb_total base paths dealt over the ranks (strong scaling), statistics summed by
            optp_allreduce_stats_dev; returns (time-to-price in ms = max over ranks, total statistics)."""
            pa = ex._native_payoffs(specs, r_, T_)
            out = torch.zeros(len(specs) * nv.NSTATS, dtype=torch.float64, device=dev)
            lo, hi = dd.shard_range(nb_total, rank, world)

            def run():
                nv.check(lib.optp_mc_path_dev(C.byref(p), pa, len(specs), n_steps, hi - lo, 1, nv.MODE_FP32_PHILOX, seed,
                                              lo, None, None, None, C.c_void_p(out.data_ptr()), sptr), "optp_mc_path_dev")
                allreduce_dev(out)
            barrier()
            ms = max_over_ranks(timed(run, reps))
            return ms, out.cpu().numpy().reshape(len(specs), nv.NSTATS)

        # cfg 4 at BASELINE size: Heston FTE, 365 steps, 1e9 effective paths, arithmetic Asian call, antithetic + CV
        hp = nv.OptpProcess()
        hp.kind = nv.PROC_HESTON_FTE
        hp.S0, hp.r, hp.q, hp.T = 100.0, 0.05, 0.0, 1.0
        hp.v0, hp.kappa, hp.theta, hp.xi, hp.rho = 0.04, 2.0, 0.04, 0.3, -0.7
        nb4 = 500_000_000
        ms4, s = path_job(hp, [ex.asian_spec(100.0, "call")], 365, nb4, 0.05, 1.0, seed=42, reps=2)
        pcv = ex.estimate(s[0], control_variate=True, expected_y=100.0)
        rate4 = 2.0 * nb4 * 365 / (ms4 * 1e-3)
        workloads.append({"workload": "cfg4 at BASELINE size: Heston FTE 365 steps x 1e9 effective paths, arithmetic Asian "
                                      "call, antithetic+CV; base paths dealt over the ranks (fixed total work)",
                          "n_paths_effective": 2 * nb4, "value": rate4, "unit": "path-steps/s (all GPUs)", "seconds": ms4 * 1e-3,
                          "fp32_pipe_frac_per_gpu": rate4 / world * OPS_PER_PATH_STEP["heston_asian"]
                          / (N_SM * FP32_LANES * 1965.0e6),
                          "price": pcv[0], "stderr": pcv[1]})
        # cfg 4, QE variant (no reference code -- parity unpinned, checked against oracle/heston_qe):
        # the scheme is meant for coarse grids, 32 steps here
        hq = nv.OptpProcess()
        C.memmove(C.byref(hq), C.byref(hp), C.sizeof(nv.OptpProcess))
        hq.kind = nv.PROC_HESTON_QE
        nbq = 1 << 23
        msq, s = path_job(hq, [ex.asian_spec(100.0, "call")], 32, nbq * world, 0.05, 1.0)
        pq = ex.estimate(s[0], control_variate=True, expected_y=100.0)
        workloads.append({"workload": "cfg4 variant: Heston QE (Andersen 2008) 32 steps, arithmetic Asian call, antithetic+CV",
                          "n_paths_effective": 2 * nbq * world, "value": 2.0 * nbq * 32 / (msq * 1e-3), "unit": "path-steps/s/GPU",
                          "ms": msq, "price": pq[0], "stderr": pq[1], "parity": "unpinned (no reference implementation)"})
        # cfg 3 at BASELINE size: LV Milstein on the Dupire table, 252 steps, 1e8 effective paths, up-and-out barrier call
        tab_file = os.path.join(ROOT, "tests", "golden", "dupire_table_252x256.npz")
        ms3 = None
        if os.path.exists(tab_file):
            z = np.load(tab_file)
            table = np.ascontiguousarray(z["table"], dtype=np.float64)
            nv.check(lib.optp_sigma_table_upload(nv.ptr(table), table.shape[0], table.shape[1]), "table upload")
            lp_ = nv.OptpProcess()
            lp_.kind = nv.PROC_LV_MILSTEIN
            lp_.S0, lp_.r, lp_.q, lp_.T, lp_.dS_bump = 100.0, 0.05, 0.02, 1.0, 0.01
            lp_.n_nodes, lp_.x0, lp_.h = table.shape[1], float(z["x0"]), float(z["h"])
            nb3 = 50_000_000
            ms3, s = path_job(lp_, [ex.barrier_spec(100.0, "call", 130.0, "up-and-out")], 252, nb3, 0.05, 1.0, seed=42)
            pb = ex.estimate(s[0])
            rate3 = 2.0 * nb3 * 252 / (ms3 * 1e-3)
            workloads.append({"workload": "cfg3 at BASELINE size: LV Milstein on the Dupire table 252x256, 252 steps x 1e8 "
                                          "effective paths, up-and-out call B=130; base paths dealt over the ranks",
                              "n_paths_effective": 2 * nb3, "value": rate3, "unit": "path-steps/s (all GPUs)", "seconds": ms3 * 1e-3,
                              "fp32_pipe_frac_per_gpu": rate3 / world * OPS_PER_PATH_STEP["lvm_barrier"]
                              / (N_SM * FP32_LANES * 1965.0e6),
                              "price": pb[0], "stderr": pb[1], "clamped_paths": float(s[0][6])})
        # cfg 1 throughput variant: terminal-only European call, 2^30 base paths per GPU
        tstats = torch.zeros(nv.NSTATS, dtype=torch.float64, device=dev)
        nb = 1 << 30
        ms = timed(lambda: nv.check(lib.optp_mc_terminal_dev(100.0, 110.0, 1.0, 0.03, 0.0, 0.2, 1, nb, 1,
                                                             nv.MODE_FP32_PHILOX, 0, 5, rank * nb, None,
                                                             C.c_void_p(tstats.data_ptr()), sptr), "terminal"))
        ts = tstats.cpu().numpy()
        pt = mc.estimator((int(ts[0]), *ts[1:6]), S0=100.0, q=0.0, T=1.0)
        rate = 2.0 * nb / (ms * 1e-3)
        workloads.append({"workload": "cfg1 contract, terminal-only euro_price_mc kernel, 2^31 effective paths per GPU",
                          "value": rate, "unit": "path-steps/s/GPU", "ms": ms,
                          "fp32_pipe_frac": rate * OPS_PER_PATH_STEP["terminal"] / (N_SM * FP32_LANES * 1965.0e6),
                          "price": pt[0], "stderr": pt[1]})
        # opt-in generator modes (default stays Philox4x32-10, 4 normals per call): headline kernel and cfg-4 kernel
        hout = torch.zeros(nv.NSTATS, dtype=torch.float64, device=dev)
        hpay = ex._native_payoffs([ex.asian_spec(100.0, "call")], 0.05, 1.0)
        nbh = 1 << 25

        def heston_step_():
            nv.check(lib.optp_mc_path_dev(C.byref(hp), hpay, 1, 365, nbh, 1, nv.MODE_FP32_PHILOX, 42, rank * nbh, None, None,
                                          None, C.c_void_p(hout.data_ptr()), sptr), "optp_mc_path_dev")
        for mode_id, label in ((1, "rng_mode 1: Philox4x32-7"), (2, "rng_mode 2: Philox4x32-10, three Box-Muller pairs "
                                                                    "of (26-bit radius, 16/16/18-bit angle) per call")):
            nv.set_option("rng_mode", mode_id)
            try:
                msr = timed(lambda: mc_step(3), reps=2)
                sr = stats_t.cpu().numpy()
                msh = timed(heston_step_, reps=2)
                sh = hout.cpu().numpy()
            finally:
                nv.set_option("rng_mode", 0)
            pr_ = ex.estimate(sr)
            ph_ = ex.estimate(sh, control_variate=True, expected_y=100.0)
            workloads.append({"workload": f"OPT-IN {label}: headline GBM+barrier kernel / cfg4 Heston FTE+Asian kernel "
                                          "(tests/test_production_scale_gpu.py::test_device_normals_quality holds the statistical price)",
                              "value": 2.0 * n_base * MC_STEPS / (msr * 1e-3), "unit": "path-steps/s/GPU", "ms": msr,
                              "price": pr_[0], "stderr": pr_[1],
                              "heston_value": 2.0 * nbh * 365 / (msh * 1e-3), "heston_price": ph_[0], "heston_stderr": ph_[1]})
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
        book_checksum = float(sum(r["price"] for r in priced[:100]))
        workloads.append({"workload": "cfg5 book: 10k contracts (European / barrier / arithmetic Asian), price_book(), "
                                      f"{2 * book_paths} effective paths each, 252 steps, items dealt over the ranks (book.shard_items), "
                                      "planned and launched slice by slice (host parsing hidden behind the kernels)",
                          "value": book_path_steps / book_s, "unit": "path-steps/s (all GPUs, end to end)",
                          "contracts_per_s": n_rows / book_s, "seconds": book_s,
                          "priced": sum(1 for r in priced if r.get("price") is not None),
                          "checksum": book_checksum})
        strong = {"what": "fixed total work on the N GPUs: seconds to price (max over ranks)", "n_gpus": world,
                  "cfg4_heston_asian_1e9_paths_s": ms4 * 1e-3, "cfg3_lv_barrier_1e8_paths_s": None if ms3 is None else ms3 * 1e-3,
                  "cfg5_book_10k_x_1e7_paths_s": book_s, "cfg5_checksum_first100": book_checksum,
                  "cfg4_price": pcv[0]}
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
        # crr_vec, many strikes on one tree (binomial.py:44-109): the reference's own batched tree call, measured
        # at 4.55 s = 220 options/s for this call in the build container
        strikes = np.linspace(60.0, 140.0, 1000)
        bn.crr_vec(100.0, strikes, 1.0, 0.03, 0.0, 0.2, "put", N=1000, american=True)
        t0 = time.perf_counter()
        for _ in range(10):
            pv = bn.crr_vec(100.0, strikes, 1.0, 0.03, 0.0, 0.2, "put", N=1000, american=True)
        vec_s = (time.perf_counter() - t0) / 10
        workloads.append({"workload": "crr_vec(S0=100, 1000 strikes in [60,140], T=1, r=.03, sigma=.2, put, N=1000, american=True) "
                                      "public API, host in / host out",
                          "value": 1000 / vec_s, "unit": "options/s (end to end)", "seconds": vec_s,
                          "checksum": float(np.sum(pv)),
                          "reference_same_call": "4.55 s (220 options/s) in the build container, SURVEY 8a row 14; the reference arm "
                                                 "times it on this box (`--impl reference`, workloads)"})
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
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        p32_nominal = N_SM * FP32_LANES * sm_mhz * 1e6      # FP32 lane-ops/s the SMs can issue at the sampled clock
        p64_nominal = N_SM * FP64_LANES * sm_mhz * 1e6
        gbm_ops = OPS_PER_PATH_STEP["gbm_barrier"]
        crr_roof = {
            # PHYSICAL figure: FP64-pipe lane-instructions the level loops of the two-zone kernel executed for this
            # batch (counted by the diagnostic build of the same kernel, optp_get_counter) per second of the timed
            # production launch, over the DFMA issue rate.  Set-up, band measurement and re-pack work between the
            # segments is NOT counted, so ncu's pipe figure (profiles/) is a few points higher.
            "bound": "fp64_pipe", "achieved": crr_rate_gpu * crr_fp64_exec / 1e12, "peak": peaks["fp64_dfma_lane_ops"] / 1e12,
            "unit": "T FP64 lane-instr/s", "frac": crr_rate_gpu * crr_fp64_exec / peaks["fp64_dfma_lane_ops"],
            "peak_nominal": p64_nominal / 1e12, "frac_nominal": crr_rate_gpu * crr_fp64_exec / p64_nominal,
            "executed_fp64_lane_instr_per_option": crr_fp64_exec, "executed_node_updates_per_option": crr_nodes_exec,
            "algorithmic_node_updates_per_option": CRR_NODES_PER_OPTION,
            # how much of the reference's work the band cuts remove (exactly: prices equal the full-width kernel's)
            "algorithmic_speedup": CRR_NODES_PER_OPTION / crr_nodes_exec if crr_nodes_exec else None,
            # north_star's ruler, for the record: ALGORITHMIC shared-memory bytes (24 B per node update of the
            # reference's level array) per second over the measured LDS.64 bandwidth.  It exceeds 1 because the
            # levels live in registers and 72 % of the node updates are provably skipped.
            "algorithmic_smem_frac": crr_rate_gpu * CRR_SMEM_BYTES_PER_OPTION / peaks["smem_lds64_bytes"],
            "traffic": 52_163_840 if n_c == 1_000_000 else None,   # ncu --set full (profiles/r2_ncu_full_bench_zone.md): dram read 49.1 MB + write 3.1 MB per launch (the write-back of the 9 MB of results is partly caught by the L2: 2.2-3.7 MB between captures)
            "peak_source": "optp_microbench DFMA issue rate measured in this run (= 148 x 64 lanes x clock to 1 %)"}
        line = {
            "metric": "mc_path_steps_per_sec", "value": mc_value, "unit": "path-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": mc_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 path state, f64 accumulators", "data": "synthetic",
            "config": {"workload": "fused GBM exact log-Euler paths, 252 steps, up-and-out barrier call "
                                   "(S0=100,K=100,B=130,r=0.05,sigma=0.2,T=1), antithetic, Philox4x32-10 in registers",
                       "base_paths_per_gpu_per_step": n_base, "effective_paths_per_step": 2 * n_base * world,
                       "l2": "flushed between steps (256 MiB write); the kernel has no HBM inputs",
                       "collective": "one optp_allreduce_stats_dev (ncclAllReduce, 8 doubles, in place on the kernel's stream) per step"
                       if world > 1 else "none"},
            "price": mc_price[0], "stderr": mc_price[1], "wall_ms_per_step": 1e3 * mc_wall / K,
            "e2e": mc_e2e,
            "roofline": {"bound": "fp32_pipe", "achieved": mc_kernel_rate * gbm_ops / 1e12,
                         # peak = what the FP32 pipes can issue: 148 SMs x 128 lanes x the SM clock sampled during the
                         # timed region (MEASURED_PEAKS.json has no FP32-pipe figure; B200_PROFILING's nominal)
                         "peak": p32_nominal / 1e12, "unit": "T lane-ops/s", "frac": mc_kernel_rate * gbm_ops / p32_nominal,
                         "traffic": 105728,  # ncu --set full of THIS launch (profiles/r2_ncu_full_bench_gbm.md): dram read 106 KB, write 0 (0.1-0.5 MB between captures: background traffic, the kernel has no HBM operands)
                         "peak_measured_ffma": peaks["fp32_ffma_lane_ops"] / 1e12,
                         "frac_vs_measured_ffma": mc_kernel_rate * gbm_ops / peaks["fp32_ffma_lane_ops"],
                         # ncu counts 11.4 executed lane-instructions per path-step (12.08e9 warp instructions per
                         # 3.382e10 path-steps): the algorithmic 13 of SURVEY App. C is slightly generous
                         "frac_executed_instr": mc_kernel_rate * 11.4 / p32_nominal,
                         # secondary bound: one MUFU per effective path-step against the measured MUFU.EX2 rate
                         "mufu_frac": mc_kernel_rate * 1.0 / peaks["mufu_ex2_ops"],
                         "ops_per_path_step": gbm_ops, "kernel_ms": mc_kernel_ms,
                         "path_steps_per_s_per_gpu": mc_kernel_rate,
                         "fp64_mode": fp64_mode,
                         "peak_source": "nominal FP32 issue rate 148 x 128 x sampled SM clock; compute-bound kernel, HBM traffic ~0"},
            "cpu_baseline": cpu_mc,
            "clocks": clocks,
            "gpu_launches": mc_launches + crr_launches,
            "gpu_launches_total_in_run": total_launches,
            "workloads": workloads,
            "peaks_measured_in_run": peaks,
            "measured_peaks_file": {k: measured.get(k) for k in ("hbm_gbs", "bf16_tflops", "sm_max_mhz")},
            "crr_detail": {
                "metric": "crr_american_N1000_options_per_sec", "value": crr_value, "unit": "options/s",
                "ms_per_step": crr_ms, "dtype": "f64",
                "config": {"workload": "BASELINE config 2: American put, N=1000, synthetic contracts "
                                       "(default_rng: S0~U(50,150), K=S0*U(.8,1.2), T~U(.1,2), sigma~U(.1,.6), r=.03)",
                           "contracts_per_gpu_per_step": n_c, "l2": "flushed between steps"},
                "e2e": crr_e2e, "roofline": crr_roof,
                "full_width": {"value": crr_full_rate * world, "unit": "options/s", "ms_per_step": crr_full_ms,
                               # 3 FP64-pipe instructions per node update (2 DFMA + 1 DSETP), every node of every level
                               "fp64_pipe_frac": crr_full_rate * 3.0 * CRR_NODES_PER_OPTION / 0.88 / peaks["fp64_dfma_lane_ops"],
                               "fp64_algorithmic_frac": crr_full_rate * CRR_FP64_OPS_PER_OPTION / peaks["fp64_dfma_lane_ops"],
                               "checksum_first8": crr_full_check,
                               "what": "optp_set_option('crr_band', 0): every node of every level updated in registers "
                                       "(88 % of the issued node updates are live)"},
                "cpu_baseline": cpu_crr, "gpu_launches": crr_launches, "checksum_first8": crr_check,
                "checksum_first8_counting_build": crr_count_check,
            },
            "strong": strong,
            # the second half of BASELINE's metric, compact, LAST (survives a truncated tail)
            "crr": {"metric": "crr_american_N1000_options_per_sec", "value": crr_value, "unit": "options/s",
                    "n_gpus": world, "ms_per_step": crr_ms, "contracts_per_gpu": n_c,
                    "e2e": {"value": crr_e2e["value"], "h2d_bytes_per_step": crr_e2e["h2d_bytes_per_step"],
                            "d2h_bytes_per_step": crr_e2e["d2h_bytes_per_step"]},
                    "roofline": {"bound": "fp64_pipe", "frac": crr_roof["frac"], "achieved": crr_roof["achieved"],
                                 "peak": crr_roof["peak"], "unit": crr_roof["unit"],
                                 "executed_node_updates_per_option": crr_nodes_exec,
                                 "algorithmic_speedup": crr_roof["algorithmic_speedup"]},
                    "full_width_options_per_s": crr_full_rate * world,
                    "cpu_baseline": None if cpu_crr is None else {k: cpu_crr[k] for k in ("value", "unit", "cores", "kind")},
                    "checksum_first8": crr_check},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
