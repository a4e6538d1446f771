#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes small fixtures (json / npz) next to this file.  Every value comes from
`optpricer` imported from /root/reference/src; nothing from this repo's product or
oracle code is used to *compute* a value, except oracle.sigma_grid.SigmaGrid, which
is the table callable handed to the reference's own milstein_local_vol_paths
("oracle = reference Milstein driven by a numpy table-callable").
numpy / scipy versions are recorded in meta.json because PCG64 streams and the SVI
fit are only stable within a version.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)

import scipy  # noqa: E402
from optpricer.core import OptionSpec  # noqa: E402
from optpricer.monte_carlo import euro_price_mc, _mc_chunk_sumstats  # noqa: E402
from optpricer.binomial import crr, crr_vec  # noqa: E402
from optpricer.processes import (  # noqa: E402
    gbm_paths, heston_paths, milstein_local_vol_paths, local_vol_paths,
    gbm_milstein_paths, sabr_paths, merton_jump_paths)
from optpricer.exotics import barrier_price, asian_price, digital_price, lookback_price  # noqa: E402
from optpricer.calibration import fit_svi_surface, dupire_local_vol_func  # noqa: E402
from oracle.sigma_grid import tabulate  # noqa: E402


def dump_json(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name)


def euro_cases():
    base = dict(S0=100.0, K=110.0, T=1.0, r=0.03, sigma=0.20, q=0.0)
    variants = [
        ("call", {}),
        ("call", {"antithetic": False}),
        ("call", {"control_variate": False}),
        ("call", {"antithetic": False, "control_variate": False}),
        ("call", {"chunk_size": 50_000}),
        ("call", {"chunk_size": 25_000}),
        ("call", {"dtype": "float32"}),
        ("put", {}),
        ("put", {"antithetic": False, "control_variate": False}),
    ]
    out = []
    for kind, kw in variants:
        k2 = dict(kw)
        if "dtype" in k2:
            k2["dtype"] = np.float32
        p, se = euro_price_mc(OptionSpec(**base), kind, n_paths=200_000, seed=1, **k2)
        out.append({"opt": base, "kind": kind, "n_paths": 200_000, "seed": 1, "kwargs": kw,
                    "price": p, "stderr": se})
    # the reference test's configuration (tests/test_monte_carlo_vs_bs.py:6-8)
    tcfg = dict(S0=100.0, K=100.0, T=1.0, r=0.03, sigma=0.25, q=0.01)
    for kind in ("call", "put"):
        p, se = euro_price_mc(OptionSpec(**tcfg), kind, n_paths=40_000, control_variate=True, seed=1)
        out.append({"opt": tcfg, "kind": kind, "n_paths": 40_000, "seed": 1, "kwargs": {},
                    "price": p, "stderr": se})
    # ragged chunking + q != 0 + odd sizes
    rag = dict(S0=95.0, K=100.0, T=0.5, r=0.01, sigma=0.35, q=0.02)
    for kind, n, cs, seed in (("call", 12_345, 5_000, 7), ("put", 1, 100_000, 3), ("put", 777, 100, 11)):
        p, se = euro_price_mc(OptionSpec(**rag), kind, n_paths=n, seed=seed, chunk_size=cs)
        out.append({"opt": rag, "kind": kind, "n_paths": n, "seed": seed, "kwargs": {"chunk_size": cs},
                    "price": p, "stderr": se})
    # chunk statistics for cfg 1 (two chunks of 100k)
    ss = np.random.SeedSequence(1).spawn(2)
    chunk_stats = [list(_mc_chunk_sumstats(100_000, kind="call", antithetic=True, seed=s, **base)) for s in ss]
    z4 = np.random.default_rng(np.random.SeedSequence(1).spawn(1)[0]).standard_normal(4).tolist()
    dump_json("euro_mc.json", {"cases": out, "cfg1_chunk_stats": chunk_stats, "cfg1_first_normals": z4})


def crr_cases():
    cases = []

    def add(S0, K, T, r, sigma, q, kind, N, am):
        v = crr(OptionSpec(S0, K, T, r, sigma, q), kind, N, american=am)
        cases.append({"S0": S0, "K": K, "T": T, "r": r, "sigma": sigma, "q": q, "kind": kind,
                      "N": N, "american": am, "price": v})

    for N in (200, 500, 1000):
        for am in (True, False):
            add(100.0, 110.0, 1.0, 0.03, 0.20, 0.0, "put", N, am)
    add(100.0, 100.0, 1.0, 0.05, 0.2, 0.0, "call", 200, False)
    add(100.0, 110.0, 1.0, 0.05, 0.3, 0.0, "put", 200, True)
    g = np.random.default_rng(2024)
    for N in (1, 2, 3, 7, 31, 32, 33, 50, 127, 128, 129, 200, 255, 256, 500, 511, 512, 1000, 1023, 1024, 1500, 2047, 2048, 3000):
        reps = 6 if N <= 512 else 2
        for _ in range(reps):
            S0 = float(g.uniform(50, 150)); K = float(S0 * g.uniform(0.8, 1.2))
            T = float(g.uniform(0.1, 2.0)); sig = float(g.uniform(0.1, 0.6))
            r = float(g.uniform(0.0, 0.08)); q = float(g.uniform(0.0, 0.04))
            kind = "call" if g.random() < 0.5 else "put"
            am = bool(g.random() < 0.6)
            try:
                add(S0, K, T, r, sig, q, kind, N, am)
            except ValueError as e:
                cases.append({"S0": S0, "K": K, "T": T, "r": r, "sigma": sig, "q": q, "kind": kind,
                              "N": N, "american": am, "error": str(e)})
    # a guaranteed p-out-of-range contract (tiny sigma, large rate, few steps)
    try:
        crr(OptionSpec(100.0, 100.0, 2.0, 0.30, 0.01, 0.0), "call", 4)
        raise SystemExit("expected ValueError")
    except ValueError as e:
        cases.append({"S0": 100.0, "K": 100.0, "T": 2.0, "r": 0.30, "sigma": 0.01, "q": 0.0,
                      "kind": "call", "N": 4, "american": False, "error": str(e)})

    # cfg 2: first 64 contracts of the synthetic batch, N=1000 American put
    g0 = np.random.default_rng(0)
    n = 1_000_000
    S0 = g0.uniform(50.0, 150.0, n); K = S0 * g0.uniform(0.8, 1.2, n)
    T = g0.uniform(0.1, 2.0, n); sig = g0.uniform(0.1, 0.6, n)
    cfg2 = []
    t0 = time.time()
    for i in range(64):
        cfg2.append(crr(OptionSpec(float(S0[i]), float(K[i]), float(T[i]), 0.03, float(sig[i]), 0.0),
                        "put", 1000, american=True))
    print("cfg2 64 contracts: %.1f s" % (time.time() - t0))

    vec = []
    for (s0, ks, T_, r_, q_, sg, kinds, N, am) in [
        (100.0, [100.0], 1.0, 0.05, 0.0, 0.2, "call", 200, False),
        (100.0, [90.0, 100.0, 110.0], 1.0, 0.05, 0.0, 0.2, "call", 200, False),
        (100.0, [110.0], 1.0, 0.05, 0.0, 0.3, "put", 200, True),
        (100.0, [80.0, 95.0, 100.0, 105.0, 130.0], 0.75, 0.04, 0.015, 0.27,
         ["call", "put", "put", "call", "put"], 333, True),
    ]:
        v = crr_vec(s0, np.array(ks), T_, r_, q_, sg, kinds, N, american=am)
        vec.append({"S0": s0, "K": ks, "T": T_, "r": r_, "q": q_, "sigma": sg, "kind": kinds, "N": N,
                    "american": am, "prices": v.tolist()})
    dump_json("crr.json", {"cases": cases, "cfg2_first64_american_put_N1000": cfg2, "crr_vec": vec})


def exotic_table(paths, K, r, T):
    """All payoff variants on one path array -> {name: [price, se]}."""
    out = {}
    for bt, kind, B in (("up-and-out", "call", 120.0), ("up-and-in", "call", 120.0),
                        ("down-and-out", "put", 85.0), ("down-and-in", "put", 85.0),
                        ("up-and-out", "put", 115.0), ("down-and-in", "call", 90.0)):
        out[f"barrier|{bt}|{kind}|{B}|0.0"] = list(barrier_price(paths, K, r, T, kind, B, bt))
    out["barrier|up-and-out|call|120.0|2.5"] = list(barrier_price(paths, K, r, T, "call", 120.0, "up-and-out", rebate=2.5))
    out["barrier|down-and-in|put|85.0|1.0"] = list(barrier_price(paths, K, r, T, "put", 85.0, "down-and-in", rebate=1.0))
    for avg in ("arithmetic", "geometric"):
        for st in ("fixed", "floating"):
            for kind in ("call", "put"):
                out[f"asian|{kind}|{avg}|{st}"] = list(asian_price(paths, K, r, T, kind, avg, st))
    for kind in ("call", "put"):
        out[f"digital|{kind}|1.0"] = list(digital_price(paths, K, r, T, kind))
        out[f"digital|{kind}|10.0"] = list(digital_price(paths, K, r, T, kind, payout=10.0))
        for st in ("floating", "fixed"):
            out[f"lookback|{kind}|{st}"] = list(lookback_price(paths, r, T, kind, K, st))
    return out


def path_sample(paths):
    """A few rows / columns of the materialised array for paths_out checks."""
    rows = sorted({0, 1, paths.shape[0] // 2, paths.shape[0] - 1})
    return {"rows": rows, "row_values": paths[rows, :].copy(),
            "col_sums": paths.sum(axis=0), "terminal": paths[-1].copy()}


def mc_path_cases():
    meta = {}
    arrays = {}
    # --- small GBM / Heston / LV cases (full per-row samples kept) ------------------
    S0, K, T, r, q, sigma = 100.0, 100.0, 1.0, 0.05, 0.0, 0.20
    for anti in (True, False):
        tag = f"gbm_small_anti{int(anti)}"
        p = gbm_paths(S0, r, q, sigma, T, 50, 2000, antithetic=anti, seed=42)
        meta[tag] = {"gen": "gbm", "args": dict(S0=S0, r=r, q=q, sigma=sigma, T=T, n_steps=50, n_paths=2000,
                                                antithetic=anti, seed=42),
                     "K": K, "r": r, "T": T, "prices": exotic_table(p, K, r, T)}
        for k_, v_ in path_sample(p).items():
            arrays[f"{tag}/{k_}"] = np.asarray(v_)
    hargs = dict(S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, T=1.0)
    for anti in (True, False):
        tag = f"heston_small_anti{int(anti)}"
        p, v = heston_paths(n_steps=73, n_paths=1500, antithetic=anti, seed=7, return_variance=True, **hargs)
        meta[tag] = {"gen": "heston", "args": dict(n_steps=73, n_paths=1500, antithetic=anti, seed=7, **hargs),
                     "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0)}
        for k_, v_ in path_sample(p).items():
            arrays[f"{tag}/{k_}"] = np.asarray(v_)
        arrays[f"{tag}/var_terminal"] = v[-1].copy()
    # stressed Heston (Feller violated -> truncation active)
    hx = dict(S0=100.0, r=0.02, q=0.01, v0=0.09, kappa=1.5, theta=0.04, xi=0.9, rho=-0.5, T=0.5)
    p, v = heston_paths(n_steps=40, n_paths=1000, antithetic=True, seed=3, return_variance=True, **hx)
    meta["heston_trunc"] = {"gen": "heston", "args": dict(n_steps=40, n_paths=1000, antithetic=True, seed=3, **hx),
                            "K": 95.0, "r": 0.02, "T": 0.5, "prices": exotic_table(p, 95.0, 0.02, 0.5),
                            "frac_v_zero": float((v == 0.0).mean())}
    for k_, v_ in path_sample(p).items():
        arrays[f"heston_trunc/{k_}"] = np.asarray(v_)

    # --- Dupire surface of the desk workflow (scripts/desk_workflow_localvol_barrier.py:52-98) ---
    S0d, rd, qd = 100.0, 0.05, 0.02
    expiries = [0.25, 0.50, 1.00]
    forwards = {Tk: S0d * np.exp((rd - qd) * Tk) for Tk in expiries}
    strikes, ivs = {}, {}
    for Tk in expiries:
        F = forwards[Tk]
        Ks = np.linspace(0.75 * F, 1.25 * F, 21)
        kk = np.log(Ks / F)
        strikes[Tk] = Ks
        ivs[Tk] = 0.20 + 0.05 * kk ** 2 - 0.02 * kk + 0.005 * np.sqrt(Tk)
    surface = fit_svi_surface(strikes, forwards, ivs)
    sig_dup = dupire_local_vol_func(surface, r=rd, q=qd)
    svi = {str(Tk): [p_.a, p_.b, p_.rho, p_.m, p_.sigma] for Tk, p_ in surface.slices.items()}
    probes = [[85.0, 0.1], [100.0, 0.1], [115.0, 0.1], [100.0, 0.5], [100.0, 0.9], [100.0, 0.25],
              [100.0, 0.2501], [100.0, 0.4999], [100.0, 0.5001], [20.0, 0.7], [300.0, 0.03], [100.0, 0.0]]
    dump_json("dupire_surface.json", {
        "S0": S0d, "r": rd, "q": qd, "svi": svi, "forwards": {str(k): v for k, v in forwards.items()},
        "probes": [[s, t, float(sig_dup(np.array([s]), t)[0])] for s, t in probes]})

    # cfg-3 table: 252 step times x 256 log-uniform nodes over ln S0 +- 1.6
    grid252 = tabulate(sig_dup, S0=S0d, T=1.0, n_steps=252, n_nodes=256, half_width=1.6)
    np.savez_compressed(os.path.join(HERE, "dupire_table_252x256.npz"),
                        table=grid252.table, x0=grid252.x0, h=grid252.h, dt=grid252.dt)
    print("wrote dupire_table_252x256.npz")

    # small LV-Milstein cases: (a) constant sigma callable, (b) table callable, (c) Dupire callable
    const = lambda S, t: 0.2 * np.ones_like(S)  # noqa: E731
    for tag, fn, extra in (("lvm_const", const, {}),):
        p = milstein_local_vol_paths(100.0, 0.05, 0.0, 1.0, 50, 2000, fn, seed=42)
        meta[tag] = {"gen": "lv_milstein", "sigma": "const0.2",
                     "args": dict(S0=100.0, r=0.05, q=0.0, T=1.0, n_steps=50, n_paths=2000, antithetic=True, seed=42),
                     "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0)}
        for k_, v_ in path_sample(p).items():
            arrays[f"{tag}/{k_}"] = np.asarray(v_)
    grid63 = tabulate(sig_dup, S0=S0d, T=1.0, n_steps=63, n_nodes=128, half_width=1.6)
    arrays["lvm_table/table"] = grid63.table
    arrays["lvm_table/grid"] = np.array([grid63.x0, grid63.h, grid63.dt])
    for anti in (True, False):
        tag = f"lvm_table_anti{int(anti)}"
        p = milstein_local_vol_paths(S0d, rd, qd, 1.0, 63, 3000, grid63, antithetic=anti, seed=42)
        meta[tag] = {"gen": "lv_milstein", "sigma": "lvm_table",
                     "args": dict(S0=S0d, r=rd, q=qd, T=1.0, n_steps=63, n_paths=3000, antithetic=anti, seed=42),
                     "K": 100.0, "r": rd, "T": 1.0, "prices": exotic_table(p, 100.0, rd, 1.0),
                     "uo130": list(barrier_price(p, 100.0, rd, 1.0, "call", 130.0, "up-and-out"))}
        for k_, v_ in path_sample(p).items():
            arrays[f"{tag}/{k_}"] = np.asarray(v_)
    # bumped dS and a log-Euler LV / gbm-milstein / sabr / merton set ("next" rows)
    p = milstein_local_vol_paths(S0d, rd, qd, 1.0, 63, 1000, grid63, seed=5, dS_bump=0.03)
    meta["lvm_table_bump3"] = {"gen": "lv_milstein", "sigma": "lvm_table", "dS_bump": 0.03,
                               "args": dict(S0=S0d, r=rd, q=qd, T=1.0, n_steps=63, n_paths=1000, antithetic=True, seed=5),
                               "K": 100.0, "r": rd, "T": 1.0, "prices": exotic_table(p, 100.0, rd, 1.0)}
    for k_, v_ in path_sample(p).items():
        arrays[f"lvm_table_bump3/{k_}"] = np.asarray(v_)
    p = local_vol_paths(S0d, rd, qd, 1.0, 63, 1000, grid63, seed=9)
    meta["lv_euler_table"] = {"gen": "lv_log_euler", "sigma": "lvm_table",
                              "args": dict(S0=S0d, r=rd, q=qd, T=1.0, n_steps=63, n_paths=1000, antithetic=True, seed=9),
                              "K": 100.0, "r": rd, "T": 1.0, "prices": exotic_table(p, 100.0, rd, 1.0)}
    for k_, v_ in path_sample(p).items():
        arrays[f"lv_euler_table/{k_}"] = np.asarray(v_)
    p = gbm_milstein_paths(100.0, 0.05, 0.0, 0.2, 1.0, 50, 1000, seed=42)
    meta["gbm_milstein"] = {"gen": "gbm_milstein",
                            "args": dict(S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0, n_steps=50, n_paths=1000, antithetic=True, seed=42),
                            "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0)}
    for k_, v_ in path_sample(p).items():
        arrays[f"gbm_milstein/{k_}"] = np.asarray(v_)
    for beta in (1.0, 0.5):
        tag = f"sabr_beta{beta}"
        sargs = dict(S0=100.0, r=0.03, q=0.0, alpha0=0.25 if beta == 1.0 else 2.5, beta=beta, nu=0.4, rho=-0.3, T=1.0)
        p = sabr_paths(n_steps=40, n_paths=1000, seed=11, **sargs)
        meta[tag] = {"gen": "sabr", "args": dict(n_steps=40, n_paths=1000, antithetic=True, seed=11, **sargs),
                     "K": 100.0, "r": 0.03, "T": 1.0, "prices": exotic_table(p, 100.0, 0.03, 1.0)}
        for k_, v_ in path_sample(p).items():
            arrays[f"{tag}/{k_}"] = np.asarray(v_)
    margs = dict(S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0)
    p = merton_jump_paths(n_steps=40, n_paths=1000, lam=0.8, mJ=-0.1, sJ=0.15, seed=13, **margs)
    meta["merton"] = {"gen": "merton", "args": dict(n_steps=40, n_paths=1000, lam=0.8, mJ=-0.1, sJ=0.15,
                                                    antithetic=True, seed=13, **margs),
                      "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0)}
    for k_, v_ in path_sample(p).items():
        arrays[f"merton/{k_}"] = np.asarray(v_)

    # --- full-size cases of the reference tests / SURVEY appendix A (prices only) --------
    t0 = time.time()
    p = gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 500, 100_000, antithetic=True, seed=42)
    meta["gbm_test_exotics_500x100k"] = {
        "gen": "gbm", "args": dict(S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0, n_steps=500, n_paths=100_000,
                                   antithetic=True, seed=42),
        "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0),
        "mean_terminal": float(p[-1].mean())}
    del p
    p = heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, 50_000, seed=42)
    meta["heston_365x50k"] = {
        "gen": "heston", "args": dict(n_steps=365, n_paths=50_000, antithetic=True, seed=42, **hargs),
        "K": 100.0, "r": 0.05, "T": 1.0, "prices": exotic_table(p, 100.0, 0.05, 1.0),
        "mean_terminal": float(p[-1].mean())}
    del p
    p = milstein_local_vol_paths(S0d, rd, qd, 1.0, 252, 100_000, sig_dup, seed=42)
    meta["desk_dupire_callable_252x100k"] = {
        "gen": "lv_milstein", "sigma": "dupire_callable",
        "args": dict(S0=S0d, r=rd, q=qd, T=1.0, n_steps=252, n_paths=100_000, antithetic=True, seed=42),
        "uo130": list(barrier_price(p, 100.0, rd, 1.0, "call", 130.0, "up-and-out")),
        "s_range": [float(p.min()), float(p.max())]}
    del p
    p = milstein_local_vol_paths(S0d, rd, qd, 1.0, 252, 100_000, grid252, seed=42)
    meta["desk_table_252x100k"] = {
        "gen": "lv_milstein", "sigma": "dupire_table_252x256",
        "args": dict(S0=S0d, r=rd, q=qd, T=1.0, n_steps=252, n_paths=100_000, antithetic=True, seed=42),
        "K": 100.0, "r": rd, "T": 1.0, "prices": exotic_table(p, 100.0, rd, 1.0),
        "uo130": list(barrier_price(p, 100.0, rd, 1.0, "call", 130.0, "up-and-out"))}
    del p
    print("full-size cases: %.1f s" % (time.time() - t0))
    dump_json("mc_paths.json", meta)
    np.savez_compressed(os.path.join(HERE, "mc_paths_samples.npz"), **arrays)
    print("wrote mc_paths_samples.npz")


if __name__ == "__main__":
    dump_json("meta.json", {"numpy": np.__version__, "scipy": scipy.__version__,
                            "python": sys.version.split()[0], "reference_version": "0.3.0",
                            "bit_generator": "PCG64"})
    euro_cases()
    crr_cases()
    mc_path_cases()
