#!/usr/bin/env python3
"""Offline model of the two-zone banded CRR planner (csrc/crr_batch.cu: crr_zone_kernel): runs the
full numpy tree for a few config-2 contracts, records the exercise boundary and the 2^-80 K top cut
per level, replays the segment planner on them and prints the layout mix and an FP64-issue estimate.
Usage: python scripts/crr_zone_model.py [n_contracts] [seg] [bz_target]"""
import os
import sys
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from bench import cfg2_contracts  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
SEG = int(sys.argv[2]) if len(sys.argv) > 2 else 48
BZT = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 1000
S0, K, T, r, q, sig, _ = cfg2_contracts(1_000_000)


def band_profile(S0, K, T, r, q, sg):
    dt = T / N
    u = np.exp(sg * np.sqrt(dt)); d = 1 / u
    disc = np.exp(-r * dt); p = (np.exp((r - q) * dt) - d) / (u - d)
    j = np.arange(N + 1)
    V = np.maximum(K - S0 * u ** (2.0 * j - N), 0.0)
    thr = K * 2.0 ** -80
    jlo = np.zeros(N + 1, int); jhi = np.zeros(N + 1, int)
    ex = (K - S0 * u ** (2.0 * j - N)) > 0
    jlo[N] = np.argmin(ex) if not ex.all() else N + 1
    live = np.nonzero(V > thr)[0]
    jhi[N] = live[-1] + 1 if live.size else 0
    for k in range(N - 1, -1, -1):
        cont = disc * (p * V[1:k + 2] + (1 - p) * V[:k + 1])
        e = K - S0 * u ** (2.0 * j[:k + 1] - k)
        exm = e > cont
        V = np.where(exm, e, cont)
        jlo[k] = np.argmin(exm) if not exm.all() else k + 1
        live = np.nonzero(V > thr)[0]
        jhi[k] = live[-1] + 1 if live.size else 0
    return jlo, jhi


def pick_bc(w):
    need = (w + 31) >> 5
    if need <= 4: return need
    if need <= 6: return 6
    if need <= 8: return 8
    return 12


layouts = Counter(); segs = []; fp64 = []; fails = Counter(); nodes_old = []
for o in range(n):
    jlo_k, jhi_k = band_profile(S0[o], K[o], T[o], r[o], q[o], sig[o])
    k = N; w0 = None; prev_drop = prev_nlev = 0; jlo_start = -1
    nseg = 0; cost = 0; fail = None
    while True:
        jlo, jhi = jlo_k[k], max(jhi_k[k], jlo_k[k])
        if jlo_start >= 0: prev_drop = max(0, jlo_start - jlo)
        if k == 0: break
        nlev = min(SEG, k)
        room = 32 * BZT
        fit = room - 4
        if prev_nlev > 0 and prev_drop > 0: fit = max(fit, (room - 8) * prev_nlev // prev_drop)
        elif prev_nlev > 0: fit = max(fit, SEG)
        nlev = max(1, min(nlev, fit))
        top = min(jhi, k + 1)
        margin = nlev + 2
        if prev_nlev > 0: margin = min(margin, (prev_drop * nlev + prev_nlev - 1) // prev_nlev + 6)
        jlo_start = jlo
        w0n = max(0, jlo - margin)
        Bz = min(3, max(1, (jlo + 2 - w0n + 31) >> 5))
        ztop = w0n + 32 * Bz
        Bc = 0
        if top > ztop:
            if top - ztop > 32 * 12: fail = "wide"; break
            Bc = pick_bc(top - ztop)
            nlev = min(nlev, k - (ztop - 1))
        # checks over the segment
        ks = np.arange(k - 1, k - nlev - 1, -1)
        if w0n > 0 and (jlo_k[ks] <= w0n).any(): fail = "bottom"; break
        if Bc > 0 and (jlo_k[ks] > ztop - 1).any(): fail = "top"; break
        layouts[(Bz, Bc)] += nlev
        cost += nlev * (2 + 3 * Bz + Bc) + 40   # FP64 instr per level + a guess for the re-pack
        nseg += 1
        k -= nlev; prev_nlev = nlev
    if fail: fails[fail] += 1
    segs.append(nseg); fp64.append(cost)
    # single-zone kernel for comparison: 3 FP64 per node over the band + 2
tot = sum(layouts.values())
print(f"contracts {n} seg {SEG} bz_target {BZT}: segments/option {np.mean(segs):.1f}  FP64 instr/option {np.mean(fp64):.0f}  fails {dict(fails)}")
for kx, v in sorted(layouts.items()): print(f"  layout Bz={kx[0]} Bc={kx[1]}: {100.0 * v / tot:5.1f} % of levels")
