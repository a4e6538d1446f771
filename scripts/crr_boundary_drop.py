#!/usr/bin/env python3
"""How fast does the early-exercise boundary of an American put fall, in lattice nodes per level?
(csrc/crr_batch.cu: kZoneSeg / kZoneSeg0 -- the two-zone kernel keeps 30 nodes below the boundary per segment.)

Runs the exact numpy lattice (binomial.py:28-36 restated) for random contracts, records the first continuation
node jlo[k] of every level and prints, for a start n0 levels from expiry, the longest stretch L over which the
boundary fell by at most 27 nodes for EVERY contract.  Usage: crr_boundary_drop.py [n_contracts] [N]"""
import sys

import numpy as np


def jlo_profile(S0, K, T, r, q, sg, N):
    dt = T / N
    u = np.exp(sg * np.sqrt(dt)); d = 1 / u
    disc = np.exp(-r * dt); p = (np.exp((r - q) * dt) - d) / (u - d)
    if not 0 < p < 1:
        return None
    j = np.arange(N + 1)
    V = np.maximum(K - S0 * u ** (2.0 * j - N), 0.0)
    jlo = np.zeros(N + 1, int)
    ex = (K - S0 * u ** (2.0 * j - N)) > 0
    jlo[N] = np.argmin(ex) if not ex.all() else N + 1
    for k in range(N - 1, -1, -1):
        cont = disc * (p * V[1:k + 2] + (1 - p) * V[:k + 1])
        e = K - S0 * u ** (2.0 * j[:k + 1] - k)
        exm = e > cont
        V = np.where(exm, e, cont)
        jlo[k] = np.argmin(exm) if not exm.all() else k + 1
    return jlo


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    g = np.random.default_rng(12)
    prof = []
    while len(prof) < n:
        S0 = g.uniform(50, 150); K = S0 * g.uniform(0.7, 1.3); T = g.uniform(0.05, 5.0)
        sg = g.uniform(0.05, 1.0); r = g.choice([0.002, 0.01, 0.03, 0.08, 0.15]); q = g.choice([0.0, 0.0, 0.01, 0.02])
        if q >= r:
            continue
        jl = jlo_profile(S0, K, T, r, q, sg, N)
        if jl is not None:
            prof.append(jl)
    for n0 in (0, 16, 32, 36, 48, 64, 96, 128, 256, 512, N - 100):
        k = N - n0
        best = None
        for L in range(20, 90):
            k2 = k - L
            if k2 <= 0:
                break
            worst = 0
            for jl in prof:
                if jl[k] <= k and jl[k2] <= k2:
                    worst = max(worst, jl[k] - jl[k2])
            if worst <= 27:
                best = (L, int(worst))
        print(f"N={N} start {n0:4d} levels from expiry: longest stretch with drop <= 27: {best}")
