"""Oracle: Cox-Ross-Rubinstein backward induction (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference src/optpricer/binomial.py:
  crr_scalar      <- crr      binomial.py:7-38   (any kind != "call" prices a put, :22-25)
  crr_shared_tree <- crr_vec  binomial.py:44-109 (kind compared with str(k) == "call", :93)
  crr_loop        <- the oracle SURVEY 8c names for a heterogeneous contract batch: a Python
                     loop over crr_scalar (the reference has no batched heterogeneous API).
Same operation order as the reference (d = 1/u, disc applied after the weighted sum,
exercise value recomputed with two pow() per node) so values are bit-identical.
A C restatement of the same recursion lives in oracle/crr_tree.c (loader: oracle/crr_c.py).
"""
from __future__ import annotations

from math import exp, sqrt

import numpy as np

P_ERROR = "Risk-neutral prob p out of (0,1); try larger N or different params."


def lattice_constants(T, r, q, sigma, N):
    if N <= 0:
        raise ValueError("N must be positive.")
    dt = T / N
    u = exp(sigma * sqrt(dt))
    d = 1.0 / u
    disc = exp(-r * dt)
    p = (exp((r - q) * dt) - d) / (u - d)
    if not (0.0 < p < 1.0):
        raise ValueError(P_ERROR)
    return u, d, disc, p


def crr_scalar(S0, K, T, r, sigma, q=0.0, kind="call", N=500, american=False):
    u, d, disc, p = lattice_constants(T, r, q, sigma, N)
    call = kind == "call"
    idx = np.arange(N + 1)
    s_end = S0 * (u ** idx) * (d ** (N - idx))
    V = np.maximum(s_end - K, 0.0) if call else np.maximum(K - s_end, 0.0)
    for level in range(N - 1, -1, -1):
        V = disc * (p * V[1:] + (1.0 - p) * V[:-1])
        if american:
            idx = np.arange(level + 1)
            s_lvl = S0 * (u ** idx) * (d ** (level - idx))
            V = np.maximum(V, s_lvl - K) if call else np.maximum(V, K - s_lvl)
    return float(V[0])


def crr_shared_tree(S0, K, T, r, q, sigma, kind, N=500, american=False):
    K = np.atleast_1d(np.asarray(K, dtype=float))
    kind = np.atleast_1d(np.asarray(kind))
    if kind.shape != K.shape:
        kind = np.broadcast_to(kind, K.shape)
    u, d, disc, p = lattice_constants(T, r, q, sigma, N)
    idx = np.arange(N + 1)
    s_end = S0 * (u ** idx) * (d ** (N - idx))
    call = np.array([str(k) == "call" for k in kind.flat], dtype=bool).reshape(K.shape)
    V = np.where(call[:, None],
                 np.maximum(s_end[None, :] - K[:, None], 0.0),
                 np.maximum(K[:, None] - s_end[None, :], 0.0))
    for level in range(N - 1, -1, -1):
        V = disc * (p * V[:, 1:] + (1.0 - p) * V[:, :-1])
        if american:
            idx = np.arange(level + 1)
            s_lvl = S0 * (u ** idx) * (d ** (level - idx))
            ex = np.where(call[:, None], s_lvl[None, :] - K[:, None], K[:, None] - s_lvl[None, :])
            V = np.maximum(V, ex)
    return V[:, 0]


def crr_loop(S0, K, T, r, q, sigma, is_call, N, american):
    """Heterogeneous batch = loop over crr_scalar; p outside (0,1) -> NaN + status 1."""
    n = len(S0)
    price = np.full(n, np.nan)
    status = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        try:
            price[i] = crr_scalar(float(S0[i]), float(K[i]), float(T[i]), float(r[i]), float(sigma[i]),
                                  float(q[i]), "call" if is_call[i] else "put", N, american)
        except ValueError:
            status[i] = 1
    return price, status


def cfg2_contracts(n, seed=0):
    """BASELINE config 2 synthetic batch: draw order S0, K-ratio, T, sigma."""
    g = np.random.default_rng(seed)
    S0 = g.uniform(50.0, 150.0, n)
    K = S0 * g.uniform(0.8, 1.2, n)
    T = g.uniform(0.1, 2.0, n)
    sigma = g.uniform(0.1, 0.6, n)
    r = np.full(n, 0.03)
    q = np.zeros(n)
    is_call = np.zeros(n, dtype=np.uint8)
    return S0, K, T, r, q, sigma, is_call
