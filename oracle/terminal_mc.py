"""Oracle: terminal-only European Monte Carlo (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference src/optpricer/monte_carlo.py:
  * chunk_sumstats   <- _mc_chunk_sumstats   monte_carlo.py:11-64
  * combine          <- _aggregate_stats     monte_carlo.py:67-74
  * estimator        <- tail of euro_price_mc monte_carlo.py:142-167
  * euro_mc          <- euro_price_mc        monte_carlo.py:77-167 (serial branch;
                        the process-pool branch yields the same chunk seeds)
The arithmetic order follows the reference exactly so results are bit-identical
on the same numpy build.
"""
from __future__ import annotations

import math

import numpy as np


def chunk_normals(seed_seq, m):
    """The fp64 normals the reference draws for one chunk (monte_carlo.py:27,35)."""
    return np.random.default_rng(seed_seq).standard_normal(m)


def chunk_seeds(seed, n_chunks):
    """Child SeedSequences (0,), (1,), ... of the root (monte_carlo.py:106,120,131)."""
    return np.random.SeedSequence(seed).spawn(n_chunks)


def plan_chunks(n_paths, chunk_size):
    """monte_carlo.py:109-114."""
    sizes, left = [], int(n_paths)
    while left > 0:
        take = min(chunk_size, left)
        sizes.append(take)
        left -= take
    return sizes


def chunk_sumstats_from_z(z, *, S0, K, T, r, q, sigma, kind, antithetic, dtype=np.float64):
    """Six sufficient statistics from given normals (monte_carlo.py:28-64)."""
    if z.size == 0:
        return (0, 0.0, 0.0, 0.0, 0.0, 0.0)
    mu = (r - q - 0.5 * sigma * sigma) * T
    sig = sigma * math.sqrt(T)
    df = math.exp(-r * T)
    z = z.astype(dtype, copy=False)
    if antithetic:
        z = np.concatenate([z, -z], axis=0)
    st = (S0 * np.exp(mu + sig * z)).astype(dtype, copy=False)
    k = kind.lower()
    if k == "call":
        pay = np.maximum(st - K, 0.0, dtype=dtype)
    elif k == "put":
        pay = np.maximum(K - st, 0.0, dtype=dtype)
    else:
        raise ValueError("kind must be 'call' or 'put'")
    x = df * pay
    y = (df * st).astype(dtype, copy=False)
    return (x.size, float(x.sum()), float((x * x).sum()),
            float(y.sum()), float((y * y).sum()), float((x * y).sum()))


def chunk_sumstats(n, *, seed, **kw):
    if n <= 0:
        return (0, 0.0, 0.0, 0.0, 0.0, 0.0)
    return chunk_sumstats_from_z(chunk_normals(seed, n), **kw)


def combine(stats):
    """Elementwise sum of 6-tuples with Python's left-to-right sum (monte_carlo.py:67-74)."""
    return tuple(sum(s[i] for s in stats) for i in range(6))


def estimator(total, *, S0, q, T, control_variate):
    """(price, stderr) from the summed statistics (monte_carlo.py:142-167)."""
    n, sx, sxx, sy, syy, sxy = total
    if n == 0:
        return float("nan"), float("nan")
    mx = sx / n
    vx = max(0.0, sxx / n - mx * mx)
    if not control_variate:
        return float(mx), float(math.sqrt(max(0.0, vx) / n))
    my = sy / n
    vy = max(0.0, syy / n - my * my)
    cxy = (sxy / n) - mx * my
    c = 0.0 if vy == 0.0 else (cxy / vy)
    ey = S0 * math.exp(-q * T)
    price = mx - c * (my - ey)
    v = vx - 2.0 * c * cxy + (c * c) * vy
    return float(price), float(math.sqrt(max(0.0, v) / n))


def euro_mc(S0, K, T, r, sigma, q, kind, *, n_paths=100_000, seed=None, chunk_size=100_000,
            antithetic=True, control_variate=True, dtype=np.float64, return_stats=False):
    sizes = plan_chunks(n_paths, chunk_size)
    seeds = chunk_seeds(seed, len(sizes))
    per_chunk = [
        chunk_sumstats(m, seed=ss, S0=S0, K=K, T=T, r=r, q=q, sigma=sigma, kind=kind,
                       antithetic=antithetic, dtype=dtype)
        for m, ss in zip(sizes, seeds)
    ]
    total = combine(per_chunk) if per_chunk else (0, 0.0, 0.0, 0.0, 0.0, 0.0)
    out = estimator(total, S0=S0, q=q, T=T, control_variate=control_variate)
    return (out, per_chunk) if return_stats else out
