"""Terminal-only European Monte Carlo on the B200 (drop-in for reference monte_carlo.py).

  _mc_chunk_sumstats <- monte_carlo.py:11-64    one chunk -> six sufficient statistics
  _aggregate_stats   <- monte_carlo.py:67-74
  euro_price_mc      <- monte_carlo.py:77-167   same signature + two optional keywords

New optional keywords (defaults keep the call a drop-in):
  mode = "fp32"   production: ONE fused launch over all n_paths, Philox4x32-10 normals generated
                  in registers, fp32 path arithmetic, fp64 accumulators.  `chunk_size`, `n_workers`
                  and `dtype` do not change the result (draws are indexed by path, not by chunk).
         "fp64"   same draws, the reference's fp64 operation order.
         "check"  deterministic check: the reference's own PCG64 draws, chunk by chunk
                  (SeedSequence children, monte_carlo.py:106,120,131) are generated with numpy on
                  the host and consumed by the fp64 kernel; results match the reference to 1e-12.
  n_gpus / rank   see optpricer_b200.distributed for the sharded variant.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import _native as nv

_MODES = {"fp32": nv.MODE_FP32_PHILOX, "fp64": nv.MODE_FP64_PHILOX, "check": nv.MODE_FP64_HOSTZ}


def _kind_flag(kind):
    k = kind.lower()
    if k == "call":
        return 1
    if k == "put":
        return 0
    raise ValueError("kind must be 'call' or 'put'")


def _seed64(seed):
    """64-bit Philox key from the user's seed; seed=None draws OS entropy (reference: processes.py:23-24)."""
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def terminal_stats(S0, K, T, r, q, sigma, is_call, n_base, *, antithetic=True, mode="fp32", seed=0,
                   path_offset=0, z=None, fp32_cast=False):
    """Raw call of optp_mc_terminal: returns the 8-slot statistics block as a numpy array."""
    lib = nv.init()
    stats = np.zeros(nv.NSTATS, dtype=np.float64)
    zc = None if z is None else nv.f64(z)
    nv.check(lib.optp_mc_terminal(float(S0), float(K), float(T), float(r), float(q), float(sigma),
                                  int(is_call), int(n_base), 1 if antithetic else 0, _MODES[mode],
                                  1 if fp32_cast else 0, int(seed), int(path_offset), nv.ptr(zc),
                                  nv.ptr(stats)), "optp_mc_terminal")
    return stats


def _mc_chunk_sumstats(n, *, S0, K, T, r, q, sigma, kind, antithetic, seed, dtype=np.float64):
    """One chunk in the reference's own draws (monte_carlo.py:11-64): (n_eff, sumX, sumX2, sumY, sumY2, sumXY)."""
    if n <= 0:
        return (0, 0.0, 0.0, 0.0, 0.0, 0.0)
    is_call = _kind_flag(kind)
    z = np.random.default_rng(seed).standard_normal(n)
    s = terminal_stats(S0, K, T, r, q, sigma, is_call, n, antithetic=antithetic, mode="check", z=z,
                       fp32_cast=(np.dtype(dtype) == np.float32))
    return (int(s[0]), float(s[1]), float(s[2]), float(s[3]), float(s[4]), float(s[5]))


def _aggregate_stats(stats_list):
    n = sum(s[0] for s in stats_list)
    return (n, sum(s[1] for s in stats_list), sum(s[2] for s in stats_list), sum(s[3] for s in stats_list),
            sum(s[4] for s in stats_list), sum(s[5] for s in stats_list))


def estimator(total, *, S0, q, T, control_variate=True, return_stderr=True):
    """(price, stderr) from the six summed statistics (monte_carlo.py:142-167)."""
    n, sumX, sumX2, sumY, sumY2, sumXY = total[:6]
    if n == 0:
        return float("nan"), float("nan")
    meanX = sumX / n
    varX = max(0.0, sumX2 / n - meanX * meanX)
    if control_variate:
        meanY = sumY / n
        varY = max(0.0, sumY2 / n - meanY * meanY)
        covXY = (sumXY / n) - meanX * meanY
        c_hat = 0.0 if varY == 0.0 else (covXY / varY)
        EY = S0 * math.exp(-q * T)
        mean_cv = meanX - c_hat * (meanY - EY)
        var_cv = varX - 2.0 * c_hat * covXY + (c_hat * c_hat) * varY
        se = math.sqrt(max(0.0, var_cv) / n)
        return (float(mean_cv), float(se)) if return_stderr else float(mean_cv)
    se = math.sqrt(max(0.0, varX) / n)
    return (float(meanX), float(se)) if return_stderr else float(meanX)


def euro_price_mc(opt, kind, *, n_paths=100_000, seed=None, chunk_size=100_000, antithetic=True,
                  control_variate=True, n_workers=1, dtype=np.float64, return_stderr=True,
                  mode="fp32"):
    """European option Monte Carlo price, returns (price, stderr) (monte_carlo.py:77-167)."""
    S0, K, T, r, sigma = opt.S0, opt.K, opt.T, opt.r, opt.sigma
    q = getattr(opt, "q", 0.0)
    if mode not in _MODES:
        raise ValueError("mode must be 'fp32', 'fp64' or 'check'")
    n_paths = int(n_paths)
    if mode == "check":
        ss_root = np.random.SeedSequence(seed)
        chunks, remaining = [], n_paths
        while remaining > 0:
            m = min(chunk_size, remaining)
            chunks.append(m)
            remaining -= m
        stats_list = []
        for m in chunks:
            ss = ss_root.spawn(1)[0]
            stats_list.append(_mc_chunk_sumstats(m, S0=S0, K=K, T=T, r=r, q=q, sigma=sigma, kind=kind,
                                                 antithetic=antithetic, seed=ss, dtype=dtype))
        total = _aggregate_stats(stats_list)
    else:
        is_call = _kind_flag(kind)
        if n_paths <= 0:
            total = (0, 0.0, 0.0, 0.0, 0.0, 0.0)
        else:
            s = terminal_stats(S0, K, T, r, q, sigma, is_call, n_paths, antithetic=antithetic, mode=mode,
                               seed=_seed64(seed))
            total = (int(s[0]), *[float(v) for v in s[1:6]])
    return estimator(total, S0=S0, q=q, T=T, control_variate=control_variate, return_stderr=return_stderr)
