"""Oracle: path-dependent payoffs and estimators (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference src/optpricer/exotics.py.  Each pricer is split into a
per-path payoff vector (so the CUDA check mode can be compared path by path)
and the shared estimator.

  discounted_mean_se <- _price_from_payoff exotics.py:18-27  (ddof=1, n = all columns)
  barrier_payoff     <- barrier_price      exotics.py:33-96  (t=0 row monitored, rebate at T)
  asian_payoff       <- asian_price        exotics.py:102-163 (t=0 row excluded)
  digital_payoff     <- digital_price      exotics.py:169-209 (strict inequalities)
  lookback_payoff    <- lookback_price     exotics.py:215-273 (t=0 row included, floating has no max(.,0))
  vanilla_payoff     <- the vanilla leg of barrier_price exotics.py:83-88
  cv_estimator       <- control variate of monte_carlo.py:54-55,150-163 applied to a
                        path payoff (no reference pricer does this; SURVEY 8c "paths with no oracle":
                        the definition restated here is the oracle for the fused path's want_cv mode)
"""
from __future__ import annotations

import math

import numpy as np

BARRIER_TYPES = ("up-and-out", "up-and-in", "down-and-out", "down-and-in")


def _kind_check(kind):
    if kind not in ("call", "put"):
        raise ValueError("kind must be 'call' or 'put'")


def discounted_mean_se(payoff, r, T):
    x = np.exp(-r * T) * payoff
    n = x.size
    mean = float(x.mean())
    se = float(x.std(ddof=1) / np.sqrt(n)) if n > 1 else 0.0
    return mean, se


def vanilla_payoff(paths, K, kind):
    _kind_check(kind)
    st = paths[-1, :]
    return np.maximum(st - K, 0.0) if kind == "call" else np.maximum(K - st, 0.0)


def barrier_payoff(paths, K, kind, barrier, barrier_type, rebate=0.0):
    if barrier_type not in BARRIER_TYPES:
        raise ValueError(f"barrier_type must be one of {set(BARRIER_TYPES)}, got {barrier_type!r}")
    if barrier_type.startswith("up"):
        hit = np.any(paths >= barrier, axis=0)
    else:
        hit = np.any(paths <= barrier, axis=0)
    van = vanilla_payoff(paths, K, kind)
    if barrier_type.endswith("out"):
        return np.where(hit, rebate, van)
    return np.where(hit, van, rebate)


def asian_payoff(paths, K, kind, average_type="arithmetic", strike_type="fixed"):
    obs = paths[1:, :]
    st = paths[-1, :]
    if average_type == "arithmetic":
        avg = obs.mean(axis=0)
    elif average_type == "geometric":
        avg = np.exp(np.log(obs).mean(axis=0))
    else:
        raise ValueError("average_type must be 'arithmetic' or 'geometric'")
    if strike_type not in ("fixed", "floating"):
        raise ValueError("strike_type must be 'fixed' or 'floating'")
    _kind_check(kind)
    if strike_type == "fixed":
        return np.maximum(avg - K, 0.0) if kind == "call" else np.maximum(K - avg, 0.0)
    return np.maximum(st - avg, 0.0) if kind == "call" else np.maximum(avg - st, 0.0)


def digital_payoff(paths, K, kind, payout=1.0):
    _kind_check(kind)
    st = paths[-1, :]
    return np.where(st > K, payout, 0.0) if kind == "call" else np.where(st < K, payout, 0.0)


def lookback_payoff(paths, kind, K=0.0, strike_type="floating"):
    hi = paths.max(axis=0)
    lo = paths.min(axis=0)
    st = paths[-1, :]
    if strike_type not in ("floating", "fixed"):
        raise ValueError("strike_type must be 'floating' or 'fixed'")
    _kind_check(kind)
    if strike_type == "floating":
        return st - lo if kind == "call" else hi - st
    return np.maximum(hi - K, 0.0) if kind == "call" else np.maximum(K - lo, 0.0)


def barrier(paths, K, r, T, kind, barrier, barrier_type, rebate=0.0):
    return discounted_mean_se(barrier_payoff(paths, K, kind, barrier, barrier_type, rebate), r, T)


def asian(paths, K, r, T, kind, average_type="arithmetic", strike_type="fixed"):
    return discounted_mean_se(asian_payoff(paths, K, kind, average_type, strike_type), r, T)


def digital(paths, K, r, T, kind, payout=1.0):
    return discounted_mean_se(digital_payoff(paths, K, kind, payout), r, T)


def lookback(paths, r, T, kind, K=0.0, strike_type="floating"):
    return discounted_mean_se(lookback_payoff(paths, kind, K, strike_type), r, T)


def six_stats(payoff, terminal, r, T):
    """(n, sumX, sumX2, sumY, sumY2, sumXY) with X = e^{-rT} payoff, Y = e^{-rT} S_T."""
    df = math.exp(-r * T)
    x = df * payoff
    y = df * terminal
    return (x.size, float(x.sum()), float((x * x).sum()),
            float(y.sum()), float((y * y).sum()), float((x * y).sum()))


def cv_estimator(stats, expected_y):
    """Control-variate (price, stderr) from six stats; formulas of monte_carlo.py:147-163."""
    n, sx, sxx, sy, syy, sxy = stats
    mx, my = sx / n, sy / n
    vx = max(0.0, sxx / n - mx * mx)
    vy = max(0.0, syy / n - my * my)
    cxy = sxy / n - mx * my
    c = 0.0 if vy == 0.0 else cxy / vy
    price = mx - c * (my - expected_y)
    var = vx - 2.0 * c * cxy + c * c * vy
    return float(price), float(math.sqrt(max(0.0, var) / n))
