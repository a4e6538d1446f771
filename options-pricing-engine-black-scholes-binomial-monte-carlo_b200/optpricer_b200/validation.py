"""The validation loops that call the hot path, batched on the B200 (reference validation.py).

  stress_test           <- validation.py:181-231  3-D shock grid; pricer="tree" is ONE optp_crr_batch launch
                           (crr defaults: N=500, European), pricer="bs" the closed form on the host
  convergence_analysis  <- validation.py:103-174  method "tree" (crr, one launch per N) and "mc"
                           (euro_price_mc, seed=42)
The PDE / FEM methods of the reference ("fdm", "fem") are outside the B200 hot path
and are declined with NotImplementedError; cross_validate and backtest_delta_hedge are host orchestration
around them and are not provided.
"""
from __future__ import annotations

from math import exp, log, sqrt
from statistics import NormalDist

import numpy as np

from .binomial import _P_ERROR, crr, crr_batch
from .core import CALL, PUT, validate_contracts
from .monte_carlo import euro_price_mc

__all__ = ["stress_test", "convergence_analysis"]

_cdf = NormalDist().cdf


def _bs_price(S0, K, T, r, q, sigma, kind):
    """black_scholes.py:10-27."""
    if T <= 0 or sigma <= 0 or S0 <= 0 or K <= 0:
        raise ValueError("S0,K,T,sigma must be positive.")
    rt = sigma * sqrt(T)
    d1 = (log(S0 / K) + (r - q + 0.5 * sigma * sigma) * T) / rt
    d2 = d1 - rt
    if kind == CALL:
        return exp(-q * T) * S0 * _cdf(d1) - exp(-r * T) * K * _cdf(d2)
    if kind == PUT:
        return exp(-r * T) * K * _cdf(-d2) - exp(-q * T) * S0 * _cdf(-d1)
    raise ValueError("kind must be 'call' or 'put'")


def stress_test(opt, kind, spot_shocks, vol_shocks, rate_shocks, *, pricer="bs"):
    """Option price across a 3-D grid of market shocks: multiplicative on S0, additive on sigma (floored
    at 1e-6) and r.  Returns an array of shape (n_spot, n_vol, n_rate)."""
    spot_shocks = np.asarray(spot_shocks, dtype=float)
    vol_shocks = np.asarray(vol_shocks, dtype=float)
    rate_shocks = np.asarray(rate_shocks, dtype=float)
    if pricer == "fdm":
        raise NotImplementedError("pricer='fdm' (reference pde.py) is outside the B200 hot path")
    if pricer not in ("bs", "tree"):
        raise ValueError(f"Unknown pricer: {pricer}")
    shape = (len(spot_shocks), len(vol_shocks), len(rate_shocks))
    S = np.broadcast_to((opt.S0 * spot_shocks)[:, None, None], shape)
    sig = np.broadcast_to(np.maximum(opt.sigma + vol_shocks, 1e-6)[None, :, None], shape)
    r = np.broadcast_to((opt.r + rate_shocks)[None, None, :], shape)
    if pricer == "bs":
        out = np.empty(shape)
        for idx in np.ndindex(*shape):
            out[idx] = _bs_price(float(S[idx]), opt.K, opt.T, float(r[idx]), opt.q, float(sig[idx]), kind)
        return out
    validate_contracts(S, opt.K, opt.T, sig)   # OptionSpec(...) per grid point, reference validation.py:223-227 (replace() re-validates)
    price, status = crr_batch(S, opt.K, opt.T, r, opt.q, sig, 1 if kind == "call" else 0, 500,
                              american=False, return_status=True)
    if status.any():
        raise ValueError(_P_ERROR)
    return price


def convergence_analysis(opt, kind, method, param_name, param_values, *, reference=None, mode="fp32"):
    """Convergence of a numerical method as a parameter varies: "n_paths" for method "mc", "N" for "tree".
    Returns params, prices, errors and the estimated order (minus the log-log slope).
    `mode` is the Monte Carlo mode of euro_price_mc ("check" reproduces the reference's draws)."""
    param_values = list(param_values)
    if reference is None:
        reference = _bs_price(opt.S0, opt.K, opt.T, opt.r, getattr(opt, "q", 0.0), opt.sigma, kind)
    prices = []
    for val in param_values:
        val = int(val)
        if method == "mc":
            p = euro_price_mc(opt, kind, n_paths=val, seed=42, return_stderr=False, mode=mode)
        elif method == "tree":
            p = crr(opt, kind, N=val)
        elif method in ("fdm", "fem"):
            raise NotImplementedError(f"method={method!r} (reference pde.py / fem.py) is outside the B200 hot path")
        else:
            raise ValueError(f"Unknown method: {method}")
        prices.append(float(p))
    errors = [abs(p - reference) for p in prices]
    order = float("nan")
    valid = [(v, e) for v, e in zip(param_values, errors) if e > 0]
    if len(valid) >= 2:
        coeffs = np.polyfit(np.log([v for v, _ in valid]), np.log([e for _, e in valid]), 1)
        order = -float(coeffs[0])
    return {"params": param_values, "prices": prices, "errors": errors, "order": order}
