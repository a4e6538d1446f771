"""Bump-and-reprice risk callers of the hot path, batched on the B200 (drop-in for the reprice loops
of reference risk.py).

  numerical_greeks  <- risk.py:27-92    8 reprices of one contract      -> ONE device batch
  scenario_grid     <- risk.py:99-139   n_spot x n_vol reprices         -> ONE device batch
  portfolio_risk    <- risk.py:146-203  9 reprices per instrument       -> ONE device batch for the book

Same signatures, bump rules and result dicts as the reference.  `pricer_func` is, as there, any callable
`f(S, K, T, r, q, sigma, kind) -> float`; when it also has a `.batch(S, K, T, r, q, sigma, kind)` method
taking equal-length sequences (TreePricer and McPricer below do), all bumped contracts are priced in one
call: one optp_crr_batch launch for trees, one optp_mc_book call for Monte Carlo, where every bumped
contract reuses the same Philox streams (common random numbers, as the reference gets by re-seeding
euro_price_mc on every call).  A plain callable is evaluated point by point exactly like the reference.
"""
from __future__ import annotations

import numpy as np

from . import monte_carlo as mc
from .binomial import _P_ERROR, crr_batch
from .core import OptionSpec

__all__ = ["TreePricer", "McPricer", "numerical_greeks", "scenario_grid", "portfolio_risk"]


class TreePricer:
    """CRR pricer with the reference's pricer_func signature; `.batch` prices many contracts in one launch."""

    def __init__(self, N=500, *, american=False):
        if N <= 0:
            raise ValueError("N must be positive.")
        self.N, self.american = int(N), bool(american)

    def batch(self, S, K, T, r, q, sigma, kind):
        is_call = np.array([1 if k == "call" else 0 for k in kind], dtype=np.uint8)   # binomial.py:22-25
        price, status = crr_batch(S, K, T, r, q, sigma, is_call, self.N, american=self.american,
                                  return_status=True)
        if status.any():
            raise ValueError(_P_ERROR)   # the reference's loop would raise on the first such contract
        return price

    def __call__(self, S, K, T, r, q, sigma, kind):
        return float(self.batch([S], [K], [T], [r], [q], [sigma], [kind])[0])


class McPricer:
    """euro_price_mc with the reference's pricer_func signature.  `.batch` runs all contracts as items of
    one optp_mc_book call with the same seed (common random numbers across bumps)."""

    def __init__(self, n_paths=100_000, *, seed=1, antithetic=True, control_variate=True, mode="fp32"):
        self.n_paths, self.seed = int(n_paths), seed
        self.antithetic, self.control_variate, self.mode = bool(antithetic), bool(control_variate), mode

    def __call__(self, S, K, T, r, q, sigma, kind):
        return mc.euro_price_mc(OptionSpec(S, K, T, r, sigma, q), kind, n_paths=self.n_paths, seed=self.seed,
                                antithetic=self.antithetic, control_variate=self.control_variate,
                                return_stderr=False, mode=self.mode)

    def batch(self, S, K, T, r, q, sigma, kind):
        n = len(S)
        if self.mode == "check" or not self.antithetic or self.n_paths <= 0:
            # host-normal mode (the reference's PCG64 draws) and the rarely used plain estimator go
            # through the scalar entry point
            return np.array([self(S[i], K[i], T[i], r[i], q[i], sigma[i], kind[i]) for i in range(n)])
        from .book import book_stats
        seed = mc._seed64(self.seed)   # one key for every item: common random numbers across the bumps
        items = []
        for i in range(n):
            OptionSpec(S[i], K[i], T[i], r[i], sigma[i], q[i])   # the reference's argument checks
            items.append(dict(kind="terminal", rows=[(i, None)], n_base=self.n_paths, seed=seed,
                              S0=float(S[i]), K=float(K[i]), T=float(T[i]), r=float(r[i]), q=float(q[i]),
                              sigma=float(sigma[i]), is_call=mc._kind_flag(kind[i])))
        stats = book_stats(items, mode=self.mode)
        out = np.empty(n)
        for i, s in enumerate(stats):
            out[i] = mc.estimator((int(s[0]), *s[1:6]), S0=items[i]["S0"], q=items[i]["q"], T=items[i]["T"],
                                  control_variate=self.control_variate, return_stderr=False)
        return out


def _price_points(pricer_func, points):
    """points: list of (S, K, T, r, q, sigma, kind) -> list of floats."""
    if not points:
        return []
    if hasattr(pricer_func, "batch"):
        cols = list(zip(*points))
        return [float(v) for v in pricer_func.batch(*[list(c) for c in cols])]
    return [float(pricer_func(*p)) for p in points]


_DT = 1.0 / 365.0


def _greek_points(S, K, T, r, q, sigma, kind, bump_pct):
    """The reprice points of risk.py:57-85, in the reference's evaluation order."""
    eps_S = bump_pct * S
    eps_v = max(bump_pct * sigma, 1e-4)
    eps_r = bump_pct
    pts = [(S, K, T, r, q, sigma, kind),
           (S + eps_S, K, T, r, q, sigma, kind), (S - eps_S, K, T, r, q, sigma, kind),
           (S, K, T, r, q, sigma + eps_v, kind), (S, K, T, r, q, max(sigma - eps_v, 1e-6), kind)]
    has_theta = T > _DT
    if has_theta:
        pts.append((S, K, T - _DT, r, q, sigma, kind))
    pts += [(S, K, T, r + eps_r, q, sigma, kind), (S, K, T, r - eps_r, q, sigma, kind)]
    return pts, (eps_S, eps_v, eps_r, has_theta)


def _greeks_from_prices(P, eps):
    eps_S, eps_v, eps_r, has_theta = eps
    P0, P_up, P_dn, P_vup, P_vdn = P[:5]
    i = 5
    theta = 0.0
    if has_theta:
        theta = (P[5] - P0) / _DT
        i = 6
    P_rup, P_rdn = P[i], P[i + 1]
    return {"delta": float((P_up - P_dn) / (2.0 * eps_S)),
            "gamma": float((P_up - 2.0 * P0 + P_dn) / (eps_S ** 2)),
            "vega": float((P_vup - P_vdn) / (2.0 * eps_v)),
            "theta": float(theta),
            "rho": float((P_rup - P_rdn) / (2.0 * eps_r))}


def numerical_greeks(pricer_func, S, K, T, r, q, sigma, kind, *, bump_pct=0.01):
    """Greeks via central finite differences on an arbitrary pricer (risk.py:27-92).
    Keys: delta, gamma, vega, theta, rho."""
    pts, eps = _greek_points(S, K, T, r, q, sigma, kind, bump_pct)
    return _greeks_from_prices(_price_points(pricer_func, pts), eps)


def scenario_grid(pricer_func, S, K, T, r, q, sigma, kind, spot_range, vol_range):
    """Evaluate a pricer across a 2-D (spot x vol) scenario grid (risk.py:99-139)."""
    spot_range = np.asarray(spot_range, dtype=float)
    vol_range = np.asarray(vol_range, dtype=float)
    pts = [(float(s), K, T, r, q, float(v), kind) for s in spot_range for v in vol_range]
    prices = np.array(_price_points(pricer_func, pts), dtype=float).reshape(len(spot_range), len(vol_range))
    return {"spot_values": spot_range.copy(), "vol_values": vol_range.copy(), "prices": prices}


def portfolio_risk(instruments, pricer_func, *, bump_pct=0.01):
    """Aggregate risk for a portfolio of instruments (risk.py:146-203): dicts with keys S, K, T, r, q,
    sigma, kind, position.  All reprices of all instruments form one device batch."""
    pts, spans = [], []
    for inst in instruments:
        p, eps = _greek_points(inst["S"], inst["K"], inst["T"], inst["r"], inst["q"], inst["sigma"],
                               inst["kind"], bump_pct)
        spans.append((len(pts), len(p), eps))
        pts += p
    prices = _price_points(pricer_func, pts)
    totals = {"delta": 0.0, "gamma": 0.0, "vega": 0.0, "theta": 0.0, "rho": 0.0}
    total_value = 0.0
    inst_greeks = []
    for inst, (lo, n, eps) in zip(instruments, spans):
        pos = inst["position"]
        g = _greeks_from_prices(prices[lo:lo + n], eps)
        price = prices[lo]   # risk.py:188-191 re-prices the unbumped contract: the same number
        scaled = {k: pos * v for k, v in g.items()}
        for k in totals:
            totals[k] += scaled[k]
        total_value += pos * price
        inst_greeks.append({**scaled, "price": pos * price})
    return {"total_delta": totals["delta"], "total_gamma": totals["gamma"], "total_vega": totals["vega"],
            "total_theta": totals["theta"], "total_rho": totals["rho"], "total_value": total_value,
            "instrument_greeks": inst_greeks}
