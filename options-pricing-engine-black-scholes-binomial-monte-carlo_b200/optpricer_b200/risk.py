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

from functools import reduce
from operator import add

import numpy as np

from . import monte_carlo as mc
from .binomial import _P_ERROR, crr_batch
from .core import OptionSpec, validate_contracts

__all__ = ["TreePricer", "McPricer", "numerical_greeks", "scenario_grid", "portfolio_risk"]


class TreePricer:
    """CRR pricer with the reference's pricer_func signature; `.batch` prices many contracts in one launch."""

    def __init__(self, N=500, *, american=False):
        if N <= 0:
            raise ValueError("N must be positive.")
        self.N, self.american = int(N), bool(american)

    def batch(self, S, K, T, r, q, sigma, kind):
        is_call = np.fromiter((k == "call" for k in kind), dtype=bool, count=len(kind))   # binomial.py:22-25
        return self.batch_flags(S, K, T, r, q, sigma, is_call)

    def batch_flags(self, S, K, T, r, q, sigma, is_call):
        """`batch` with the kinds already reduced to flags (1 = call): the callers below bump every contract eight
        times, and 160k string compares cost more than pricing 160k trees."""
        validate_contracts(S, K, T, sigma)   # the OptionSpec(...) a reference pricer_func builds per point
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


def _price_points(pricer_func, cols, kinds):
    """cols: six equal-length float arrays (S, K, T, r, q, sigma); kinds: list of str -> float array."""
    n = len(kinds)
    if n == 0:
        return np.empty(0)
    if hasattr(pricer_func, "batch"):
        return np.asarray(pricer_func.batch(*cols, kinds), dtype=float)
    return np.array([float(pricer_func(*[float(c[i]) for c in cols], kinds[i])) for i in range(n)])


_DT = 1.0 / 365.0
_SLOTS = 8   # P0, S up, S down, vol up, vol down, T - 1 day, r up, r down (risk.py:57-85, in that order)


def _greek_points(S, K, T, r, q, sigma, bump_pct):
    """The reprice points of risk.py:57-85 for n contracts at once: six (n, 8) arrays, the bump sizes and
    the theta mask.  Contracts with T <= 1 day have no theta point (risk.py:74-79); their slot repeats the
    unbumped contract and is ignored."""
    S, K, T, r, q, sigma = [np.asarray(x, dtype=float).reshape(-1) for x in (S, K, T, r, q, sigma)]
    eps_S = bump_pct * S
    eps_v = np.maximum(bump_pct * sigma, 1e-4)
    eps_r = bump_pct
    has_theta = T > _DT
    rep = lambda x: np.repeat(x[:, None], _SLOTS, axis=1)  # noqa: E731
    Sm, Km, Tm, rm, qm, vm = rep(S), rep(K), rep(T), rep(r), rep(q), rep(sigma)
    Sm[:, 1] = S + eps_S
    Sm[:, 2] = S - eps_S
    vm[:, 3] = sigma + eps_v
    vm[:, 4] = np.maximum(sigma - eps_v, 1e-6)
    Tm[:, 5] = np.where(has_theta, T - _DT, T)
    rm[:, 6] = r + eps_r
    rm[:, 7] = r - eps_r
    return (Sm, Km, Tm, rm, qm, vm), (eps_S, eps_v, eps_r, has_theta)


def _greeks_from_prices(P, eps):
    """P: (n, 8) prices -> dict of (n,) arrays."""
    eps_S, eps_v, eps_r, has_theta = eps
    return {"delta": (P[:, 1] - P[:, 2]) / (2.0 * eps_S),
            "gamma": (P[:, 1] - 2.0 * P[:, 0] + P[:, 2]) / (eps_S ** 2),
            "vega": (P[:, 3] - P[:, 4]) / (2.0 * eps_v),
            "theta": np.where(has_theta, (P[:, 5] - P[:, 0]) / _DT, 0.0),
            "rho": (P[:, 6] - P[:, 7]) / (2.0 * eps_r)}


def _reprice(pricer_func, S, K, T, r, q, sigma, kinds, bump_pct):
    mats, eps = _greek_points(S, K, T, r, q, sigma, bump_pct)
    if not hasattr(pricer_func, "batch"):
        # a plain callable is evaluated exactly where the reference evaluates it (no theta point for
        # contracts inside their last day)
        P = np.empty(mats[0].shape)
        for i in range(P.shape[0]):
            for j in range(_SLOTS):
                if j == 5 and not eps[3][i]:
                    P[i, j] = P[i, 0]
                else:
                    P[i, j] = float(pricer_func(*[float(m[i, j]) for m in mats], kinds[i]))
        return P, eps
    if hasattr(pricer_func, "batch_flags"):
        flags = np.repeat(np.fromiter((k == "call" for k in kinds), dtype=bool, count=len(kinds)), _SLOTS)
        P = np.asarray(pricer_func.batch_flags(*[m.reshape(-1) for m in mats], flags), dtype=float)
        return P.reshape(mats[0].shape), eps
    flat_kinds = [k for k in kinds for _ in range(_SLOTS)]
    P = _price_points(pricer_func, [m.reshape(-1) for m in mats], flat_kinds).reshape(mats[0].shape)
    return P, eps


def numerical_greeks(pricer_func, S, K, T, r, q, sigma, kind, *, bump_pct=0.01):
    """Greeks via central finite differences on an arbitrary pricer (risk.py:27-92).
    Keys: delta, gamma, vega, theta, rho."""
    P, eps = _reprice(pricer_func, [S], [K], [T], [r], [q], [sigma], [kind], bump_pct)
    return {k: float(v[0]) for k, v in _greeks_from_prices(P, eps).items()}


def scenario_grid(pricer_func, S, K, T, r, q, sigma, kind, spot_range, vol_range):
    """Evaluate a pricer across a 2-D (spot x vol) scenario grid (risk.py:99-139)."""
    spot_range = np.asarray(spot_range, dtype=float)
    vol_range = np.asarray(vol_range, dtype=float)
    shape = (len(spot_range), len(vol_range))
    n = shape[0] * shape[1]
    Sg = np.broadcast_to(spot_range[:, None], shape).reshape(-1)
    vg = np.broadcast_to(vol_range[None, :], shape).reshape(-1)
    cols = [Sg, np.full(n, float(K)), np.full(n, float(T)), np.full(n, float(r)), np.full(n, float(q)), vg]
    prices = _price_points(pricer_func, cols, [kind] * n).reshape(shape)
    return {"spot_values": spot_range.copy(), "vol_values": vol_range.copy(), "prices": prices}


def portfolio_risk(instruments, pricer_func, *, bump_pct=0.01):
    """Aggregate risk for a portfolio of instruments (risk.py:146-203): dicts with keys S, K, T, r, q,
    sigma, kind, position.  All reprices of all instruments form one device batch."""
    get = lambda key: np.array([inst[key] for inst in instruments], dtype=float)  # noqa: E731
    kinds = [inst["kind"] for inst in instruments]
    pos = get("position")
    if len(instruments):
        P, eps = _reprice(pricer_func, get("S"), get("K"), get("T"), get("r"), get("q"), get("sigma"), kinds,
                          bump_pct)
        g = _greeks_from_prices(P, eps)
    else:
        P, g = np.empty((0, _SLOTS)), {k: np.empty(0) for k in ("delta", "gamma", "vega", "theta", "rho")}
    names = ("delta", "gamma", "vega", "theta", "rho")
    scaled_cols = [(pos * g[k]).tolist() for k in names]
    values = (pos * P[:, 0]).tolist()   # risk.py:188-191 re-prices the unbumped contract: the same number
    # summed in instrument order from 0.0 with plain additions, like the reference's loop (builtin sum() compensates
    # since Python 3.12, numpy sums pairwise: both would differ from it in the last digits)
    totals = {k: reduce(add, col, 0.0) for k, col in zip(names, scaled_cols)}
    total_value = reduce(add, values, 0.0)
    inst_greeks = [{"delta": a, "gamma": b, "vega": c, "theta": d, "rho": e, "price": v}
                   for a, b, c, d, e, v in zip(*scaled_cols, values)]
    return {"total_delta": totals["delta"], "total_gamma": totals["gamma"], "total_vega": totals["vega"],
            "total_theta": totals["theta"], "total_rho": totals["rho"], "total_value": total_value,
            "instrument_greeks": inst_greeks}
