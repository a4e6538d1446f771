"""Path-dependent payoff pricers on the B200 (drop-in for reference exotics.py).

  _price_from_payoff <- exotics.py:18-27   (discount, mean, ddof=1 standard error)
  barrier_price      <- exotics.py:33-96
  asian_price        <- exotics.py:102-163
  digital_price      <- exotics.py:169-209
  lookback_price     <- exotics.py:215-273

`paths` may be a LazyPaths from optpricer_b200.processes -- then generator + payoff + reduction
run as ONE fused kernel and no path is ever stored -- or any (n_steps+1, n_paths) array, which
is reduced on the device by optp_payoff_from_paths.  Argument checks, their order and their
messages follow the reference.

New: price_payoffs(paths, specs) evaluates up to 16 payoffs on the same path set in one launch
(scripts/price_book.py:80,92-93 seeds every row alike -- common random numbers), and every
pricer accepts control_variate=True to apply the reference's control variate Y = e^{-rT} S_T
(monte_carlo.py:54-55,150-163) to a path payoff (the reference has it for euro_price_mc only).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as nv
from .processes import LazyPaths

_VALID_BARRIERS = {"up-and-out", "up-and-in", "down-and-out", "down-and-in"}
_KIND_ERR = "kind must be 'call' or 'put'"


def _kind_flag(kind):
    if kind == "call":
        return nv.F_CALL
    if kind == "put":
        return 0
    raise ValueError(_KIND_ERR)


# ---- payoff descriptors -------------------------------------------------------------------
def barrier_spec(K, kind, barrier, barrier_type, rebate=0.0):
    if barrier_type not in _VALID_BARRIERS:
        raise ValueError(f"barrier_type must be one of {_VALID_BARRIERS}, got {barrier_type!r}")
    flags = _kind_flag(kind)
    if barrier_type.startswith("up"):
        flags |= nv.F_UP
    if barrier_type.endswith("in"):
        flags |= nv.F_IN
    return dict(kind=nv.PAY_BARRIER, flags=flags, K=K, barrier=barrier, rebate=rebate)


def asian_spec(K, kind, average_type="arithmetic", strike_type="fixed"):
    if average_type not in ("arithmetic", "geometric"):
        raise ValueError("average_type must be 'arithmetic' or 'geometric'")
    if strike_type not in ("fixed", "floating"):
        raise ValueError("strike_type must be 'fixed' or 'floating'")
    flags = _kind_flag(kind)
    if average_type == "geometric":
        flags |= nv.F_GEOMETRIC
    if strike_type == "floating":
        flags |= nv.F_FLOATING
    return dict(kind=nv.PAY_ASIAN, flags=flags, K=K)


def digital_spec(K, kind, payout=1.0):
    return dict(kind=nv.PAY_DIGITAL, flags=_kind_flag(kind), K=K, payout=payout)


def lookback_spec(kind, K=0.0, strike_type="floating"):
    if strike_type not in ("floating", "fixed"):
        raise ValueError("strike_type must be 'floating' or 'fixed'")
    flags = _kind_flag(kind)
    if strike_type == "floating":
        flags |= nv.F_FLOATING
    return dict(kind=nv.PAY_LOOKBACK, flags=flags, K=K)


def vanilla_spec(K, kind):
    return dict(kind=nv.PAY_VANILLA, flags=_kind_flag(kind), K=K)


def _native_payoffs(specs, r, T):
    arr = (nv.OptpPayoff * len(specs))()
    for i, s in enumerate(specs):
        arr[i].kind = s["kind"]
        arr[i].flags = s.get("flags", 0)
        arr[i].K = float(s.get("K", 0.0))
        arr[i].barrier = float(s.get("barrier", 0.0))
        arr[i].rebate = float(s.get("rebate", 0.0))
        arr[i].payout = float(s.get("payout", 1.0))
        arr[i].r_disc = float(r)
        arr[i].T_disc = float(T)
    return arr


# ---- estimators ------------------------------------------------------------------------------
def estimate(stats, *, control_variate=False, expected_y=None):
    """(price, stderr) from one statistics block.  Plain: exotics.py:22-27 (ddof=1, n = all columns).
    control_variate: formulas of monte_carlo.py:147-163 (ddof=0) with E[Y] = expected_y."""
    n, sx, sxx, sy, syy, sxy = (float(v) for v in stats[:6])
    if n <= 0:
        return float("nan"), float("nan")
    mean = sx / n
    if not control_variate:
        if n <= 1:
            return mean, 0.0
        var = max(0.0, (sxx - n * mean * mean) / (n - 1.0))
        return mean, math.sqrt(var) / math.sqrt(n)
    my = sy / n
    vx = max(0.0, sxx / n - mean * mean)
    vy = max(0.0, syy / n - my * my)
    cxy = sxy / n - mean * my
    c = 0.0 if vy == 0.0 else cxy / vy
    price = mean - c * (my - expected_y)
    var = vx - 2.0 * c * cxy + c * c * vy
    return float(price), math.sqrt(max(0.0, var) / n)


def _price_from_payoff(payoff, r, T):
    """Discount, average and compute the standard error of a payoff vector (exotics.py:18-27)."""
    payoff = np.asarray(payoff, dtype=np.float64)
    X = np.exp(-r * T) * payoff
    n = X.size
    price = float(X.mean())
    stderr = float(X.std(ddof=1) / np.sqrt(n)) if n > 1 else 0.0
    return price, stderr


# ---- device calls -------------------------------------------------------------------------------
def path_stats(paths, specs, r, T, *, want_payoffs=False, want_terminal=False, n_base=None,
               path_offset=0):
    """Run the fused kernel for a LazyPaths and return (stats[n_specs, 8], payoffs|None, terminal|None).
    n_base / path_offset select a sub-range of base paths (multi-GPU sharding)."""
    lib = nv.init()
    if len(specs) > nv.MAX_PAYOFFS or not specs:
        raise ValueError(f"between 1 and {nv.MAX_PAYOFFS} payoffs per call")
    pay = _native_payoffs(specs, r, T)
    nb = paths.n_paths if n_base is None else int(n_base)
    n_eff = 2 * nb if paths.antithetic else nb
    stats = np.zeros((len(specs), nv.NSTATS), dtype=np.float64)
    pout = np.empty((len(specs), n_eff), dtype=np.float64) if want_payoffs else None
    tout = np.empty(n_eff, dtype=np.float64) if want_terminal else None
    z = None
    if paths.rng == "numpy":
        if path_offset or nb != paths.n_paths:
            raise ValueError("rng='numpy' (check mode) does not shard")
        z = paths.host_normals()
    p = paths.native_process()
    nv.check(lib.optp_mc_path(C.byref(p), pay, len(specs), paths.n_steps, nb, int(paths.antithetic),
                              paths.native_mode(), paths.seed, int(path_offset), nv.ptr(z), nv.ptr(pout),
                              nv.ptr(tout), nv.ptr(stats)), "optp_mc_path")
    _warn_if_clamped(paths, stats)
    return stats, pout, tout


CLAMP_WARN_FRACTION = 1e-4


def _warn_if_clamped(paths, stats):
    """Local-vol runs: statistics slot 6 counts the effective paths (fp64 kernels: lookups) that left the
    tabulated ln S range, where sigma is the edge value instead of the reference's sigma_loc(S_t, t)
    (processes.py:226,361).  A negligible count is the design; anything else must not pass silently."""
    if paths.table is None or stats[0][0] <= 0:
        return
    frac = float(stats[0][6]) / float(stats[0][0])
    if frac > CLAMP_WARN_FRACTION:
        import warnings
        warnings.warn(f"{frac:.2%} of the paths left the tabulated sigma(S,t) range "
                      f"[{math.exp(paths.table.x0):.4g}, "
                      f"{math.exp(paths.table.x0 + paths.table.h * (paths.table.table.shape[1] - 1)):.4g}] and saw the "
                      "edge volatility there; pass a wider sigma_grid=(n_nodes, half_width)", RuntimeWarning,
                      stacklevel=3)


def array_stats(paths, spec, r, T, *, want_payoffs=False):
    """Reduce a caller-supplied (n_rows, n_cols) array on the device (the literal exotics.py contract)."""
    lib = nv.init()
    a = nv.f64(paths)
    if a.ndim != 2:
        raise ValueError("paths must be a 2-D array (n_steps+1, n_paths)")
    pay = _native_payoffs([spec], r, T)
    stats = np.zeros(nv.NSTATS, dtype=np.float64)
    pout = np.empty(a.shape[1], dtype=np.float64) if want_payoffs else None
    nv.check(lib.optp_payoff_from_paths(nv.ptr(a), a.shape[0], a.shape[1], pay, nv.ptr(pout), nv.ptr(stats)),
             "optp_payoff_from_paths")
    return stats, pout


def _expected_y(paths, r, T):
    """E[e^{-rT} S_T] of the DISCRETISED process the kernel simulates (monte_carlo.py:157 has the exact GBM
    value S0 e^{-qT}).  Exact and log-Euler schemes (GBM, Heston FTE, LV log-Euler, Merton, SABR beta = 1) keep
    the continuous mean S0 e^{(r-q)T}; the Euler / Milstein schemes in S (gbm_milstein, LV Milstein, SABR
    beta < 1) have E[S_T] = S0 (1 + (r-q) dt)^n before their floors -- using the continuous mean there would
    bias the control-variate estimator by c (E_disc[Y] - EY), which does not shrink with the path count.
    Heston QE without the martingale correction has no closed-form discrete mean: declined."""
    pr = paths.proc
    kind = pr["kind"]
    rq = pr["r"] - pr["q"]
    if kind == nv.PROC_HESTON_QE:
        raise ValueError("control_variate needs E[S_T] of the discretised process; the QE scheme has it only "
                         "with the martingale correction (heston_paths(..., scheme='qe', martingale_correction=True))")
    euler_in_S = kind in (nv.PROC_GBM_MILSTEIN, nv.PROC_LV_MILSTEIN) or (kind == nv.PROC_SABR and pr["kappa"] != 1.0)
    if euler_in_S:
        growth = (1.0 + rq * pr["T"] / paths.n_steps) ** paths.n_steps
    else:
        growth = math.exp(rq * pr["T"])
    return pr["S0"] * growth * math.exp(-r * T)


def price_payoffs(paths, specs, r, T, *, control_variate=False):
    """Price several payoffs on one path set; returns a list of (price, stderr)."""
    if isinstance(paths, LazyPaths):
        stats, _, _ = path_stats(paths, specs, r, T)
        ey = _expected_y(paths, r, T) if control_variate else None
        return [estimate(s, control_variate=control_variate, expected_y=ey) for s in stats]
    if control_variate:
        raise ValueError("control_variate needs a LazyPaths (E[S_T] of a bare array is unknown)")
    return [estimate(array_stats(paths, s, r, T)[0]) for s in specs]


def _one(paths, spec, r, T, control_variate):
    return price_payoffs(paths, [spec], r, T, control_variate=control_variate)[0]


# ---- the reference's four pricers ---------------------------------------------------------------
def barrier_price(paths, K, r, T, kind, barrier, barrier_type, rebate=0.0, *, control_variate=False):
    """European barrier option, discrete monitoring incl. t=0, rebate paid at T (exotics.py:33-96)."""
    return _one(paths, barrier_spec(K, kind, barrier, barrier_type, rebate), r, T, control_variate)


def asian_price(paths, K, r, T, kind, average_type="arithmetic", strike_type="fixed", *,
                control_variate=False):
    """Asian option, average over t_1..t_N (t=0 excluded) (exotics.py:102-163)."""
    return _one(paths, asian_spec(K, kind, average_type, strike_type), r, T, control_variate)


def digital_price(paths, K, r, T, kind, payout=1.0, *, control_variate=False):
    """Cash-or-nothing digital, strict inequality (exotics.py:169-209)."""
    return _one(paths, digital_spec(K, kind, payout), r, T, control_variate)


def lookback_price(paths, r, T, kind, K=0.0, strike_type="floating", *, control_variate=False):
    """Lookback option, extrema incl. t=0 (exotics.py:215-273)."""
    return _one(paths, lookback_spec(kind, K, strike_type), r, T, control_variate)
