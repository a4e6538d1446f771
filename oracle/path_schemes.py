"""Oracle: path generators (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference src/optpricer/processes.py.  Every generator is split into
(1) the normal draws exactly as the reference takes them from PCG64 and
(2) a pure function of those draws, so that the CUDA deterministic-check mode
and this oracle consume the *same* Z arrays.

  draw_normals        <- processes.py:23-24,48 / 138-139 / 353 (draw order, C-order)
  mirror              <- processes.py:49-50, 140-142           (columns [0,n) base, [n,2n) = -Z)
  gbm_exact           <- gbm_paths                 processes.py:30-56
  heston_fte          <- heston_paths              processes.py:117-162
  heston_qe           <- (no reference code: Andersen 2008 QE, parity unpinned)
  lv_milstein         <- milstein_local_vol_paths  processes.py:311-381
  lv_log_euler        <- local_vol_paths           processes.py:223-258
  gbm_milstein        <- gbm_milstein_paths        processes.py:264-305
  sabr_lognormal_vol  <- sabr_paths                processes.py:168-217
  merton_jumps        <- merton_jump_paths         processes.py:62-111
All return (n_steps+1, n_cols) float64 arrays with the t=0 row.
"""
from __future__ import annotations

import numpy as np


def draw_normals(seed, n_steps, n_paths, n_arrays=1):
    """n_arrays consecutive (n_steps, n_paths) standard-normal blocks from default_rng(seed)."""
    g = np.random.default_rng(seed)
    return [g.standard_normal((n_steps, n_paths)) for _ in range(n_arrays)]


def mirror(z, antithetic):
    return np.concatenate([z, -z], axis=1) if antithetic else z


def _with_s0_row(S0, body):
    top = np.full((1, body.shape[1]), S0, dtype=body.dtype)
    return np.vstack([top, body])


def gbm_exact(z_base, *, S0, r, q, sigma, T, antithetic=True):
    n_steps = z_base.shape[0]
    dt = T / n_steps
    drift = (r - q - 0.5 * sigma * sigma) * dt
    vol = sigma * np.sqrt(dt)
    z = mirror(z_base, antithetic)
    logs = np.cumsum(drift + vol * z, axis=0)
    return _with_s0_row(S0, S0 * np.exp(logs))


def heston_fte(z2_base, zp_base, *, S0, r, q, v0, kappa, theta, xi, rho, T,
               antithetic=True, return_variance=False):
    n_steps = z2_base.shape[0]
    dt = T / n_steps
    sdt = np.sqrt(dt)
    z2 = mirror(z2_base, antithetic)
    zp = mirror(zp_base, antithetic)
    z1 = rho * z2 + np.sqrt(max(0.0, 1.0 - rho * rho)) * zp
    cols = z1.shape[1]
    S = np.empty((n_steps + 1, cols))
    V = np.empty_like(S)
    S[0] = S0
    v = np.full(cols, max(v0, 0.0))
    V[0] = v
    for i in range(n_steps):
        vp = np.maximum(v, 0.0)
        S[i + 1] = S[i] * np.exp((r - q - 0.5 * vp) * dt + np.sqrt(vp) * sdt * z1[i])
        v = v + kappa * (theta - vp) * dt + xi * np.sqrt(vp) * sdt * z2[i]
        v = np.maximum(v, 0.0)
        V[i + 1] = v
    return (S, V) if return_variance else S


def heston_qe(z2_base, zp_base, *, S0, r, q, v0, kappa, theta, xi, rho, T, antithetic=True,
              psi_c=1.5, return_variance=False, martingale_correction=False):
    """Heston quadratic-exponential scheme -- PARITY UNPINNED: the reference has no QE code
    (processes.py:117-162 is full-truncation Euler only; README.md:42 merely mentions QE).  This
    restates Andersen, "Simple and efficient simulation of the Heston stochastic volatility model"
    (J. Comp. Finance 11(3), 2008), sec. 3.2.4 / 4.2: moment-matched quadratic (psi <= psi_c) or
    exponential (psi > psi_c) variance update, central discretisation gamma1 = gamma2 = 1/2 of the
    log-price.  It consumes the same two normal streams as heston_fte (Z2 drives the variance -- its uniform
    Phi(Z2) feeds the exponential branch -- Zp the price).
    martingale_correction=True replaces K0 by Andersen's K0* (sec. 4.3.2, eqs. 4.14-4.16):
        K0* = -ln M - (K1 + K3/2) v,   M = E[exp(A v') | v],   A = K2 + K4/2   (K4 = K3 here),
        quadratic branch:   M = exp(A b^2 a / (1 - 2 A a)) / sqrt(1 - 2 A a)     (needs A < 1/(2a))
        exponential branch: M = p + beta (1 - p) / (beta - A)                    (needs A < beta)
    so that E[S' e^{-(r-q) dt} | S, v] = S exactly; where the regularity condition fails (it cannot for
    rho <= 0) the uncorrected K0 is kept for that path and step."""
    from scipy.special import erfc
    n_steps = z2_base.shape[0]
    dt = T / n_steps
    z2 = mirror(z2_base, antithetic)
    zp = mirror(zp_base, antithetic)
    E = np.exp(-kappa * dt)
    c1 = xi * xi * E * (1.0 - E) / kappa
    c2 = theta * xi * xi * (1.0 - E) ** 2 / (2.0 * kappa)
    K0 = -rho * kappa * theta * dt / xi
    K1 = 0.5 * dt * (kappa * rho / xi - 0.5) - rho / xi
    K2 = 0.5 * dt * (kappa * rho / xi - 0.5) + rho / xi
    K3 = 0.5 * dt * (1.0 - rho * rho)
    cols = z2.shape[1]
    X = np.empty((n_steps + 1, cols))
    V = np.empty_like(X)
    X[0] = np.log(S0)
    v = np.full(cols, max(v0, 0.0))
    V[0] = v
    rsqrt2 = 0.7071067811865476
    for i in range(n_steps):
        m = theta + (v - theta) * E
        s2 = v * c1 + c2
        psi = s2 / (m * m)
        t = 2.0 / psi
        b2 = np.maximum(t - 1.0 + np.sqrt(t) * np.sqrt(np.maximum(t - 1.0, 0.0)), 0.0)
        a = m / (1.0 + b2)
        vq = a * (np.sqrt(b2) + z2[i]) ** 2
        p = (psi - 1.0) / (psi + 1.0)
        beta = (1.0 - p) / m
        u = 0.5 * erfc(-z2[i] * rsqrt2)
        one_minus_u = 0.5 * erfc(z2[i] * rsqrt2)
        with np.errstate(divide="ignore", invalid="ignore"):
            ve = np.where(u <= p, 0.0, np.log((1.0 - p) / one_minus_u) / beta)
        vn = np.where(psi <= psi_c, vq, ve)
        k0 = K0 + K1 * v
        if martingale_correction:
            A = K2 + 0.5 * K3
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                den_q = 1.0 - 2.0 * A * a
                lnM_q = A * b2 * a / den_q - 0.5 * np.log(den_q)
                den_e = beta - A
                lnM_e = np.log(p + beta * (1.0 - p) / den_e)
            quad = psi <= psi_c
            ok = np.where(quad, den_q > 0.0, den_e > 0.0)
            lnM = np.where(quad, lnM_q, lnM_e)
            k0 = np.where(ok, -lnM - 0.5 * K3 * v, k0)
        X[i + 1] = X[i] + (r - q) * dt + k0 + K2 * vn + np.sqrt(np.maximum(K3 * (v + vn), 0.0)) * zp[i]
        v = vn
        V[i + 1] = v
    S = np.exp(X)
    S[0] = S0
    return (S, V) if return_variance else S


def lv_milstein(z_base, sigma_loc, *, S0, r, q, T, antithetic=True, dS_bump=0.01):
    n_steps = z_base.shape[0]
    dt = T / n_steps
    sdt = np.sqrt(dt)
    z = mirror(z_base, antithetic)
    S = np.empty((n_steps + 1, z.shape[1]))
    S[0] = S0
    for i in range(n_steps):
        t = i * dt
        zi, s = z[i], S[i]
        sig = np.clip(np.asarray(sigma_loc(s, t), dtype=float), 1e-8, 10.0)
        eps = dS_bump * s
        s_up = s + eps
        s_dn = np.maximum(s - eps, 1e-10)
        sig_up = np.asarray(sigma_loc(s_up, t), dtype=float)
        sig_dn = np.asarray(sigma_loc(s_dn, t), dtype=float)
        da = (sig_up * s_up - sig_dn * s_dn) / (s_up - s_dn)
        a = sig * s
        S[i + 1] = s + (r - q) * s * dt + a * sdt * zi + 0.5 * a * da * (zi ** 2 - 1.0) * dt
        S[i + 1] = np.maximum(S[i + 1], 1e-10)
    return S


def lv_log_euler(z_base, sigma_loc, *, S0, r, q, T, antithetic=True):
    n_steps = z_base.shape[0]
    dt = T / n_steps
    sdt = np.sqrt(dt)
    z = mirror(z_base, antithetic)
    S = np.empty((n_steps + 1, z.shape[1]))
    S[0] = S0
    for i in range(n_steps):
        sig = np.clip(np.asarray(sigma_loc(S[i], i * dt), dtype=float), 0.0, np.inf)
        S[i + 1] = S[i] * np.exp((r - q - 0.5 * sig * sig) * dt + sig * sdt * z[i])
    return S


def gbm_milstein(z_base, *, S0, r, q, sigma, T, antithetic=True):
    n_steps = z_base.shape[0]
    dt = T / n_steps
    sdt = np.sqrt(dt)
    z = mirror(z_base, antithetic)
    S = np.empty((n_steps + 1, z.shape[1]))
    S[0] = S0
    for i in range(n_steps):
        zi, s = z[i], S[i]
        S[i + 1] = (s + (r - q) * s * dt + sigma * s * sdt * zi
                    + 0.5 * sigma ** 2 * s * (zi ** 2 - 1.0) * dt)
        S[i + 1] = np.maximum(S[i + 1], 1e-10)
    return S


def sabr_lognormal_vol(z2_base, zp_base, *, S0, r, q, alpha0, beta, nu, rho, T, antithetic=True):
    n_steps = z2_base.shape[0]
    dt = T / n_steps
    sdt = np.sqrt(dt)
    z2 = mirror(z2_base, antithetic)
    zp = mirror(zp_base, antithetic)
    z1 = rho * z2 + np.sqrt(max(0.0, 1.0 - rho * rho)) * zp
    cols = z1.shape[1]
    S = np.empty((n_steps + 1, cols))
    S[0] = S0
    vol = np.full(cols, alpha0, dtype=float)
    for i in range(n_steps):
        vol *= np.exp(nu * sdt * z2[i] - 0.5 * (nu * nu) * dt)
        if beta == 1.0:
            S[i + 1] = S[i] * np.exp((r - q - 0.5 * vol ** 2) * dt + vol * sdt * z1[i])
        else:
            S[i + 1] = S[i] + (r - q) * S[i] * dt + vol * (S[i] ** beta) * sdt * z1[i]
            S[i + 1] = np.maximum(S[i + 1], 1e-12)
    return S


def merton_draws(seed, n_steps, n_paths, lam, T):
    """Z, Poisson counts, jump normals in the reference's draw order (processes.py:89-93)."""
    g = np.random.default_rng(seed)
    dt = T / n_steps
    z = g.standard_normal((n_steps, n_paths))
    k = g.poisson(lam * dt, size=(n_steps, n_paths))
    zj = g.standard_normal(size=(n_steps, n_paths))
    return z, k, zj


def merton_jumps(z_base, k_base, zj_base, *, S0, r, q, sigma, T, lam, mJ, sJ, antithetic=True):
    n_steps = z_base.shape[0]
    dt = T / n_steps
    kappa = np.exp(mJ + 0.5 * sJ * sJ) - 1.0
    drift = (r - q - 0.5 * sigma * sigma - lam * kappa) * dt
    vol = sigma * np.sqrt(dt)
    if antithetic:
        z = np.concatenate([z_base, -z_base], axis=1)
        k = np.concatenate([k_base, k_base], axis=1)
        zj = np.concatenate([zj_base, -zj_base], axis=1)
    else:
        z, k, zj = z_base, k_base, zj_base
    jumps = mJ * k + sJ * np.sqrt(k) * zj
    logs = np.cumsum(drift + vol * z + jumps, axis=0)
    return _with_s0_row(S0, S0 * np.exp(logs))
