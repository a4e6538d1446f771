"""Path generators on the B200 (drop-in for reference processes.py).

The reference generators return a materialised (n_steps+1, n_eff) float64 array that the
payoff functions of exotics.py then reduce.  At production sizes that array cannot exist
(252 x 1e8 x 8 B = 202 GB), so the generators here return a `LazyPaths`: process parameters +
seed + shape.  Handing it to barrier_price / asian_price / lookback_price / digital_price runs
ONE fused kernel (paths live in registers only); `np.asarray(paths)`, indexing or any numpy
use materialises the array through the fp64 kernel, so small-size code keeps working.

  gbm_paths                 <- processes.py:30-56
  heston_paths              <- processes.py:117-162   (full-truncation Euler; the reference has no QE)
  milstein_local_vol_paths  <- processes.py:311-381
  local_vol_paths           <- processes.py:223-258
  gbm_milstein_paths        <- processes.py:264-305
  sabr_paths                <- processes.py:168-217   (fp64 kernels only)
  merton_jump_paths         <- processes.py:62-111

New optional keywords (defaults keep every call a drop-in):
  rng  = "philox"  counter-based Philox4x32-10 on the device (draw = f(seed, path, step))
         "numpy"   deterministic check: default_rng(seed).standard_normal((n_steps, n_paths)) is
                   drawn on the host exactly as the reference does (processes.py:48,138-139,353)
                   and consumed by the fp64 kernels -> 1e-12 parity with the reference
  mode = "fp32" | "fp64"   arithmetic of the fused kernels with rng="philox"
  sigma_grid = (n_nodes, half_width)  tabulation of sigma_loc for the local-vol generators; the default
         None means (256, 1.6) around ln S0, widened (same node spacing) when 6 standard deviations of ln S_T
         at the surface's median volatility would not fit
The kernels count the effective paths that left the tabulated range (statistics slot 6); exotics.path_stats
warns when that fraction is not negligible -- outside the table sigma is the edge value, a departure from the
reference's sigma_loc(S_t, t) (processes.py:226,361).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nv
from .monte_carlo import _seed64

_POSITIVE = "n_steps and n_paths must be positive."


class SigmaTable:
    """sigma(S,t) tabulated at the exact step times t_i = i*T/n_steps (no interpolation in t) on a
    uniform ln S grid; linear in ln S inside, flat outside.  This is the device-side statement of
    the `sigma_loc(S_array, t) -> array` contract (processes.py:226,361-369)."""

    def __init__(self, table, x0, h, dt):
        self.table = np.ascontiguousarray(table, dtype=np.float64)
        if self.table.ndim != 2 or self.table.shape[1] < 2:
            raise ValueError("sigma table must be (n_steps, n_nodes>=2)")
        self.x0, self.h, self.dt = float(x0), float(h), float(dt)

    def check_against(self, S0, T, n_steps):
        """A ready-made table must belong to this run: rows at the step times i*T/n_steps, grid around S0."""
        if self.table.shape[0] < n_steps:
            raise ValueError("sigma table has fewer rows than n_steps")
        dt = T / n_steps
        if abs(self.dt - dt) > 1e-12 * dt:
            raise ValueError(f"sigma table was tabulated for dt={self.dt!r}, this run has T/n_steps={dt!r}: "
                             "its rows would be applied at the wrong times")
        x = float(np.log(S0))
        if not (self.x0 < x < self.x0 + self.h * (self.table.shape[1] - 1)):
            raise ValueError(f"S0={S0!r} lies outside the table's grid "
                             f"[{float(np.exp(self.x0))!r}, {float(np.exp(self.x0 + self.h * (self.table.shape[1] - 1)))!r}]")

    @staticmethod
    def auto_grid(sigma_loc, S0, T, n_steps):
        """(n_nodes, half_width): 256 nodes over ln S0 +- 1.6 (the desk workflow's paths span 19.7-307.6,
        SURVEY 8d cfg 3), widened at the same node spacing when 6 sigma sqrt(T) at the surface's median
        at-the-money volatility exceeds 1.6 -- so that long maturities / high volatilities do not run into the
        flat extrapolation."""
        dt = T / n_steps
        probe_t = [i * dt for i in range(0, n_steps, max(1, n_steps // 16))]
        at = [float(np.asarray(sigma_loc(np.array([float(S0)]), t), dtype=float)[0]) for t in probe_t]
        med = float(np.median(np.clip(at, 1e-8, 10.0)))
        half = max(1.6, 6.0 * med * float(np.sqrt(T)))
        n_nodes = int(np.ceil(255 * half / 1.6)) + 1
        # rows are staged in shared memory four at a time, twice (double buffering): 1024 nodes x 16 B is what fits;
        # beyond half_width 6.4 the node spacing grows instead
        return min(n_nodes, 1024), half

    @classmethod
    def from_callable(cls, sigma_loc, *, S0, T, n_steps, n_nodes=256, half_width=1.6):
        dt = T / n_steps
        x0 = float(np.log(S0) - half_width)
        h = 2.0 * half_width / (n_nodes - 1)
        nodes = np.exp(x0 + h * np.arange(n_nodes))
        table = np.empty((n_steps, n_nodes))
        for i in range(n_steps):
            table[i] = np.asarray(sigma_loc(nodes, i * dt), dtype=float)
        return cls(table, x0, h, dt)


    @classmethod
    def from_svi_surface(cls, surface, *, S0, T, n_steps, n_nodes=256, half_width=1.6, dT=1e-4,
                         forward_curve=None):
        """Tabulate the Dupire local volatility of a raw-SVI surface on the device (optp_dupire_table):
        the same values as `from_callable(dupire_local_vol_func(surface, r, q), ...)` without the
        n_steps host evaluations of the reference's callable (calibration.py:268-390).

        `surface` is the reference's VolSurface (or anything with `.slices` = {expiry: object with
        a, b, rho, m, sigma}); the forward curve is taken from it (`forward_curve=` overrides).  The
        reference falls back to F = mean(S nodes) when the surface has no curve (calibration.py:317-320);
        so does this."""
        slices = dict(sorted(dict(surface.slices).items()))
        if not slices:
            raise ValueError("At least one SVI slice is required.")
        fwd = forward_curve if forward_curve is not None else getattr(surface, "_forward_curve", None)
        dt = T / n_steps
        x0 = float(np.log(S0) - half_width)
        h = 2.0 * half_width / (n_nodes - 1)
        if not fwd:
            fwd = {1.0: float(np.exp(x0 + h * np.arange(n_nodes)).mean())}
        fwd = dict(sorted(fwd.items()))
        exp_arr = np.array(list(slices), dtype=np.float64)
        svi = np.array([[p.a, p.b, p.rho, p.m, p.sigma] for p in slices.values()], dtype=np.float64)
        fT = np.array(list(fwd), dtype=np.float64)
        fF = np.array(list(fwd.values()), dtype=np.float64)
        table = np.empty((int(n_steps), int(n_nodes)), dtype=np.float64)
        lib = nv.init()
        nv.check(lib.optp_dupire_table(nv.ptr(exp_arr), nv.ptr(svi), len(exp_arr), nv.ptr(fT), nv.ptr(fF), len(fT),
                                       float(T), int(n_steps), x0, h, int(n_nodes), float(dT), nv.ptr(table)),
                 "optp_dupire_table")
        return cls(table, x0, h, dt)


class LazyPaths:
    """A path set described by (process, n_steps, n_paths, antithetic, seed); see module docstring."""

    def __init__(self, proc, n_steps, n_paths, antithetic, seed, rng, mode, table=None):
        if rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        if mode not in ("fp32", "fp64"):
            raise ValueError("mode must be 'fp32' or 'fp64'")
        self.proc = dict(proc)
        self.n_steps, self.n_paths = int(n_steps), int(n_paths)
        self.antithetic = bool(antithetic)
        self.seed_arg = seed
        self.seed = _seed64(seed)
        self.rng, self.mode = rng, mode
        self.table = table
        self._array = None
        self._var = None

    # -- shape protocol -----------------------------------------------------------------
    @property
    def n_eff(self):
        return 2 * self.n_paths if self.antithetic else self.n_paths

    @property
    def shape(self):
        return (self.n_steps + 1, self.n_eff)

    ndim = 2
    dtype = np.dtype(np.float64)

    def __len__(self):
        return self.n_steps + 1

    def __repr__(self):
        return f"LazyPaths(kind={self.proc['kind']}, shape={self.shape}, rng={self.rng!r}, mode={self.mode!r})"

    # -- native descriptors ----------------------------------------------------------------
    def n_streams(self):
        if self.proc["kind"] == nv.PROC_MERTON:
            return 3
        return 2 if self.proc["kind"] in (nv.PROC_HESTON_FTE, nv.PROC_SABR, nv.PROC_HESTON_QE,
                                          nv.PROC_HESTON_QE_M) else 1

    def native_process(self):
        p = nv.OptpProcess()
        for k, v in self.proc.items():
            setattr(p, k, v)
        if self.table is not None:
            p.n_nodes = self.table.table.shape[1]
            p.x0, p.h = self.table.x0, self.table.h
            p.sigma_table = self.table.table.ctypes.data_as(C.c_void_p).value
        return p

    def host_normals(self):
        """The reference's draws: n_streams consecutive (n_steps, n_paths) blocks (Appendix B.3)."""
        g = np.random.default_rng(self.seed_arg)
        if self.proc["kind"] == nv.PROC_MERTON:   # Z, Poisson counts, ZJ (processes.py:89-93)
            shape = (self.n_steps, self.n_paths)
            z = g.standard_normal(shape)
            k = g.poisson(self.proc["v0"] * self.proc["T"] / self.n_steps, size=shape).astype(np.float64)
            zj = g.standard_normal(size=shape)
            return np.ascontiguousarray(np.stack([z, k, zj], axis=0))
        blocks = [g.standard_normal((self.n_steps, self.n_paths)) for _ in range(self.n_streams())]
        return np.ascontiguousarray(np.stack(blocks, axis=0))

    def native_mode(self):
        if self.rng == "numpy":
            return nv.MODE_FP64_HOSTZ
        return nv.MODE_FP32_PHILOX if self.mode == "fp32" else nv.MODE_FP64_PHILOX

    # -- materialisation ----------------------------------------------------------------
    def materialize(self, return_variance=False):
        if self._array is None or (return_variance and self._var is None):
            lib = nv.init()
            out = np.empty(self.shape, dtype=np.float64)
            var = np.empty(self.shape, dtype=np.float64) if return_variance else None
            mode = nv.MODE_FP64_HOSTZ if self.rng == "numpy" else nv.MODE_FP64_PHILOX
            z = self.host_normals() if self.rng == "numpy" else None
            p = self.native_process()
            nv.check(lib.optp_paths(C.byref(p), self.n_steps, self.n_paths, int(self.antithetic), mode,
                                    self.seed, 0, nv.ptr(z), nv.ptr(out), nv.ptr(var)), "optp_paths")
            self._array, self._var = out, var if return_variance else self._var
        return (self._array, self._var) if return_variance else self._array

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __iter__(self):
        return iter(self.materialize())


def _check_sizes(n_steps, n_paths):
    if n_steps <= 0 or n_paths <= 0:
        raise ValueError(_POSITIVE)


def gbm_paths(S0, r, q, sigma, T, n_steps, n_paths, *, antithetic=True, seed=None,
              rng="philox", mode="fp32"):
    """Exact log-Euler GBM (processes.py:30-56)."""
    _check_sizes(n_steps, n_paths)
    proc = dict(kind=nv.PROC_GBM, S0=S0, r=r, q=q, T=T, sigma=sigma)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode)


def gbm_milstein_paths(S0, r, q, sigma, T, n_steps, n_paths, *, antithetic=True, seed=None,
                       rng="philox", mode="fp32"):
    """Milstein GBM with the 1e-10 floor (processes.py:264-305)."""
    _check_sizes(n_steps, n_paths)
    proc = dict(kind=nv.PROC_GBM_MILSTEIN, S0=S0, r=r, q=q, T=T, sigma=sigma)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode)


def heston_paths(S0, r, q, v0, kappa, theta, xi, rho, T, n_steps, n_paths, *, antithetic=True,
                 seed=None, return_variance=False, rng="philox", mode="fp32", scheme="fte",
                 martingale_correction=False):
    """Heston paths (processes.py:117-162).  return_variance=True materialises.
    scheme="fte": the reference's full-truncation Euler (default, drop-in).
    scheme="qe":  Andersen's quadratic-exponential scheme -- NEW, the reference has no such code
                  (parity unpinned; checked against oracle/path_schemes.py:heston_qe).
                  martingale_correction=True uses Andersen's K0* (2008, sec. 4.3.2), which makes
                  E[S_T] = S0 e^{(r-q)T} exact for the discretised process -- what control_variate=True needs."""
    _check_sizes(n_steps, n_paths)
    if not (-1.0 <= rho <= 1.0):
        raise ValueError("rho must be in [-1, 1].")
    if scheme not in ("fte", "qe"):
        raise ValueError("scheme must be 'fte' or 'qe'")
    if scheme == "qe" and not (xi > 0 and kappa > 0 and theta > 0):
        raise ValueError("the QE scheme needs xi > 0, kappa > 0 and theta > 0")
    if martingale_correction and scheme != "qe":
        raise ValueError("martingale_correction belongs to scheme='qe'")
    kind = nv.PROC_HESTON_FTE if scheme == "fte" else (nv.PROC_HESTON_QE_M if martingale_correction else nv.PROC_HESTON_QE)
    proc = dict(kind=kind, S0=S0, r=r, q=q, T=T, v0=v0, kappa=kappa, theta=theta, xi=xi, rho=rho)
    lp = LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode)
    if return_variance:
        return lp.materialize(return_variance=True)
    return lp


def sabr_paths(S0, r, q, alpha0, beta, nu, rho, T, n_steps, n_paths, *, antithetic=True, seed=None,
               rng="philox", mode="fp32"):
    """SABR with lognormal vol (processes.py:168-217): exact lognormal volatility, log-Euler price step for
    beta = 1, Euler step with the reference's 1e-12 floor otherwise."""
    _check_sizes(n_steps, n_paths)
    if not (0.0 <= beta <= 1.0):
        raise ValueError("beta must be in [0, 1].")
    if alpha0 <= 0 or nu < 0:
        raise ValueError("alpha0 must be >0, nu >= 0.")
    if not (-1.0 <= rho <= 1.0):
        raise ValueError("rho must be in [-1, 1].")
    proc = dict(kind=nv.PROC_SABR, S0=S0, r=r, q=q, T=T, v0=alpha0, kappa=beta, xi=nu, rho=rho)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode)


def merton_jump_paths(S0, r, q, sigma, T, n_steps, n_paths, *, lam, mJ, sJ, antithetic=True, seed=None,
                      rng="philox", mode="fp32"):
    """Merton jump diffusion: exact GBM step + compound Poisson jump in log space (processes.py:62-111).
    The antithetic mirror negates Z and ZJ and re-uses the Poisson counts, as the reference does."""
    _check_sizes(n_steps, n_paths)
    if lam < 0 or sJ < 0:
        raise ValueError("lam and sJ must be non-negative.")
    proc = dict(kind=nv.PROC_MERTON, S0=S0, r=r, q=q, T=T, sigma=sigma, v0=lam, kappa=mJ, theta=sJ)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode)


def _table_for(sigma_loc, S0, T, n_steps, sigma_grid):
    if isinstance(sigma_loc, SigmaTable):
        sigma_loc.check_against(S0, T, n_steps)
        return sigma_loc
    n_nodes, half_width = sigma_grid if sigma_grid is not None else SigmaTable.auto_grid(sigma_loc, S0, T, n_steps)
    return SigmaTable.from_callable(sigma_loc, S0=S0, T=T, n_steps=n_steps, n_nodes=n_nodes,
                                    half_width=half_width)


def milstein_local_vol_paths(S0, r, q, T, n_steps, n_paths, sigma_loc, *, antithetic=True, seed=None,
                             dS_bump=0.01, rng="philox", mode="fp32", sigma_grid=None):
    """Local-vol Milstein (processes.py:311-381).  `sigma_loc` is the reference's callable (it is
    tabulated once on the host at the step times) or a ready SigmaTable."""
    _check_sizes(n_steps, n_paths)
    table = _table_for(sigma_loc, S0, T, n_steps, sigma_grid)
    proc = dict(kind=nv.PROC_LV_MILSTEIN, S0=S0, r=r, q=q, T=T, dS_bump=dS_bump)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode, table=table)


def local_vol_paths(S0, r, q, T, n_steps, n_paths, sigma_loc, *, antithetic=True, seed=None,
                    rng="philox", mode="fp32", sigma_grid=None):
    """Local-vol log-Euler (processes.py:223-258)."""
    _check_sizes(n_steps, n_paths)
    table = _table_for(sigma_loc, S0, T, n_steps, sigma_grid)
    proc = dict(kind=nv.PROC_LV_EULER, S0=S0, r=r, q=q, T=T)
    return LazyPaths(proc, n_steps, n_paths, antithetic, seed, rng, mode, table=table)
