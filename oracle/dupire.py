"""Oracle: Dupire sigma(S,t) callable from raw-SVI slices (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the evaluation half of reference src/optpricer/calibration.py so the GPU
box (which has no /root/reference) can still run the *callable* CPU baseline and
re-tabulate the cfg-3 surface:
  svi_w / svi_dw / svi_d2w <- SVIParams.total_var/dw_dk/d2w_dk2   calibration.py:37-66
  Surface.forward          <- VolSurface._get_forward             calibration.py:108-122
  Surface.iv_from_logm     <- VolSurface.iv_from_logm             calibration.py:125-148
  Surface.local_vol        <- dupire_local_vol                    calibration.py:268-359
The SVI *fit* (scipy least_squares, calibration.py:166-229) is not restated: fitted
parameters are stored in tests/golden/dupire_surface.json by make_golden.py.
"""
from __future__ import annotations

import numpy as np


def svi_w(p, k):
    a, b, rho, m, s = p
    km = k - m
    return a + b * (rho * km + np.sqrt(km * km + s * s))


def svi_dw(p, k):
    a, b, rho, m, s = p
    x = k - m
    return b * (rho + x / np.sqrt(x * x + s ** 2))


def svi_d2w(p, k):
    a, b, rho, m, s = p
    x = k - m
    return b * s ** 2 / (x * x + s ** 2) ** 1.5


class Surface:
    """slices: {expiry: (a, b, rho, m, sigma)}; forwards: {expiry: F}."""

    def __init__(self, slices, forwards):
        self.exp = np.array(sorted(slices), dtype=float)
        self.par = {float(T): tuple(float(v) for v in slices[T]) for T in slices}
        self.fwd = {float(T): float(F) for T, F in forwards.items()}

    def _iv_slice(self, T, k):
        return np.sqrt(np.maximum(svi_w(self.par[T], k), 0.0) / T)

    def forward(self, T):
        if T in self.fwd:
            return self.fwd[T]
        ts = np.array(sorted(self.fwd), dtype=float)
        fs = np.array([self.fwd[t] for t in sorted(self.fwd)], dtype=float)
        if len(ts) == 1:
            return float(fs[0])
        return float(np.interp(T, ts, fs))

    def iv_from_logm(self, k, T):
        k = np.asarray(k, dtype=float)
        if T in self.par:
            return self._iv_slice(T, k)
        i = np.searchsorted(self.exp, T)
        if i == 0:
            return self._iv_slice(float(self.exp[0]), k)
        if i >= len(self.exp):
            return self._iv_slice(float(self.exp[-1]), k)
        lo, hi = float(self.exp[i - 1]), float(self.exp[i])
        w_lo = svi_w(self.par[lo], k) * lo
        w_hi = svi_w(self.par[hi], k) * hi
        al = (T - lo) / (hi - lo)
        wT = (1 - al) * w_lo + al * w_hi
        return np.sqrt(np.maximum(wT, 0.0) / T)

    def local_vol(self, S, t, dT=1e-4):
        S = np.asarray(S, dtype=float)
        t = max(t, 1e-8)
        k = np.log(S / self.forward(t))
        i = int(np.searchsorted(self.exp, t))
        i = max(0, min(i, len(self.exp) - 1))
        par = self.par[float(self.exp[i])]
        w = np.maximum(svi_w(par, k), 1e-12)
        dw = svi_dw(par, k)
        d2w = svi_d2w(par, k)
        t_up = t + dT
        t_dn = max(t - dT, 1e-8)
        w_up = self.iv_from_logm(k, t_up) ** 2 * t_up
        w_dn = self.iv_from_logm(k, t_dn) ** 2 * t_dn
        num = np.maximum((w_up - w_dn) / (t_up - t_dn), 1e-12)
        den = (1.0 - (k / w) * dw + 0.25 * (-0.25 - 1.0 / w + (k / w) ** 2) * dw ** 2 + 0.5 * d2w)
        den = np.maximum(den, 1e-8)
        return np.clip(np.sqrt(np.maximum(num / den, 0.0)), 0.01, 5.0)

    def sigma_loc(self):
        return lambda S, t: np.asarray(self.local_vol(S, t), dtype=float)
