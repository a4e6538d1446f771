"""Oracle: tabulated sigma(S,t) callable (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference has no tabulated surface: sigma_loc is a Python closure evaluated on
every call (calibration.py:386-390).  The contract it must satisfy is
`sigma_loc(S_array, t_scalar) -> array of the same shape` (processes.py:226,233-234,
361,368-369).  The B200 path tabulates sigma at the exact step times t_i = i*T/n_steps
(no interpolation in t) on a log-uniform S grid and interpolates
linearly in ln S inside the kernel.  This module is the numpy statement of that same
lookup, so that `oracle.path_schemes.lv_milstein(z, SigmaGrid(...))` -- i.e. the
reference's Milstein recursion driven by this callable -- is the oracle for the kernel.

Lookup (identical in csrc/mc_path.cuh, check mode):
    pos = (ln S - x0) * inv_h ;  pos = min(max(pos, 0), n-1)
    j   = min(floor(pos), n-2) ;  w = pos - j
    sigma = tab[i, j] + w * (tab[i, j+1] - tab[i, j])
"""
from __future__ import annotations

import numpy as np


class SigmaGrid:
    def __init__(self, table, x0, h, dt):
        self.table = np.ascontiguousarray(table, dtype=np.float64)  # (n_steps, n_nodes)
        self.x0 = float(x0)
        self.h = float(h)
        self.inv_h = 1.0 / float(h)
        self.dt = float(dt)

    @property
    def nodes(self):
        return np.exp(self.x0 + self.h * np.arange(self.table.shape[1]))

    def row_of(self, t):
        i = int(round(t / self.dt))
        if abs(t - i * self.dt) > 1e-9 * max(1.0, abs(t)) or not (0 <= i < self.table.shape[0]):
            raise ValueError(f"t={t!r} is not one of the tabulated step times")
        return i

    def __call__(self, S, t):
        row = self.table[self.row_of(t)]
        n = row.shape[0]
        pos = (np.log(np.asarray(S, dtype=float)) - self.x0) * self.inv_h
        pos = np.minimum(np.maximum(pos, 0.0), float(n - 1))
        j = np.minimum(np.floor(pos), float(n - 2))
        w = pos - j
        j = j.astype(np.int64)
        lo = row[j]
        return lo + w * (row[j + 1] - lo)


def tabulate(sigma_loc, *, S0, T, n_steps, n_nodes=256, half_width=1.6):
    """Evaluate a sigma(S,t) callable at step times i*dt (i < n_steps) x log-uniform S nodes."""
    dt = T / n_steps
    x0 = float(np.log(S0) - half_width)
    h = 2.0 * half_width / (n_nodes - 1)
    nodes = np.exp(x0 + h * np.arange(n_nodes))
    tab = np.empty((n_steps, n_nodes))
    for i in range(n_steps):
        tab[i] = np.asarray(sigma_loc(nodes, i * dt), dtype=float)
    return SigmaGrid(tab, x0, h, dt)
