"""Boundary data model of the hot path: OptionSpec, CALL, PUT (reference core.py:12-34,112-113)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class OptionSpec:
    """Single-option container (same fields, order, defaults and validation messages as the
    reference's OptionSpec, core.py:13-34)."""
    S0: float
    K: float
    T: float
    r: float
    sigma: float
    q: float = 0.0

    def __post_init__(self):
        if self.S0 <= 0:
            raise ValueError(f"S0 must be positive, got {self.S0}")
        if self.K <= 0:
            raise ValueError(f"K must be positive, got {self.K}")
        if self.T <= 0:
            raise ValueError(f"T must be positive, got {self.T}")
        if self.sigma <= 0:
            raise ValueError(f"sigma must be positive, got {self.sigma}")


def validate_contracts(S0, K, T, sigma):
    """OptionSpec's checks (core.py:27-34 of the reference) for whole columns at once: raises the
    reference's message for the FIRST offending contract, checking S0, K, T, sigma in that order -- what a
    loop constructing OptionSpec(...) per contract would raise."""
    import numpy as np
    cols = [np.asarray(x, dtype=float) for x in (S0, K, T, sigma)]
    cols = np.broadcast_arrays(*cols)
    bad = np.zeros(cols[0].shape, dtype=bool)
    for c in cols:
        bad |= ~(c > 0)          # NaN fails like the reference's `<= 0` would not; a NaN contract is rejected here
    if bad.any():
        i = int(np.flatnonzero(bad.reshape(-1))[0])
        s0, k, t, sg = (float(c.reshape(-1)[i]) for c in cols)
        OptionSpec(S0=s0, K=k, T=t, r=0.0, sigma=sg)      # raises the reference's text
        raise ValueError(f"contract {i} is not a valid option (S0={s0}, K={k}, T={t}, sigma={sg})")


CALL = "call"
PUT = "put"
