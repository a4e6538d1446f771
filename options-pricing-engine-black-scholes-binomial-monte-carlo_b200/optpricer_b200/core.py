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


CALL = "call"
PUT = "put"
