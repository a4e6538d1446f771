"""optpricer_b200 -- B200-native drop-in for the two data-parallel hot paths of optpricer 0.3.0:
Monte Carlo path simulation + payoff reduction and batched CRR backward induction.

Same names and signatures as the reference (src/optpricer/__init__.py:5-44 for the re-exports;
the generators are imported from `.processes` exactly as reference callers do).  Everything
executes in hand-written sm_100a CUDA behind the C ABI of include/optp_b200.h; there is no CPU
fallback.
"""
from .core import OptionSpec, CALL, PUT
from .monte_carlo import euro_price_mc
from .binomial import crr, crr_vec, crr_batch
from .exotics import barrier_price, asian_price, digital_price, lookback_price, price_payoffs
from .processes import (gbm_paths, heston_paths, milstein_local_vol_paths, local_vol_paths,
                        gbm_milstein_paths, sabr_paths, merton_jump_paths, SigmaTable, LazyPaths)
from .risk import TreePricer, McPricer, numerical_greeks, scenario_grid, portfolio_risk
from .validation import stress_test, convergence_analysis
from ._native import NativeUnavailable

__version__ = "0.1.0"

__all__ = [
    "OptionSpec", "CALL", "PUT", "euro_price_mc", "crr", "crr_vec", "crr_batch",
    "barrier_price", "asian_price", "digital_price", "lookback_price", "price_payoffs",
    "gbm_paths", "heston_paths", "milstein_local_vol_paths", "local_vol_paths",
    "gbm_milstein_paths", "sabr_paths", "merton_jump_paths", "SigmaTable", "LazyPaths", "NativeUnavailable",
    "TreePricer", "McPricer", "numerical_greeks", "scenario_grid", "portfolio_risk", "stress_test",
    "convergence_analysis",
]
