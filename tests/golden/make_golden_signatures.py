#!/usr/bin/env python3
"""Record the signatures (parameter names, kinds, defaults) of the reference functions the B200 package
mirrors, from the UNMODIFIED reference:  PYTHONPATH=/root/reference/src python tests/golden/make_golden_signatures.py
-> tests/golden/signatures.json (checked by tests/test_layout_cpu.py::test_signatures_mirror_reference)."""
import inspect
import json
import os
import sys

sys.path.insert(0, os.environ.get("OPTPRICER_REF", "/root/reference/src"))
import numpy as np  # noqa: E402
from optpricer import binomial, exotics, monte_carlo, processes, risk, validation  # noqa: E402

FUNCS = {
    "monte_carlo": (monte_carlo, ["euro_price_mc"]),
    "processes": (processes, ["gbm_paths", "heston_paths", "milstein_local_vol_paths", "local_vol_paths",
                              "gbm_milstein_paths", "sabr_paths", "merton_jump_paths"]),
    "exotics": (exotics, ["barrier_price", "asian_price", "digital_price", "lookback_price"]),
    "binomial": (binomial, ["crr", "crr_vec"]),
    "risk": (risk, ["numerical_greeks", "scenario_grid", "portfolio_risk"]),
    "validation": (validation, ["stress_test", "convergence_analysis"]),
}


def enc(v):
    if v is inspect.Parameter.empty:
        return "<required>"
    if v is np.float64:
        return "np.float64"
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    if isinstance(v, (list, tuple)):
        return [enc(x) for x in v]
    return repr(v)


out = {}
for mod, (m, names) in FUNCS.items():
    for n in names:
        ps = inspect.signature(getattr(m, n)).parameters
        out[f"{mod}.{n}"] = [[p.name, p.kind.name, enc(p.default)] for p in ps.values()]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "signatures.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(out), "signatures")
