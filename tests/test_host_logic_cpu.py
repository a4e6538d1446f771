"""Host-side logic of the product that needs no GPU: the estimators (monte_carlo.py:142-167 and its
path-payoff analogue), the chunk planning of check mode and the statistics aggregation -- fed with the
oracle's statistics of the reference's own draws and compared with the prices the unmodified reference
produced (tests/golden/euro_mc.json, mc_paths.json)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))

from golden_io import load_json, same_numpy_as_golden  # noqa: E402
from oracle import terminal_mc, path_schemes, payoffs  # noqa: E402
from optpricer_b200 import monte_carlo as mc, exotics as ex  # noqa: E402

TOL = 1e-12 if same_numpy_as_golden() else 1e-9


def close(a, b, tol=TOL):
    return abs(a - b) <= tol * max(1.0, abs(b))


def test_terminal_estimator_reproduces_reference_prices():
    for c in load_json("euro_mc.json")["cases"]:
        o, kw = c["opt"], dict(c["kwargs"])
        if kw.get("dtype") == "float32" or kw.get("n_workers", 1) != 1:
            continue
        n_paths, chunk = c["n_paths"], kw.get("chunk_size", 100_000)
        anti, cv = kw.get("antithetic", True), kw.get("control_variate", True)
        sizes = terminal_mc.plan_chunks(n_paths, chunk)
        seeds = terminal_mc.chunk_seeds(c["seed"], len(sizes))
        stats = [terminal_mc.chunk_sumstats(m, seed=s, S0=o["S0"], K=o["K"], T=o["T"], r=o["r"], q=o["q"],
                                            sigma=o["sigma"], kind=c["kind"], antithetic=anti)
                 for m, s in zip(sizes, seeds)]
        total = mc._aggregate_stats(stats)                      # product code
        price, se = mc.estimator(total, S0=o["S0"], q=o["q"], T=o["T"], control_variate=cv)
        assert close(price, c["price"]) and close(se, c["stderr"]), c
    assert all(np.isnan(v) for v in mc.estimator((0, 0, 0, 0, 0, 0), S0=100.0, q=0.0, T=1.0))
    assert isinstance(mc.estimator((4, 8.0, 20.0, 400.0, 40_000.0, 800.0), S0=100.0, q=0.0, T=1.0,
                                   return_stderr=False), float)


def test_path_payoff_estimator_reproduces_reference_prices():
    """exotics.estimate on the eight-slot statistics block == _price_from_payoff (exotics.py:18-27)."""
    meta = load_json("mc_paths.json")["gbm_small_anti1"]
    a = dict(meta["args"])
    (z,) = path_schemes.draw_normals(a.pop("seed"), a["n_steps"], a["n_paths"])
    anti = a.pop("antithetic", True)
    a.pop("n_steps"); a.pop("n_paths")
    paths = path_schemes.gbm_exact(z, antithetic=anti, **a)
    K, r, T = meta["K"], meta["r"], meta["T"]
    df = np.exp(-r * T)
    checked = 0
    for key, (ref_price, ref_se) in meta["prices"].items():
        f = key.split("|")
        if f[0] != "asian":
            continue
        x = payoffs.asian_payoff(paths, K, f[1], f[2], f[3])
        X, Y = df * x, df * paths[-1]
        block = np.array([X.size, X.sum(), (X * X).sum(), Y.sum(), (Y * Y).sum(), (X * Y).sum(), 0.0, 0.0])
        price, se = ex.estimate(block)                          # product code
        assert close(price, ref_price) and close(se, ref_se, 1e-9), key
        checked += 1
    assert checked >= 4
