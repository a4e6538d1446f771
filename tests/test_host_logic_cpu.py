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


# ---- round 2 host logic (no device needed) ---------------------------------------------------------
def test_validate_contracts_raises_the_reference_text_for_the_first_offender():
    import numpy as np
    import pytest
    from optpricer_b200.core import validate_contracts
    validate_contracts([100.0, 90.0], [100.0, 95.0], 1.0, 0.2)
    with pytest.raises(ValueError, match="K must be positive, got 0.0"):
        validate_contracts([100.0, 90.0, -1.0], [100.0, 0.0, 95.0], 1.0, 0.2)       # contract 1 fails before contract 2
    with pytest.raises(ValueError, match="S0 must be positive, got -1.0"):
        validate_contracts([100.0, -1.0], [100.0, 0.0], 1.0, 0.2)                  # S0 is checked before K
    with pytest.raises(ValueError, match="sigma must be positive"):
        validate_contracts(100.0, 100.0, 1.0, np.array([0.2, 0.0]))
    with pytest.raises(ValueError):
        validate_contracts(np.nan, 100.0, 1.0, 0.2)


def test_sigma_table_belongs_to_its_run():
    import math
    import numpy as np
    import pytest
    from optpricer_b200.processes import SigmaTable
    tab = SigmaTable.from_callable(lambda S, t: 0.2 + 0.0 * S, S0=100.0, T=1.0, n_steps=12, n_nodes=64, half_width=1.6)
    tab.check_against(100.0, 1.0, 12)
    tab.check_against(150.0, 0.5, 6)                      # same dt, S0 still inside the grid
    with pytest.raises(ValueError, match="fewer rows"):
        tab.check_against(100.0, 2.0, 24)
    with pytest.raises(ValueError, match="tabulated for dt"):
        tab.check_against(100.0, 2.0, 12)
    with pytest.raises(ValueError, match="outside the table"):
        tab.check_against(100.0 * math.exp(1.7), 1.0, 12)
    assert SigmaTable.auto_grid(lambda S, t: 0.25 + 0.0 * S, 100.0, 1.0, 50) == (256, 1.6)
    n_nodes, half = SigmaTable.auto_grid(lambda S, t: 0.5 + 0.0 * S, 100.0, 4.0, 50)
    assert half == 6.0 and n_nodes == math.ceil(255 * 6.0 / 1.6) + 1
    spiky = lambda S, t: np.where(np.isclose(t, 0.5), 5.0, 0.2) + 0.0 * S   # noqa: E731  (one bad slice: the median ignores it)
    assert SigmaTable.auto_grid(spiky, 100.0, 1.0, 48) == (256, 1.6)


def test_control_variate_mean_per_scheme():
    import math
    import pytest
    from optpricer_b200 import exotics as ex, processes as pr
    gbm = pr.gbm_paths(100.0, 0.08, 0.02, 0.2, 2.0, 10, 5, seed=1)
    assert ex._expected_y(gbm, 0.08, 2.0) == pytest.approx(100.0 * math.exp(-0.02 * 2.0))          # monte_carlo.py:157
    mil = pr.gbm_milstein_paths(100.0, 0.08, 0.02, 0.2, 2.0, 10, 5, seed=1)
    assert ex._expected_y(mil, 0.08, 2.0) == pytest.approx(100.0 * (1.0 + 0.06 * 0.2) ** 10 * math.exp(-0.16))
    sabr1 = pr.sabr_paths(100.0, 0.08, 0.02, 0.2, 1.0, 0.4, -0.3, 2.0, 10, 5, seed=1)
    sabr5 = pr.sabr_paths(100.0, 0.08, 0.02, 2.0, 0.5, 0.4, -0.3, 2.0, 10, 5, seed=1)
    assert ex._expected_y(sabr1, 0.08, 2.0) == pytest.approx(100.0 * math.exp(-0.04))
    assert ex._expected_y(sabr5, 0.08, 2.0) == pytest.approx(100.0 * 1.012 ** 10 * math.exp(-0.16))
    qe = pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 8, 5, seed=1, scheme="qe")
    with pytest.raises(ValueError, match="martingale correction"):
        ex._expected_y(qe, 0.05, 1.0)
    qem = pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 8, 5, seed=1, scheme="qe", martingale_correction=True)
    assert ex._expected_y(qem, 0.05, 1.0) == pytest.approx(100.0)
    with pytest.raises(ValueError, match="belongs to scheme='qe'"):
        pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 8, 5, martingale_correction=True)
