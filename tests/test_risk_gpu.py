"""GPU parity of the batched bump-and-reprice callers (optpricer_b200.risk / .validation) against
tests/golden/risk.json (unmodified reference: risk.py, validation.py) and the oracle restatement.
Tolerances: prices 1e-10 relative (CRR) / 1e-12 (check-mode Monte Carlo); Greeks are differences of such
prices divided by the bump, so 1e-7 absolute + 1e-7 relative."""
import numpy as np
import pytest

from golden_io import load_json
from oracle import risk_callers as rc

pytestmark = pytest.mark.gpu

G = load_json("risk.json")


@pytest.fixture(scope="module")
def op():
    import optpricer_b200 as o
    from optpricer_b200 import _native
    _native.init()
    return o


def _args(c):
    return (c["S"], c["K"], c["T"], c["r"], c["q"], c["sigma"], c["kind"])


def _pricer(op, c, **kw):
    if c["pricer"] == "tree":
        return op.TreePricer(c["N"], american=c["american"])
    return op.McPricer(c["n_paths"], seed=c["seed"], **kw)


def _close(a, b):
    return abs(a - b) <= 1e-7 + 1e-7 * abs(b)


def test_numerical_greeks_tree_and_check_mode_mc_match_reference(op):
    from optpricer_b200 import _native as nv
    for c in G["greeks"]:
        before = nv.launch_count()
        got = op.numerical_greeks(_pricer(op, c, mode="check"), *_args(c), bump_pct=c["bump_pct"])
        if c["pricer"] == "tree":
            # ONE batch: banded pass + its rigorous-schedule retry pass + European pass of the never-exercised calls +
            # full-width fallback pass
            assert nv.launch_count() - before <= 4
        assert set(got) == set(c["result"])
        for k, v in c["result"].items():
            assert _close(got[k], v), (c, k, got[k], v)


def test_numerical_greeks_plain_callable_is_evaluated_point_by_point(op):
    c = G["greeks"][0]
    f = lambda S, K, T, r, q, sigma, kind: op.crr(op.OptionSpec(S, K, T, r, sigma, q), kind, c["N"],  # noqa: E731
                                                  american=c["american"])
    got = op.numerical_greeks(f, *_args(c), bump_pct=c["bump_pct"])
    batched = op.numerical_greeks(op.TreePricer(c["N"], american=c["american"]), *_args(c), bump_pct=c["bump_pct"])
    assert got == batched       # a contract's price does not depend on the batch it is priced in


def test_mc_pricer_batch_uses_common_random_numbers(op):
    """Philox mode: bumped contracts share streams, so the finite-difference Greeks are smooth; compare
    with the closed-form Black-Scholes Greeks of the same contract."""
    from math import exp, log, sqrt
    from statistics import NormalDist
    nd = NormalDist()
    S, K, T, r, q, sigma = 100.0, 105.0, 1.0, 0.04, 0.01, 0.25
    pr = op.McPricer(2_000_000, seed=11)
    g = op.numerical_greeks(pr, S, K, T, r, q, sigma, "call")
    d1 = (log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * sqrt(T))
    delta = exp(-q * T) * nd.cdf(d1)
    gamma = exp(-q * T) * nd.pdf(d1) / (S * sigma * sqrt(T))
    vega = S * exp(-q * T) * nd.pdf(d1) * sqrt(T)
    assert abs(g["delta"] - delta) < 2e-3 and abs(g["gamma"] - gamma) < 2e-4 and abs(g["vega"] - vega) < 0.2
    one = pr(S, K, T, r, q, sigma, "call")
    many = pr.batch([S, S], [K, K], [T, T], [r, r], [q, q], [sigma, sigma], ["call", "call"])
    assert abs(many[0] - one) <= 1e-9 * one and many[0] == many[1]


def test_scenario_grid_and_portfolio_risk_match_reference(op):
    for c in G["scenario"]:
        got = op.scenario_grid(_pricer(op, c), *_args(c), c["spot_range"], c["vol_range"])
        np.testing.assert_allclose(got["prices"], c["prices"], rtol=1e-10, atol=1e-10)
        np.testing.assert_array_equal(got["spot_values"], c["spot_range"])
        np.testing.assert_array_equal(got["vol_values"], c["vol_range"])
    for c in G["portfolio"]:
        got = op.portfolio_risk(c["instruments"], _pricer(op, c), bump_pct=c["bump_pct"])
        for k, v in c["result"].items():
            if k == "instrument_greeks":
                for a, b in zip(got[k], v):
                    assert all(abs(a[kk] - b[kk]) <= 1e-6 + 1e-7 * abs(b[kk]) for kk in b), (a, b)
            else:
                assert abs(got[k] - v) <= 1e-6 + 1e-7 * abs(v), (k, got[k], v)


def test_stress_test_and_convergence_match_reference(op):
    for c in G["stress"]:
        opt = op.OptionSpec(c["S0"], c["K"], c["T"], c["r"], c["sigma"], c["q"])
        got = op.stress_test(opt, c["kind"], c["spot_shocks"], c["vol_shocks"], c["rate_shocks"], pricer="tree")
        assert got.shape == (len(c["spot_shocks"]), len(c["vol_shocks"]), len(c["rate_shocks"]))
        np.testing.assert_allclose(got, c["prices"], rtol=1e-10, atol=1e-10)
        bs = op.stress_test(opt, c["kind"], c["spot_shocks"], c["vol_shocks"], c["rate_shocks"])
        assert np.max(np.abs(bs - got)) < 0.05          # N=500 trees sit within a few cents of Black-Scholes
    with pytest.raises(ValueError, match="Unknown pricer"):
        op.stress_test(opt, "call", [1.0], [0.0], [0.0], pricer="nope")
    for c in G["convergence"]:
        opt = op.OptionSpec(c["S0"], c["K"], c["T"], c["r"], c["sigma"], c["q"])
        got = op.convergence_analysis(opt, c["kind"], c["method"], c["param_name"], c["params"], mode="check")
        np.testing.assert_allclose(got["prices"], c["prices"], rtol=1e-10)
        np.testing.assert_allclose(got["errors"], c["errors"], rtol=1e-6, atol=1e-11)
        assert abs(got["order"] - c["order"]) <= 1e-6
    with pytest.raises(ValueError, match="Unknown method"):
        op.convergence_analysis(opt, "call", "nope", "N", [10])


def test_oracle_and_product_agree_on_a_fresh_contract(op):
    args = (120.0, 100.0, 0.6, 0.05, 0.0, 0.35, "put")
    ref = rc.numerical_greeks(rc.tree_pricer(300, True), *args, bump_pct=0.02)
    got = op.numerical_greeks(op.TreePricer(300, american=True), *args, bump_pct=0.02)
    for k in ref:
        assert _close(got[k], ref[k]), (k, got[k], ref[k])
