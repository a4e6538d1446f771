"""Pins the numpy oracle (oracle/) to fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact where the numpy build matches the one that
generated the fixtures, 1e-13 relative otherwise."""
import numpy as np
import pytest

from oracle import terminal_mc, path_schemes, payoffs, crr_tree, sigma_grid, dupire
from golden_io import load_json, load_npz, dupire_grid_arrays, same_numpy_as_golden

EXACT = same_numpy_as_golden()


def close(a, b):
    if EXACT:
        return a == b or (a != a and b != b)
    return abs(a - b) <= 1e-13 * max(1.0, abs(b))


def test_euro_mc_cases():
    g = load_json("euro_mc.json")
    for c in g["cases"]:
        kw = dict(c["kwargs"])
        if kw.get("dtype") == "float32":
            kw["dtype"] = np.float32
        price, se = terminal_mc.euro_mc(**c["opt"], kind=c["kind"], n_paths=c["n_paths"], seed=c["seed"], **kw)
        assert close(price, c["price"]) and close(se, c["stderr"]), c
    (p, se), chunks = terminal_mc.euro_mc(100.0, 110.0, 1.0, 0.03, 0.2, 0.0, "call", n_paths=200_000, seed=1,
                                          return_stats=True)
    for mine, ref in zip(chunks, g["cfg1_chunk_stats"]):
        assert all(close(float(a), float(b)) for a, b in zip(mine, ref))
    z = terminal_mc.chunk_normals(terminal_mc.chunk_seeds(1, 1)[0], 4)
    assert z.tolist() == g["cfg1_first_normals"] or not EXACT


def test_euro_mc_empty_is_nan():
    p, se = terminal_mc.euro_mc(100.0, 110.0, 1.0, 0.03, 0.2, 0.0, "call", n_paths=0, seed=1)
    assert p != p and se != se


def test_crr_cases():
    g = load_json("crr.json")
    for c in g["cases"]:
        args = (c["S0"], c["K"], c["T"], c["r"], c["sigma"], c["q"], c["kind"], c["N"], c["american"])
        if "error" in c:
            with pytest.raises(ValueError) as e:
                crr_tree.crr_scalar(*args)
            assert str(e.value) == c["error"]
        elif c["N"] <= 1024:  # keep the CPU suite short; larger N are exercised on the GPU box
            assert close(crr_tree.crr_scalar(*args), c["price"]), c
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(1_000_000)
    for i in range(8):
        v = crr_tree.crr_scalar(S0[i], K[i], T[i], 0.03, sig[i], 0.0, "put", 1000, True)
        assert close(v, g["cfg2_first64_american_put_N1000"][i])
    for c in g["crr_vec"]:
        v = crr_tree.crr_shared_tree(c["S0"], np.array(c["K"]), c["T"], c["r"], c["q"], c["sigma"], c["kind"],
                                     c["N"], c["american"])
        assert v.shape == (len(c["K"]),)
        assert all(close(float(a), b) for a, b in zip(v, c["prices"]))


def test_crr_c_restatement_against_reference_fixtures():
    """oracle/crr_tree.c: the reference-order loop to 1e-13, the table-based loop to 1e-11."""
    from oracle import crr_c
    g = load_json("crr.json")
    ok = [c for c in g["cases"] if "error" not in c and c["N"] <= 1024]
    bad = [c for c in g["cases"] if "error" in c and c["N"] > 0]
    for american in (True, False):
        for N in sorted({c["N"] for c in ok}):
            grp = [c for c in ok + bad if c["N"] == N and c["american"] == american]
            if not grp:
                continue
            cols = [[c[k] for c in grp] for k in ("S0", "K", "T", "r", "q", "sigma")]
            call = [1 if c["kind"] == "call" else 0 for c in grp]
            for fn, tol in ((crr_c.crr_loop_ref, 1e-13), (crr_c.crr_loop_fast, 1e-11)):
                if fn is crr_c.crr_loop_ref and N > 300:
                    continue   # two pow() per node: keep the CPU suite short
                price, status = fn(*cols, call, N, american)
                for c, p, st in zip(grp, price, status):
                    if "error" in c:
                        assert st == 1 and p != p, c
                    else:
                        assert st == 0 and abs(p - c["price"]) <= tol * max(1.0, abs(c["price"])), (c, p)
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(1_000_000)
    price, _ = crr_c.crr_loop_fast(S0[:64], K[:64], T[:64], r[:64], q[:64], sig[:64], is_call[:64], 1000, True)
    np.testing.assert_allclose(price, g["cfg2_first64_american_put_N1000"], rtol=1e-11)
    with pytest.raises(ValueError, match="N must be positive."):
        crr_c.crr_loop_ref([100.0], [100.0], [1.0], [0.05], [0.0], [0.2], [1], 0, False)


def _grid_from_samples(npz):
    x0, h, dt = npz["lvm_table/grid"]
    return sigma_grid.SigmaGrid(npz["lvm_table/table"], x0, h, dt)


def _make_paths(tag, meta, npz):
    a = dict(meta["args"])
    gen = meta["gen"]
    seed, n_steps, n_paths, anti = a.pop("seed"), a.pop("n_steps"), a.pop("n_paths"), a.pop("antithetic")
    if gen == "gbm":
        (z,) = path_schemes.draw_normals(seed, n_steps, n_paths)
        return path_schemes.gbm_exact(z, antithetic=anti, **a)
    if gen == "heston":
        z2, zp = path_schemes.draw_normals(seed, n_steps, n_paths, 2)
        return path_schemes.heston_fte(z2, zp, antithetic=anti, **a)
    if gen == "sabr":
        z2, zp = path_schemes.draw_normals(seed, n_steps, n_paths, 2)
        return path_schemes.sabr_lognormal_vol(z2, zp, antithetic=anti, **a)
    if gen == "gbm_milstein":
        (z,) = path_schemes.draw_normals(seed, n_steps, n_paths)
        return path_schemes.gbm_milstein(z, antithetic=anti, **a)
    if gen == "merton":
        z, k, zj = path_schemes.merton_draws(seed, n_steps, n_paths, a["lam"], a["T"])
        return path_schemes.merton_jumps(z, k, zj, antithetic=anti, **a)
    if meta["sigma"] == "const0.2":
        fn = lambda S, t: 0.2 * np.ones_like(S)  # noqa: E731
    elif meta["sigma"] == "lvm_table":
        fn = _grid_from_samples(npz)
    else:
        raise KeyError(meta["sigma"])
    (z,) = path_schemes.draw_normals(seed, n_steps, n_paths)
    if gen == "lv_milstein":
        return path_schemes.lv_milstein(z, fn, antithetic=anti, dS_bump=meta.get("dS_bump", 0.01), **a)
    return path_schemes.lv_log_euler(z, fn, antithetic=anti, **a)


def price_by_key(paths, key, K, r, T):
    f = key.split("|")
    if f[0] == "barrier":
        return payoffs.barrier(paths, K, r, T, f[2], float(f[3]), f[1], float(f[4]))
    if f[0] == "asian":
        return payoffs.asian(paths, K, r, T, f[1], f[2], f[3])
    if f[0] == "digital":
        return payoffs.digital(paths, K, r, T, f[1], float(f[2]))
    if f[0] == "lookback":
        return payoffs.lookback(paths, r, T, f[1], K, f[2])
    raise KeyError(key)


SMALL = ["gbm_small_anti1", "gbm_small_anti0", "heston_small_anti1", "heston_small_anti0", "heston_trunc",
         "lvm_const", "lvm_table_anti1", "lvm_table_anti0", "lvm_table_bump3", "lv_euler_table",
         "gbm_milstein", "sabr_beta1.0", "sabr_beta0.5", "merton"]


@pytest.mark.parametrize("tag", SMALL)
def test_small_path_cases(tag):
    meta = load_json("mc_paths.json")[tag]
    npz = load_npz("mc_paths_samples.npz")
    paths = _make_paths(tag, meta, npz)
    rows = npz[f"{tag}/rows"]
    ref_rows = npz[f"{tag}/row_values"]
    if EXACT:
        assert np.array_equal(paths[rows, :], ref_rows)
        assert np.array_equal(paths.sum(axis=0), npz[f"{tag}/col_sums"])
    else:
        np.testing.assert_allclose(paths[rows, :], ref_rows, rtol=1e-12)
    for key, (price, se) in meta["prices"].items():
        p, s = price_by_key(paths, key, meta["K"], meta["r"], meta["T"])
        assert close(p, price) and close(s, se), (tag, key)


def test_dupire_restatement_and_table():
    g = load_json("dupire_surface.json")
    surf = dupire.Surface({float(k): v for k, v in g["svi"].items()}, {float(k): v for k, v in g["forwards"].items()})
    fn = surf.sigma_loc()
    for s, t, v in g["probes"]:
        assert close(float(fn(np.array([s]), t)[0]), v)
    table, x0, h, dt = dupire_grid_arrays()
    grid = sigma_grid.tabulate(fn, S0=g["S0"], T=1.0, n_steps=252, n_nodes=256, half_width=1.6)
    assert close(grid.x0, x0) and close(grid.h, h)
    if EXACT:
        assert np.array_equal(grid.table, table)
    else:
        np.testing.assert_allclose(grid.table, table, rtol=1e-12)
    # node values are reproduced exactly by the lookup, and the callable clamps outside the grid
    nodes = grid.nodes
    np.testing.assert_allclose(grid(nodes, 10 * dt), table[10], rtol=1e-10)
    assert grid(np.array([1e-3]), 0.0)[0] == table[0, 0]
    assert grid(np.array([1e6]), 0.0)[0] == table[0, -1]
    with pytest.raises(ValueError):
        grid(nodes, 0.5 * dt)


def test_cv_estimator_matches_terminal_estimator():
    # the CV definition on a path payoff degenerates to euro_price_mc's when the payoff is vanilla
    (z,) = path_schemes.draw_normals(3, 1, 5000)
    paths = path_schemes.gbm_exact(z, S0=100.0, r=0.03, q=0.0, sigma=0.2, T=1.0)
    st = payoffs.six_stats(payoffs.vanilla_payoff(paths, 110.0, "call"), paths[-1], 0.03, 1.0)
    a = payoffs.cv_estimator(st, 100.0)
    b = terminal_mc.estimator(st, S0=100.0, q=0.0, T=1.0, control_variate=True)
    assert a == pytest.approx(b, rel=1e-15)


def heston_call_semi_analytic(S0, K, T, r, q, v0, kappa, theta, xi, rho):
    """Heston (1993) call by numerical inversion of the characteristic function (Albrecher et al.
    'little trap' form).  Independent yardstick for the QE oracle, which has no reference code."""
    u = np.linspace(1e-8, 200.0, 200_001)
    x = np.log(S0)

    def cf(phi):
        d = np.sqrt((rho * xi * 1j * phi - kappa) ** 2 + xi * xi * (1j * phi + phi * phi))
        g = (kappa - rho * xi * 1j * phi - d) / (kappa - rho * xi * 1j * phi + d)
        C = (r - q) * 1j * phi * T + kappa * theta / (xi * xi) * (
            (kappa - rho * xi * 1j * phi - d) * T - 2.0 * np.log((1.0 - g * np.exp(-d * T)) / (1.0 - g)))
        D = (kappa - rho * xi * 1j * phi - d) / (xi * xi) * (1.0 - np.exp(-d * T)) / (1.0 - g * np.exp(-d * T))
        return np.exp(C + D * v0 + 1j * phi * x)

    k = np.log(K)
    f1 = np.real(np.exp(-1j * u * k) * cf(u - 1j) / (1j * u * cf(-1j)))
    f2 = np.real(np.exp(-1j * u * k) * cf(u) / (1j * u))
    P1 = 0.5 + np.trapezoid(f1, u) / np.pi
    P2 = 0.5 + np.trapezoid(f2, u) / np.pi
    return S0 * np.exp(-q * T) * P1 - K * np.exp(-r * T) * P2


def test_heston_qe_oracle_against_semi_analytic_price():
    """QE is 'parity unpinned' (no reference code); pin the restatement to the Heston closed form."""
    cases = [dict(S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, T=1.0),
             dict(S0=100.0, r=0.02, q=0.01, v0=0.09, kappa=1.5, theta=0.04, xi=0.9, rho=-0.5, T=0.5)]
    for c, n_steps in zip(cases, (16, 16)):
        exact = heston_call_semi_analytic(c["S0"], 100.0, c["T"], c["r"], c["q"], c["v0"], c["kappa"], c["theta"],
                                          c["xi"], c["rho"])
        z2, zp = path_schemes.draw_normals(11, n_steps, 200_000, 2)
        paths = path_schemes.heston_qe(z2, zp, **c)
        price, se = payoffs.discounted_mean_se(payoffs.vanilla_payoff(paths, 100.0, "call"), c["r"], c["T"])
        assert abs(price - exact) <= 4.0 * se + 0.02, (c, price, exact, se)
        # discounted terminal price is a martingale up to the scheme's (small) drift bias
        fwd = c["S0"] * np.exp((c["r"] - c["q"]) * c["T"])
        assert abs(paths[-1].mean() - fwd) <= 4.0 * paths[-1].std() / np.sqrt(paths.shape[1]) + 0.05
    # Black-Scholes limit of the closed form itself (xi -> 0): 10.4506 (reference tests/test_black_scholes.py:4-7)
    bs = heston_call_semi_analytic(100.0, 100.0, 1.0, 0.05, 0.0, 0.04, 1.0, 0.04, 1e-4, 0.0)
    assert abs(bs - 10.4506) < 2e-3
