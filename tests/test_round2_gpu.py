"""GPU tests of the round-2 additions behind the C ABI (run with -m gpu):
local-vol items in optp_mc_book (table handles), the executed-work counters of the two-zone CRR kernel,
the collective for a world of one, the sigma-table guards, the per-scheme control-variate mean, and the
crr_vec many-strikes call the reference times at 220 options/s (binomial.py:44-109)."""
import ctypes as C
import math
import warnings

import numpy as np
import pytest

from golden_io import dupire_grid_arrays
from oracle import crr_tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def op():
    import optpricer_b200 as o
    from optpricer_b200 import _native
    _native.init()
    return o


def test_book_local_vol_rows_equal_the_single_call_api(op):
    """scripts/desk_workflow_localvol_barrier.py:136-148 (Milstein local-vol paths, barrier + vanilla legs from
    the same paths) as rows of ONE optp_mc_book call, next to a bumped-surface reprice and plain GBM rows."""
    from optpricer_b200 import book as bk, exotics as ex, processes as pr
    tab = pr.SigmaTable(*dupire_grid_arrays())
    bumped = pr.SigmaTable(tab.table * 1.01, tab.x0, tab.h, tab.dt)
    n, steps = 200_000, 252
    base = dict(S0=100.0, K=100.0, T=1.0, r=0.05, q=0.02, sigma=0.2, kind="call", method="exotic", n_steps=steps,
                n_paths=n, seed=42)
    rows = [
        dict(base, id="lv_barrier", exotic_type="barrier", barrier=130.0, barrier_type="up-and-out", sigma_table=tab),
        dict(base, id="lv_digital", exotic_type="digital", sigma_table=tab),
        dict(base, id="lv_barrier_bumped", exotic_type="barrier", barrier=130.0, barrier_type="up-and-out",
             sigma_table=bumped),
        dict(base, id="gbm_barrier", exotic_type="barrier", barrier=130.0, barrier_type="up-and-out"),
        dict(base, id="bad_table", exotic_type="barrier", barrier=130.0, barrier_type="up-and-out", sigma_table=tab,
             n_steps=100),
    ]
    got = bk.price_book(rows)
    lp = pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, steps, n, tab, seed=42)
    want = ex.price_payoffs(lp, [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.digital_spec(100.0, "call")],
                            0.05, 1.0)
    assert (got[0]["price"], got[0]["stderr"]) == want[0]
    assert (got[1]["price"], got[1]["stderr"]) == want[1]
    lpb = pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, steps, n, bumped, seed=42)
    wantb = ex.barrier_price(lpb, 100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
    assert (got[2]["price"], got[2]["stderr"]) == wantb and wantb[0] != want[0][0]
    wantg = ex.barrier_price(pr.gbm_paths(100.0, 0.05, 0.02, 0.2, 1.0, steps, n, seed=42), 100.0, 0.05, 1.0, "call",
                             130.0, "up-and-out")
    assert (got[3]["price"], got[3]["stderr"]) == wantg
    assert got[4]["price"] is None and "sigma_table does not match" in got[4]["error"]


def test_sigma_table_handles(op):
    from optpricer_b200 import _native as nv
    lib = nv.init()
    table = np.full((8, 16), 0.2)
    hs = []
    for _ in range(3):
        h = C.c_int32()
        nv.check(lib.optp_sigma_table_create(nv.ptr(table), 8, 16, C.byref(h)), "create")
        hs.append(h.value)
    assert len(set(hs)) == 3 and min(hs) >= 1
    for h in hs:
        assert lib.optp_sigma_table_destroy(h) == 0
    assert lib.optp_sigma_table_destroy(hs[0]) > 0           # already gone: argument error, not a crash
    assert lib.optp_sigma_table_destroy(0) > 0
    # a local-vol book item must name a live handle
    it = (nv.OptpBookItem * 1)()
    it[0].n_payoffs, it[0].n_steps, it[0].n_base, it[0].antithetic = 1, 8, 1000, 1
    it[0].proc.kind, it[0].proc.S0, it[0].proc.T, it[0].proc.n_nodes = nv.PROC_LV_MILSTEIN, 100.0, 1.0, 16
    it[0].proc.x0, it[0].proc.h, it[0].proc.dS_bump = math.log(100.0) - 1.0, 2.0 / 15, 0.01
    from optpricer_b200 import exotics as ex
    pay = ex._native_payoffs([ex.vanilla_spec(100.0, "call")], 0.05, 1.0)
    out = np.zeros(nv.NSTATS)
    assert lib.optp_mc_book(it, 1, pay, 1, nv.ptr(out)) > 0
    assert "table handle" in nv.last_error()


def test_crr_executed_work_counters(op):
    """bench.py's physical roofline figure: the counting build of the two-zone kernel prices identically and
    reports what its level loops executed."""
    from optpricer_b200 import _native as nv
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(100_000)
    plain = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True)
    nv.set_option("crr_count_ops", 1)
    try:
        counted = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True)
        fp64 = nv.get_counter("crr_fp64_lane_ops") / S0.size
        nodes = nv.get_counter("crr_node_updates") / S0.size
    finally:
        nv.set_option("crr_count_ops", 0)
    np.testing.assert_array_equal(counted, plain)
    assert 60_000 < nodes < 500_500            # the band cuts skip most of the 500 500 node updates ...
    assert nodes < fp64 < 3.0 * nodes + 2.0 * 1000 * 32   # ... and most executed nodes cost one DFMA
    nv.set_option("crr_count_ops", 1)          # setting the option resets the counters
    assert nv.get_counter("crr_node_updates") == 0
    nv.set_option("crr_count_ops", 0)
    with pytest.raises(ValueError):
        nv.get_counter("no_such_counter")


def test_collective_for_a_world_of_one(op):
    from optpricer_b200 import _native as nv
    lib = nv.init()
    assert lib.optp_comm_init(0, 1, None) == 0
    r_, w_ = C.c_int(-1), C.c_int(-1)
    assert lib.optp_comm_info(C.byref(r_), C.byref(w_)) == 0 and (r_.value, w_.value) == (0, 1)
    a = np.arange(24, dtype=np.float64)
    b = a.copy()
    assert lib.optp_allreduce_stats(nv.ptr(b), b.size) == 0
    np.testing.assert_array_equal(a, b)
    assert lib.optp_comm_init(3, 2, None) > 0       # rank outside the world
    assert lib.optp_comm_destroy() == 0


def test_sigma_table_guards_and_clamp_warning(op):
    from optpricer_b200 import exotics as ex, processes as pr
    tab = pr.SigmaTable(*dupire_grid_arrays())
    with pytest.raises(ValueError, match="tabulated for dt"):
        pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 2.0, 252, 1000, tab, seed=1)        # another T
    with pytest.raises(ValueError, match="outside the table"):
        pr.milstein_local_vol_paths(1000.0, 0.05, 0.02, 1.0, 252, 1000, tab, seed=1)       # another S0
    flat = lambda S, t: 0.6 + 0.0 * S  # noqa: E731
    # default grid: widened from +-1.6 to 6 sigma sqrt(T) = 5.09 at the same node spacing
    n_nodes, half = pr.SigmaTable.auto_grid(flat, 100.0, 2.0, 100)
    assert half == pytest.approx(6 * 0.6 * math.sqrt(2.0)) and n_nodes == math.ceil(255 * half / 1.6) + 1
    assert pr.SigmaTable.auto_grid(lambda S, t: 0.2 + 0.0 * S, 100.0, 1.0, 252) == (256, 1.6)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        lp = pr.milstein_local_vol_paths(100.0, 0.05, 0.0, 2.0, 100, 200_000, flat, seed=3)
        wide = ex.barrier_price(lp, 100.0, 0.05, 2.0, "call", 400.0, "up-and-out")          # no warning
    # a deliberately narrow table: several per cent of the paths leave it -> RuntimeWarning, price moves
    with pytest.warns(RuntimeWarning, match="left the tabulated"):
        lpn = pr.milstein_local_vol_paths(100.0, 0.05, 0.0, 2.0, 100, 200_000, flat, seed=3, sigma_grid=(64, 0.8))
        ex.barrier_price(lpn, 100.0, 0.05, 2.0, "call", 400.0, "up-and-out")
    assert wide[0] > 0


def test_control_variate_uses_the_discrete_mean_of_the_scheme(op):
    """Milstein / Euler-in-S schemes have E[S_T] = S0 (1 + (r-q) dt)^n, not S0 e^{(r-q)T}; with the
    continuous mean the CV estimator carries a bias that does not shrink with the path count."""
    from optpricer_b200 import exotics as ex, processes as pr
    lp = pr.gbm_milstein_paths(100.0, 0.12, 0.0, 0.2, 1.0, 4, 20_000_000, seed=9)
    spec = [ex.vanilla_spec(100.0, "call")]
    stats, _, _ = ex.path_stats(lp, spec, 0.12, 1.0)
    plain = ex.estimate(stats[0])
    cv = ex.price_payoffs(lp, spec, 0.12, 1.0, control_variate=True)[0]
    assert cv[1] < 0.6 * plain[1]
    assert abs(cv[0] - plain[0]) <= 4.0 * plain[1], (cv, plain)
    wrong = ex.estimate(stats[0], control_variate=True, expected_y=100.0)      # the continuous-time mean
    assert abs(wrong[0] - plain[0]) > 10.0 * plain[1], (wrong, plain)
    assert ex._expected_y(lp, 0.12, 1.0) == pytest.approx(100.0 * 1.03 ** 4 * math.exp(-0.12))
    qe = pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 8, 1000, seed=1, scheme="qe")
    with pytest.raises(ValueError, match="martingale correction"):
        ex.asian_price(qe, 100.0, 0.05, 1.0, "call", control_variate=True)


def test_crr_european_sum_at_its_depth_limit_with_extreme_probabilities(op):
    """N = 20000 is where crr_euro_sum_kernel stops; p close to 0 and to 1 against the backward
    induction (optp_set_option("crr_euro_sum", 0))."""
    from optpricer_b200 import _native as nv
    S0 = np.array([100.0, 100.0, 100.0, 100.0, 50.0, 150.0])
    K = np.array([100.0, 90.0, 110.0, 100.0, 55.0, 140.0])
    T = np.array([1.0, 2.0, 0.5, 1.0, 1.0, 1.0])
    # p = (e^{(r-q)dt} - d)/(u - d): large |r - q| with small sigma pushes p to the ends of (0, 1)
    r = np.array([0.03, 1.20, -1.00, 0.03, 2.0, 0.0])
    q = np.array([0.00, 0.00, 0.00, 0.00, 0.0, 1.8])
    sig = np.array([0.2, 0.02, 0.02, 0.9, 0.03, 0.03])
    is_call = np.array([1, 1, 0, 0, 1, 0], dtype=np.uint8)
    fast, st = op.crr_batch(S0, K, T, r, q, sig, is_call, 20000, american=False, return_status=True)
    nv.set_option("crr_euro_sum", 0)
    try:
        slow, st2 = op.crr_batch(S0, K, T, r, q, sig, is_call, 20000, american=False, return_status=True)
    finally:
        nv.set_option("crr_euro_sum", 1)
    np.testing.assert_array_equal(st, st2)
    ok = st == 0
    assert ok.sum() >= 4
    np.testing.assert_allclose(fast[ok], slow[ok], rtol=1e-10, atol=1e-10)


def test_crr_vec_many_strikes_on_one_tree(op):
    """binomial.py:44-109 with 1000 strikes, N = 1000, American (the call bench.py times): every strike against
    the scalar path, a sample against the oracle's shared-tree restatement."""
    strikes = np.linspace(60.0, 140.0, 1000)
    got = op.crr_vec(100.0, strikes, 1.0, 0.03, 0.0, 0.2, "put", N=1000, american=True)
    assert got.shape == (1000,)
    pick = [0, 137, 500, 731, 999]
    want = crr_tree.crr_shared_tree(100.0, strikes[pick], 1.0, 0.03, 0.0, 0.2, "put", N=1000, american=True)
    np.testing.assert_allclose(got[pick], want, rtol=1e-10, atol=1e-10)
    assert np.all(np.diff(got) > 0)                      # a put's price rises with the strike
    one = op.crr(op.OptionSpec(100.0, float(strikes[500]), 1.0, 0.03, 0.2), "put", 1000, american=True)
    assert one == got[500]
    kinds = np.where(np.arange(1000) % 2 == 0, "call", "put")
    mixed = op.crr_vec(100.0, strikes, 1.0, 0.03, 0.01, 0.2, kinds, N=1000, american=True)
    want = crr_tree.crr_shared_tree(100.0, strikes[pick], 1.0, 0.03, 0.01, 0.2, kinds[pick], N=1000, american=True)
    np.testing.assert_allclose(mixed[pick], want, rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------ Heston QE (no reference code: parity unpinned)
QE_CASES = [dict(S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, T=1.0),
            dict(S0=100.0, r=0.02, q=0.01, v0=0.09, kappa=1.5, theta=0.04, xi=0.9, rho=-0.5, T=0.5)]


def _qe_grid():
    """20 (kappa, theta, xi, v0, dt) points: both branches of the scheme, Feller-satisfying (2 kappa theta >= xi^2)
    and Feller-violating sets, v0 from far below to far above theta."""
    g = []
    for kappa, theta, xi in ((2.0, 0.04, 0.3), (1.5, 0.04, 0.9), (0.5, 0.09, 1.0), (4.0, 0.02, 0.5), (0.3, 0.04, 0.6)):
        for v0, dt in ((0.25 * theta, 1.0 / 12), (theta, 1.0 / 50), (4.0 * theta, 1.0 / 4), (0.01 * theta, 1.0 / 12)):
            g.append((kappa, theta, xi, v0, dt))
    return g


@pytest.mark.parametrize("block", [0, 1, 2, 3])
def test_heston_qe_variance_step_has_the_exact_cir_moments(op, block):
    """One QE step of the variance from a fixed v0 against the exact conditional mean and variance of the CIR
    process (Andersen 2008, eqs. 17-18) -- the two moments the scheme is built to match -- over a 20-point grid
    (5 per block); 4e7 effective paths per point, bar 1e-3 relative (or 4 sampling standard errors if wider)."""
    from optpricer_b200 import processes as pr
    n = 20_000_000
    for kappa, theta, xi, v0, dt in _qe_grid()[5 * block:5 * block + 5]:
        _, V = pr.heston_paths(100.0, 0.03, 0.0, v0, kappa, theta, xi, -0.5, dt, 1, n, seed=17, scheme="qe",
                               return_variance=True)
        v1 = V[1]
        E = math.exp(-kappa * dt)
        m = theta + (v0 - theta) * E
        s2 = v0 * xi * xi * E * (1.0 - E) / kappa + theta * xi * xi * (1.0 - E) ** 2 / (2.0 * kappa)
        assert (v1 >= 0.0).all()
        mean = float(v1.mean())
        d = v1 - mean
        var = float((d * d).mean())
        m4 = float((d ** 4).mean())
        se_mean = math.sqrt(var / v1.size)
        se_var = math.sqrt(max(m4 - var * var, 0.0) / v1.size)
        feller = 2.0 * kappa * theta >= xi * xi
        assert abs(mean - m) <= max(1e-3 * m, 4.0 * se_mean), (kappa, theta, xi, v0, dt, feller, mean, m, se_mean)
        assert abs(var - s2) <= max(1e-3 * s2, 4.0 * se_var), (kappa, theta, xi, v0, dt, feller, var, s2, se_var)
        del V, v1, d


def test_heston_qe_martingale_correction(op):
    """K0* (Andersen 2008, sec. 4.3.2): with coarse steps, strong correlation and high vol-of-vol the plain scheme
    misses E[S_T] by many standard errors; the corrected one does not -- fp64 and fp32 kernels; the fp64 kernel
    equals the numpy restatement on the same normals; and the control variate becomes available."""
    from optpricer_b200 import exotics as ex, processes as pr
    from oracle import path_schemes, payoffs
    c = dict(S0=100.0, r=0.0, q=0.0, v0=0.04, kappa=0.5, theta=0.04, xi=1.0, rho=-0.9, T=10.0)
    spec = [ex.vanilla_spec(100.0, "call")]

    def mean_y(mart, mode):
        lp = pr.heston_paths(n_steps=10, n_paths=4_000_000, seed=2, scheme="qe", martingale_correction=mart, mode=mode, **c)
        s, _, _ = ex.path_stats(lp, spec, 0.0, c["T"])
        n, sy, syy = s[0][0], s[0][3], s[0][4]
        my = sy / n
        return my, math.sqrt(max(syy / n - my * my, 0.0) / n)

    for mode in ("fp32", "fp64"):
        plain, se = mean_y(False, mode)
        fixed, se2 = mean_y(True, mode)
        assert abs(plain - 100.0) > 10.0 * se, (mode, plain, se)
        assert abs(fixed - 100.0) <= 4.0 * se2, (mode, fixed, se2)
    # check mode against the restatement, both branches active
    for case in (QE_CASES[1], c):
        lp = pr.heston_paths(n_steps=12, n_paths=3000, seed=5, rng="numpy", scheme="qe", martingale_correction=True, **case)
        z2, zp = path_schemes.draw_normals(5, 12, 3000, 2)
        ref = path_schemes.heston_qe(z2, zp, martingale_correction=True, **case)
        np.testing.assert_allclose(np.asarray(lp), ref, rtol=1e-11)
        plain = path_schemes.heston_qe(z2, zp, martingale_correction=False, **case)
        assert np.max(np.abs(ref[-1] / plain[-1] - 1.0)) > 1e-6
    # fp32 tracks fp64 on the same draws, and the control variate is accepted (and helps)
    lp32 = pr.heston_paths(n_steps=32, n_paths=400_000, seed=3, scheme="qe", martingale_correction=True, **QE_CASES[1])
    lp64 = pr.heston_paths(n_steps=32, n_paths=400_000, seed=3, scheme="qe", martingale_correction=True, mode="fp64", **QE_CASES[1])
    a = ex.price_payoffs(lp32, spec, 0.02, 0.5)[0]
    b = ex.price_payoffs(lp64, spec, 0.02, 0.5)[0]
    assert abs(a[0] - b[0]) <= 0.1 * b[1], (a, b)
    cv = ex.price_payoffs(lp32, spec, 0.02, 0.5, control_variate=True)[0]
    assert cv[1] < 0.8 * a[1] and abs(cv[0] - a[0]) <= 4.0 * a[1]


@pytest.mark.parametrize("case", [0, 1])
@pytest.mark.parametrize("n_steps", [64, 128])
def test_heston_qe_against_the_semi_analytic_price_without_an_allowance(op, case, n_steps):
    """Martingale-corrected QE at 64 / 128 steps, 1.6e7 effective paths: the vanilla call within 4 of the run's own
    standard errors of the semi-analytic Heston price -- no discretisation allowance (round 1 granted 0.02)."""
    from optpricer_b200 import exotics as ex, processes as pr
    from test_oracle_cpu import heston_call_semi_analytic
    c = QE_CASES[case]
    exact = heston_call_semi_analytic(c["S0"], 100.0, c["T"], c["r"], c["q"], c["v0"], c["kappa"], c["theta"], c["xi"], c["rho"])
    lp = pr.heston_paths(n_steps=n_steps, n_paths=8_000_000, seed=11 + n_steps, scheme="qe", martingale_correction=True, **c)
    plain = ex.price_payoffs(lp, [ex.vanilla_spec(100.0, "call")], c["r"], c["T"])[0]
    cv = ex.price_payoffs(lp, [ex.vanilla_spec(100.0, "call")], c["r"], c["T"], control_variate=True)[0]
    assert abs(plain[0] - exact) <= 4.0 * plain[1], (plain, exact)
    assert abs(cv[0] - exact) <= 4.0 * cv[1] + 1e-3, (cv, exact)


def test_dev_launches_on_several_streams_overlap_safely(op):
    """The *_dev Monte Carlo entry points take the slots of the reduction scratch round-robin: launches queued on
    different streams (no ordering between them) give the statistics of the same launches queued on one stream."""
    import torch
    from optpricer_b200 import _native as nv, exotics as ex
    lib = nv.init()
    dev = torch.device("cuda", nv.default_device())
    proc = nv.OptpProcess()
    proc.kind = nv.PROC_GBM
    proc.S0, proc.r, proc.q, proc.T, proc.sigma = 100.0, 0.05, 0.0, 1.0, 0.2
    pay = ex._native_payoffs([ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.asian_spec(100.0, "call")], 0.05, 1.0)
    n_jobs = 6

    def run(streams):
        outs = [torch.zeros(2 * nv.NSTATS, dtype=torch.float64, device=dev) for _ in range(n_jobs)]
        for rep in range(3):
            for j in range(n_jobs):
                st = streams[j % len(streams)]
                nv.check(lib.optp_mc_path_dev(C.byref(proc), pay, 2, 64, 300_000 + 1000 * j, 1, nv.MODE_FP32_PHILOX, 7 + j, 0,
                                              None, None, None, C.c_void_p(outs[j].data_ptr()), C.c_void_p(st.cuda_stream)),
                         "optp_mc_path_dev")
        torch.cuda.synchronize()
        return [o.cpu().numpy() for o in outs]

    one = run([torch.cuda.Stream(device=dev)])
    many = run([torch.cuda.Stream(device=dev) for _ in range(n_jobs)])
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("N", [250, 1000, 1500])
def test_crr_measured_segment_schedule_under_random_stress(op, N):
    """The two-zone kernel walks 32 + 64 + 64 ... levels per segment on the MEASURED speed of the exercise boundary
    (csrc/crr_batch.cu: kZoneSeg); exactness rests on its per-level checks, and contracts that fail them are retried
    on the rigorous schedule.  Wide random parameters (r dt down to 1e-8, 1.5 volatility, 10 years, deep in and out of
    the money, dividends, 30 % calls): every price equals the full-width kernel's, statuses are identical, and no
    eligible put is left to the full-width kernel at N >= 1000."""
    from optpricer_b200 import _native as nv
    n = 60_000
    g = np.random.default_rng(2026 + N)
    S0 = g.uniform(20.0, 200.0, n)
    K = S0 * np.exp(g.uniform(-0.6, 0.6, n))
    T = np.exp(g.uniform(np.log(0.01), np.log(10.0), n))
    sig = np.exp(g.uniform(np.log(0.03), np.log(1.5), n))
    r = np.exp(g.uniform(np.log(0.001), np.log(0.3), n))
    q = np.where(g.random(n) < 0.5, 0.0, g.uniform(0.0, 0.12, n))
    is_call = (g.random(n) < 0.3).astype(np.uint8)
    band, st = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    nv.set_option("crr_band_only", 1)
    try:
        _, st_only = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    finally:
        nv.set_option("crr_band_only", 0)
    nv.set_option("crr_band", 0)
    try:
        full, st_full = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
    finally:
        nv.set_option("crr_band", 1)
    np.testing.assert_array_equal(st, st_full)
    ok = st == 0
    assert ok.sum() > 0.99 * n
    np.testing.assert_allclose(band[ok], full[ok], rtol=1e-10, atol=1e-10)
    assert np.max(np.abs(band[ok] - full[ok]) / np.maximum(np.abs(full[ok]), 1e-8 * K[ok])) < 5e-10
    puts = ok & (is_call == 0) & (r > q)
    left = int(((st_only == 2) & puts).sum())
    assert left <= (0 if N >= 1000 else 40), left


def test_crr_drifting_window_regimes(op):
    """The regimes of the drifting-window walk (csrc/crr_batch.cu, OPTP_ZONE_DRIFT) against the full-width kernel, in
    RELATIVE terms down to prices of 1e-30 (an absolute bar hides a band that was lost on the way to level 0):
    deep out-of-the-money contracts with q > 0, whose last levels run in the regime without exercise nodes (tree above
    S1 = K(1-disc)/(1-e^{-q dt})) and whose window origin goes negative before that; odd tree depths (one closed-form
    level at the end); trees narrower than Z; and all of it mixed into one batch with ordinary contracts, so that
    every contract meets warp mates with wider and narrower bands."""
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(77)
    n = 24_000
    S0 = g.uniform(20, 300, n)
    K = S0 * np.exp(g.uniform(-1.3, 1.3, n))
    T = np.exp(g.uniform(np.log(0.01), np.log(3.0), n))
    sig = g.uniform(0.05, 0.9, n)
    r = g.uniform(0.0005, 0.12, n)
    q = np.where(g.random(n) < 0.25, 0.0, g.uniform(0.0, 0.1, n))
    is_call = (g.random(n) < 0.3).astype(np.uint8)
    for N in (1000, 999, 257, 31, 16, 3, 2, 1):
        nv.set_option("crr_band", 0)
        try:
            full, st_f = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
        finally:
            nv.set_option("crr_band", 1)
        zone, st = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
        np.testing.assert_array_equal(st, st_f)
        ok = st == 0
        big = ok & (full > 1e-30)
        rel = np.abs(zone[big] - full[big]) / full[big]
        # values below 2^-80 K = 8e-25 K are cut on the way, a few thousand of them per walk: prices
        # above 1e-10 K agree to 1e-9 relative, prices down to 1e-18 K to 1e-4, and anything above 1e-21 K is present
        firm = full[big] > 1e-10 * K[big]
        assert rel[firm].max() < 1e-9, (N, rel[firm].max())
        faint = full[big] > 1e-18 * K[big]
        assert rel[faint].max() < 1e-4, (N, rel[faint].max())
        live = full[big] > 1e-21 * K[big]
        assert np.all(zone[big][live] > 0.0), N
        np.testing.assert_allclose(zone[ok], full[ok], rtol=1e-10, atol=1e-10)
