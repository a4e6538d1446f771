"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the public Python
API of optpricer_b200, i.e. through the C ABI of liboptp_b200.so, and is compared with
  * the golden fixtures produced by the unmodified reference (tests/golden/), and
  * the numpy oracle (oracle/) on the same seeded inputs.
Tolerances (BASELINE.json north_star):
  check mode (reference's own PCG64 draws, fp64)   1e-12 relative on prices and per-path payoffs
  on-device Philox (fp32 and fp64 arithmetic)      |price - ref| <= 3 reference standard errors
  CRR                                              1e-10 relative
"""
import math
import os

import numpy as np
import pytest

from golden_io import load_json, load_npz, dupire_grid_arrays
from oracle import terminal_mc, path_schemes, payoffs, crr_tree, sigma_grid

pytestmark = pytest.mark.gpu

REL_CHECK = 1e-12
REL_SE = 1e-9
REL_CRR = 1e-10


@pytest.fixture(scope="module")
def op():
    import optpricer_b200 as o
    from optpricer_b200 import _native
    _native.init()
    assert os.path.basename(_native.LIB_PATH).startswith("liboptp_")   # (OPTP_LIB: a variant build of the same library)
    return o


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


# ------------------------------------------------------------------------------ euro_price_mc
def _np_dtype(kw):
    kw = dict(kw)
    if kw.get("dtype") == "float32":
        kw["dtype"] = np.float32
    return kw


def test_euro_mc_check_mode_matches_reference(op):
    g = load_json("euro_mc.json")
    for c in g["cases"]:
        kw = _np_dtype(c["kwargs"])
        tol = 2e-6 if kw.get("dtype") is np.float32 else REL_CHECK
        price, se = op.euro_price_mc(op.OptionSpec(**c["opt"]), c["kind"], n_paths=c["n_paths"], seed=c["seed"],
                                     mode="check", **kw)
        assert rel(price, c["price"]) <= tol, (c, price)
        assert rel(se, c["stderr"]) <= max(tol, REL_SE), (c, se)


def test_euro_mc_chunk_statistics(op):
    from optpricer_b200.monte_carlo import _mc_chunk_sumstats
    g = load_json("euro_mc.json")
    base = dict(S0=100.0, K=110.0, T=1.0, r=0.03, sigma=0.20, q=0.0)
    for ss, ref in zip(np.random.SeedSequence(1).spawn(2), g["cfg1_chunk_stats"]):
        got = _mc_chunk_sumstats(100_000, kind="call", antithetic=True, seed=ss, **base)
        assert got[0] == ref[0]
        for a, b in zip(got[1:], ref[1:]):
            assert rel(a, b) <= REL_CHECK
    assert _mc_chunk_sumstats(0, kind="call", antithetic=True, seed=1, **base) == (0, 0.0, 0.0, 0.0, 0.0, 0.0)


@pytest.mark.parametrize("mode", ["fp32", "fp64"])
def test_euro_mc_philox_within_three_reference_se(op, mode):
    g = load_json("euro_mc.json")
    for c in g["cases"]:
        if c["n_paths"] < 10_000:
            continue  # the reference's stderr of a 1- or 777-path run is not a usable yardstick
        kw = {k: v for k, v in c["kwargs"].items() if k in ("antithetic", "control_variate")}
        for seed in (c["seed"], 12345):
            # 16x the reference's paths: the bar is literally 3 reference standard errors
            price, se = op.euro_price_mc(op.OptionSpec(**c["opt"]), c["kind"], n_paths=16 * c["n_paths"], seed=seed,
                                         mode=mode, **kw)
            assert abs(price - c["price"]) <= 3.0 * c["stderr"] + 1e-12, (c, price, se)
            if c["n_paths"] > 1000:
                assert 0.8 < 4.0 * se / c["stderr"] < 1.25


def test_euro_mc_fp32_tracks_fp64_on_same_draws(op):
    opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
    for kind in ("call", "put"):
        for anti in (True, False):
            a = op.euro_price_mc(opt, kind, n_paths=1_000_003, seed=9, mode="fp32", antithetic=anti)
            b = op.euro_price_mc(opt, kind, n_paths=1_000_003, seed=9, mode="fp64", antithetic=anti)
            assert abs(a[0] - b[0]) <= 0.02 * b[1], (a, b)
            assert rel(a[1], b[1]) < 1e-4


def test_euro_mc_large_run_converges_to_black_scholes(op):
    # cfg 1 contract at 2e8 base paths: |MC - BS| must be a few standard errors of ~1e-4
    opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
    bs_call = 5.293398058044907
    p, se = op.euro_price_mc(opt, "call", n_paths=200_000_000, seed=3)
    assert se < 4e-4
    assert abs(p - bs_call) <= 4.0 * se + 2e-5


def test_euro_mc_edge_cases(op):
    opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
    for mode in ("fp32", "fp64", "check"):
        p, se = op.euro_price_mc(opt, "call", n_paths=0, seed=1, mode=mode)
        assert p != p and se != se
        p1 = op.euro_price_mc(opt, "call", n_paths=1, seed=1, mode=mode, return_stderr=False)
        assert isinstance(p1, float)
    with pytest.raises(ValueError, match="kind must be 'call' or 'put'"):
        op.euro_price_mc(opt, "straddle", n_paths=10, seed=1)
    # sharding invariance: statistics of [0,n) equal the sum over disjoint path ranges
    from optpricer_b200.monte_carlo import terminal_stats
    args = (100.0, 110.0, 1.0, 0.03, 0.0, 0.2, 1)
    whole = terminal_stats(*args, 1_000_001, seed=5)
    parts = [terminal_stats(*args, n, seed=5, path_offset=o) for o, n in ((0, 333_333), (333_333, 7), (333_340, 666_661))]
    np.testing.assert_allclose(sum(parts)[:6], whole[:6], rtol=2e-6)
    whole64 = terminal_stats(*args, 100_001, seed=5, mode="fp64")
    parts64 = [terminal_stats(*args, n, seed=5, mode="fp64", path_offset=o) for o, n in ((0, 50_001), (50_001, 50_000))]
    np.testing.assert_allclose(sum(parts64)[:6], whole64[:6], rtol=1e-12)


# ------------------------------------------------------------------------------ CRR
def test_crr_golden_cases(op):
    g = load_json("crr.json")
    for c in g["cases"]:
        opt = op.OptionSpec(c["S0"], c["K"], c["T"], c["r"], c["sigma"], c["q"])
        if "error" in c:
            with pytest.raises(ValueError) as e:
                op.crr(opt, c["kind"], c["N"], american=c["american"])
            assert str(e.value) == c["error"]
        else:
            v = op.crr(opt, c["kind"], c["N"], american=c["american"])
            assert rel(v, c["price"]) <= REL_CRR, (c, v)
    with pytest.raises(ValueError, match="N must be positive."):
        op.crr(op.OptionSpec(100, 100, 1, 0.05, 0.2), "call", 0)


def test_crr_vec_golden_cases(op):
    for c in load_json("crr.json")["crr_vec"]:
        v = op.crr_vec(c["S0"], np.array(c["K"]), c["T"], c["r"], c["q"], c["sigma"], c["kind"], c["N"],
                       american=c["american"])
        assert v.shape == (len(c["K"]),)
        np.testing.assert_allclose(v, c["prices"], rtol=REL_CRR, atol=REL_CRR)


def test_crr_cfg2_batch_against_reference_and_oracle(op):
    g = load_json("crr.json")
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(1_000_000)
    n = 4096
    price, status = op.crr_batch(S0[:n], K[:n], T[:n], r[:n], q[:n], sig[:n], is_call[:n], 1000, american=True,
                                 return_status=True)
    assert not status.any()
    ref64 = np.array(g["cfg2_first64_american_put_N1000"])
    np.testing.assert_allclose(price[:64], ref64, rtol=REL_CRR, atol=REL_CRR)
    oracle_p, _ = crr_tree.crr_loop(S0[64:96], K[64:96], T[64:96], r[64:96], q[64:96], sig[64:96], is_call[64:96],
                                    1000, True)
    np.testing.assert_allclose(price[64:96], oracle_p, rtol=REL_CRR, atol=REL_CRR)
    # American put >= European put >= 0 and >= intrinsic, for the whole slice
    euro = op.crr_batch(S0[:n], K[:n], T[:n], r[:n], q[:n], sig[:n], is_call[:n], 1000, american=False)
    assert np.all(price >= euro - 1e-12) and np.all(price >= np.maximum(K[:n] - S0[:n], 0.0) - 1e-12)
    # the result of a contract does not depend on its position or on the batch it is in
    again = op.crr_batch(S0[100:229], K[100:229], T[100:229], r[100:229], q[100:229], sig[100:229], is_call[100:229],
                         1000, american=True)
    np.testing.assert_array_equal(again, price[100:229])


def test_crr_cfg2_sample_against_c_oracle(op):
    """8192 config-2 contracts at N = 1000 (American and European) against the table-based C
    restatement (oracle/crr_tree.c, itself pinned to the reference fixtures at 1e-11)."""
    from oracle import crr_c
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(1_000_000)
    sl = slice(500_000, 508_192)
    cols = [a[sl] for a in (S0, K, T, r, q, sig, is_call)]
    for american in (True, False):
        want, st_w = crr_c.crr_loop_fast(*cols, 1000, american)
        got, st_g = op.crr_batch(*cols, 1000, american=american, return_status=True)
        np.testing.assert_array_equal(st_g, st_w)
        np.testing.assert_allclose(got, want, rtol=REL_CRR, atol=REL_CRR)


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 255, 256, 257, 1000, 1023, 1024, 1500, 4000, 9000])
def test_crr_ragged_sizes_against_oracle(op, N):
    g = np.random.default_rng(N)
    n = 9
    S0 = g.uniform(50, 150, n); K = S0 * g.uniform(0.8, 1.2, n); T = g.uniform(0.1, 2.0, n)
    sig = g.uniform(0.1, 0.6, n); r = g.uniform(0.0, 0.08, n); q = g.uniform(0.0, 0.04, n)
    is_call = (g.random(n) < 0.5).astype(np.uint8)
    for american in (True, False):
        ref, ref_status = crr_tree.crr_loop(S0, K, T, r, q, sig, is_call, N, american)
        got, status = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=american, return_status=True)
        np.testing.assert_array_equal(status, ref_status)
        ok = ref_status == 0
        np.testing.assert_allclose(got[ok], ref[ok], rtol=REL_CRR, atol=REL_CRR)
        assert np.all(np.isnan(got[~ok]))


def test_crr_bad_contracts_do_not_disturb_their_warp_mates(op):
    """Four options share a warp in the two-zone kernel: a NaN / non-positive / p-out-of-range contract in
    the middle of a batch gets NaN + status 1 and every other contract keeps, bit for bit, the price it has
    in a clean batch."""
    g = np.random.default_rng(8)
    n = 40_000                                     # large enough for four options per warp
    S0 = g.uniform(50, 150, n); K = S0 * g.uniform(0.8, 1.2, n); T = g.uniform(0.1, 2.0, n)
    sig = g.uniform(0.1, 0.6, n); r = np.full(n, 0.03); q = g.choice([0.0, 0.01, 0.05], n)
    is_call = (g.random(n) < 0.3).astype(np.uint8)
    clean = op.crr_batch(S0, K, T, r, q, sig, is_call, 400, american=True)
    bad = np.arange(n) % 7 == 3
    S0b, sigb, Tb, rb = S0.copy(), sig.copy(), T.copy(), r.copy()
    kinds = np.arange(n) % 4
    S0b[bad & (kinds == 0)] = np.nan
    sigb[bad & (kinds == 1)] = 1e-6                # p outside (0, 1)
    Tb[bad & (kinds == 2)] = np.inf
    rb[bad & (kinds == 3)] = 40.0                  # p outside (0, 1)
    got, st = op.crr_batch(S0b, K, Tb, rb, q, sigb, is_call, 400, american=True, return_status=True)
    assert np.all(st[bad] == 1) and np.all(np.isnan(got[bad]))
    assert np.all(st[~bad] == 0)
    np.testing.assert_array_equal(got[~bad], clean[~bad])


def test_crr_european_expectation_kernel_equals_backward_induction(op):
    """European trees are priced as the discounted binomial expectation of the terminal payoff (O(N) per
    option); `crr_euro_sum = 0` restores the backward induction.  Both against each other over the corners of
    parameter space and several depths, and against the C restatement of binomial.py:7-38 (1e-10)."""
    from oracle import crr_c
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(17)
    n = 20_000
    S0 = g.uniform(20, 300, n); K = S0 * np.exp(g.uniform(-1.2, 1.2, n)); T = np.exp(g.uniform(np.log(0.01), np.log(5.0), n))
    sig = g.uniform(0.05, 0.9, n); r = g.uniform(-0.01, 0.12, n); q = np.where(g.random(n) < 0.5, 0.0, g.uniform(0.0, 0.1, n))
    is_call = (g.random(n) < 0.5).astype(np.uint8)
    worst = 0.0
    for N in (1, 2, 5, 64, 257, 1000, 1023, 1024, 4000):
        m = n if N <= 1024 else 400
        cols = [c[:m] for c in (S0, K, T, r, q, sig, is_call)]
        fast, st_f = op.crr_batch(*cols, N, return_status=True)
        nv.set_option("crr_euro_sum", 0)
        try:
            tree, st_t = op.crr_batch(*cols, N, return_status=True)
        finally:
            nv.set_option("crr_euro_sum", 1)
        np.testing.assert_array_equal(st_f, st_t)
        ok = st_t == 0
        np.testing.assert_allclose(fast[ok], tree[ok], rtol=REL_CRR, atol=REL_CRR)
        assert np.all(np.isnan(fast[~ok]))
        worst = max(worst, float(np.max(np.abs(fast[ok] - tree[ok]) / np.maximum(1.0, np.abs(tree[ok])))))
    assert worst < 2e-11, worst
    want, st_w = crr_c.crr_loop_ref(S0[:256], K[:256], T[:256], r[:256], q[:256], sig[:256], is_call[:256], 500, False)
    got, st_g = op.crr_batch(S0[:256], K[:256], T[:256], r[:256], q[:256], sig[:256], is_call[:256], 500, return_status=True)
    np.testing.assert_array_equal(st_g, st_w)
    np.testing.assert_allclose(got[st_w == 0], want[st_w == 0], rtol=REL_CRR, atol=REL_CRR)


def test_crr_never_exercised_contracts_take_the_european_recursion(op):
    """American calls with q <= 0 < r and American puts with r <= 0 <= q are never exercised early
    (binomial.py:30-36 always keeps the continuation value): they are routed to the European kernel and
    must equal both the American full-width kernel and the European price bit for bit / to 1e-10."""
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(31)
    n = 2000
    S0 = g.uniform(50, 150, n); K = S0 * g.uniform(0.8, 1.2, n); T = g.uniform(0.1, 2.0, n); sig = g.uniform(0.1, 0.6, n)
    is_call = (np.arange(n) % 2).astype(np.uint8)
    r = np.where(is_call == 1, g.uniform(0.001, 0.08, n), -g.uniform(0.0, 0.02, n))
    r[2:40:2] = 0.0                                  # puts at r = 0 exactly
    q = np.where(is_call == 1, -g.uniform(0.0, 0.02, n), g.uniform(0.0, 0.05, n))
    amer, st = op.crr_batch(S0, K, T, r, q, sig, is_call, 500, american=True, return_status=True)
    euro = op.crr_batch(S0, K, T, r, q, sig, is_call, 500, american=False)
    assert np.all(st == 0)
    np.testing.assert_array_equal(amer, euro)
    nv.set_option("crr_band", 0)
    try:
        full = op.crr_batch(S0, K, T, r, q, sig, is_call, 500, american=True)
    finally:
        nv.set_option("crr_band", 1)
    np.testing.assert_allclose(amer, full, rtol=REL_CRR, atol=REL_CRR)


def test_crr_trees_deeper_than_shared_memory(op):
    """N > 9685: the one-option-per-CTA kernel keeps its level arrays in global memory; American puts still
    take the two-zone passes first.  Against the C restatement (oracle/crr_tree.c)."""
    from oracle import crr_c
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(12)
    n = 6
    S0 = g.uniform(80, 120, n); K = S0 * g.uniform(0.9, 1.1, n); T = g.uniform(0.3, 1.5, n)
    sig = g.uniform(0.15, 0.5, n); r = np.full(n, 0.04); q = np.array([0.0, 0.01, 0.0, 0.02, 0.0, 0.06])
    is_call = np.array([0, 0, 1, 1, 0, 0], dtype=np.uint8)
    N = 12_000
    for american in (True, False):
        want, st_w = crr_c.crr_loop_fast(S0, K, T, r, q, sig, is_call, N, american)
        got, st_g = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=american, return_status=True)
        np.testing.assert_array_equal(st_g, st_w)
        np.testing.assert_allclose(got, want, rtol=REL_CRR, atol=REL_CRR)
    nv.set_option("crr_band", 0)
    try:
        full = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True)
    finally:
        nv.set_option("crr_band", 1)
    banded = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True)
    np.testing.assert_allclose(full, banded, rtol=REL_CRR, atol=REL_CRR)


def test_crr_banded_kernel_equals_full_width_kernel(op):
    """The banded American-put kernels (exercise-region / zero-region cuts; two-zone kernel with 1, 2
    and 4 options per warp, and the earlier single-zone kernel) against the full-width kernel on the
    same contracts, including the corners of parameter space: deep in / out of the money, tiny and
    long maturities, low and high volatility, q > 0, r <= q (declined -> fallback)."""
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(99)
    n = 30_000
    S0 = g.uniform(20, 300, n); K = S0 * np.exp(g.uniform(-1.2, 1.2, n)); T = np.exp(g.uniform(np.log(0.01), np.log(5.0), n))
    sig = g.uniform(0.05, 0.9, n); r = g.uniform(0.0005, 0.12, n); q = np.where(g.random(n) < 0.5, 0.0, g.uniform(0.0, 0.1, n))
    is_call = (g.random(n) < 0.2).astype(np.uint8)
    try:
        for N in (1000, 257, 64, 29, 28, 5):
            nv.set_option("crr_band", 0)
            full, st_f = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
            nv.set_option("crr_band", 1)
            ok = st_f == 0
            zone = {}
            for G in (1, 2, 4):
                nv.set_option("crr_zone_group", G)
                zone[G], st_b = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
                np.testing.assert_array_equal(st_b, st_f)
                np.testing.assert_allclose(zone[G][ok], full[ok], rtol=REL_CRR, atol=REL_CRR)
            # the arithmetic applied to a contract does not depend on the lane layout or on its warp mates
            np.testing.assert_array_equal(zone[1][ok], zone[2][ok])
            np.testing.assert_array_equal(zone[1][ok], zone[4][ok])
            nv.set_option("crr_zone_group", 4)
            perm = g.permutation(n)
            shuffled = op.crr_batch(S0[perm], K[perm], T[perm], r[perm], q[perm], sig[perm], is_call[perm], N, american=True)
            np.testing.assert_array_equal(shuffled[ok[perm]], zone[4][perm][ok[perm]])
            nv.set_option("crr_zone_group", 0)
            nv.set_option("crr_band_kernel", 1)   # the single-zone banded kernel
            single, st_s = op.crr_batch(S0, K, T, r, q, sig, is_call, N, american=True, return_status=True)
            nv.set_option("crr_band_kernel", 2)
            np.testing.assert_array_equal(st_s, st_f)
            np.testing.assert_allclose(single[ok], full[ok], rtol=REL_CRR, atol=REL_CRR)
    finally:
        nv.set_option("crr_band", 1)
        nv.set_option("crr_band_kernel", 2)
        nv.set_option("crr_zone_group", 0)
    # how much of a typical book the banded kernel actually takes (config-2 contracts: all of it)
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(20_000)
    nv.set_option("crr_band_only", 1)
    try:
        _, st = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True, return_status=True)
        assert np.count_nonzero(st == 2) == 0
    finally:
        nv.set_option("crr_band_only", 0)


def test_crr_dividend_calls_and_high_yield_puts_take_the_banded_kernel(op):
    """American calls with q > 0 are priced as their dual puts (C(S0,K,r,q) = P(K,S0,q,r) node by node
    on a CRR lattice) and puts with q > r start from the boundary S1 ~ K r/q below the strike: both go
    through the two-zone kernel and must match the C restatement of binomial.py:7-38 at 1e-10."""
    from oracle import crr_c
    from optpricer_b200 import _native as nv
    g = np.random.default_rng(2024)
    n = 4096
    S0 = g.uniform(50, 150, n); K = S0 * g.uniform(0.8, 1.2, n); T = g.uniform(0.1, 2.0, n)
    sig = g.uniform(0.1, 0.6, n)
    r = g.uniform(0.0, 0.08, n); q = g.uniform(0.002, 0.08, n)
    is_call = (np.arange(n) % 2).astype(np.uint8)
    r[:64] = 0.0                      # calls with r = 0 (dual put with q' = 0); puts with r = 0 are never exercised
    want, st_w = crr_c.crr_loop_fast(S0, K, T, r, q, sig, is_call, 1000, True)
    got, st_g = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True, return_status=True)
    np.testing.assert_array_equal(st_g, st_w)
    np.testing.assert_allclose(got, want, rtol=REL_CRR, atol=REL_CRR)
    # a few of them against the reference-order loop (two pow() per node) as well
    ref, _ = crr_c.crr_loop_ref(S0[:96], K[:96], T[:96], r[:96], q[:96], sig[:96], is_call[:96], 1000, True)
    np.testing.assert_allclose(got[:96], ref, rtol=REL_CRR, atol=REL_CRR)
    # ... and the banded kernel really took them (status 2 = declined, left to the full-width kernel)
    nv.set_option("crr_band_only", 1)
    try:
        _, st = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True, return_status=True)
    finally:
        nv.set_option("crr_band_only", 0)
    calls = is_call == 1
    hy_puts = (~calls) & (q > r) & (r > 1e-3)
    assert np.count_nonzero(st[calls] == 0) >= 0.9 * np.count_nonzero(calls)
    assert np.count_nonzero(st[hy_puts] == 0) >= 0.8 * np.count_nonzero(hy_puts)


def test_crr_batch_flags_bad_contracts_without_aborting(op):
    S0 = np.array([100.0, 100.0, 100.0]); K = np.array([100.0, 100.0, 90.0])
    T = np.array([1.0, 2.0, 1.0]); r = np.array([0.05, 0.30, 0.05]); q = np.zeros(3)
    sig = np.array([0.2, 0.01, 0.2])
    price, status = op.crr_batch(S0, K, T, r, q, sig, [1, 1, 0], 4, return_status=True)
    assert status.tolist() == [0, 1, 0] and math.isnan(price[1])
    assert rel(price[0], crr_tree.crr_scalar(100.0, 100.0, 1.0, 0.05, 0.2, 0.0, "call", 4)) <= REL_CRR
    empty = op.crr_batch(np.empty(0), np.empty(0), np.empty(0), np.empty(0), np.empty(0), np.empty(0), np.empty(0), 10)
    assert empty.shape == (0,)


# ------------------------------------------------------------------------------ paths + payoffs
def _spec(op, key, K):
    from optpricer_b200 import exotics as ex
    f = key.split("|")
    if f[0] == "barrier":
        return ex.barrier_spec(K, f[2], float(f[3]), f[1], float(f[4]))
    if f[0] == "asian":
        return ex.asian_spec(K, f[1], f[2], f[3])
    if f[0] == "digital":
        return ex.digital_spec(K, f[1], float(f[2]))
    if f[0] == "lookback":
        return ex.lookback_spec(f[1], K, f[2])
    raise KeyError(key)


def _oracle_payoff(paths, key, K):
    f = key.split("|")
    if f[0] == "barrier":
        return payoffs.barrier_payoff(paths, K, f[2], float(f[3]), f[1], float(f[4]))
    if f[0] == "asian":
        return payoffs.asian_payoff(paths, K, f[1], f[2], f[3])
    if f[0] == "digital":
        return payoffs.digital_payoff(paths, K, f[1], float(f[2]))
    return payoffs.lookback_payoff(paths, f[1], K, f[2])


def _table(meta, npz):
    if meta.get("sigma") == "lvm_table":
        x0, h, dt = npz["lvm_table/grid"]
        return npz["lvm_table/table"], float(x0), float(h), float(dt)
    if meta.get("sigma") == "dupire_table_252x256":
        return dupire_grid_arrays()
    if meta.get("sigma") == "const0.2":
        n_steps = meta["args"]["n_steps"]
        return np.full((n_steps, 64), 0.2), math.log(meta["args"]["S0"]) - 1.6, 3.2 / 63, meta["args"]["T"] / n_steps
    return None


def _lazy(op, meta, npz, rng, mode="fp64", seed=None, n_mult=1):
    from optpricer_b200 import processes as pr
    a = dict(meta["args"])
    a["n_paths"] *= n_mult
    if seed is not None:
        a["seed"] = seed
    gen = meta["gen"]
    kw = dict(rng=rng, mode=mode)
    if gen == "gbm":
        return pr.gbm_paths(**a, **kw)
    if gen == "heston":
        return pr.heston_paths(**a, **kw)
    if gen == "gbm_milstein":
        return pr.gbm_milstein_paths(**a, **kw)
    if gen == "sabr":
        return pr.sabr_paths(**a, **kw)
    if gen == "merton":
        return pr.merton_jump_paths(**a, **kw)
    tab = pr.SigmaTable(*_table(meta, npz))
    if gen == "lv_milstein":
        return pr.milstein_local_vol_paths(sigma_loc=tab, dS_bump=meta.get("dS_bump", 0.01), **a, **kw)
    return pr.local_vol_paths(sigma_loc=tab, **a, **kw)


def _oracle_paths(meta, npz):
    a = dict(meta["args"])
    seed, n_steps, n_paths, anti = a.pop("seed"), a.pop("n_steps"), a.pop("n_paths"), a.pop("antithetic")
    gen = meta["gen"]
    if gen == "gbm":
        (z,) = path_schemes.draw_normals(seed, n_steps, n_paths)
        return path_schemes.gbm_exact(z, antithetic=anti, **a)
    if gen == "heston":
        z2, zp = path_schemes.draw_normals(seed, n_steps, n_paths, 2)
        return path_schemes.heston_fte(z2, zp, antithetic=anti, **a)
    if gen == "sabr":
        z2, zp = path_schemes.draw_normals(seed, n_steps, n_paths, 2)
        return path_schemes.sabr_lognormal_vol(z2, zp, antithetic=anti, **a)
    if gen == "merton":
        z, k, zj = path_schemes.merton_draws(seed, n_steps, n_paths, a["lam"], a["T"])
        return path_schemes.merton_jumps(z, k, zj, antithetic=anti, **a)
    (z,) = path_schemes.draw_normals(seed, n_steps, n_paths)
    if gen == "gbm_milstein":
        return path_schemes.gbm_milstein(z, antithetic=anti, **a)
    tab = sigma_grid.SigmaGrid(*_table(meta, npz))
    if gen == "lv_milstein":
        return path_schemes.lv_milstein(z, tab, antithetic=anti, dS_bump=meta.get("dS_bump", 0.01), **a)
    return path_schemes.lv_log_euler(z, tab, antithetic=anti, **a)


SMALL = ["gbm_small_anti1", "gbm_small_anti0", "heston_small_anti1", "heston_small_anti0", "heston_trunc",
         "lvm_const", "lvm_table_anti1", "lvm_table_anti0", "lvm_table_bump3", "lv_euler_table",
         "gbm_milstein", "sabr_beta1.0", "sabr_beta0.5", "merton"]


@pytest.mark.parametrize("tag", SMALL)
def test_check_mode_small_cases(op, tag):
    """Reference's own normals -> materialised paths, per-path payoffs and prices at 1e-12."""
    from optpricer_b200 import exotics as ex
    meta = load_json("mc_paths.json")[tag]
    npz = load_npz("mc_paths_samples.npz")
    lp = _lazy(op, meta, npz, rng="numpy")
    ref_paths = _oracle_paths(meta, npz)
    arr = np.asarray(lp)
    assert arr.shape == ref_paths.shape == lp.shape
    np.testing.assert_allclose(arr, ref_paths, rtol=REL_CHECK)
    np.testing.assert_allclose(arr[npz[f"{tag}/rows"], :], npz[f"{tag}/row_values"], rtol=REL_CHECK)
    if meta["gen"] == "heston" and f"{tag}/var_terminal" in npz:
        _, var = lp.materialize(return_variance=True)
        np.testing.assert_allclose(var[-1], npz[f"{tag}/var_terminal"], rtol=1e-11, atol=1e-15)
    K, r, T = meta["K"], meta["r"], meta["T"]
    keys = list(meta["prices"])
    for chunk in (keys[:12], keys[12:]):
        specs = [_spec(op, k, K) for k in chunk]
        stats, pay, term = ex.path_stats(lp, specs, r, T, want_payoffs=True, want_terminal=True)
        np.testing.assert_allclose(term, ref_paths[-1], rtol=REL_CHECK)
        for i, k in enumerate(chunk):
            want = _oracle_payoff(ref_paths, k, K)
            np.testing.assert_allclose(pay[i], want, rtol=REL_CHECK, atol=1e-12, err_msg=k)
            price, se = ex.estimate(stats[i])
            assert rel(price, meta["prices"][k][0]) <= REL_CHECK, (tag, k)
            assert rel(se, meta["prices"][k][1]) <= REL_SE, (tag, k)
            assert stats[i][0] == ref_paths.shape[1]
    # the array contract of exotics.py: reduce a caller-supplied ndarray on the device
    for k in (keys[0], keys[8], keys[13], keys[18], keys[23]):
        stats, pay = ex.array_stats(ref_paths, _spec(op, k, K), r, T, want_payoffs=True)
        np.testing.assert_allclose(pay, _oracle_payoff(ref_paths, k, K), rtol=REL_CHECK, atol=1e-10)
        assert rel(ex.estimate(stats)[0], meta["prices"][k][0]) <= REL_CHECK


BIG = ["gbm_test_exotics_500x100k", "heston_365x50k", "desk_table_252x100k"]


@pytest.mark.parametrize("tag", BIG)
def test_check_mode_reference_test_fixtures(op, tag):
    """The reference tests' own configurations (tests/test_exotics.py:22-28, desk workflow) at full size."""
    from optpricer_b200 import exotics as ex
    meta = load_json("mc_paths.json")[tag]
    lp = _lazy(op, meta, None, rng="numpy")
    K, r, T = meta["K"], meta["r"], meta["T"]
    keys = list(meta["prices"])
    got = []
    for chunk in (keys[:12], keys[12:]):
        got += ex.price_payoffs(lp, [_spec(op, k, K) for k in chunk], r, T)
    for k, (price, se) in zip(keys, got):
        assert rel(price, meta["prices"][k][0]) <= REL_CHECK, (tag, k, price)
        assert rel(se, meta["prices"][k][1]) <= REL_SE, (tag, k, se)
    if "uo130" in meta:
        p, se = op.barrier_price(lp, 100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
        assert rel(p, meta["uo130"][0]) <= REL_CHECK and rel(se, meta["uo130"][1]) <= REL_SE


@pytest.mark.parametrize("tag", BIG + ["gbm_small_anti0", "heston_small_anti0", "lvm_table_anti0", "lv_euler_table",
                                       "gbm_milstein", "merton", "sabr_beta1.0", "sabr_beta0.5"])
@pytest.mark.parametrize("mode", ["fp32", "fp64"])
def test_philox_within_three_reference_se(op, tag, mode):
    from optpricer_b200 import exotics as ex
    meta = load_json("mc_paths.json")[tag]
    npz = load_npz("mc_paths_samples.npz")
    K, r, T = meta["K"], meta["r"], meta["T"]
    keys = list(meta["prices"])
    # the device run uses 16x the reference's path count, so the difference is dominated by the
    # reference's own sampling error and the bar is literally "3 reference standard errors"
    for seed in (2024, 77):
        lp = _lazy(op, meta, npz, rng="philox", mode=mode, seed=seed, n_mult=16)
        got = []
        for chunk in (keys[:12], keys[12:]):
            got += ex.price_payoffs(lp, [_spec(op, k, K) for k in chunk], r, T)
        for k, (price, se) in zip(keys, got):
            ref_p, ref_se = meta["prices"][k]
            assert abs(price - ref_p) <= 3.0 * ref_se + 1e-9, (tag, mode, seed, k, price, ref_p, ref_se)
            if ref_se > 0 and lp.n_paths >= 16 * 50_000:
                assert 0.85 < 4.0 * se / ref_se < 1.18, (tag, k, se, ref_se)


@pytest.mark.parametrize("gen", ["gbm", "heston", "lvm", "lve", "gbmm", "merton", "sabr1", "sabr05", "sabr0"])
def test_fp32_tracks_fp64_on_same_philox_draws(op, gen):
    """Production fp32 arithmetic vs the reference operation order in fp64, SAME device normals:
    the price difference must be a small fraction of the standard error."""
    from optpricer_b200 import exotics as ex, processes as pr
    table = pr.SigmaTable(*dupire_grid_arrays())
    n = 200_000

    def make(mode):
        if gen == "gbm":
            return pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, n, seed=11, mode=mode)
        if gen == "heston":
            return pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, n, seed=11, mode=mode)
        if gen == "lvm":
            return pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, n, table, seed=11, mode=mode)
        if gen == "lve":
            return pr.local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, n, table, seed=11, mode=mode)
        if gen == "merton":
            return pr.merton_jump_paths(100.0, 0.05, 0.0, 0.2, 1.0, 100, n, lam=0.8, mJ=-0.1, sJ=0.15, seed=11, mode=mode)
        if gen.startswith("sabr"):   # alpha0 scaled so that the price volatility is ~20 % at S = 100 for every beta
            beta = {"sabr1": 1.0, "sabr05": 0.5, "sabr0": 0.0}[gen]
            return pr.sabr_paths(100.0, 0.05, 0.0, 0.2 * 100.0 ** (1.0 - beta), beta, 0.4, -0.3, 1.0, 200, n, seed=11, mode=mode)
        return pr.gbm_milstein_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, n, seed=11, mode=mode)

    specs = [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.barrier_spec(100.0, "put", 85.0, "down-and-in"),
             ex.asian_spec(100.0, "call"), ex.asian_spec(100.0, "put", "geometric", "floating"),
             ex.lookback_spec("call"), ex.lookback_spec("put", 100.0, "fixed"), ex.vanilla_spec(100.0, "call"),
             ex.digital_spec(100.0, "call")]
    a = ex.price_payoffs(make("fp32"), specs, 0.05, 1.0)
    b = ex.price_payoffs(make("fp64"), specs, 0.05, 1.0)
    for (pa, sa), (pb, sb) in zip(a, b):
        assert abs(pa - pb) <= 0.05 * sb, (gen, pa, pb, sb)
        assert rel(sa, sb) < 2e-3


def test_heston_fp32_scaled_normals_cover_negative_and_zero_vol_of_vol(op):
    """The fp32 Heston kernel takes its normals pre-scaled by |xi| sqrt(dt): xi < 0 runs as (-xi, -rho) -- the
    same process with Z2 relabelled -Z2, so it must reproduce the (-xi, -rho) run exactly and agree with the
    fp64 kernel (processes.py:153-160 literally, a different but equally valid sample) statistically; xi = 0
    keeps a representable scale and stays path-identical to the fp64 kernel."""
    from optpricer_b200 import exotics as ex, processes as pr
    n = 200_000
    specs = [ex.asian_spec(100.0, "call"), ex.barrier_spec(100.0, "put", 85.0, "down-and-in"), ex.vanilla_spec(100.0, "call")]

    def run(xi, rho, mode):
        lp = pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, xi, rho, 1.0, 100, n, seed=5, mode=mode)
        return ex.price_payoffs(lp, specs, 0.05, 1.0)

    neg, pos = run(-0.3, 0.7, "fp32"), run(0.3, -0.7, "fp32")
    for (a, sa), (b, sb) in zip(neg, pos):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)) and abs(sa - sb) <= 1e-6 * sb
    neg64 = run(-0.3, 0.7, "fp64")
    for (a, sa), (b, sb) in zip(neg, neg64):
        assert abs(a - b) <= 4.0 * math.hypot(sa, sb), (a, b, sa, sb)
        assert rel(sa, sb) < 2e-2
    zero, zero64 = run(0.0, -0.7, "fp32"), run(0.0, -0.7, "fp64")
    for (a, sa), (b, sb) in zip(zero, zero64):
        assert abs(a - b) <= 0.05 * sb and rel(sa, sb) < 2e-3


def test_size_independent_properties_at_scale(op):
    """BASELINE-scale path counts: identities that hold path by path, whatever the size."""
    from optpricer_b200 import exotics as ex, processes as pr
    n = 20_000_000
    lp = pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, n, seed=1)
    specs = [ex.barrier_spec(100.0, "call", 120.0, "up-and-out"), ex.barrier_spec(100.0, "call", 120.0, "up-and-in"),
             ex.vanilla_spec(100.0, "call"), ex.asian_spec(100.0, "call", "arithmetic"),
             ex.asian_spec(100.0, "call", "geometric"), ex.lookback_spec("call", 100.0, "fixed")]
    stats, _, _ = ex.path_stats(lp, specs, 0.05, 1.0)
    assert all(s[0] == 2 * n for s in stats)
    # knock-in + knock-out == vanilla (exotics tests: tests/test_exotics.py:35-49), here to fp64 rounding
    assert abs(stats[0][1] + stats[1][1] - stats[2][1]) <= 1e-9 * stats[2][1]
    assert stats[3][1] >= stats[4][1]                      # arithmetic >= geometric
    assert stats[5][1] >= stats[2][1]                      # fixed lookback call >= vanilla call
    bs_call = 10.450583572185565
    p, se = ex.estimate(stats[2])
    assert abs(p - bs_call) <= 4 * se + 1e-4
    # E[e^{-rT} S_T] = S0 (martingale check of the generator + reduction)
    assert abs(stats[2][3] / stats[2][0] - 100.0) < 4 * 100.0 * 0.2 / math.sqrt(2 * n) + 1e-4
    # sharding invariance: disjoint Philox path ranges add up to the single-device statistics
    a, _, _ = ex.path_stats(lp, specs, 0.05, 1.0, n_base=n // 2 + 3, path_offset=0)
    b, _, _ = ex.path_stats(lp, specs, 0.05, 1.0, n_base=n - (n // 2 + 3), path_offset=n // 2 + 3)
    np.testing.assert_allclose((a + b)[:, :6], stats[:, :6], rtol=1e-9)


def test_control_variate_on_path_payoff(op):
    from optpricer_b200 import exotics as ex, processes as pr
    meta = load_json("mc_paths.json")["heston_small_anti1"]
    lp = _lazy(op, meta, None, rng="numpy")
    ref_paths = _oracle_paths(meta, None)
    pay = payoffs.asian_payoff(ref_paths, 100.0, "call")
    want = payoffs.cv_estimator(payoffs.six_stats(pay, ref_paths[-1], 0.05, 1.0), 100.0)
    got = op.asian_price(lp, 100.0, 0.05, 1.0, "call", control_variate=True)
    assert rel(got[0], want[0]) <= 1e-11 and rel(got[1], want[1]) <= REL_SE
    plain = op.asian_price(lp, 100.0, 0.05, 1.0, "call")
    assert got[1] < plain[1]


def test_argument_errors_match_reference_texts(op):
    from optpricer_b200 import processes as pr
    with pytest.raises(ValueError, match="n_steps and n_paths must be positive."):
        pr.gbm_paths(100, 0.05, 0, 0.2, 1, 0, 10)
    with pytest.raises(ValueError, match=r"rho must be in \[-1, 1\]."):
        pr.heston_paths(100, 0.05, 0, 0.04, 2, 0.04, 0.3, -1.5, 1, 10, 10)
    lp = pr.gbm_paths(100, 0.05, 0, 0.2, 1, 10, 10, seed=1)
    with pytest.raises(ValueError, match="barrier_type must be one of"):
        op.barrier_price(lp, 100, 0.05, 1, "call", 120, "sideways")
    with pytest.raises(ValueError, match="kind must be 'call' or 'put'"):
        op.barrier_price(lp, 100, 0.05, 1, "both", 120, "up-and-out")
    with pytest.raises(ValueError, match="average_type must be 'arithmetic' or 'geometric'"):
        op.asian_price(lp, 100, 0.05, 1, "call", "harmonic")
    with pytest.raises(ValueError, match="strike_type must be 'fixed' or 'floating'"):
        op.asian_price(lp, 100, 0.05, 1, "call", "arithmetic", "odd")
    with pytest.raises(ValueError, match="strike_type must be 'floating' or 'fixed'"):
        op.lookback_price(lp, 0.05, 1, "call", 0.0, "odd")
    assert lp.shape == (11, 20) and np.asarray(lp).shape == (11, 20)
    assert np.all(np.asarray(lp)[0] == 100.0)


# ------------------------------------------------------------------------------ book pricing
def test_price_book_equals_row_by_row_api(op):
    """optp_mc_book (all items back to back, one copy) against the single-call public API on the same
    Philox keys: identical statistics, hence identical prices; per-row failures do not abort."""
    from optpricer_b200 import book, processes as pr
    rows = [
        dict(id="a", S0=100, K=110, T=0.5, r=0.05, sigma=0.20, q=0.0, kind="call", method="bs"),
        dict(id="b", S0=100, K=95, T=1.0, r=0.05, sigma=0.25, q=0.01, kind="put", method="mc"),
        dict(id="c", S0=100, K=105, T=0.5, r=0.05, sigma=0.20, q=0.0, kind="call", method="binomial"),
        dict(id="d", S0=100, K=100, T=1.0, r=0.05, sigma=0.20, q=0.0, kind="call", method="exotic",
             exotic_type="barrier", barrier=120, barrier_type="up-and-out", n_steps=100, n_paths=50_000),
        dict(id="e", S0=100, K=100, T=1.0, r=0.05, sigma=0.20, q=0.0, kind="put", method="exotic",
             exotic_type="asian", n_steps=100, n_paths=50_000),
        dict(id="f", S0=100, K=100, T=1.0, r=0.05, sigma=0.20, q=0.0, kind="put", method="binomial", american="true"),
        dict(id="g", S0=100, K=100, T=1.0, r=0.05, sigma=0.20, kind="call", method="fem"),
        dict(id="h", S0=100, K=100, T=2.0, r=0.03, sigma=0.30, q=0.0, kind="call", method="exotic",
             exotic_type="lookback", strike_type="fixed", n_steps=50, n_paths=20_000),
        dict(id="i", S0=100, K=100, T=2.0, r=0.03, sigma=0.30, q=0.0, kind="call", method="exotic",
             exotic_type="digital", payout=10.0, n_steps=50, n_paths=20_000),
    ]
    out = book.price_book(rows, mc_paths=200_000)
    assert [r["id"] for r in out] == list("abcdefghi")
    assert out[6]["price"] is None and "Unknown method" in out[6]["error"]
    assert out[0]["price"] == pytest.approx(2.9057, abs=2e-3)
    want_b = op.euro_price_mc(op.OptionSpec(100, 95, 1.0, 0.05, 0.25, 0.01), "put", n_paths=200_000, seed=1)
    assert out[1]["price"] == pytest.approx(want_b[0], rel=1e-12) and out[1]["stderr"] == pytest.approx(want_b[1], rel=1e-9)
    assert out[2]["price"] == pytest.approx(op.crr(op.OptionSpec(100, 105, 0.5, 0.05, 0.2), "call", 500), rel=1e-12)
    assert out[5]["price"] == pytest.approx(op.crr(op.OptionSpec(100, 100, 1.0, 0.05, 0.2), "put", 500, american=True),
                                            rel=1e-12)
    lp = pr.gbm_paths(100, 0.05, 0.0, 0.2, 1.0, 100, 50_000, seed=1)
    assert out[3]["price"] == pytest.approx(op.barrier_price(lp, 100, 0.05, 1.0, "call", 120, "up-and-out")[0], rel=1e-12)
    assert out[4]["price"] == pytest.approx(op.asian_price(lp, 100, 0.05, 1.0, "put")[0], rel=1e-12)
    lp2 = pr.gbm_paths(100, 0.03, 0.0, 0.3, 2.0, 50, 20_000, seed=1)
    assert out[7]["price"] == pytest.approx(op.lookback_price(lp2, 0.03, 2.0, "call", 100, "fixed")[0], rel=1e-12)
    assert out[8]["price"] == pytest.approx(op.digital_price(lp2, 100, 0.03, 2.0, "call", 10.0)[0], rel=1e-12)
    assert out[8]["stderr"] == pytest.approx(op.digital_price(lp2, 100, 0.03, 2.0, "call", 10.0)[1], rel=1e-9)


def test_synthetic_book_prices_are_sane(op):
    from optpricer_b200 import book
    rows = book.synthetic_book(300, n_paths=20_000, exotic_steps=64)
    out = book.price_book(rows, mc_paths=20_000)
    assert all(r["price"] is not None and r["price"] >= 0.0 for r in out)
    for row, res in zip(rows, out):
        if row["method"] == "mc":
            bs = book._bs_price(row["S0"], row["K"], row["T"], row["r"], row["q"], row["sigma"], row["kind"])
            assert abs(res["price"] - bs) <= 5 * res["stderr"] + 1e-3


# ------------------------------------------------------------------------------ Heston QE (unpinned scheme)
QE_CASES = [dict(S0=100.0, r=0.05, q=0.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, T=1.0),
            dict(S0=100.0, r=0.02, q=0.01, v0=0.09, kappa=1.5, theta=0.04, xi=0.9, rho=-0.5, T=0.5)]


@pytest.mark.parametrize("case", [0, 1])
def test_heston_qe_check_mode_matches_oracle(op, case):
    """No reference code exists for QE; the fp64 kernel is checked against the numpy restatement of
    Andersen (2008) on the same normals (case 1 violates Feller: the exponential branch is active)."""
    from optpricer_b200 import exotics as ex, processes as pr
    c = QE_CASES[case]
    lp = pr.heston_paths(n_steps=24, n_paths=4000, seed=5, rng="numpy", scheme="qe", **c)
    z2, zp = path_schemes.draw_normals(5, 24, 4000, 2)
    ref, ref_v = path_schemes.heston_qe(z2, zp, return_variance=True, **c)
    got, got_v = lp.materialize(return_variance=True)
    np.testing.assert_allclose(got, ref, rtol=1e-11)
    np.testing.assert_allclose(got_v, ref_v, rtol=1e-9, atol=1e-14)
    if case == 1:
        assert (ref_v == 0.0).any()          # the exponential branch produced exact zeros
    specs = [ex.asian_spec(100.0, "call"), ex.barrier_spec(100.0, "put", 85.0, "down-and-in"), ex.vanilla_spec(100.0, "call")]
    stats, pay, _ = ex.path_stats(lp, specs, c["r"], c["T"], want_payoffs=True)
    np.testing.assert_allclose(pay[0], payoffs.asian_payoff(ref, 100.0, "call"), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(pay[1], payoffs.barrier_payoff(ref, 100.0, "put", 85.0, "down-and-in"), rtol=1e-10, atol=1e-10)
    want = payoffs.discounted_mean_se(payoffs.vanilla_payoff(ref, 100.0, "call"), c["r"], c["T"])
    assert rel(ex.estimate(stats[2])[0], want[0]) <= 1e-11


@pytest.mark.parametrize("case", [0, 1])
def test_heston_qe_production_kernel(op, case):
    """fp32 QE kernel: tracks the fp64 kernel on the same Philox draws, and prices a vanilla call
    within 4 SE (+ the scheme's discretisation bias allowance) of the semi-analytic Heston price."""
    from optpricer_b200 import exotics as ex, processes as pr
    from test_oracle_cpu import heston_call_semi_analytic
    c = QE_CASES[case]
    specs = [ex.vanilla_spec(100.0, "call"), ex.asian_spec(100.0, "call"), ex.lookback_spec("put")]
    a = ex.price_payoffs(pr.heston_paths(n_steps=32, n_paths=400_000, seed=3, scheme="qe", mode="fp32", **c), specs, c["r"], c["T"])
    b = ex.price_payoffs(pr.heston_paths(n_steps=32, n_paths=400_000, seed=3, scheme="qe", mode="fp64", **c), specs, c["r"], c["T"])
    for (pa, sa), (pb, sb) in zip(a, b):
        assert abs(pa - pb) <= 0.1 * sb, (pa, pb, sb)
    exact = heston_call_semi_analytic(c["S0"], 100.0, c["T"], c["r"], c["q"], c["v0"], c["kappa"], c["theta"], c["xi"], c["rho"])
    big = ex.price_payoffs(pr.heston_paths(n_steps=32, n_paths=4_000_000, seed=8, scheme="qe", **c), specs[:1], c["r"], c["T"])[0]
    assert abs(big[0] - exact) <= 4.0 * big[1] + 0.02, (big, exact)


# ------------------------------------------------------------------------------ BASELINE full sizes
def test_full_size_config2_crr_batch(op):
    """BASELINE config 2 at full size: 1 000 000 American puts, N = 1000.  Size-independent
    properties + the reference fixture for the head of the batch."""
    g = load_json("crr.json")
    S0, K, T, r, q, sig, is_call = crr_tree.cfg2_contracts(1_000_000)
    price, status = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=True, return_status=True)
    assert price.shape == (1_000_000,) and not status.any() and np.isfinite(price).all()
    np.testing.assert_allclose(price[:64], np.array(g["cfg2_first64_american_put_N1000"]), rtol=REL_CRR, atol=REL_CRR)
    assert np.all(price >= np.maximum(K - S0, 0.0) - 1e-12)          # never below intrinsic
    assert np.all(price <= K + 1e-12)                                 # a put is worth at most K
    euro = op.crr_batch(S0, K, T, r, q, sig, is_call, 1000, american=False)
    assert np.all(price >= euro - 1e-10)                               # early exercise never hurts
    # permutation invariance: a contract's price does not depend on where it sits in the batch
    perm = np.random.default_rng(1).permutation(1_000_000)
    again = op.crr_batch(S0[perm], K[perm], T[perm], r[perm], q[perm], sig[perm], is_call[perm], 1000, american=True)
    np.testing.assert_array_equal(again, price[perm])


def test_full_size_config3_local_vol_barrier(op):
    """BASELINE config 3: LV Milstein on the Dupire table, 252 steps, 1e8 effective paths."""
    from optpricer_b200 import exotics as ex, processes as pr
    tab = pr.SigmaTable(*dupire_grid_arrays())
    n = 50_000_000
    lp = pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, n, tab, seed=42)
    specs = [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.barrier_spec(100.0, "call", 130.0, "up-and-in"),
             ex.vanilla_spec(100.0, "call")]
    stats, _, _ = ex.path_stats(lp, specs, 0.05, 1.0)
    assert all(s[0] == 2 * n for s in stats)
    assert abs(stats[0][1] + stats[1][1] - stats[2][1]) <= 1e-9 * stats[2][1]        # in + out == vanilla, path by path
    fwd_disc = 100.0 * math.exp(-0.02)                                               # E[e^{-rT} S_T] = S0 e^{-qT}
    my = stats[2][3] / stats[2][0]
    sy = math.sqrt(max(stats[2][4] / stats[2][0] - my * my, 0.0) / stats[2][0])
    assert abs(my - fwd_disc) <= 4 * sy + 2e-4 * fwd_disc    # + the Milstein scheme's weak error at 252 steps
    want = load_json("mc_paths.json")["desk_table_252x100k"]["uo130"]
    p, se = ex.estimate(stats[0])
    assert se < 1e-3 and abs(p - want[0]) <= 3 * want[1]
    # sharding over 8 ranks reproduces the single-device statistics
    parts = sum(ex.path_stats(lp, specs[:1], 0.05, 1.0, n_base=hi - lo, path_offset=lo)[0]
                for lo, hi in [(k * n // 8, (k + 1) * n // 8) for k in range(8)])
    np.testing.assert_allclose(parts[0][:6], stats[0][:6], rtol=1e-9)


def test_full_size_config4_heston_asian(op):
    """BASELINE config 4: Heston full-truncation Euler, 365 steps, 1e9 effective paths, arithmetic
    Asian call with antithetic variates and the reference's control variate."""
    from optpricer_b200 import exotics as ex, processes as pr
    n = 500_000_000
    lp = pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, n, seed=42)
    stats, _, _ = ex.path_stats(lp, [ex.asian_spec(100.0, "call"), ex.asian_spec(100.0, "call", "geometric")], 0.05, 1.0)
    assert stats[0][0] == 2 * n == 1_000_000_000
    plain = ex.estimate(stats[0])
    cv = ex.estimate(stats[0], control_variate=True, expected_y=100.0)
    want = load_json("mc_paths.json")["heston_365x50k"]["prices"]["asian|call|arithmetic|fixed"]
    assert abs(plain[0] - want[0]) <= 3 * want[1] and abs(cv[0] - want[0]) <= 3 * want[1]
    assert cv[1] < plain[1] < 3e-4                                    # the control variate helps
    assert stats[0][1] >= stats[1][1]                                 # arithmetic >= geometric
    assert abs(stats[0][3] / stats[0][0] - 100.0) < 5e-3              # martingale: E[e^{-rT} S_T] = S0


def test_full_size_config5_book(op):
    """BASELINE config 5 at full size on ONE device: 10 000 contracts (European / barrier / arithmetic
    Asian, price_book.py row schema), 1e7 effective paths each.  European rows against the closed form,
    every row against a bounded re-run with another seed, plus the orderings the payoffs guarantee."""
    from optpricer_b200 import book as bk
    rows = bk.synthetic_book(10_000, n_paths=5_000_000)
    out = bk.price_book(rows, mc_paths=5_000_000)
    assert len(out) == 10_000 and all(r.get("price") is not None and math.isfinite(r["price"]) for r in out)
    n_euro = 0
    for row, res in zip(rows, out):
        if row["method"] == "mc":
            want = bk._bs_price(row["S0"], row["K"], row["T"], row["r"], row["q"], row["sigma"], row["kind"])
            assert abs(res["price"] - want) <= 5 * res["stderr"] + 1e-4, (row, res, want)
            n_euro += 1
        else:
            vanilla = bk._bs_price(row["S0"], row["K"], row["T"], row["r"], row["q"], row["sigma"], row["kind"])
            if row["exotic_type"] == "barrier":                      # knock-out <= vanilla
                assert -1e-9 <= res["price"] <= vanilla + 5 * res["stderr"] + 1e-3, (row, res, vanilla)
            assert res["stderr"] < 0.02
    assert 3000 < n_euro < 3700
    # an independent, smaller run (other Philox key) of a sample of rows agrees within combined errors
    pick = list(range(0, 10_000, 500))
    again = bk.price_book([dict(rows[i], n_paths=200_000) for i in pick], mc_paths=200_000, seed=77)
    for i, b in zip(pick, again):
        a = out[i]
        assert abs(a["price"] - b["price"]) <= 5 * math.hypot(a["stderr"], b["stderr"]) + 1e-4, (rows[i], a, b)


# ------------------------------------------------------------------------------ device Dupire tabulation
class _Slice:
    def __init__(self, a, b, rho, m, sigma):
        self.a, self.b, self.rho, self.m, self.sigma = a, b, rho, m, sigma


class _Surface:
    """Duck-typed stand-in for the reference's VolSurface (calibration.py:72-148): .slices and the
    private forward curve, built from the fitted parameters stored by make_golden.py."""
    def __init__(self, js, with_forwards=True):
        self.slices = {float(T): _Slice(*p) for T, p in js["svi"].items()}
        self._forward_curve = {float(T): F for T, F in js["forwards"].items()} if with_forwards else {}


def test_dupire_table_on_device_matches_reference_and_oracle(op):
    """optp_dupire_table against (a) the 252 x 256 table the UNMODIFIED reference's callable produced
    (tests/golden/dupire_table_252x256.npz) and (b) the oracle restatement on other grids, including
    the reference's no-forward-curve fallback.  The central difference in T amplifies last-ulp
    differences of exp/log by ~1e4, hence 1e-9 relative."""
    from optpricer_b200 import processes as pr
    from oracle import dupire
    js = load_json("dupire_surface.json")
    table, x0, h, dt = dupire_grid_arrays()
    got = pr.SigmaTable.from_svi_surface(_Surface(js), S0=js["S0"], T=1.0, n_steps=252)
    assert got.table.shape == table.shape and got.x0 == x0 and got.h == h and got.dt == dt
    np.testing.assert_allclose(got.table, table, rtol=1e-9, atol=0)
    slices = {float(T): tuple(p) for T, p in js["svi"].items()}
    fwds = {float(T): F for T, F in js["forwards"].items()}
    for (T, n_steps, n_nodes, hw, with_fwd) in ((0.6, 37, 64, 1.0, True), (2.5, 50, 33, 2.0, True), (1.0, 16, 40, 1.2, False)):
        dev = pr.SigmaTable.from_svi_surface(_Surface(js, with_fwd), S0=95.0, T=T, n_steps=n_steps, n_nodes=n_nodes,
                                             half_width=hw)
        nodes = np.exp(dev.x0 + dev.h * np.arange(n_nodes))
        surf = dupire.Surface(slices, fwds if with_fwd else {1.0: float(nodes.mean())})
        want = sigma_grid.tabulate(surf.sigma_loc(), S0=95.0, T=T, n_steps=n_steps, n_nodes=n_nodes, half_width=hw)
        np.testing.assert_allclose(dev.table, want.table, rtol=1e-9, atol=0)
    # the device-tabulated surface prices like the host-tabulated one (same Philox draws)
    host = pr.SigmaTable(table, x0, h, dt)
    a = op.barrier_price(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 200_000, got, seed=5),
                         100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
    b = op.barrier_price(pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, 200_000, host, seed=5),
                         100.0, 0.05, 1.0, "call", 130.0, "up-and-out")
    assert abs(a[0] - b[0]) <= 1e-6 * b[0]


# ------------------------------------------------------------------------------ host threads
def test_concurrent_host_threads_get_the_serial_results(op):
    """ctypes releases the GIL around a call and the library owns one context: calls from several
    Python threads must be serialised by the library (include/optp_b200.h, "threads"), never
    interleaved on the shared scratch."""
    from concurrent.futures import ThreadPoolExecutor
    opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
    g = np.random.default_rng(11)
    n = 3000
    S0 = g.uniform(50, 150, n)
    K = S0 * g.uniform(0.8, 1.2, n)
    T = g.uniform(0.1, 2.0, n)
    sig = g.uniform(0.1, 0.6, n)

    def mc(seed):
        return op.euro_price_mc(opt, "call", n_paths=300_000, seed=seed)

    def path(seed):
        p = op.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 64, 200_000, seed=seed)
        return op.barrier_price(p, 100.0, 0.05, 1.0, "call", 130.0, "up-and-out")

    def tree(lo):
        s = slice(lo, lo + 1000)
        return op.crr_batch(S0[s], K[s], T[s], 0.03, 0.0, sig[s], 0, N=300, american=True)

    jobs = [(mc, 1), (path, 2), (tree, 0), (mc, 3), (path, 4), (tree, 1000), (mc, 5), (tree, 2000), (path, 6)] * 3
    serial = [f(x) for f, x in jobs]
    with ThreadPoolExecutor(max_workers=6) as ex:
        futs = [ex.submit(f, x) for f, x in jobs]
        threaded = [f.result() for f in futs]
    for a, b in zip(serial, threaded):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
