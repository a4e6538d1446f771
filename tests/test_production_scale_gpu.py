"""Production (fp32 arithmetic, on-device Philox) path pinned at BASELINE path counts, in the product's
OWN standard errors (run with -m gpu on the B200 box).

The reference computes in float64 (processes.py:48-55, monte_carlo.py:35-42) on PCG64 draws; the
production kernels compute the path state in fp32 with MUFU approximations on Philox/Box-Muller normals.
The small-size tests bound the difference by the REFERENCE's standard error (50k-100k paths); at
BASELINE sizes (1e8-1e9 effective paths) the product's own standard error is 30-100x smaller, so:

  (a) fp32 vs fp64 kernels on the SAME Philox draws at >= 1e8 effective paths: the price may move by at
      most half of the run's own standard error, the standard error itself by 1e-3 relative;
  (b) the device normals themselves at 2^32 (and more) draws: moments, tail mass, pair and serial
      products against the exact N(0,1) values, in sampling standard errors -- for the default generator
      mode and for the two opt-in throughput modes;
  (c) tests/test_generator_cpu.py holds the Philox known-answer test (no device needed);
  (d) tests/test_multigpu_gpu.py holds the 2-rank NCCL run.

Every run appends its numbers to gpurun_out/scale_parity.jsonl when that directory exists (kept under
profiles/ per round)."""
import json
import math
import os

import numpy as np
import pytest

from golden_io import dupire_grid_arrays

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def op():
    import optpricer_b200 as o
    from optpricer_b200 import _native
    _native.init()
    return o


def _log(rec):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "scale_parity.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")


# ------------------------------------------------------------------ (a) fp32 vs fp64, same draws, 1e8 paths
N_BASE = 50_000_000   # x2 antithetic mirrors = 1e8 effective paths (BASELINE cfg 3 size; cfg 5 legs are 1e7)


def _cases(op):
    from optpricer_b200 import exotics as ex, processes as pr
    tab = pr.SigmaTable(*dupire_grid_arrays())
    return {
        # cfg 5 leg: GBM, 252 steps, up-and-out barrier call (+ the Asian leg and the vanilla on the same paths)
        "gbm_barrier": (lambda mode: pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, N_BASE, seed=20261018, mode=mode),
                        [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.asian_spec(100.0, "call"),
                         ex.vanilla_spec(100.0, "call"), ex.barrier_spec(100.0, "put", 85.0, "down-and-out")],
                        0.05, 1.0, 100.0),
        # cfg 3: LV Milstein on the Dupire table, 252 steps, up-and-out barrier call B = 130
        "lvm_barrier": (lambda mode: pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, N_BASE, tab, seed=42,
                                                                 mode=mode),
                        [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.vanilla_spec(100.0, "call")],
                        0.05, 1.0, 100.0 * math.exp(-0.02)),
        # cfg 4: Heston FTE, 365 steps, arithmetic Asian call, antithetic + control variate
        "heston_asian": (lambda mode: pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, N_BASE,
                                                      seed=42, mode=mode),
                         [ex.asian_spec(100.0, "call"), ex.asian_spec(100.0, "put", "geometric"),
                          ex.vanilla_spec(100.0, "call")],
                         0.05, 1.0, 100.0),
    }


@pytest.mark.parametrize("case", ["gbm_barrier", "lvm_barrier", "heston_asian"])
def test_fp32_tracks_fp64_at_baseline_scale(op, case):
    from optpricer_b200 import exotics as ex
    make, specs, r, T, ey = _cases(op)[case]
    s32, _, _ = ex.path_stats(make("fp32"), specs, r, T)
    s64, _, _ = ex.path_stats(make("fp64"), specs, r, T)
    assert s32[0][0] == s64[0][0] == 2 * N_BASE
    rec = {"test": "fp32_vs_fp64_same_draws", "case": case, "n_eff": 2 * N_BASE, "payoffs": []}
    worst = 0.0
    for i in range(len(specs)):
        for cv in (False, True):
            p32, e32 = ex.estimate(s32[i], control_variate=cv, expected_y=ey)
            p64, e64 = ex.estimate(s64[i], control_variate=cv, expected_y=ey)
            rec["payoffs"].append({"i": i, "cv": cv, "fp32": [p32, e32], "fp64": [p64, e64],
                                   "dprice_in_se": (p32 - p64) / e64, "dse_rel": e32 / e64 - 1.0})
            worst = max(worst, abs(p32 - p64) / e64)
    # the martingale E[e^{-rT} S_T] of both arithmetics, in the run's own standard error of Y
    for tag, s in (("fp32", s32[0]), ("fp64", s64[0])):
        n, sy, syy = s[0], s[3], s[4]
        my = sy / n
        rec["mean_Y_" + tag] = my
        rec["se_Y_" + tag] = math.sqrt(max(syy / n - my * my, 0.0) / n)
    rec["worst_dprice_in_se"] = worst
    _log(rec)
    for e in rec["payoffs"]:
        assert abs(e["dprice_in_se"]) <= 0.5, (case, e)
        assert abs(e["dse_rel"]) <= 1e-3, (case, e)
    # fp32 E[Y] agrees with fp64 E[Y] far inside the sampling error of Y (same draws)
    assert abs(rec["mean_Y_fp32"] - rec["mean_Y_fp64"]) <= 0.25 * rec["se_Y_fp64"], rec


# ------------------------------------------------------------------ (b) the device normals
def _tail(k):
    return math.erfc(k / math.sqrt(2.0))   # P(|Z| > k)


@pytest.mark.parametrize("rng_mode", [0, 1, 2])
def test_device_normals_quality(op, rng_mode):
    """2^30 Philox calls = 2^32 normals (6.4e9 in the six-normal mode), drawn exactly as the path kernels draw
    them: counter (path, block), 63 blocks = 252 steps per path."""
    from optpricer_b200 import _native as nv
    n_blocks = 64
    n_paths = 1 << 24
    s = nv.normal_stats(20261018, n_paths, n_blocks, rng_mode=rng_mode)
    n = s[0]
    per_call = 6 if rng_mode == 2 else 4
    assert n == n_paths * n_blocks * per_call >= 2**32
    mean, m2, m3, m4 = s[1] / n, s[2] / n, s[3] / n, s[4] / n
    rec = {"test": "device_normals", "rng_mode": rng_mode, "n": n, "mean": mean, "var": m2, "m3": m3, "m4": m4,
           "mean_in_se": mean * math.sqrt(n), "var_in_se": (m2 - 1.0) / math.sqrt(2.0 / n),
           "m3_in_se": m3 / math.sqrt(15.0 / n), "m4_in_se": (m4 - 3.0) / math.sqrt(96.0 / n),
           "tails": {}, "max_abs": s[12]}
    for k, cnt in zip((3, 4, 5, 6), s[5:9]):
        want = n * _tail(k)
        rec["tails"][str(k)] = {"count": cnt, "expected": want, "in_sd": (cnt - want) / math.sqrt(want)}
    rec["pair_in_se"] = s[9] / math.sqrt(s[13])        # Var(r^2 sin cos) = E[r^4]/8 = 1
    rec["cross_in_se"] = s[10] / math.sqrt(s[14])      # independent pairs: Var = 1
    rec["lag_in_se"] = s[11] / math.sqrt(s[15])
    _log(rec)
    assert abs(rec["mean_in_se"]) <= 4.0, rec
    assert abs(rec["var_in_se"]) <= 4.0, rec
    assert abs(rec["m3_in_se"]) <= 4.0, rec
    assert abs(rec["m4_in_se"]) <= 4.0, rec
    assert abs(rec["tails"]["3"]["in_sd"]) <= 4.5 and abs(rec["tails"]["4"]["in_sd"]) <= 4.5, rec
    assert abs(rec["tails"]["5"]["in_sd"]) <= 4.5, rec
    if rng_mode == 2:
        # 26-bit radii end at sqrt(2*27*ln 2) = 6.12 sigma: the mass beyond 6 sigma is (knowingly) thinner
        assert s[12] <= 6.13 and rec["tails"]["6"]["count"] <= rec["tails"]["6"]["expected"] + 4.5 * math.sqrt(
            rec["tails"]["6"]["expected"]), rec
    else:
        assert abs(rec["tails"]["6"]["in_sd"]) <= 4.5, rec
        assert 6.0 < s[12] <= 6.8, rec            # 32-bit radii end at sqrt(2*33*ln 2) = 6.76 sigma
    assert abs(rec["pair_in_se"]) <= 4.0 and abs(rec["cross_in_se"]) <= 4.0 and abs(rec["lag_in_se"]) <= 4.0, rec


def test_device_normals_are_a_function_of_seed_path_block_only(op):
    """Disjoint path ranges add up to the whole (the multi-GPU sharding contract), another seed is another
    sample."""
    from optpricer_b200 import _native as nv
    whole = nv.normal_stats(7, 1 << 16, 8)
    a = nv.normal_stats(7, 40_000, 8)
    b = nv.normal_stats(7, (1 << 16) - 40_000, 8, path_offset=40_000)
    np.testing.assert_allclose((a + b)[:12], whole[:12], rtol=1e-12, atol=1e-7)
    assert max(a[12], b[12]) == whole[12]
    other = nv.normal_stats(8, 1 << 16, 8)
    assert abs(other[1] - whole[1]) > 1.0


@pytest.mark.parametrize("rng_mode", [1, 2])
def test_opt_in_generator_modes_price_within_their_own_standard_error(op, rng_mode):
    """rng_mode 1 / 2 (fp32 GBM and Heston FTE path kernels) are different samples of the same law: prices agree with
    the default generator within the combined standard error at 2e7 effective paths, and the default is restored."""
    from optpricer_b200 import _native as nv, exotics as ex, processes as pr
    specs = [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.asian_spec(100.0, "call"),
             ex.vanilla_spec(100.0, "call"), ex.lookback_spec("put", 100.0, "fixed")]
    jobs = [pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, n_steps, 10_000_000, seed=3) for n_steps in (252, 50, 7)]
    jobs += [pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, n_steps, 10_000_000, seed=4)
             for n_steps in (365, 50, 7)]          # 7: a partial last call in every mode
    for lp in jobs:
        base = ex.price_payoffs(lp, specs, 0.05, 1.0)
        nv.set_option("rng_mode", rng_mode)
        try:
            alt = ex.price_payoffs(lp, specs, 0.05, 1.0)
        finally:
            nv.set_option("rng_mode", 0)
        again = ex.price_payoffs(lp, specs, 0.05, 1.0)
        assert again == base
        for (pa, sa), (pb, sb) in zip(alt, base):
            assert pa != pb
            assert abs(pa - pb) <= 4.0 * math.hypot(sa, sb), (rng_mode, lp, pa, pb, sa, sb)
            assert abs(sa / sb - 1.0) < 5e-3
