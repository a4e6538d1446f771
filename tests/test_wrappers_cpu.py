"""CPU tests of the array plumbing in front of the C ABI (optpricer_b200/binomial.py, risk.py): a stand-in for the
library records what optp_crr_batch is handed -- column order, broadcasting, the flag column, output shape -- so the
wrappers are covered without a device.  (The stand-in lives in this test; the product has no CPU path.)"""
import ctypes as C

import numpy as np
import pytest

from oracle import risk_callers


class _RecordingLib:
    """optp_crr_batch that reads its columns back through the pointers and 'prices' S0 + 2 K + 1000 flag."""

    def __init__(self):
        self.calls = []

    def optp_crr_batch(self, S0, K, T, r, q, sigma, is_call, n, N, american, price, status):
        col = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).copy()   # noqa: E731
        cols = [col(p) for p in (S0, K, T, r, q, sigma)]
        flag = np.ctypeslib.as_array(C.cast(is_call, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        self.calls.append(dict(cols=cols, flag=flag, n=n, N=N, american=american))
        out = np.ctypeslib.as_array(C.cast(price, C.POINTER(C.c_double)), shape=(n,))
        out[:] = cols[0] + 2.0 * cols[1] + 1000.0 * flag
        return 0


@pytest.fixture()
def lib(monkeypatch):
    from optpricer_b200 import _native as nv
    fake = _RecordingLib()
    monkeypatch.setattr(nv, "init", lambda device=None: fake)
    return fake


def test_crr_batch_broadcasts_columns_and_keeps_the_shape(lib):
    from optpricer_b200 import binomial as bn
    K = np.array([[90.0, 100.0, 110.0], [95.0, 105.0, 115.0]])
    out = bn.crr_batch(100.0, K, 1.0, 0.03, 0.0, 0.2, 0, 250, american=True)
    call = lib.calls[-1]
    assert out.shape == K.shape and call["n"] == 6 and call["N"] == 250 and call["american"] == 1
    np.testing.assert_array_equal(call["cols"][0], np.full(6, 100.0))
    np.testing.assert_array_equal(call["cols"][1], K.reshape(-1))
    for c, v in zip(call["cols"][2:], (1.0, 0.03, 0.0, 0.2)):
        np.testing.assert_array_equal(c, np.full(6, v))
    np.testing.assert_array_equal(call["flag"], np.zeros(6, np.uint8))
    np.testing.assert_array_equal(out, 100.0 + 2.0 * K)


@pytest.mark.parametrize("flags", [np.array([True, False, True]), np.array([1, 0, 2], dtype=np.int64),
                                   np.array([1.0, 0.0, 0.5]), np.array([1, 0, 255], dtype=np.uint8), [1, 0, 1]])
def test_crr_batch_flag_column_is_nonzero_means_call(lib, flags):
    from optpricer_b200 import binomial as bn
    out, st = bn.crr_batch([100.0, 101.0, 102.0], 100.0, 1.0, 0.03, 0.0, 0.2, flags, 100, return_status=True)
    call = lib.calls[-1]
    assert call["flag"].dtype == np.uint8
    np.testing.assert_array_equal(call["flag"], [1, 0, 1])
    np.testing.assert_array_equal(out, np.array([100.0, 101.0, 102.0]) + 200.0 + 1000.0 * np.array([1, 0, 1]))
    assert st.shape == (3,) and not st.any()


def test_crr_batch_strided_and_non_float_inputs_are_made_contiguous(lib):
    from optpricer_b200 import binomial as bn
    S = np.arange(20, dtype=np.int32)[::2] + 90           # strided view of another dtype
    out = bn.crr_batch(S, np.float32(100.0), 1.0, 0.03, 0.0, 0.2, np.ones(10, bool)[::1], 64)
    np.testing.assert_array_equal(lib.calls[-1]["cols"][0], S.astype(float))
    np.testing.assert_array_equal(out, S + 200.0 + 1000.0)
    assert bn.crr_batch(np.empty(0), 100.0, 1.0, 0.03, 0.0, 0.2, 1, 64).shape == (0,) and len(lib.calls) == 1


def test_crr_vec_kinds(lib):
    from optpricer_b200 import binomial as bn
    K = np.array([90.0, 100.0, 110.0, 120.0])
    bn.crr_vec(100.0, K, 1.0, 0.03, 0.0, 0.2, np.array(["call", "put", "Call", "call"]), N=50)
    np.testing.assert_array_equal(lib.calls[-1]["flag"], [1, 0, 0, 1])       # anything but "call" is a put
    bn.crr_vec(100.0, K, 1.0, 0.03, 0.0, 0.2, "call", N=50)
    np.testing.assert_array_equal(lib.calls[-1]["flag"], [1, 1, 1, 1])
    out = bn.crr_vec(100.0, 95.0, 1.0, 0.03, 0.0, 0.2, "put", N=50)
    assert out.shape == (1,)                                                  # atleast_1d(K), binomial.py:44-109


def test_portfolio_risk_hands_one_batch_with_flags_and_sums_like_the_reference_loop(lib):
    from optpricer_b200 import risk as rk
    g = np.random.default_rng(3)
    inst = [dict(S=float(s), K=float(k), T=float(t), r=0.03, q=0.01, sigma=float(v), kind=("call" if i % 3 == 0 else "put"),
                 position=float(p))
            for i, (s, k, t, v, p) in enumerate(zip(g.uniform(50, 150, 40), g.uniform(50, 150, 40), g.uniform(0.1, 2, 40),
                                                    g.uniform(0.1, 0.6, 40), g.integers(-9, 9, 40)))]
    pricer = rk.TreePricer(100, american=True)
    out = rk.portfolio_risk(inst, pricer)
    assert len(lib.calls) == 1 and lib.calls[0]["n"] == 40 * 8
    np.testing.assert_array_equal(lib.calls[0]["flag"], np.repeat([1 if i % 3 == 0 else 0 for i in range(40)], 8))
    # the same numbers through the oracle's restatement of risk.py:146-203 driven by the same stand-in pricer
    def scalar_pricer(S, K, T, r, q, sigma, kind):
        return S + 2.0 * K + (1000.0 if kind == "call" else 0.0)
    ref = risk_callers.portfolio_risk(inst, scalar_pricer)
    for key in ("total_delta", "total_gamma", "total_vega", "total_theta", "total_rho", "total_value"):
        assert out[key] == ref[key], key                                      # bit for bit: same order of additions
    assert list(out["instrument_greeks"][0]) == list(ref["instrument_greeks"][0])
    for a, b in zip(out["instrument_greeks"], ref["instrument_greeks"]):
        assert a == b


def test_python_constants_mirror_the_header():
    """Every OPTP_PROC_* / OPTP_PAY_* / OPTP_F_* / OPTP_MODE_* value of include/optp_b200.h against the constants the
    wrappers pass (optpricer_b200/_native.py): a renumbering on one side only would still load and run."""
    import os
    import re
    from optpricer_b200 import _native as nv
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "optp_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define OPTP_([A-Z0-9_]+) (\d+)u?\b", text)}
    assert defs["ABI_VERSION"] == nv.ABI_VERSION and defs["NSTATS"] == nv.NSTATS
    seen = 0
    for name, value in defs.items():
        if name.split("_")[0] in ("PROC", "PAY", "F", "MODE"):
            assert getattr(nv, name) == value, name
            seen += 1
    assert seen == 9 + 5 + 5 + 3


def test_payoff_descriptors_follow_the_reference_arguments():
    """exotics.py:33-96, 102-163, 169-209, 215-273: every argument combination of the four payoff pricers lands in the
    kind / flags / level fields the kernels read (mc_kernels.cuh eval_payoff), and bad arguments raise before any call."""
    from optpricer_b200 import _native as nv, exotics as ex
    for kind, kf in (("call", nv.F_CALL), ("put", 0)):
        for bt in ("up-and-out", "up-and-in", "down-and-out", "down-and-in"):
            s = ex.barrier_spec(100.0, kind, 120.0, bt, rebate=1.5)
            want = kf | (nv.F_UP if bt.startswith("up") else 0) | (nv.F_IN if bt.endswith("in") else 0)
            assert (s["kind"], s["flags"], s["K"], s["barrier"], s["rebate"]) == (nv.PAY_BARRIER, want, 100.0, 120.0, 1.5)
        for avg in ("arithmetic", "geometric"):
            for st in ("fixed", "floating"):
                s = ex.asian_spec(100.0, kind, avg, st)
                want = kf | (nv.F_GEOMETRIC if avg == "geometric" else 0) | (nv.F_FLOATING if st == "floating" else 0)
                assert (s["kind"], s["flags"]) == (nv.PAY_ASIAN, want)
        for st in ("floating", "fixed"):
            s = ex.lookback_spec(kind, 95.0, st)
            assert (s["kind"], s["flags"], s["K"]) == (nv.PAY_LOOKBACK, kf | (nv.F_FLOATING if st == "floating" else 0), 95.0)
        s = ex.digital_spec(100.0, kind, payout=7.0)
        assert (s["kind"], s["flags"], s["payout"]) == (nv.PAY_DIGITAL, kf, 7.0)
        assert ex.vanilla_spec(100.0, kind)["kind"] == nv.PAY_VANILLA
    arr = ex._native_payoffs([ex.barrier_spec(100.0, "call", 130.0, "up-and-out", 2.0), ex.digital_spec(90.0, "put", 3.0)],
                             0.05, 2.0)
    assert (arr[0].kind, arr[0].flags, arr[0].K, arr[0].barrier, arr[0].rebate) == (nv.PAY_BARRIER, nv.F_CALL | nv.F_UP, 100.0, 130.0, 2.0)
    assert (arr[1].kind, arr[1].flags, arr[1].K, arr[1].payout) == (nv.PAY_DIGITAL, 0, 90.0, 3.0)
    assert all((a.r_disc, a.T_disc) == (0.05, 2.0) for a in arr)
    for bad in (lambda: ex.barrier_spec(100.0, "call", 120.0, "sideways"), lambda: ex.asian_spec(100.0, "call", "harmonic"),
                lambda: ex.asian_spec(100.0, "call", "arithmetic", "sticky"), lambda: ex.lookback_spec("call", 0.0, "sticky"),
                lambda: ex.digital_spec(100.0, "straddle")):
        with pytest.raises(ValueError):
            bad()
