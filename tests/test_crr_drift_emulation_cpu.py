"""CPU check of the ALGORITHM behind the drifting-window CRR walk (csrc/crr_batch.cu, OPTP_ZONE_DRIFT): the
option-level emulation in scripts/crr_drift_emu.py -- window placement 13 nodes under the exercise boundary, A / B
levels, the tree extended below node 0, the margin above the live top, the per-level bottom and top tests -- against the
plain lattice of binomial.py:7-38 (restated in the same file and pinned here to oracle/crr_tree.py)."""
import importlib.util
import os

import numpy as np

from oracle import crr_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("crr_drift_emu", os.path.join(ROOT, "scripts", "crr_drift_emu.py"))
emu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(emu)


def _contracts(n, seed):
    g = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        S0 = g.uniform(50, 150); K = S0 * g.uniform(0.6, 1.5); T = g.uniform(0.05, 3.0)
        sg = g.uniform(0.08, 0.8); r = float(g.choice([0.005, 0.03, 0.1])); q = float(g.choice([0.0, 0.0, 0.02]))
        if q < r:
            out.append((S0, K, T, r, q, sg))
    return out


def test_emulated_lattice_is_the_oracle_lattice():
    for S0, K, T, r, q, sg in _contracts(5, 1):
        a = emu.exact(S0, K, T, r, q, sg, 150)
        b = crr_tree.crr_scalar(S0, K, T, r, sg, q, "put", 150, True)
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b))


def test_drifting_walk_reproduces_the_full_lattice():
    worst, walked = 0.0, 0
    for N, seg0, seg in ((300, 32, 64), (301, 32, 64), (120, 24, 24)):
        for S0, K, T, r, q, sg in _contracts(12, N):
            price, why = emu.drift_walk(S0, K, T, r, q, sg, N, seg0=seg0, seg=seg, maxM=24)
            if price is None:          # a failed per-level test sends the contract to another pass: never a wrong price
                assert why in ("bottom", "top", "bottom0", "bottom_m", "wide")
                continue
            ref = emu.exact(S0, K, T, r, q, sg, N)
            worst = max(worst, abs(price - ref) / max(abs(ref), 1e-8 * K))
            walked += 1
    assert walked >= 30 and worst < 1e-12, (walked, worst)


def test_regime_without_exercise_nodes():
    """q > 0, spot far above the strike: near level 0 the whole tree lies above S1 = K(1-disc)/(1-e^{-q dt}), no node can
    choose exercise, and the walk starts its window at node 0 without the per-level tests (csrc/crr_batch.cu: k_pin)."""
    g = np.random.default_rng(11)
    pinned_seen, worst = 0, 0.0
    for _ in range(16):
        S0 = g.uniform(80, 160); K = S0 * g.uniform(0.2, 0.45); T = g.uniform(0.1, 0.5)
        sg = g.uniform(0.3, 0.8); r = g.uniform(0.04, 0.11); q = r * g.uniform(0.8, 0.98)
        N = 400
        dt = T / N
        S1 = K * (-np.expm1(-r * dt)) / (-np.expm1(-q * dt))
        k_pin = int(np.ceil(-np.log(S1 / S0) / (sg * np.sqrt(dt)) - 0.01))
        pinned_seen += k_pin > 64
        price, why = emu.drift_walk(S0, K, T, r, q, sg, N, seg0=32, seg=64, maxM=24)
        assert price is not None, why
        ref = emu.exact(S0, K, T, r, q, sg, N)
        if ref > 1e-18 * K:                 # (values below 2^-80 K are cut on the way)
            worst = max(worst, abs(price - ref) / ref)
        else:
            assert price <= 1e-17 * K
    assert pinned_seen >= 8 and worst < 1e-6, (pinned_seen, worst)
