"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/optp_b200.h declares, the host package mirrors the reference's names, nothing in the
product imports the oracle, and without a GPU the compute entry points fail loudly."""
import ast
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")


@pytest.fixture(scope="module")
def native():
    import __graft_entry__  # noqa: F401  (puts the package on sys.path)
    from optpricer_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import build_native
        build_native.build()
    return _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "optp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(optp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/optp_b200.h but not exported"
    assert sorted(native.EXPORTS) == declared
    assert lib.optp_abi_version() == native.ABI_VERSION == 7


def test_struct_layout_matches_header(native, tmp_path):
    """sizeof / offsetof as a C compiler sees include/optp_b200.h vs the ctypes mirror."""
    import ctypes as C
    import shutil
    import subprocess
    assert C.sizeof(native.OptpProcess) == 8 + 14 * 8
    assert C.sizeof(native.OptpPayoff) == 8 + 6 * 8
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "optp_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %d\\n", sizeof(optp_process), '
                   'offsetof(optp_process, sigma), offsetof(optp_process, dS_bump), '
                   'offsetof(optp_process, sigma_table), offsetof(optp_process, h), sizeof(optp_payoff), '
                   'offsetof(optp_payoff, r_disc), OPTP_NSTATS);return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    P, Y = native.OptpProcess, native.OptpPayoff
    assert got == [C.sizeof(P), P.sigma.offset, P.dS_bump.offset, P.sigma_table.offset, P.h.offset, C.sizeof(Y),
                   Y.r_disc.offset, native.NSTATS]


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import optpricer_b200 as op
    with pytest.raises(native.NativeUnavailable):
        op.crr(op.OptionSpec(100, 100, 1, 0.05, 0.2), "call", 10)
    with pytest.raises(native.NativeUnavailable):
        op.euro_price_mc(op.OptionSpec(100, 100, 1, 0.05, 0.2), "call", n_paths=10, seed=1)
    with pytest.raises(native.NativeUnavailable):
        op.barrier_price(op.gbm_paths(100, 0.05, 0, 0.2, 1, 5, 5, seed=1), 100, 0.05, 1, "call", 120, "up-and-out")
    # argument errors of the reference are raised before any device work
    with pytest.raises(ValueError, match="N must be positive."):
        op.crr(op.OptionSpec(100, 100, 1, 0.05, 0.2), "call", 0)
    with pytest.raises(ValueError, match="S0 must be positive, got -1"):
        op.OptionSpec(-1, 100, 1, 0.05, 0.2)
    with pytest.raises(ValueError, match="n_steps and n_paths must be positive."):
        op.gbm_paths(100, 0.05, 0, 0.2, 1, 5, 0)


def test_product_never_imports_oracle():
    offenders = []
    for base, _, files in os.walk(PKG):
        for f in files:
            path = os.path.join(base, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    if any(n == "oracle" or n.startswith("oracle.") for n in names):
                        offenders.append(path)
            elif f.endswith((".cu", ".cuh", ".h")):
                code = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)   # comments may cite the oracle
                code = re.sub(r"//[^\n]*", "", code)
                if "oracle" in code:
                    offenders.append(path)
    assert not offenders, offenders


def test_public_names_mirror_reference(native):
    import inspect
    import optpricer_b200 as op
    from optpricer_b200 import processes, exotics, binomial, monte_carlo
    sig = lambda f: list(inspect.signature(f).parameters)  # noqa: E731
    assert sig(monte_carlo.euro_price_mc)[:10] == ["opt", "kind", "n_paths", "seed", "chunk_size", "antithetic",
                                                   "control_variate", "n_workers", "dtype", "return_stderr"]
    assert sig(processes.gbm_paths)[:9] == ["S0", "r", "q", "sigma", "T", "n_steps", "n_paths", "antithetic", "seed"]
    assert sig(processes.heston_paths)[:14] == ["S0", "r", "q", "v0", "kappa", "theta", "xi", "rho", "T", "n_steps",
                                                "n_paths", "antithetic", "seed", "return_variance"]
    assert sig(processes.milstein_local_vol_paths)[:10] == ["S0", "r", "q", "T", "n_steps", "n_paths", "sigma_loc",
                                                            "antithetic", "seed", "dS_bump"]
    assert sig(exotics.barrier_price)[:8] == ["paths", "K", "r", "T", "kind", "barrier", "barrier_type", "rebate"]
    assert sig(exotics.asian_price)[:7] == ["paths", "K", "r", "T", "kind", "average_type", "strike_type"]
    assert sig(exotics.lookback_price)[:6] == ["paths", "r", "T", "kind", "K", "strike_type"]
    assert sig(exotics.digital_price)[:6] == ["paths", "K", "r", "T", "kind", "payout"]
    assert sig(binomial.crr) == ["opt", "kind", "N", "american"]
    assert sig(binomial.crr_vec) == ["S0", "K", "T", "r", "q", "sigma", "kind", "N", "american"]
    assert op.CALL == "call" and op.PUT == "put"


def test_sigma_table_tabulation_matches_oracle(native):
    from optpricer_b200.processes import SigmaTable
    from oracle import sigma_grid
    fn = lambda S, t: 0.2 + 0.1 * np.exp(-S / 100.0) + 0.05 * t  # noqa: E731
    a = SigmaTable.from_callable(fn, S0=100.0, T=1.0, n_steps=12, n_nodes=64, half_width=1.6)
    b = sigma_grid.tabulate(fn, S0=100.0, T=1.0, n_steps=12, n_nodes=64, half_width=1.6)
    assert a.x0 == b.x0 and a.h == b.h and np.array_equal(a.table, b.table)


def test_lazy_paths_shape_protocol(native):
    from optpricer_b200 import processes as pr
    lp = pr.gbm_paths(100, 0.05, 0, 0.2, 1, 500, 100_000, antithetic=True, seed=42)
    assert lp.shape == (501, 200_000) and len(lp) == 501 and lp.ndim == 2
    assert pr.gbm_paths(100, 0.05, 0, 0.2, 1, 10, 7, antithetic=False, seed=1).shape == (11, 7)
    z = pr.heston_paths(100, .05, 0, .04, 2, .04, .3, -.7, 1, 3, 4, seed=5, rng="numpy").host_normals()
    g = np.random.default_rng(5)
    assert np.array_equal(z[0], g.standard_normal((3, 4))) and np.array_equal(z[1], g.standard_normal((3, 4)))


def test_signatures_mirror_reference(native):
    """Every parameter of every mirrored reference function -- name, kind (positional / keyword-only) and
    default -- in the reference's order; the B200 package may only APPEND keyword-only options.  Golden:
    tests/golden/signatures.json, recorded from the unmodified reference (make_golden_signatures.py)."""
    import importlib
    import inspect
    import json
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "signatures.json")) as f:
        golden = json.load(f)

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

    assert len(golden) == 19
    for key, ref in golden.items():
        mod, name = key.split(".")
        fn = getattr(importlib.import_module("optpricer_b200." + mod), name)
        mine = [[p.name, p.kind.name, enc(p.default)] for p in inspect.signature(fn).parameters.values()]
        assert mine[:len(ref)] == ref, (key, ref, mine)
        for extra in mine[len(ref):]:
            assert extra[1] == "KEYWORD_ONLY" and extra[2] != "<required>", (key, extra)


def test_argument_errors_mirror_reference(native):
    """Invalid arguments raise the reference's exception type and text BEFORE any native code runs (no GPU is
    needed to get the error).  Golden: tests/golden/errors.json, recorded by replaying golden_io.error_cases
    against the unmodified reference (make_golden_errors.py)."""
    import json
    import optpricer_b200
    from golden_io import error_cases
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "errors.json")) as f:
        golden = json.load(f)

    def norm(msg):  # the reference prints a set of allowed values: its order changes from process to process
        return re.sub(r"\{([^}]*)\}", lambda m: "{" + ", ".join(sorted(x.strip() for x in m.group(1).split(","))) + "}", msg)

    cases = error_cases(optpricer_b200)
    assert {n for n, _ in cases} == set(golden) and all(v is not None for v in golden.values())
    for name, call in cases:
        with pytest.raises(Exception) as info:
            call()
        assert [type(info.value).__name__, norm(str(info.value))] == [golden[name][0], norm(golden[name][1])], name
