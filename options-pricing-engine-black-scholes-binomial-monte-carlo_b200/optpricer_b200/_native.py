"""ctypes binding of liboptp_b200.so (C ABI: include/optp_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is
usable, every compute entry point raises NativeUnavailable.  Build the library with
`python options-pricing-engine-black-scholes-binomial-monte-carlo_b200/build_native.py`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

# OPTP_LIB: another build of the same library (A/B runs of kernel variants, scripts/crr_ab.py)
LIB_PATH = os.environ.get("OPTP_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib",
                                                      "liboptp_b200.so")
ABI_VERSION = 7
NSTATS = 8
MAX_PAYOFFS = 16

# process kinds / payoff kinds / flags / modes -- mirror include/optp_b200.h
(PROC_GBM, PROC_HESTON_FTE, PROC_LV_MILSTEIN, PROC_LV_EULER, PROC_GBM_MILSTEIN, PROC_SABR, PROC_HESTON_QE,
 PROC_MERTON, PROC_HESTON_QE_M) = range(9)
PAY_VANILLA, PAY_DIGITAL, PAY_BARRIER, PAY_ASIAN, PAY_LOOKBACK = range(5)
F_CALL, F_UP, F_IN, F_GEOMETRIC, F_FLOATING = 1, 2, 4, 8, 16
MODE_FP32_PHILOX, MODE_FP64_PHILOX, MODE_FP64_HOSTZ = 0, 1, 2

EXPORTS = (
    "optp_abi_version", "optp_init", "optp_shutdown", "optp_last_error", "optp_device_count",
    "optp_launch_count", "optp_set_option", "optp_get_counter", "optp_host_alloc", "optp_host_free",
    "optp_mc_terminal", "optp_mc_terminal_dev", "optp_mc_path", "optp_mc_path_dev", "optp_mc_book",
    "optp_sigma_table_upload", "optp_dupire_table", "optp_paths", "optp_payoff_from_paths",
    "optp_crr_batch", "optp_crr_batch_dev", "optp_microbench",
    "optp_sigma_table_create", "optp_sigma_table_destroy",
    "optp_comm_unique_id", "optp_comm_init", "optp_comm_destroy", "optp_comm_info",
    "optp_allreduce_stats", "optp_allreduce_stats_dev",
    "optp_philox4x32_10", "optp_normal_stats",
)


class NativeUnavailable(RuntimeError):
    """The CUDA library or a CUDA device is missing; there is no CPU fallback."""


class OptpProcess(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_nodes", C.c_int32),
                ("S0", C.c_double), ("r", C.c_double), ("q", C.c_double), ("T", C.c_double),
                ("sigma", C.c_double),
                ("v0", C.c_double), ("kappa", C.c_double), ("theta", C.c_double),
                ("xi", C.c_double), ("rho", C.c_double),
                ("dS_bump", C.c_double),
                ("sigma_table", C.c_void_p), ("x0", C.c_double), ("h", C.c_double)]


class OptpPayoff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32),
                ("K", C.c_double), ("barrier", C.c_double), ("rebate", C.c_double),
                ("payout", C.c_double), ("r_disc", C.c_double), ("T_disc", C.c_double)]


class OptpBookItem(C.Structure):
    _fields_ = [("terminal_only", C.c_int32), ("n_payoffs", C.c_int32),
                ("first_payoff", C.c_uint32), ("n_steps", C.c_uint32),
                ("n_base", C.c_uint64), ("seed", C.c_uint64), ("path_offset", C.c_uint64),
                ("antithetic", C.c_int32), ("mode", C.c_int32),
                ("is_call", C.c_int32), ("table_handle", C.c_int32),
                ("K", C.c_double), ("proc", OptpProcess)]


_lock = threading.Lock()
_lib = None
_device = None

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def _declare(lib):
    d, i, u64, u32, vp = C.c_double, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p
    lib.optp_abi_version.restype = i
    lib.optp_init.argtypes = [i]
    lib.optp_last_error.restype = C.c_char_p
    lib.optp_launch_count.restype = u64
    lib.optp_set_option.argtypes = [C.c_char_p, C.c_long]
    lib.optp_get_counter.argtypes = [C.c_char_p, C.POINTER(u64)]
    lib.optp_host_alloc.argtypes = [C.POINTER(vp), u64]
    lib.optp_host_free.argtypes = [vp]
    term = [d, d, d, d, d, d, i, u64, i, i, i, u64, u64, vp, vp]
    lib.optp_mc_terminal.argtypes = term
    lib.optp_mc_terminal_dev.argtypes = term + [vp]
    path = [C.POINTER(OptpProcess), C.POINTER(OptpPayoff), u32, u32, u64, i, i, u64, u64, vp, vp, vp, vp]
    lib.optp_mc_path.argtypes = path
    lib.optp_mc_path_dev.argtypes = path + [vp]
    lib.optp_mc_book.argtypes = [C.POINTER(OptpBookItem), u64, C.POINTER(OptpPayoff), u64, vp]
    lib.optp_sigma_table_upload.argtypes = [vp, u32, u32]
    lib.optp_dupire_table.argtypes = [vp, vp, u32, vp, vp, u32, d, u32, d, d, u32, d, vp]
    lib.optp_paths.argtypes = [C.POINTER(OptpProcess), u32, u64, i, i, u64, u64, vp, vp, vp]
    lib.optp_payoff_from_paths.argtypes = [vp, u64, u64, C.POINTER(OptpPayoff), vp, vp]
    crr = [vp, vp, vp, vp, vp, vp, vp, u64, u32, i, vp, vp]
    lib.optp_crr_batch.argtypes = crr
    lib.optp_crr_batch_dev.argtypes = crr + [vp]
    lib.optp_microbench.argtypes = [i, _dp, _dp]
    lib.optp_sigma_table_create.argtypes = [vp, u32, u32, C.POINTER(C.c_int32)]
    lib.optp_sigma_table_destroy.argtypes = [C.c_int32]
    lib.optp_comm_unique_id.argtypes = [C.c_char_p]
    lib.optp_comm_init.argtypes = [i, i, C.c_char_p]
    lib.optp_comm_info.argtypes = [C.POINTER(i), C.POINTER(i)]
    lib.optp_allreduce_stats.argtypes = [vp, u64]
    lib.optp_allreduce_stats_dev.argtypes = [vp, u64, vp]
    lib.optp_philox4x32_10.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    lib.optp_normal_stats.argtypes = [u64, u64, u64, u32, i, _dp]
    for name in EXPORTS:
        if name not in ("optp_abi_version", "optp_last_error", "optp_launch_count"):
            getattr(lib, name).restype = i


def load():
    """Load the shared library (no device is touched)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeUnavailable(
                    f"{LIB_PATH} is missing: build it with build_native.py (nvcc, sm_100a). "
                    "optpricer_b200 has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            _declare(lib)
            if lib.optp_abi_version() != ABI_VERSION:
                raise NativeUnavailable(
                    f"ABI mismatch: library {lib.optp_abi_version()} vs binding {ABI_VERSION}; rebuild")
            _lib = lib
    return _lib


def last_error():
    return load().optp_last_error().decode("utf-8", "replace")


def check(status, what=""):
    if status == 0:
        return
    msg = last_error()
    if status > 0:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise NativeUnavailable(f"{what or 'optp'} failed with CUDA status {-status}: {msg}")


def default_device():
    return int(os.environ.get("LOCAL_RANK", "0"))


def init(device=None):
    """Create the per-device context (idempotent); returns the library handle."""
    global _device
    lib = load()
    dev = default_device() if device is None else int(device)
    if _device != dev:
        if lib.optp_device_count() <= 0:
            raise NativeUnavailable("no CUDA device visible: optpricer_b200 has no CPU fallback")
        check(lib.optp_init(dev), "optp_init")
        _device = dev
    return lib


def shutdown():
    global _device
    if _lib is not None:
        _lib.optp_shutdown()
    _device = None


def set_option(key, value):
    check(load().optp_set_option(key.encode(), int(value)), "optp_set_option")


def get_counter(key):
    v = C.c_uint64()
    check(load().optp_get_counter(key.encode(), C.byref(v)), "optp_get_counter")
    return int(v.value)


def launch_count():
    return int(load().optp_launch_count())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def philox4x32_10(ctr, key):
    """The library's Philox4x32-10 on the host (known-answer tests); needs no device."""
    lib = load()
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    out = (C.c_uint32 * 4)()
    check(lib.optp_philox4x32_10(c, k, out), "optp_philox4x32_10")
    return tuple(int(x) for x in out)


def normal_stats(seed, n_paths, n_blocks, *, rng_mode=0, path_offset=0):
    """Moments, tail counts and pair products of the device normals (see include/optp_b200.h)."""
    lib = init()
    out = (C.c_double * 16)()
    check(lib.optp_normal_stats(int(seed), int(path_offset), int(n_paths), int(n_blocks), int(rng_mode), out),
          "optp_normal_stats")
    return np.array(out[:], dtype=np.float64)


def microbench(which):
    lib = init()
    v, mhz = C.c_double(), C.c_double()
    check(lib.optp_microbench(int(which), C.byref(v), C.byref(mhz)), "optp_microbench")
    return v.value, mhz.value
