"""CRR binomial pricers on the B200 (drop-in for reference binomial.py).

  crr        <- crr       binomial.py:7-38    (scalar contract; one-element device batch)
  crr_vec    <- crr_vec   binomial.py:44-109  (many (K, kind) on one shared tree)
  crr_batch  -- new: heterogeneous contract batch, the shape BASELINE config 2 prices
               (the reference can only loop `crr` over it)
All three run the same CUDA kernel (csrc/crr_batch.cu) through optp_crr_batch.
"""
from __future__ import annotations

import ctypes as C
from math import exp, sqrt

import numpy as np

from . import _native as nv
from .core import CALL

_P_ERROR = "Risk-neutral prob p out of (0,1); try larger N or different params."


def _check_lattice(T, r, q, sigma, N):
    """The reference's argument checks, in its order and with its texts (binomial.py:9-17)."""
    if N <= 0:
        raise ValueError("N must be positive.")
    dt = T / N
    u = exp(sigma * sqrt(dt))
    d = 1.0 / u
    p = (exp((r - q) * dt) - d) / (u - d)
    if not (0.0 < p < 1.0):
        raise ValueError(_P_ERROR)


def crr_batch(S0, K, T, r, q, sigma, is_call, N=500, *, american=False, return_status=False):
    """Price n independent contracts (arrays broadcast to a common length) on N-step CRR trees.

    Contracts whose risk-neutral probability falls outside (0,1) get NaN and status 1
    (binomial.py:16-17 raises; a batch must not abort).  Returns prices (and statuses).
    """
    if N <= 0:
        raise ValueError("N must be positive.")
    # (no pass over the data unless a column has to be broadcast or converted: at 1M contracts a detour of the flag
    # column through float64 cost 2.5 ms of a 22 ms call)
    vals = [np.asarray(x, dtype=np.float64) for x in (S0, K, T, r, q, sigma)]
    flag = np.asarray(is_call)
    shape = np.broadcast_shapes(*[v.shape for v in vals], flag.shape)
    flat = [np.ascontiguousarray(np.broadcast_to(v, shape).reshape(-1)) for v in vals]
    call = np.ascontiguousarray(np.broadcast_to(flag != 0, shape).reshape(-1)).view(np.uint8)
    n = flat[0].size
    price = np.empty(n, dtype=np.float64)
    status = np.zeros(n, dtype=np.uint8)
    if n:
        lib = nv.init()
        nv.check(lib.optp_crr_batch(*[nv.ptr(c) for c in flat], nv.ptr(call), n, int(N),
                                    1 if american else 0, nv.ptr(price), nv.ptr(status)),
                 "optp_crr_batch")
    price = price.reshape(shape)
    status = status.reshape(shape)
    return (price, status) if return_status else price


def crr(opt, kind=CALL, N=500, *, american=False):
    """Cox-Ross-Rubinstein price of one option (binomial.py:7-38).

    As in the reference, any `kind` other than "call" prices a put (binomial.py:22-25)."""
    _check_lattice(opt.T, opt.r, opt.q, opt.sigma, N)
    # a batch of one, without the array plumbing of crr_batch: the six "columns" are consecutive doubles
    cols = np.array([opt.S0, opt.K, opt.T, opt.r, opt.q, opt.sigma, 0.0], dtype=np.float64)
    call = np.array([1 if kind == "call" else 0], dtype=np.uint8)
    status = np.zeros(1, dtype=np.uint8)
    base = cols.ctypes.data
    lib = nv.init()
    nv.check(lib.optp_crr_batch(*[C.c_void_p(base + 8 * i) for i in range(6)], nv.ptr(call), 1, int(N),
                                1 if american else 0, C.c_void_p(base + 48), nv.ptr(status)), "optp_crr_batch")
    if status[0] != 0:   # the device's own p test disagreed with the host's (binomial.py:16-17): never return its NaN
        raise ValueError(_P_ERROR)
    return float(cols[6])


def crr_vec(S0, K, T, r, q, sigma, kind, N=500, *, american=False):
    """Many (K, kind) pairs on one shared tree (binomial.py:44-109); returns shape of atleast_1d(K)."""
    K = np.atleast_1d(np.asarray(K, dtype=float))
    kind = np.atleast_1d(np.asarray(kind))
    if kind.shape != K.shape:
        kind = np.broadcast_to(kind, K.shape)
    _check_lattice(T, r, q, sigma, N)
    is_call = (kind.astype(str) == "call").astype(np.uint8).reshape(K.shape)   # (anything else prices a put)
    return crr_batch(S0, K, T, r, q, sigma, is_call, N, american=american)
