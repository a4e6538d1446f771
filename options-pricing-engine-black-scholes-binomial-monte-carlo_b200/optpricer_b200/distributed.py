"""Single-node multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed).

Monte Carlo: the reference's only parallelism is chunk-level data parallelism with additive
sufficient statistics (monte_carlo.py:126-139, :67-74).  Here every rank simulates a contiguous
range of BASE path indices -- Philox draws are a function of (seed, path index, step), so the
union over ranks is exactly the single-GPU path set -- and ONE all-reduce (NCCL over NVLink on
the GPU box, gloo in the CPU tests) sums the statistics blocks of all priced payoffs at once.
CRR: contracts are independent; ranks price contiguous slices and the prices are all-gathered.

`stats_fn` / `price_fn` are injectable so the host logic is testable without a GPU (tests use
the numpy oracle there); the defaults call the CUDA library and raise if it is unavailable.
"""
from __future__ import annotations

import numpy as np

from . import _native as nv


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_range(n, rank, size):
    """Contiguous [lo, hi) of n items for `rank` of `size`; sizes differ by at most one."""
    base, extra = divmod(int(n), int(size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_comm_world = 1   # size of the library's own NCCL communicator (optp_comm_init), 1 = none


def init_comm():
    """Create the library's NCCL communicator over the ranks of the default process group
    (optp_comm_unique_id on rank 0 -> broadcast -> optp_comm_init everywhere).  After this the statistics
    blocks are reduced by optp_allreduce_stats -- pinned staging, ncclAllReduce and the copy back inside ONE
    C call -- instead of a numpy -> torch -> NCCL -> numpy round trip.  Idempotent; a no-op for a world of
    one or a CPU (gloo) process group."""
    global _comm_world
    rank, size = world()
    if size == 1 or _comm_world == size:
        return _comm_world > 1
    import ctypes as C
    import torch.distributed as dist
    if dist.get_backend() != "nccl":
        return False
    lib = nv.init()
    buf = C.create_string_buffer(128)
    if rank == 0:
        nv.check(lib.optp_comm_unique_id(buf), "optp_comm_unique_id")
    box = [bytes(buf.raw)]
    import torch
    dist.broadcast_object_list(box, src=0, device=torch.device("cuda", nv.default_device()))
    nv.check(lib.optp_comm_init(rank, size, box[0]), "optp_comm_init")
    _comm_world = size
    return True


def allreduce_sum(block):
    """Sum a float64 numpy block over all ranks (one collective for the whole block): through the C ABI
    (optp_allreduce_stats) on the GPU box, through torch.distributed (gloo) in the CPU tests."""
    rank, size = world()
    block = np.ascontiguousarray(block, dtype=np.float64)
    if size == 1:
        return block
    if init_comm():
        out = block.copy()
        nv.check(nv.load().optp_allreduce_stats(nv.ptr(out), out.size), "optp_allreduce_stats")
        return out
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(block.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(nv.default_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allgather_concat(part, counts):
    """Concatenate per-rank 1-D float64 arrays of the given lengths, in rank order."""
    rank, size = world()
    if size == 1:
        return np.asarray(part)
    import torch
    import torch.distributed as dist
    m = max(counts)
    pad = np.zeros(m, dtype=np.float64)
    pad[:len(part)] = part
    t = torch.from_numpy(pad)
    if dist.get_backend() == "nccl":
        t = t.cuda(nv.default_device())
    out = [torch.empty_like(t) for _ in range(size)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(out, counts)])


# ---------------------------------------------------------------------------------- Monte Carlo
def euro_price_mc(opt, kind, *, n_paths=100_000, seed=None, antithetic=True, control_variate=True,
                  return_stderr=True, mode="fp32", stats_fn=None):
    """euro_price_mc (monte_carlo.py:77-167) with the base paths sharded over the process group."""
    from . import monte_carlo as mc
    rank, size = world()
    if seed is None and size > 1:
        raise ValueError("a sharded run needs an explicit seed (every rank must use the same key)")
    q = getattr(opt, "q", 0.0)
    is_call = mc._kind_flag(kind)
    lo, hi = shard_range(n_paths, rank, size)
    fn = stats_fn or (lambda n, off: mc.terminal_stats(opt.S0, opt.K, opt.T, opt.r, q, opt.sigma, is_call, n,
                                                       antithetic=antithetic, mode=mode, seed=mc._seed64(seed),
                                                       path_offset=off))
    local = np.zeros(nv.NSTATS) if hi == lo else np.asarray(fn(hi - lo, lo), dtype=np.float64)
    total = allreduce_sum(local)
    return mc.estimator((int(total[0]), *total[1:6]), S0=opt.S0, q=q, T=opt.T, control_variate=control_variate,
                        return_stderr=return_stderr)


def price_payoffs(paths, specs, r, T, *, control_variate=False, stats_fn=None):
    """exotics.price_payoffs with the base paths of a LazyPaths sharded over the process group."""
    from . import exotics as ex
    rank, size = world()
    lo, hi = shard_range(paths.n_paths, rank, size)
    fn = stats_fn or (lambda n, off: ex.path_stats(paths, specs, r, T, n_base=n, path_offset=off)[0])
    local = (np.zeros((len(specs), nv.NSTATS)) if hi == lo
             else np.asarray(fn(hi - lo, lo), dtype=np.float64).reshape(len(specs), nv.NSTATS))
    total = allreduce_sum(local)
    ey = ex._expected_y(paths, r, T) if control_variate else None
    return [ex.estimate(s, control_variate=control_variate, expected_y=ey) for s in total]


# ---------------------------------------------------------------------------------- CRR
def crr_batch(S0, K, T, r, q, sigma, is_call, N=500, *, american=False, gather=True, price_fn=None):
    """binomial.crr_batch with the contracts sharded over the process group (no data-path
    collective; the all-gather only assembles the result when gather=True)."""
    from . import binomial as bn
    rank, size = world()
    cols = [np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=np.float64), np.shape(S0)).reshape(-1))
            for c in (S0, K, T, r, q, sigma, is_call)]
    n = cols[0].size
    lo, hi = shard_range(n, rank, size)
    fn = price_fn or (lambda *a: bn.crr_batch(*a, N, american=american))
    part = np.asarray(fn(*[c[lo:hi] for c in cols]), dtype=np.float64) if hi > lo else np.empty(0)
    if not gather:
        return part
    counts = [shard_range(n, k, size)[1] - shard_range(n, k, size)[0] for k in range(size)]
    return allgather_concat(part, counts)
