"""Book pricing on the B200: the batched replacement of the serial row loop of the reference's
scripts/price_book.py (:45-118 `_price_row`, :139-145 the loop).

Rows use the reference's CSV schema (price_book.py:9-18): id,S0,K,T,r,sigma,q,kind,method and, for
exotics, exotic_type,barrier,barrier_type,rebate,average_type,strike_type,payout,n_steps,n_paths,
american.  Methods on the hot path:
  binomial  all rows of one `american` flag become ONE optp_crr_batch call (reference: crr, N=500)
  mc        euro_price_mc (reference: n_paths=100_000, seed=1)          -> one terminal-kernel item
  exotic    gbm_paths(..., antithetic=True, seed=1) + one payoff pricer -> one fused path-kernel
            item; rows that share (S0, r, q, sigma, T, n_steps, n_paths) -- the script seeds every
            row with seed=1, i.e. common random numbers -- share ONE simulated path set and are
            evaluated as several payoffs of the same launch (up to 16 per launch)
            a row may carry `sigma_table` (a processes.SigmaTable; Python rows only): it is then simulated
            with milstein_local_vol_paths(..., seed=1) on that table instead of gbm_paths -- the desk
            workflow's local-vol barrier + vanilla legs (scripts/desk_workflow_localvol_barrier.py:136-148)
            and its bumped-surface reprices batch into the same optp_mc_book call, one table handle each
All Monte Carlo items of the book are launched back to back by optp_mc_book and their statistics
return in one device-to-host copy; with a torch.distributed process group the items (or, for a book of
few items, the base paths of every item) are sharded over the ranks and ONE all-reduce sums the whole
book's statistics (shard_items).
`method == "bs"` rows are closed-form Black-Scholes (not a hot path; priced on the host so a mixed
book still comes back complete).  As in the reference, a failing row yields
{"id", "price": None, "error": str} and does not abort the book.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as nv
from . import exotics as ex
from . import monte_carlo as mc
from .binomial import _check_lattice
from .core import OptionSpec


def _f(row, key, default=None):
    v = row.get(key, default)
    if v is None or (isinstance(v, str) and v.strip() == ""):
        if default is None:
            raise KeyError(key)
        v = default
    return float(v)


def _s(row, key, default=""):
    v = row.get(key, default)
    return (default if v is None else str(v)).strip().lower()


def _bs_price(S0, K, T, r, q, sigma, kind):
    """Closed-form Black-Scholes with continuous yield (reference black_scholes_vec.py:37-60)."""
    d1 = (math.log(S0 / K) + (r - q + 0.5 * sigma * sigma) * T) / (sigma * math.sqrt(T))
    d2 = d1 - sigma * math.sqrt(T)
    N = lambda x: 0.5 * math.erfc(-x / math.sqrt(2.0))  # noqa: E731
    if kind == "call":
        return S0 * math.exp(-q * T) * N(d1) - K * math.exp(-r * T) * N(d2)
    if kind == "put":
        return K * math.exp(-r * T) * N(-d2) - S0 * math.exp(-q * T) * N(-d1)
    raise ValueError("kind must be 'call' or 'put'")


def _exotic_spec(row, K, kind):
    et = _s(row, "exotic_type")
    if et == "barrier":
        return ex.barrier_spec(K, kind, _f(row, "barrier"), _s(row, "barrier_type"), _f(row, "rebate", 0.0))
    if et == "asian":
        return ex.asian_spec(K, kind, _s(row, "average_type", "arithmetic"), _s(row, "strike_type", "fixed"))
    if et == "digital":
        return ex.digital_spec(K, kind, _f(row, "payout", 1.0))
    if et == "lookback":
        return ex.lookback_spec(kind, K, _s(row, "strike_type", "floating"))
    raise ValueError(f"Unknown exotic_type: {et!r}")


def plan_book(rows, *, tree_steps=500, mc_paths=100_000, exotic_steps=500, exotic_paths=100_000, seed=1):
    """Parse rows into device work.  Returns (results, crr_groups, items) where `results` already
    holds the host-priced / failed rows and the rest are filled in by price_book."""
    results = [None] * len(rows)
    crr_groups = {}      # american flag -> list of (row index, S0, K, T, r, q, sigma, is_call)
    items = _plan_rows(rows, 0, len(rows), results, crr_groups, tree_steps=tree_steps, mc_paths=mc_paths,
                       exotic_steps=exotic_steps, exotic_paths=exotic_paths, seed=seed)
    return results, crr_groups, items


def _plan_rows(rows, lo, hi, results, crr_groups, *, tree_steps, mc_paths, exotic_steps, exotic_paths, seed):
    """plan_book for rows[lo:hi] (row indices stay global); returns the Monte Carlo items of the slice."""
    items = []           # dicts: kind "terminal" | "path", rows [(row index, payoff spec)], params
    path_groups = {}     # (S0, r, q, sigma, T, n_steps, n_paths) -> item (<= 16 payoffs each)
    for i in range(lo, hi):
        row = rows[i]
        rid = row.get("id", "")
        try:
            S0, K, T, r, sigma = (_f(row, k) for k in ("S0", "K", "T", "r", "sigma"))
            q = _f(row, "q", 0.0)
            kind, method = _s(row, "kind"), _s(row, "method")
            if method == "bs":
                results[i] = {"id": rid, "price": _bs_price(S0, K, T, r, q, sigma, kind), "stderr": None}
            elif method == "binomial":
                OptionSpec(S0=S0, K=K, T=T, r=r, sigma=sigma, q=q)
                _check_lattice(T, r, q, sigma, tree_steps)
                american = _s(row, "american", "false") == "true"
                crr_groups.setdefault(american, []).append((i, S0, K, T, r, q, sigma, 1 if kind == "call" else 0))
            elif method == "mc":
                OptionSpec(S0=S0, K=K, T=T, r=r, sigma=sigma, q=q)
                items.append(dict(kind="terminal", rows=[(i, None)], S0=S0, K=K, T=T, r=r, q=q, sigma=sigma,
                                  is_call=mc._kind_flag(kind), n_base=int(mc_paths), seed=seed))
            elif method == "exotic":
                n_steps = int(_f(row, "n_steps", exotic_steps))
                n_paths = int(_f(row, "n_paths", exotic_paths))
                if n_steps <= 0 or n_paths <= 0:
                    raise ValueError("n_steps and n_paths must be positive.")
                spec = _exotic_spec(row, K, kind)
                table = row.get("sigma_table")
                if table is not None and (table.table.shape[0] < n_steps
                                          or abs(table.dt - T / n_steps) > 1e-12 * table.dt):
                    raise ValueError("sigma_table does not match T / n_steps of the row")
                row_seed = int(row["seed"]) if row.get("seed") not in (None, "") else seed   # (the script: always 1)
                key = (S0, r, q, sigma if table is None else id(table), T, n_steps, n_paths, row_seed)
                it = path_groups.get(key)
                if it is None or len(it["rows"]) >= nv.MAX_PAYOFFS:
                    it = dict(kind="path", rows=[], S0=S0, r=r, q=q, sigma=sigma, T=T, n_steps=n_steps,
                              n_base=n_paths, seed=row_seed, table=table,
                              dS_bump=float(row.get("dS_bump", 0.01)))
                    path_groups[key] = it
                    items.append(it)
                it["rows"].append((i, spec))
            else:
                raise ValueError(f"Unknown method: {method!r}")
        except Exception as e:  # per-row failure, as scripts/price_book.py:139-145
            results[i] = {"id": rid, "price": None, "error": str(e)}
    return items


def shard_items(items, rank, world, load=None):
    """This rank's share of every item as a list of base-path ranges [(lo, hi), ...], one per item.

    A book with many items (>= 4 per rank) is sharded BY ITEM: each item runs whole on one rank, so a rank
    fills and launches 1/world of the work descriptors and no kernel shrinks to a launch-bound size; items
    are dealt out longest-first to the least loaded rank (cost = base paths x steps), identically on every
    rank.  A book with few items is sharded BY PATHS (every item on every rank, disjoint ranges of its base
    paths).  Either way the statistics are additive, so ONE all-reduce over the (n_blocks, 8) array -- blocks
    a rank did not touch stay zero -- gives the whole book (monte_carlo.py:67-74).
    `load` (a list of `world` running totals, updated in place) carries the ranks' loads from one slice of a
    book to the next (price_book plans and launches a large book slice by slice)."""
    import heapq
    from .distributed import shard_range
    if world <= 1:
        return [(0, it["n_base"]) for it in items]
    if len(items) < 4 * world:
        return [shard_range(it["n_base"], rank, world) for it in items]
    cost = [it["n_base"] * (1 if it["kind"] == "terminal" else it["n_steps"]) for it in items]
    order = sorted(range(len(items)), key=lambda i: (-cost[i], i))
    if load is None:
        load = [0] * world
    heap = [(load[k], k) for k in range(world)]
    heapq.heapify(heap)
    owner = [0] * len(items)
    for i in order:
        ld, r = heapq.heappop(heap)      # least loaded rank, lowest rank on ties
        owner[i] = r
        heapq.heappush(heap, (ld + cost[i], r))
    for ld, r in heap:
        load[r] = ld
    return [(0, it["n_base"]) if owner[i] == rank else (0, 0) for i, it in enumerate(items)]


def book_stats(items, *, mode="fp32", rank=0, world=1, load=None):
    """Run this rank's share (shard_items) of all Monte Carlo items through optp_mc_book.
    Returns the (n_blocks, 8) statistics array (item order, payoff order inside an item)."""
    return book_run(book_prepare(items, mode=mode, rank=rank, world=world, load=load))


def book_prepare(items, *, mode="fp32", rank=0, world=1, load=None):
    """Host half of book_stats: deal the items (shard_items), fill the optp_book_item / optp_payoff arrays of this
    rank's share and stage the sigma tables its local-vol items name.  price_book runs this for slice k+1 while
    a worker thread is inside optp_mc_book for slice k, so the device never waits for descriptors."""
    shards = shard_items(items, rank, world, load)
    n_blocks = sum(len(it["rows"]) for it in items)
    prep = dict(items=items, n_blocks=n_blocks, live=0, handles={})
    if not items:
        return prep
    lib = nv.init()
    arr = (nv.OptpBookItem * len(items))()
    specs = []
    live = 0
    handles = prep["handles"]   # id(SigmaTable) -> device table handle, one per distinct table of this rank's items
    try:
        return _fill_descriptors(lib, prep, items, shards, arr, specs, live, handles, mode)
    except BaseException:
        for h in handles.values():   # nothing will run this slice: do not leave its tables staged
            lib.optp_sigma_table_destroy(h)
        handles.clear()
        raise


def _fill_descriptors(lib, prep, items, shards, arr, specs, live, handles, mode):
    skip = []
    for it, (lo, hi) in zip(items, shards):
        a = arr[live]
        if hi == lo:
            skip.append(True)    # this rank has no paths of the item; its blocks stay zero
            continue
        skip.append(False)
        live += 1
        a.terminal_only = 1 if it["kind"] == "terminal" else 0
        a.n_base, a.seed, a.path_offset = hi - lo, mc._seed64(it["seed"]), lo
        a.antithetic, a.mode = 1, mc._MODES[mode]
        a.proc.kind = nv.PROC_GBM
        a.proc.S0, a.proc.r, a.proc.q, a.proc.T, a.proc.sigma = it["S0"], it["r"], it["q"], it["T"], it["sigma"]
        tb = it.get("table")
        if tb is not None:
            if id(tb) not in handles:
                h = C.c_int32()
                nv.check(lib.optp_sigma_table_create(nv.ptr(tb.table), tb.table.shape[0], tb.table.shape[1],
                                                     C.byref(h)), "optp_sigma_table_create")
                handles[id(tb)] = h.value
            a.proc.kind = nv.PROC_LV_MILSTEIN
            a.proc.n_nodes, a.proc.x0, a.proc.h, a.proc.dS_bump = tb.table.shape[1], tb.x0, tb.h, it["dS_bump"]
            a.table_handle = handles[id(tb)]
        if it["kind"] == "terminal":
            a.K, a.is_call, a.n_steps, a.n_payoffs, a.first_payoff = it["K"], it["is_call"], 1, 1, 0
        else:
            a.n_steps, a.n_payoffs, a.first_payoff = it["n_steps"], len(it["rows"]), len(specs)
            specs.extend((s, it["r"], it["T"]) for _, s in it["rows"])
    pay = (nv.OptpPayoff * max(1, len(specs)))()
    for j, (s, r, T) in enumerate(specs):
        one = ex._native_payoffs([s], r, T)[0]
        C.memmove(C.byref(pay[j]), C.byref(one), C.sizeof(nv.OptpPayoff))
    prep.update(arr=arr, pay=pay, n_specs=len(specs), live=live, skip=skip)
    return prep


def book_run(prep):
    """Device half of book_stats: ONE optp_mc_book call (blocking, GIL released) and the scatter of its blocks
    into item order; destroys the table handles of the slice."""
    items = prep["items"]
    stats = np.zeros((prep["n_blocks"], nv.NSTATS), dtype=np.float64)
    if not items:
        return stats
    lib = nv.init()
    try:
        live_blocks = sum(len(it["rows"]) for it, sk in zip(items, prep["skip"]) if not sk)
        out = np.zeros((live_blocks, nv.NSTATS), dtype=np.float64)
        if prep["live"]:
            nv.check(lib.optp_mc_book(prep["arr"], prep["live"], prep["pay"], prep["n_specs"], nv.ptr(out)),
                     "optp_mc_book")
        src = dst = 0
        for it, sk in zip(items, prep["skip"]):
            n = len(it["rows"])
            if not sk:
                stats[dst:dst + n] = out[src:src + n]
                src += n
            dst += n
        return stats
    finally:
        for h in prep["handles"].values():
            lib.optp_sigma_table_destroy(h)


def estimate_rows(items, stats, control_variate=True):
    """(price, stderr) of every row of `items` from the (n_blocks, 8) statistics, vectorised: terminal rows with
    the euro_price_mc estimator (monte_carlo.py:142-167: ddof=0, control variate Y = e^{-rT} S_T), path rows with
    exotics._price_from_payoff's (exotics.py:22-27: ddof=1).  Same formulas, same operation order as
    monte_carlo.estimator / exotics.estimate, evaluated on arrays.  Returns (row indices, prices, stderrs)."""
    idx, term, ey = [], [], []
    for it in items:
        t = it["kind"] == "terminal"
        e = it["S0"] * math.exp(-it["q"] * it["T"]) if t else 0.0
        for (i, _) in it["rows"]:
            idx.append(i)
            term.append(t)
            ey.append(e)
    n, sx, sxx, sy, syy, sxy = (stats[:, k] for k in range(6))
    term = np.array(term, dtype=bool)
    ey = np.array(ey, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = sx / n
        # path rows: ddof = 1
        var1 = np.maximum(0.0, (sxx - n * mean * mean) / (n - 1.0))
        p_price, p_se = mean, np.where(n > 1, np.sqrt(var1) / np.sqrt(n), 0.0)
        # terminal rows: ddof = 0 (+ control variate)
        varX = np.maximum(0.0, sxx / n - mean * mean)
        if control_variate:
            meanY = sy / n
            varY = np.maximum(0.0, syy / n - meanY * meanY)
            cov = (sxy / n) - mean * meanY
            c_hat = np.where(varY == 0.0, 0.0, cov / np.where(varY == 0.0, 1.0, varY))
            t_price = mean - c_hat * (meanY - ey)
            var_cv = varX - 2.0 * c_hat * cov + (c_hat * c_hat) * varY
            t_se = np.sqrt(np.maximum(0.0, var_cv) / n)
        else:
            t_price, t_se = mean, np.sqrt(np.maximum(0.0, varX) / n)
    price = np.where(term, t_price, p_price)
    se = np.where(term, t_se, p_se)
    empty = n <= 0
    price = np.where(empty, np.nan, price)
    se = np.where(empty, np.nan, se)
    return idx, price, se


def price_book(rows, *, tree_steps=500, mc_paths=100_000, exotic_steps=500, exotic_paths=100_000, seed=1,
               mode="fp32", control_variate=True, sharded=None, slice_rows=1024):
    """Price a portfolio (list of dict rows).  Returns one result dict per row, in row order:
    {"id", "price", "stderr"} or {"id", "price": None, "error"}.

    `sharded`: None = shard over the default torch.distributed process group when one is initialised.
    `control_variate` applies to the "mc" rows only (euro_price_mc's default, monte_carlo.py:84).
    A book of more than 2*slice_rows rows is planned and launched slice by slice: a worker thread runs the
    (blocking, GIL-free) optp_mc_book call of slice k while this thread parses slice k+1, so the device
    starts after the first slice is parsed and the host work hides behind the kernels."""
    import concurrent.futures as cf
    from . import distributed as dd
    from .binomial import crr_batch
    rank, world = dd.world() if sharded in (None, True) else (0, 1)
    n_rows = len(rows)
    results = [None] * n_rows
    crr_groups = {}
    step = n_rows if n_rows <= 2 * slice_rows else int(slice_rows)
    load = [0] * world
    slices = []      # (items, future of the slice's statistics)
    with cf.ThreadPoolExecutor(max_workers=1) as pool:
        for lo in range(0, n_rows, max(step, 1)):
            items = _plan_rows(rows, lo, min(lo + step, n_rows), results, crr_groups, tree_steps=tree_steps,
                               mc_paths=mc_paths, exotic_steps=exotic_steps, exotic_paths=exotic_paths, seed=seed)
            if items:   # descriptors are filled HERE, while the worker thread is inside the previous slice's call
                prep = book_prepare(items, mode=mode, rank=rank, world=world, load=load)
                slices.append((items, pool.submit(book_run, prep)))
        # --- trees: one batch per american flag (queues behind the slices on the device) ---------------
        for american, group in crr_groups.items():
            cols = [np.array([g[c] for g in group], dtype=np.float64) for c in range(1, 8)]
            if world > 1:
                prices = dd.crr_batch(*cols, tree_steps, american=american)
            else:
                prices = crr_batch(*cols, tree_steps, american=american)
            for g, p in zip(group, prices):
                results[g[0]] = {"id": rows[g[0]].get("id", ""), "price": float(p), "stderr": None}
        # --- Monte Carlo: one all-reduce for the whole book, vectorised estimators ------------------------
        if slices:
            stats = np.concatenate([f.result() for _, f in slices], axis=0)
            if world > 1:
                stats = dd.allreduce_sum(stats)
            all_items = [it for items, _ in slices for it in items]
            idx, price, se = estimate_rows(all_items, stats, control_variate)
            for i, p, e in zip(idx, price.tolist(), se.tolist()):
                results[i] = {"id": rows[i].get("id", ""), "price": p, "stderr": e}
    return results


def synthetic_book(n_rows, seed=5, *, exotic_steps=252, n_paths=5_000_000):
    """BASELINE config 5: S0=100, K~U(80,120), T in {.25,.5,1,2}, r=.03, q=0,
    sigma~U(.15,.45), call/put 50/50, one third European ("mc"), one third barrier (up-and-out /
    down-and-out at B = S0*(1 +- U(.1,.4))), one third arithmetic Asian; n_paths base paths each."""
    g = np.random.default_rng(seed)
    rows = []
    for i in range(n_rows):
        K = float(g.uniform(80.0, 120.0))
        T = float(g.choice([0.25, 0.5, 1.0, 2.0]))
        sigma = float(g.uniform(0.15, 0.45))
        kind = "call" if g.random() < 0.5 else "put"
        typ = int(g.integers(0, 3))
        row = dict(id=i, S0=100.0, K=K, T=T, r=0.03, sigma=sigma, q=0.0, kind=kind)
        if typ == 0:
            row.update(method="mc")
        elif typ == 1:
            up = g.random() < 0.5
            B = 100.0 * (1.0 + float(g.uniform(0.1, 0.4))) if up else 100.0 * (1.0 - float(g.uniform(0.1, 0.4)))
            row.update(method="exotic", exotic_type="barrier", barrier=B,
                       barrier_type="up-and-out" if up else "down-and-out", n_steps=exotic_steps, n_paths=n_paths)
        else:
            row.update(method="exotic", exotic_type="asian", average_type="arithmetic", strike_type="fixed",
                       n_steps=exotic_steps, n_paths=n_paths)
        rows.append(row)
    return rows
