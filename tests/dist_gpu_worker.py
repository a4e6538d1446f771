"""Worker of tests/test_multigpu_gpu.py: one process per GPU (torchrun), NCCL process group.

Checks, on every rank, that the sharded entry points of optpricer_b200.distributed reproduce the
single-device results computed by the same rank -- the reference's own multi-worker mode is bit-identical
to its serial mode (monte_carlo.py:126-139 with the chunk seeds of :120,131; _aggregate_stats :67-74):
  * Monte Carlo statistics blocks: sharded == single to 1e-9 relative (fp64 sums in another order),
  * CRR prices and item-sharded book prices: bit-identical,
  * the C-ABI collective (optp_allreduce_stats / _dev) == torch.distributed.all_reduce.
Rank 0 writes a JSON report to argv[1]."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import optpricer_b200 as op
    from optpricer_b200 import _native as nv, book as bk, distributed as dd, exotics as ex, processes as pr
    from optpricer_b200 import binomial as bn
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_io import dupire_grid_arrays
    lib = nv.init(local)
    report = {"world": size, "checks": []}

    def note(name, ok, **kw):
        report["checks"].append(dict(name=name, ok=bool(ok), **kw))

    assert dd.init_comm()
    r_, w_ = C.c_int(), C.c_int()
    lib.optp_comm_info(C.byref(r_), C.byref(w_))
    note("comm_info", (r_.value, w_.value) == (rank, size))

    # --- the collective itself: host and device variants against torch's NCCL all-reduce
    g = np.random.default_rng(100 + rank)
    block = g.standard_normal((37, nv.NSTATS))
    t = torch.from_numpy(block.copy()).cuda()
    dist.all_reduce(t)
    want = t.cpu().numpy()
    got_host = dd.allreduce_sum(block)
    note("allreduce_host_vs_torch", np.array_equal(got_host, want), maxdiff=float(np.abs(got_host - want).max()))
    td = torch.from_numpy(block.copy()).cuda()
    st = torch.cuda.current_stream()
    nv.check(lib.optp_allreduce_stats_dev(C.c_void_p(td.data_ptr()), td.numel(), C.c_void_p(st.cuda_stream)),
             "optp_allreduce_stats_dev")
    torch.cuda.synchronize()
    note("allreduce_dev_vs_torch", np.array_equal(td.cpu().numpy(), want))

    # --- Monte Carlo: sharded statistics == single-device statistics
    tab = pr.SigmaTable(*dupire_grid_arrays())
    n = 4_000_001   # odd: ranks get different counts
    sets = {
        "gbm": (pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 252, n, seed=5),
                [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.asian_spec(100.0, "call"),
                 ex.lookback_spec("call"), ex.vanilla_spec(100.0, "put")]),
        "heston": (pr.heston_paths(100.0, 0.05, 0.0, 0.04, 2.0, 0.04, 0.3, -0.7, 1.0, 365, n // 4, seed=6),
                   [ex.asian_spec(100.0, "call"), ex.vanilla_spec(100.0, "call")]),
        "lvm": (pr.milstein_local_vol_paths(100.0, 0.05, 0.02, 1.0, 252, n // 4, tab, seed=7),
                [ex.barrier_spec(100.0, "call", 130.0, "up-and-out"), ex.vanilla_spec(100.0, "call")]),
        "gbm_fp64": (pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 50, n // 8, seed=8, mode="fp64"),
                     [ex.barrier_spec(100.0, "call", 130.0, "up-and-out")]),
    }
    for name, (lp, specs) in sets.items():
        single, _, _ = ex.path_stats(lp, specs, 0.05, 1.0)
        lo, hi = dd.shard_range(lp.n_paths, rank, size)
        local_stats, _, _ = ex.path_stats(lp, specs, 0.05, 1.0, n_base=hi - lo, path_offset=lo)
        total = dd.allreduce_sum(local_stats)
        relerr = float(np.max(np.abs(total[:, :7] - single[:, :7]) / np.maximum(np.abs(single[:, :7]), 1e-300)))
        note("mc_stats_" + name, relerr <= 1e-9 and np.array_equal(total[:, 0], single[:, 0]), relerr=relerr)
        sharded = dd.price_payoffs(lp, specs, 0.05, 1.0, control_variate=True)
        alone = ex.price_payoffs(lp, specs, 0.05, 1.0, control_variate=True)
        d = max(abs(a[0] - b[0]) / max(1.0, abs(b[0])) for a, b in zip(sharded, alone))
        note("mc_price_" + name, d <= 1e-9, reldiff=d)
    opt = op.OptionSpec(100.0, 110.0, 1.0, 0.03, 0.2)
    a = dd.euro_price_mc(opt, "call", n_paths=3_000_001, seed=1)
    b = op.euro_price_mc(opt, "call", n_paths=3_000_001, seed=1)
    # (the terminal kernel keeps fp32 partial sums over 64 effective paths before folding them into fp64, and which
    # paths share a partial sum depends on the grid: sharded and single agree to ~1e-9, not to fp64 rounding; the
    # control-variate variance is a difference of nearly equal numbers, so the standard error sees it amplified)
    note("euro_price_mc", abs(a[0] - b[0]) <= 1e-8 * b[0] and abs(a[1] - b[1]) <= 1e-7 * b[1], sharded=a, single=b)

    # --- CRR: contracts shard with no exchange; bit-identical prices
    gg = np.random.default_rng(0)
    m = 200_003
    S0 = gg.uniform(50.0, 150.0, m)
    K = S0 * gg.uniform(0.8, 1.2, m)
    T = gg.uniform(0.1, 2.0, m)
    sig = gg.uniform(0.1, 0.6, m)
    is_call = (gg.random(m) < 0.3).astype(np.uint8)
    q = np.where(is_call == 1, 0.01, 0.0)
    sharded = dd.crr_batch(S0, K, T, 0.03, q, sig, is_call, 1000, american=True)
    alone = bn.crr_batch(S0, K, T, 0.03, q, sig, is_call, 1000, american=True)
    note("crr_bit_identical", np.array_equal(sharded, alone), n=m)

    # --- book: item-sharded (many items) bit-identical; path-sharded (few items) to 1e-9
    rows = bk.synthetic_book(120, n_paths=200_000)
    for r in rows[:10]:
        r2 = dict(r, id=f"tree{r['id']}", method="binomial", american="true")
        rows.append(r2)
    sharded = bk.price_book(rows, mc_paths=200_000, tree_steps=500)
    alone = bk.price_book(rows, mc_paths=200_000, tree_steps=500, sharded=False)
    same = all(x["price"] == y["price"] and x.get("stderr") == y.get("stderr") for x, y in zip(sharded, alone))
    note("book_item_sharded_bit_identical", same, rows=len(rows))
    few = rows[:3]
    sharded = bk.price_book(few, mc_paths=1_000_000)
    alone = bk.price_book(few, mc_paths=1_000_000, sharded=False)
    d = max(abs(x["price"] - y["price"]) / max(1.0, abs(y["price"])) for x, y in zip(sharded, alone))
    note("book_path_sharded", d <= 1e-9, reldiff=d)

    ok = all(c["ok"] for c in report["checks"])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["all_ranks_ok"] = bool(flag.item())
    report["launches_rank0"] = nv.launch_count()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(report, f, indent=1)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
