"""world_size-2 gloo tests of the sharding logic (optpricer_b200.distributed): the statistics of
disjoint path ranges / contract slices, summed by ONE all-reduce (gathered for CRR), reproduce the
single-process result.  The numpy oracle stands in for the CUDA kernels via the injectable
stats_fn / price_fn, so the host logic is exercised without a GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_terminal_stats(n, off, seed=11):
    """Counter-style stand-in: draw i of the global stream belongs to base path i."""
    from oracle import terminal_mc
    z = np.random.default_rng(seed).standard_normal(1000)[off:off + n]
    s = terminal_mc.chunk_sumstats_from_z(z, S0=100.0, K=105.0, T=1.0, r=0.03, q=0.0, sigma=0.2, kind="call",
                                          antithetic=True)
    return np.array([*s, 0.0, 0.0])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import optpricer_b200 as op
    from optpricer_b200 import distributed as dd, processes as pr, exotics as ex
    from oracle import crr_tree, path_schemes, payoffs
    out = {}
    opt = op.OptionSpec(100.0, 105.0, 1.0, 0.03, 0.2)
    out["euro"] = dd.euro_price_mc(opt, "call", n_paths=1000, seed=11, stats_fn=_oracle_terminal_stats)
    # path payoffs: 2 specs, 301 base paths (ragged split), stats from the oracle on a fixed Z
    (z,) = path_schemes.draw_normals(3, 20, 301)
    paths_kw = dict(S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0)

    def stats_fn(n, off):
        p = path_schemes.gbm_exact(z[:, off:off + n], **paths_kw)
        rows = []
        for pay in (payoffs.barrier_payoff(p, 100.0, "call", 120.0, "up-and-out"), payoffs.asian_payoff(p, 100.0, "put")):
            rows.append([*payoffs.six_stats(pay, p[-1], 0.05, 1.0), 0.0, 0.0])
        return np.array(rows)
    lp = pr.gbm_paths(100.0, 0.05, 0.0, 0.2, 1.0, 20, 301, seed=3)
    specs = [ex.barrier_spec(100.0, "call", 120.0, "up-and-out"), ex.asian_spec(100.0, "put")]
    out["paths"] = dd.price_payoffs(lp, specs, 0.05, 1.0, stats_fn=stats_fn)
    out["paths_cv"] = dd.price_payoffs(lp, specs, 0.05, 1.0, control_variate=True, stats_fn=stats_fn)
    # CRR: 7 contracts over 2 ranks
    S0, K, T, r, qq, sig, is_call = crr_tree.cfg2_contracts(7)
    price_fn = lambda *a: crr_tree.crr_loop(*a[:6], a[6].astype(np.uint8), 50, True)[0]  # noqa: E731
    out["crr"] = dd.crr_batch(S0, K, T, r, qq, sig, is_call, 50, american=True, price_fn=price_fn)
    out["range"] = dd.shard_range(7, rank, world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process():
    import torch.multiprocessing as mp
    from oracle import terminal_mc, crr_tree, path_schemes, payoffs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth
    z = np.random.default_rng(11).standard_normal(1000)
    s = terminal_mc.chunk_sumstats_from_z(z, S0=100.0, K=105.0, T=1.0, r=0.03, q=0.0, sigma=0.2, kind="call",
                                          antithetic=True)
    want_euro = terminal_mc.estimator(s, S0=100.0, q=0.0, T=1.0, control_variate=True)
    (zz,) = path_schemes.draw_normals(3, 20, 301)
    # sharded columns are [base_lo.., mirrors_lo..] per rank: same multiset of paths as the full set
    p = path_schemes.gbm_exact(zz, S0=100.0, r=0.05, q=0.0, sigma=0.2, T=1.0)
    want_b = payoffs.barrier(p, 100.0, 0.05, 1.0, "call", 120.0, "up-and-out")
    want_a = payoffs.asian(p, 100.0, 0.05, 1.0, "put")
    want_cv = payoffs.cv_estimator(payoffs.six_stats(payoffs.asian_payoff(p, 100.0, "put"), p[-1], 0.05, 1.0),
                                   100.0 * np.exp(0.05) * np.exp(-0.05))
    S0, K, T, r, qq, sig, is_call = crr_tree.cfg2_contracts(7)
    want_crr, _ = crr_tree.crr_loop(S0, K, T, r, qq, sig, is_call, 50, True)
    for rank in (0, 1):
        out = results[rank]
        assert out["euro"] == pytest.approx(want_euro, rel=1e-12)
        assert out["paths"][0] == pytest.approx(want_b, rel=1e-11)
        assert out["paths"][1] == pytest.approx(want_a, rel=1e-11)
        assert out["paths_cv"][1] == pytest.approx(want_cv, rel=1e-10)
        np.testing.assert_allclose(out["crr"], want_crr, rtol=0, atol=0)
    assert results[0]["range"] == (0, 4) and results[1]["range"] == (4, 7)


def test_shard_ranges_partition():
    sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))
    from optpricer_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1_000_003):
        for size in (1, 2, 3, 8):
            spans = [shard_range(n, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
