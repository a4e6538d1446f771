"""Host-side logic of the book pricer (optpricer_b200.book): row parsing in the reference's CSV
schema (scripts/price_book.py:9-18), grouping into device work, per-row failures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"))

from optpricer_b200 import book  # noqa: E402


ROWS = [
    dict(id="1", S0="100", K="110", T="0.5", r="0.05", sigma="0.20", q="0.0", kind="call", method="bs"),
    dict(id="2", S0="100", K="95", T="1.0", r="0.05", sigma="0.25", q="0.01", kind="put", method="mc"),
    dict(id="3", S0="100", K="105", T="0.5", r="0.05", sigma="0.20", q="0.0", kind="call", method="binomial"),
    dict(id="4", S0="100", K="100", T="1.0", r="0.05", sigma="0.20", q="0.0", kind="call", method="exotic",
         exotic_type="barrier", barrier="120", barrier_type="up-and-out"),
    dict(id="5", S0="100", K="100", T="1.0", r="0.05", sigma="0.20", q="0.0", kind="put", method="exotic",
         exotic_type="asian"),
    dict(id="6", S0="100", K="100", T="1.0", r="0.05", sigma="0.20", q="0.0", kind="put", method="binomial",
         american="true"),
    dict(id="7", S0="100", K="100", T="1.0", r="0.05", sigma="0.20", kind="call", method="fem"),
    dict(id="8", S0="-1", K="100", T="1.0", r="0.05", sigma="0.20", kind="call", method="mc"),
    dict(id="9", S0="100", K="100", T="1.0", r="0.05", sigma="0.20", kind="call", method="exotic",
         exotic_type="barrier", barrier="120", barrier_type="sideways"),
    dict(id="10", S0="100", K="100", T="1.0", r="0.05", sigma="0.30", q="0.0", kind="call", method="exotic",
         exotic_type="lookback", n_steps="50", n_paths="1000"),
]


def test_plan_groups_rows_and_keeps_failures():
    results, crr_groups, items = book.plan_book(ROWS)
    # closed-form row priced on the host (BS call 100/110/0.5/0.05/0.2 = 2.9057...)
    assert results[0]["price"] == pytest.approx(2.9057, abs=2e-3)
    # failures carry the reference's messages and do not abort the book
    assert results[6] == {"id": "7", "price": None, "error": "Unknown method: 'fem'"}
    assert results[7]["error"] == "S0 must be positive, got -1.0"
    assert "barrier_type must be one of" in results[8]["error"]
    assert sorted(crr_groups) == [False, True] and len(crr_groups[False]) == 1 and crr_groups[True][0][0] == 5
    kinds = [(it["kind"], [r for r, _ in it["rows"]]) for it in items]
    # rows 4 and 5 share (S0, r, q, sigma, T, n_steps, n_paths) -> ONE simulated path set, two payoffs
    assert kinds == [("terminal", [1]), ("path", [3, 4]), ("path", [9])]
    assert items[1]["n_steps"] == 500 and items[1]["n_base"] == 100_000 and items[2]["n_steps"] == 50
    assert all(results[i] is None for i in (1, 2, 3, 4, 5, 9))


def test_more_than_sixteen_payoffs_split_into_launches():
    rows = [dict(id=i, S0=100, K=90 + i, T=1.0, r=0.05, sigma=0.2, kind="call", method="exotic", exotic_type="digital")
            for i in range(40)]
    _, _, items = book.plan_book(rows)
    assert [len(it["rows"]) for it in items] == [16, 16, 8]


def test_synthetic_book_composition():
    rows = book.synthetic_book(3000, n_paths=1000)
    methods = [r["method"] for r in rows]
    assert 800 < methods.count("mc") < 1200
    ex = [r for r in rows if r["method"] == "exotic"]
    assert {r["exotic_type"] for r in ex} == {"barrier", "asian"}
    for r in ex:
        if r["exotic_type"] == "barrier":
            assert (r["barrier"] > 100) == (r["barrier_type"] == "up-and-out")
    assert rows == book.synthetic_book(3000, n_paths=1000)


def test_shard_items_by_item_for_large_books_and_by_paths_for_small_ones():
    _, _, items = book.plan_book(book.synthetic_book(400, n_paths=1000), mc_paths=5000)
    world = 8
    shares = [book.shard_items(items, r, world) for r in range(world)]
    cost = [it["n_base"] * (1 if it["kind"] == "terminal" else it["n_steps"]) for it in items]
    loads = []
    for sh in shares:
        assert len(sh) == len(items)
        loads.append(sum(c for c, (lo, hi) in zip(cost, sh) if hi > lo))
    # every item runs whole on exactly one rank, and the ranks are balanced to within one item
    for i, it in enumerate(items):
        owners = [r for r in range(world) if shares[r][i] != (0, 0)]
        assert len(owners) == 1 and shares[owners[0]][i] == (0, it["n_base"])
    assert max(loads) - min(loads) <= max(cost)
    # few items: every rank takes a disjoint range of every item's base paths
    few = items[:5]
    parts = [book.shard_items(few, r, world) for r in range(world)]
    for i, it in enumerate(few):
        edges = [parts[r][i] for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == it["n_base"]
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
    assert book.shard_items(items, 0, 1) == [(0, it["n_base"]) for it in items]


def test_vectorised_estimators_equal_the_scalar_ones_bit_for_bit():
    """price_book evaluates the estimators of monte_carlo.py:142-167 / exotics.py:22-27 on arrays; the numbers
    must be the ones the row-by-row API returns."""
    import numpy as np
    from optpricer_b200 import exotics as ex, monte_carlo as mc
    g = np.random.default_rng(3)
    items, blocks = [], []
    for j in range(200):
        n = float(g.integers(2, 10**7))
        x = g.uniform(0.0, 20.0)
        y = g.uniform(50.0, 150.0)
        sx, sy = n * x, n * y
        sxx = n * (x * x + g.uniform(0.0, 30.0))
        syy = n * (y * y + g.uniform(0.0, 400.0))
        sxy = n * (x * y + g.uniform(-20.0, 60.0))
        if j % 2:
            items.append(dict(kind="terminal", rows=[(j, None)], S0=g.uniform(50, 150), q=g.uniform(0, 0.05), T=g.uniform(0.1, 2)))
        else:
            items.append(dict(kind="path", rows=[(j, None)], S0=100.0, q=0.0, T=1.0))
        blocks.append([n, sx, sxx, sy, syy, sxy, 0.0, 0.0])
    blocks.append([1.0, 3.0, 9.0, 100.0, 1e4, 300.0, 0.0, 0.0])      # n = 1: stderr 0 for a path row
    items.append(dict(kind="path", rows=[(200, None)], S0=100.0, q=0.0, T=1.0))
    blocks.append([0.0] * 8)                                          # n = 0: NaN
    items.append(dict(kind="terminal", rows=[(201, None)], S0=100.0, q=0.0, T=1.0))
    stats = np.array(blocks)
    for cv in (True, False):
        idx, price, se = book.estimate_rows(items, stats, cv)
        assert idx == list(range(202))
        for it, s, p, e in zip(items, stats, price, se):
            if it["kind"] == "terminal":
                want = mc.estimator((int(s[0]), *s[1:6]), S0=it["S0"], q=it["q"], T=it["T"], control_variate=cv)
            else:
                want = ex.estimate(s)
            assert (p == want[0] or (np.isnan(p) and np.isnan(want[0]))), (it, p, want)
            assert (e == want[1] or (np.isnan(e) and np.isnan(want[1]))), (it, e, want)


def test_item_sharding_carries_the_load_between_slices():
    items = [dict(kind="path", rows=[(i, None)], n_base=1000 + 10 * (i % 7), n_steps=252) for i in range(64)]
    whole = [book.shard_items(items, r, 4) for r in range(4)]
    # every item on exactly one rank
    for i in range(64):
        assert sum(1 for r in range(4) if whole[r][i][1] > 0) == 1
    loads = [sum((hi - lo) * 252 for lo, hi in whole[r]) for r in range(4)]
    assert max(loads) - min(loads) <= 252 * 1060
    # two slices with a carried load balance as well as one pass, identically on every rank
    per_rank = []
    for r in range(4):
        load = [0] * 4
        a = book.shard_items(items[:32], r, 4, load)
        b = book.shard_items(items[32:], r, 4, load)
        per_rank.append((a + b, list(load)))
    assert all(pr[1] == per_rank[0][1] for pr in per_rank)
    assert max(per_rank[0][1]) - min(per_rank[0][1]) <= 252 * 1060
    for i in range(64):
        assert sum(1 for r in range(4) if per_rank[r][0][i][1] > 0) == 1
