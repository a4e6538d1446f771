#!/usr/bin/env python3
"""Option-level emulation of the drifting-window two-zone CRR walk of csrc/crr_batch.cu (OPTP_ZONE_DRIFT) in plain
unscaled numpy arithmetic: it tests the LOGIC -- window placement under the exercise boundary, A / B levels, nodes
below node 0, the margin above the live top, the per-level bottom and top tests, the final read -- not the rounding,
and it is where the schedule (Z width, nodes below the boundary, segment lengths) was chosen before any GPU time was
spent.  `exact` restates binomial.py:7-38 for a put.
Usage: python scripts/crr_drift_emu.py N n_contracts cfg2|wide ["seg0=32,seg=64,..."]   (tests/test_crr_drift_emulation_cpu.py)"""
import sys, numpy as np

def exact(S0, K, T, r, q, sg, N):
    dt = T / N; u = np.exp(sg * np.sqrt(dt)); d = 1 / u
    disc = np.exp(-r * dt); p = (np.exp((r - q) * dt) - d) / (u - d)
    j = np.arange(N + 1)
    V = np.maximum(K - S0 * u ** (2.0 * j - N), 0.0)
    for k in range(N - 1, -1, -1):
        cont = disc * (p * V[1:k + 2] + (1 - p) * V[:k + 1])
        V = np.maximum(cont, K - S0 * u ** (2.0 * j[:k + 1] - k))
    return V[0]

def drift_walk(S0, K, T, r, q, sg, N, ZW=16, P=13, seg0=32, seg=44, maxM=6, stats=None, CH=32):
    dt = T / N; u = np.exp(sg * np.sqrt(dt)); d = 1 / u
    disc = np.exp(-r * dt); p = (np.exp((r - q) * dt) - d) / (u - d)
    a, b = disc * p, disc * (1 - p)
    ln_u = np.log(u); thr = K * 2.0 ** -80
    E = lambda j, k: K - S0 * np.exp(ln_u * (2.0 * j - k))
    jz = 0.5 * (N + np.log(K / S0) / ln_u)
    # q > 0: a node can only choose exercise if S < S1 = K(1-disc)/(1-e^{-q dt}); on the levels k < k_pin even node 0
    # lies above S1 (the kernel's regime without exercise nodes: window from node 0, no per-level tests)
    k_pin = 0
    if q > 0.0:
        S1 = K * (-np.expm1(-r * dt)) / (-np.expm1(-q * dt))
        k_pin = int(min(max(np.ceil(-np.log(S1 / S0) / ln_u - 0.01), 0.0), N + 1))
    w0 = int(min(np.floor(jz) - (ZW - 4), N + 1 - ZW))
    k = N; M = 0
    W = ZW
    idx = w0 + np.arange(W)
    v = np.maximum(E(idx, N), 0.0)
    exm = E(idx, N) > 0
    nodes_z = nodes_c = 0; nseg = 0
    def measure(v, exm, w0, W, k):
        cont = ~exm[:ZW]
        jlo = w0 + int(np.argmax(cont)) if cont.any() else w0 + ZW
        idx = w0 + np.arange(W)
        live = (v > thr) & (idx <= k)
        jhi = w0 + (int(np.nonzero(live)[0].max()) + 1 if live.any() else 0)
        jhi = max(jhi, jlo)
        return jlo, jhi
    jlo, jhi = measure(v, exm, w0, W, k)
    if jlo <= w0: return None, 'bottom0'
    while k > 0:
        nlev = min(seg0 if k == N else seg, k)
        nB = nlev >> 1
        top = min(jhi, k + 1)
        pinned = k < k_pin
        w0n = 0 if pinned else jlo - P
        need = 0
        if top + nB > w0n + ZW:
            need = (top + nB - (w0n + ZW) + CH - 1) // CH
        if need > maxM: return None, 'wide'
        Wn = ZW + CH * need
        # re-pack
        idxn = w0n + np.arange(Wn)
        jr = idxn - w0
        vn = np.zeros(Wn)
        below = jr < 0
        vn[below] = E(idxn[below], k)
        inside = (~below) & (jr < W) & (idxn < top)
        vn[inside] = v[jr[inside]]
        v = vn; w = w0n; W = Wn; M = need
        D = k - w0n - (ZW - 1)
        n_chk = 0 if pinned else int(np.clip(2 * D, 0, nlev))
        ok = True
        for t in range(1, nlev + 1):
            kk = k - t
            if t & 1:   # A step: origin unchanged
                up = np.append(v[1:], 0.0)
                cont = a * up + b * v
            else:       # B step: origin - 1
                eb = K if pinned else E(w - 1, kk + 1)   # (pinned: anything above the tree's values, never used by it)
                low = np.append(eb, v[:-1])
                cont = a * v + b * low
                w -= 1
            idx = w + np.arange(W)
            ex = E(idx, kk)
            exz = ex[:ZW] > cont[:ZW]
            v = cont.copy()
            v[:ZW] = np.where(exz, ex[:ZW], cont[:ZW])
            if not exz[0] and not pinned: ok = False; why = 'bottom'
            if exz[ZW - 1] and t <= n_chk: ok = False; why = 'top'
            nodes_z += ZW; nodes_c += W - ZW
        if not ok: return None, why
        k -= nlev; w0 = w; nseg += 1
        idx = w0 + np.arange(W)
        exm = np.zeros(W, bool); exm[:ZW] = v[:ZW] == E(idx[:ZW], k)
        jlo, jhi = measure(v, exm, w0, W, k)
        if jlo <= w0 and k >= k_pin: return None, 'bottom_m'
    if stats is not None: stats.append((nodes_z, nodes_c, nseg))
    pos = -w0
    if pos < 0: return K - S0, 'ok'
    if pos >= W: return 0.0, 'ok'
    return v[pos], 'ok'

if __name__ == '__main__':
    N = int(sys.argv[1]); n = int(sys.argv[2]); mode = sys.argv[3]
    kw = eval('dict(%s)' % (sys.argv[4] if len(sys.argv) > 4 else ''))
    g = np.random.default_rng(0)
    fails = {}; worst = 0; stats = []
    for i in range(n):
        if mode == 'cfg2':
            S0 = g.uniform(50, 150); K = S0 * g.uniform(0.8, 1.2); T = g.uniform(0.1, 2.0); sg = g.uniform(0.1, 0.6); r = 0.03; q = 0.0
        else:
            S0 = g.uniform(50, 150); K = S0 * g.uniform(0.5, 1.6); T = g.uniform(0.02, 6.0)
            sg = g.uniform(0.04, 1.2); r = g.choice([0.002, 0.01, 0.03, 0.08, 0.2]); q = g.choice([0.0, 0.0, 0.01, 0.02])
            if q >= r: continue
        dtt = T / N; uu = np.exp(sg * np.sqrt(dtt)); pp = (np.exp((r - q) * dtt) - 1 / uu) / (uu - 1 / uu)
        if not 0 < pp < 1: continue
        pr, why = drift_walk(S0, K, T, r, q, sg, N, stats=stats, **kw)
        if pr is None:
            fails[why] = fails.get(why, 0) + 1; continue
        ex = exact(S0, K, T, r, q, sg, N)
        err = abs(pr - ex) / max(abs(ex), 1e-8 * K)
        if err > worst: worst = err; print('worst so far', err, pr, ex, (S0, K, T, r, q, sg))
    st = np.array(stats)
    print('fails', fails, 'worst rel err', worst, 'priced', len(stats))
    print('mean Z nodes %.0f C nodes %.0f segs %.1f' % tuple(st.mean(0)))
