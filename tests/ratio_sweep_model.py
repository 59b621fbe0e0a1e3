"""numpy model of the ratio-driven match-only sweep of sfm_opencv_b200/csrc/match_knn.cu (kPrune) --
TEST INFRASTRUCTURE: it restates the kernel's DECISION RULE (which columns a row may skip, which rows
are handed to the exact recheck), not the reference; the reference side of every comparison is
oracle/matching.py (NViewReconstuct.cpp:873-913).

What is modelled, per query row and per column half of every 128-column train tile (one epilogue
thread of the kernel):
  * the first `cold_tiles` tiles are inserted unfiltered;
  * afterwards an 8-column group is looked at only if  min|t|^2(group) - 2 max(q.t)(group) < bound
    for the row (`warp_rows` = 1) or for some row of its 32-row warp (`warp_rows` = 32, as the kernel
    does: the extra inserts only add knowledge);
  * the bound is tightened after every insert with the second best of the current 1024-column window
    and recomputed when a window closes from the row's joint (best, second) over both column halves
    (`partner`: the other half's pair as published at this close, at the previous one, or not at all --
    the kernel reads it without synchronisation) and the ratio test:
        F (fails with what is known)  bound = ratio^2 * D0     (only a passing new nearest neighbour matters)
        P (passes with what is known) bound = D0 / ratio^2     (better neighbours and what makes it fail)
  * a row that ends in state P with D0 / ratio^2 above the smallest F bound it ever skipped under is
    flagged and recomputed exactly (recheck_rows_kernel).
The claim the tests check: after the recheck, the two filter passes of match_features give the same
match list, distances and min_dist as on the exact kNN table.
"""
from __future__ import annotations

import numpy as np

TILE = 128
GROUP = 8
IDX_BITS = 20
INF_KEY = np.int64(1) << 60
K_NONE = (1 << 21) - 1


def _pack(v, i):
    return v.astype(np.int64) * (1 << IDX_BITS) + i


def _val(key):
    return key >> IDX_BITS


def _top2(keys):
    """Two smallest DISTINCT keys per row (a key names a column: inserting a column twice changes
    nothing, as in the kernel's knock-out insert)."""
    first = keys.min(1)
    second = np.where(keys > first[:, None], keys, INF_KEY).min(1)
    return np.stack([first, second], 1)


def ratio_sweep(q_u8, t_u8, ratio=0.6, cold_tiles=2, win_tiles=8, warp_rows=32, partner="fresh", seed=0,
                recheck=True, seed_group=False):
    """Returns (d2[Nq,2] int64 squared distances, idx[Nq,2], flagged[Nq] bool, inserted_groups)."""
    qi, ti = q_u8.astype(np.int64), t_u8.astype(np.int64)
    nq, nt = len(qi), len(ti)
    assert nt >= 2 and nt < (1 << IDX_BITS)
    rng = np.random.default_rng(seed)
    tn, qn = (ti * ti).sum(1), (qi * qi).sum(1)
    dots = qi @ ti.T
    val = tn[None, :] - 2 * dots
    ntiles = -(-nt // TILE)
    ratio2 = np.float32(ratio * ratio * (1.0 + 1e-5))
    inv_ratio2 = np.float32((1.0 + 1e-5) / (ratio * ratio))

    G = np.full((nq, 2, 2), INF_KEY)       # [row, part, rank]: finished windows
    W = np.full((nq, 2, 2), INF_KEY)       # current window
    bv = np.full((nq, 2), np.int64(1) << 40)
    smin = np.full((nq, 2), np.int64(1) << 40)
    pub = np.full((nq, 2, 2), np.int64(K_NONE))     # what each half has published (values)
    pub_prev = pub.copy()
    has_pub = np.zeros((nq, 2), bool)
    has_prev = np.zeros((nq, 2), bool)
    inserted = 0

    def close():
        nonlocal pub, pub_prev, has_pub, has_prev
        for p in range(2):
            G[:, p] = _top2(np.concatenate([G[:, p], W[:, p]], 1))
            W[:, p] = INF_KEY
        own = np.minimum(_val(G), K_NONE)                            # [row, part, rank]
        new_pub, new_has = own.copy(), np.ones((nq, 2), bool)
        for p in range(2):
            o = 1 - p
            if partner == "fresh":
                seen, ok = new_pub[:, o], np.ones(nq, bool)
            elif partner == "stale":
                seen, ok = pub[:, o], has_pub[:, o]
            elif partner == "none":
                seen, ok = pub[:, o], np.zeros(nq, bool)
            else:                                                    # per row and close: any of the three
                c = rng.integers(0, 3, nq)
                seen = np.where((c == 0)[:, None], new_pub[:, o], pub[:, o])
                ok = np.where(c == 0, True, np.where(c == 1, has_pub[:, o], False))
            j = np.sort(np.concatenate([own[:, p], np.where(ok[:, None], seen, K_NONE)], 1), 1)
            j1, j2 = j[:, 0], j[:, 1]
            bound = _val(G[:, p, 1])
            known = j2 < K_NONE
            d0 = (j1 + qn).astype(np.float32)
            d1 = (j2 + qn).astype(np.float32)
            fails = d0 > ratio2 * d1
            tf = (ratio2 * d0).astype(np.int64) + 1 - qn
            tp = (d0 * inv_ratio2).astype(np.int64) + 2 - qn
            b2 = np.minimum(bound, j2 + 1)
            b2 = np.where(fails, np.minimum(b2, tf), np.minimum(b2, tp))
            smin[:, p] = np.where(known & fails, np.minimum(smin[:, p], tf), smin[:, p])
            bv[:, p] = np.where(known, b2, bound)
        pub_prev, has_prev = pub, has_pub
        pub, has_pub = new_pub, new_has

    if seed_group:
        # SFM_PRUNE_SEED: no unfiltered tiles at all.  The first 8 columns of the thread's share of tile 0
        # give it two known columns, the bound follows from them alone (no exchange with the other half),
        # and the filtered sweep starts at tile 0 (re-inserting a column is idempotent: keys are unique)
        cold_tiles = 0
        for p in range(2):
            c0 = p * 64
            c1 = min(c0 + GROUP, nt)
            if c1 > c0:
                keys = _pack(val[:, c0:c1], np.arange(c0, c1)[None, :])
                W[:, p] = _top2(np.concatenate([W[:, p], keys], 1))
            v1, v2 = np.minimum(_val(W[:, p, 0]), K_NONE), np.minimum(_val(W[:, p, 1]), K_NONE)
            known = v2 < K_NONE
            d0, d1 = (v1 + qn).astype(np.float32), (v2 + qn).astype(np.float32)
            fails = d0 > ratio2 * d1
            tf = (ratio2 * d0).astype(np.int64) + 1 - qn
            tp = (d0 * inv_ratio2).astype(np.int64) + 2 - qn
            smin[:, p] = np.where(known & fails, np.minimum(smin[:, p], tf), smin[:, p])
            bv[:, p] = np.where(known, np.minimum(v2, np.where(fails, tf, tp)), v2)
    for tile in range(ntiles):
        cold = tile < cold_tiles
        for p in range(2):
            for g in range(TILE // 2 // GROUP):
                c0 = tile * TILE + p * 64 + g * GROUP
                c1 = min(c0 + GROUP, nt)
                if c1 <= c0:
                    continue
                if cold:
                    ins = np.ones(nq, bool)
                else:
                    lb = tn[c0:c1].min() - 2 * dots[:, c0:c1].max(1)
                    ins = lb < bv[:, p]
                    if warp_rows > 1 and ins.any():
                        pad = (-nq) % warp_rows
                        w = np.concatenate([ins, np.zeros(pad, bool)]).reshape(-1, warp_rows).any(1)
                        ins = np.repeat(w, warp_rows)[:nq]
                if not ins.any():
                    continue
                inserted += int(ins.sum())
                keys = _pack(val[ins, c0:c1], np.arange(c0, c1)[None, :])
                W[ins, p] = _top2(np.concatenate([W[ins, p], keys], 1))
                if not cold:
                    bv[ins, p] = np.minimum(bv[ins, p], _val(W[ins, p, 1]))
        last = tile + 1 == ntiles
        if last or (tile + 1) % win_tiles == 0 or tile + 1 == min(cold_tiles, ntiles):
            close()

    fin = _top2(G.reshape(nq, 4))
    v, idx = _val(fin), fin & ((1 << IDX_BITS) - 1)
    d0 = (v[:, 0] + qn).astype(np.float32)
    d1 = (v[:, 1] + qn).astype(np.float32)
    s = smin.min(1)
    flagged = (d0 <= ratio2 * d1) & ((d0 * inv_ratio2).astype(np.int64) + 2 - qn > s)
    d2 = v + qn[:, None]
    if recheck and flagged.any():           # recheck_rows_kernel: exact top-2 of the whole train image
        ex = _top2(_pack(val[flagged], np.arange(nt)[None, :]))
        d2[flagged] = _val(ex) + qn[flagged, None]
        idx[flagged] = ex & ((1 << IDX_BITS) - 1)
    return d2, idx.astype(np.int32), flagged, inserted


def undecidable_case(nq=64, nt=1024, seed=3, mid=600, near=900):
    """Rows the sweep cannot decide: every train row is exactly D = 51 200 away (state F after the cold
    tiles, bound ratio^2 D = 18 432), except `mid` at 25 600 -- skipped under that bound -- and `near` at
    14 400, which is looked at, passes against the known second (D) and fails against `mid`: only the
    exact recheck of the row can tell.  Returns (query, train); every query row is such a row."""
    rng = np.random.default_rng(seed)
    q = np.zeros((nq, 128), np.uint8)
    q[:, :64] = 100
    t = np.zeros((nt, 128), np.uint8)
    t[:, :64] = 100
    for r in range(nt):
        t[r, 64 + rng.choice(64, 32, replace=False)] = 40     # 32 * 1600
    t[mid, 64:] = 0
    t[mid, 64:64 + 16] = 40
    t[near, 64:] = 0
    t[near, 64:64 + 9] = 40
    return q, t
