"""Development tool (CPU, numpy): replay of the kNN-2 epilogue's group filter on one image pair.

Counts, per epilogue warp (32 query rows x one 64-column half of every 128-column tile), how many
8-column groups pass the filter for SOME row of the warp ("hit groups" -- each costs the whole warp an
exact insert), under different column orders / bound-sharing rules.  Used to decide which epilogue
variants are worth GPU time; DESIGN.md 4.1 quotes its numbers.

  python tools/sim_filter.py [n_desc] [order]      order: orig | hub | hubwin
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle.synth import image_bank  # noqa: E402  (input generator only)


def replay(q, t, cold_tiles=8, win_tiles=8, share_every=8, exact_norm=False, ratio_aware=False, verbose=True):
    """q, t uint8 [n,128].  Returns dict of counters."""
    nq, nt = len(q), len(t)
    qi, ti = q.astype(np.int64), t.astype(np.int64)
    tn = (ti * ti).sum(1)
    qn = (qi * qi).sum(1)
    ntiles = nt // 128
    INF = np.int64(1 << 40)
    # state per (row, chalf)
    m1 = np.full((nq, 2), INF)
    m2 = np.full((nq, 2), INF)
    bv = np.full((nq, 2), INF)
    hit_groups = 0
    row_hits = 0
    groups_total = 0
    true_updates = 0
    per_tile = []
    dots_t = (qi @ ti.T)  # [nq, nt] int64 (8192^2 * 8 B = 512 MB: fine here)
    val = tn[None, :] - 2 * dots_t          # value = |t|^2 - 2 q.t
    del dots_t
    for tile in range(ntiles):
        cold = tile < cold_tiles
        hg_tile = 0
        for ch in range(2):
            c0 = tile * 128 + ch * 64
            for g in range(8):
                cols = slice(c0 + g * 8, c0 + g * 8 + 8)
                v = val[:, cols]                                  # [nq, 8]
                if exact_norm:
                    gbound = v.min(1)
                else:
                    gbound = tn[cols].min() - 2 * ((tn[cols][None, :] - v) // 2).max(1)
                passed = gbound < bv[:, ch]
                if cold:
                    hit_w = np.ones(nq // 32, bool)
                else:
                    hit_w = passed.reshape(-1, 32).any(1)
                    hit_groups += int(hit_w.sum())
                    row_hits += int(passed.sum())
                    groups_total += nq // 32
                    hg_tile += int(hit_w.sum())
                hit_r = np.repeat(hit_w, 32)
                if hit_r.any():
                    vv = v[hit_r]
                    a1, a2 = m1[hit_r, ch], m2[hit_r, ch]
                    allv = np.concatenate([vv, a1[:, None], a2[:, None]], 1)
                    allv.sort(1)
                    if not cold:
                        true_updates += int((allv[:, 1] < a2).sum())
                    m1[hit_r, ch] = allv[:, 0]
                    m2[hit_r, ch] = allv[:, 1]
                    if not cold:
                        nb = np.minimum(bv[hit_r, ch], allv[:, 1])
                        if ratio_aware:
                            pass
                        bv[hit_r, ch] = nb
        per_tile.append(hg_tile)
        if (tile + 1) % share_every == 0 or tile + 1 == cold_tiles:
            # joint bound over both column halves: second smallest of the four
            allv = np.concatenate([m1, m2], 1)
            allv.sort(1)
            joint = allv[:, 1]
            b = np.minimum(m2, joint[:, None] + 1)
            if ratio_aware:
                # rows that fail the ratio test for sure need only a new best
                d0 = np.sqrt((allv[:, 0] + qn).astype(np.float64))
                d1 = np.sqrt((allv[:, 1] + qn).astype(np.float64))
                fail = d0 > 0.6 * d1 * (1 + 1e-6)
                bb = np.where(fail, allv[:, 0], joint + 1)
                b = np.minimum(b, bb[:, None])
            bv = b.copy() if tile + 1 >= cold_tiles else bv
    hot_tiles = ntiles - cold_tiles
    res = dict(hit_groups_per_warp_tile=hit_groups / max(1, groups_total) * 8,
               rows_per_hit=row_hits / max(1, hit_groups),
               row_hits_per_thread=row_hits / (nq * 2),
               true_updates_per_thread=true_updates / (nq * 2), hot_tiles=hot_tiles)
    if verbose:
        print({k: round(float(x), 3) for k, x in res.items()})
    return res, per_tile


def hub_order(t, mu, mode):
    ti = t.astype(np.int64)
    score = ti @ mu - 0.5 * (ti * ti).sum(1)
    order = np.argsort(-score, kind="stable")
    if mode == "hubwin":           # only the first 1024 by score, the rest in original order
        first = np.sort(order[:1024])
        rest = np.sort(order[1024:])
        return np.concatenate([first, rest])
    # full sort by score into 1024-windows, each window internally in original order
    out = []
    for w in range(0, len(order), 1024):
        out.append(np.sort(order[w:w + 1024]))
    return np.concatenate(out)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    order = sys.argv[2] if len(sys.argv) > 2 else "orig"
    bank = image_bank(3, n)
    q, t = bank[0], bank[2]
    if order != "orig":
        mu = t.astype(np.float64).mean(0)
        t = t[hub_order(t, mu, order)]
    for kw in (dict(), dict(exact_norm=True), dict(share_every=4), dict(ratio_aware=True)):
        print(order, kw, end=" ")
        replay(q, t, **kw)
