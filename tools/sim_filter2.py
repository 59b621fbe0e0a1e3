"""Development tool (CPU, numpy): replay of the round-2 kNN-2 epilogue schedule on one image pair.

Schedule of a work item: a COLD pass over the first c train tiles that only tracks, per thread, the
smallest chunk bound  max|t|^2(chunk) - 2 max(q.t)(chunk)  (an upper bound on the value of one real
column), a joint bound over the two column halves of a row, then the filtered sweep over tiles
c..ntiles-1 followed by tiles 0..c-1 again.  Counts per epilogue warp-tile (32 rows x 64 columns):
32-column chunks that pass the chunk-level test for some row, 8-column groups that pass, and the same
for the round-1 schedule (first window unfiltered) for comparison.

  python tools/sim_filter2.py [n_desc] [cold_tiles]
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle.synth import image_bank  # noqa: E402  (input generator only)


def replay(q, t, cold_tiles=8, resweep=True, match_only=True, ratio=0.6, chunk_first=True):
    nq, nt = len(q), len(t)
    qi, ti = q.astype(np.int64), t.astype(np.int64)
    tn = (ti * ti).sum(1)
    qn = (qi * qi).sum(1)
    ntiles = nt // 128
    c = min(cold_tiles, ntiles)
    INF = np.int64(1 << 40)
    dots = qi @ ti.T
    val = tn[None, :] - 2 * dots
    m1 = np.full((nq, 2), INF)
    m2 = np.full((nq, 2), INF)
    # ---- cold pass
    B = np.full((nq, 2), INF)
    if resweep:
        for tile in range(c):
            for ch in range(2):
                for k in range(2):
                    cols = slice(tile * 128 + ch * 64 + k * 32, tile * 128 + ch * 64 + k * 32 + 32)
                    B[:, ch] = np.minimum(B[:, ch], tn[cols].max() - 2 * dots[:, cols].max(1))
        cold_b = B.max(1) + 1
        seq = list(range(c, ntiles)) + list(range(c))
    else:
        # round-1 schedule: the first window is inserted unfiltered
        for tile in range(c):
            for ch in range(2):
                v = val[:, tile * 128 + ch * 64: tile * 128 + ch * 64 + 64]
                allv = np.concatenate([v, m1[:, ch:ch + 1], m2[:, ch:ch + 1]], 1)
                allv.partition(1, axis=1)
                allv = np.sort(allv[:, :2], 1)
                m1[:, ch], m2[:, ch] = allv[:, 0], allv[:, 1]
        cold_b = np.full(nq, INF)
        seq = list(range(c, ntiles))
    bv = np.empty((nq, 2), np.int64)

    def window_end():
        allv = np.sort(np.concatenate([m1, m2], 1), 1)
        j1, j2 = allv[:, 0], allv[:, 1]
        e = np.minimum(cold_b, j2 + 1)
        if match_only:
            ok = j2 < INF
            d0 = (j1 + qn).astype(np.float64)
            d1 = (j2 + qn).astype(np.float64)
            fail = ok & (d0 > ratio * ratio * (1 + 1e-5) * d1)
            e = np.where(fail, np.minimum(e, j1), e)
        for ch in range(2):
            bv[:, ch] = np.minimum(e, m2[:, ch] + 1)

    window_end()
    chunks = chunk_hits = group_hits = groups_tested = units = 0
    for n, tile in enumerate(seq):
        for ch in range(2):
            units += nq // 32
            for k in range(2):
                c0 = tile * 128 + ch * 64 + k * 32
                d32 = dots[:, c0:c0 + 32]
                gmax = d32.reshape(nq, 4, 8).max(2)
                n8 = tn[c0:c0 + 32].reshape(4, 8).min(1)
                chunks += nq // 32
                if chunk_first:
                    pc = (n8.min() - 2 * gmax.max(1) < bv[:, ch]).reshape(-1, 32).any(1)
                    chunk_hits += int(pc.sum())
                else:
                    pc = np.ones(nq // 32, bool)
                if not pc.any():
                    continue
                groups_tested += 4 * int(pc.sum())
                for g in range(4):
                    hw = ((n8[g] - 2 * gmax[:, g] < bv[:, ch]).reshape(-1, 32).any(1)) & pc
                    group_hits += int(hw.sum())
                    if not hw.any():
                        continue
                    r = np.repeat(hw, 32)
                    v = val[r, c0 + 8 * g:c0 + 8 * g + 8]
                    allv = np.sort(np.concatenate([v, m1[r, ch, None], m2[r, ch, None]], 1), 1)
                    m1[r, ch], m2[r, ch] = allv[:, 0], allv[:, 1]
                    bv[r, ch] = np.minimum(bv[r, ch], allv[:, 1] + 1)
        if (n + 1) % 8 == 0:
            window_end()
    # exactness of the replayed schedule itself (in match-only mode only for rows that pass the ratio test)
    true = np.sort(val, 1)[:, :2]
    got = np.sort(np.concatenate([m1, m2], 1), 1)[:, :2]
    passes = np.sqrt((true[:, 0] + qn).astype(np.float64)) <= ratio * np.sqrt((true[:, 1] + qn).astype(np.float64))
    chk = passes if match_only else np.ones(nq, bool)
    assert np.array_equal(true[chk], got[chk]), "replayed schedule is not exact"
    return dict(chunk_hit_frac=chunk_hits / max(1, chunks), group_hits_per_unit=group_hits / units,
                group_tests_per_unit=groups_tested / units, hot_tiles=len(seq))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    bank = image_bank(3, n)
    for qa, ta in ((0, 2), (0, 1)):
        q, t = bank[qa], bank[ta]
        for kw in (dict(resweep=False, chunk_first=False), dict(resweep=False), dict(cold_tiles=8), dict(cold_tiles=4),
                   dict(cold_tiles=16), dict(cold_tiles=8, match_only=False)):
            r = replay(q, t, **kw)
            print((qa, ta), kw, {k: round(float(x), 3) for k, x in r.items()}, flush=True)
