"""CPU: pair / point sharding over ranks incl. a world_size-2 gloo run of
the final host gather.  The per-rank 'matcher' here is the CPU oracle -- this tests the host
logic around the C ABI (shard, gather, order), not the kernels."""
import os
import socket

import numpy as np
import pytest

from sfm_opencv_b200 import sharding as S


def test_shard_pairs_covers_in_order():
    n_desc = [8192] * 20
    pairs = [(i, j) for i in range(20) for j in range(i + 1, 20)]
    for world in (1, 2, 3, 4, 8):
        rngs = S.shard_pairs(pairs, n_desc, world)
        assert len(rngs) == world and rngs[0][0] == 0 and rngs[-1][1] == len(pairs)
        for a, b in zip(rngs, rngs[1:]):
            assert a[1] == b[0]
        sizes = [e - s for s, e in rngs]
        assert max(sizes) - min(sizes) <= 1            # equal-cost pairs -> equal counts


def test_shard_pairs_balances_uneven_cost():
    n_desc = [16717, 16440, 7823, 1244, 838]            # desktop SIFT sizes
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    cost = S.pair_cost(n_desc, pairs)
    rngs = S.shard_pairs(pairs, n_desc, 2)
    c = [cost[s:e].sum() for s, e in rngs]
    assert sum(c) == cost.sum()
    # the first pair alone is ~45 % of the work: the best contiguous cut is right after it
    assert rngs[0] == (0, 1)
    assert S.shard_pairs([], n_desc, 3) == [(0, 0)] * 3
    assert S.shard_pairs(pairs[:1], n_desc, 4)[-1][1] == 1


def test_shard_query_rows_covers_in_blocks():
    from sfm_opencv_b200.sharding import shard_query_rows
    for n, w in ((65536, 8), (1000, 3), (255, 4), (0, 2), (257, 2)):
        r = shard_query_rows(n, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(s % 256 == 0 for s, _ in r if s < n)


def test_shard_range():
    assert S.shard_range(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert S.shard_range(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import matching as M
    from oracle import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bank = synth.image_bank(4, 200, seed0=90)
    pairs = M.all_pairs(4)
    s, e = S.shard_pairs(pairs, [len(b) for b in bank], world)[rank]
    local = []
    for a, b in pairs[s:e]:
        m, d, _, _, _ = M.match_features(bank[a], bank[b])
        out = np.zeros(len(m), [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                                ("distance", "<f4")])
        out["queryIdx"], out["trainIdx"], out["distance"] = m[:, 0], m[:, 1], d
        local.append(out)
    full = S.gather_match_lists(local, s, e, len(pairs))
    if rank == 0:
        q.put([(x["queryIdx"].tolist(), x["trainIdx"].tolist()) for x in full])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gather_world2_gloo_equals_single_rank():
    import torch.multiprocessing as mp
    from oracle import matching as M
    from oracle import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bank = synth.image_bank(4, 200, seed0=90)
    pairs = M.all_pairs(4)
    assert len(got) == len(pairs)
    for (a, b), (qi, ti) in zip(pairs, got):
        m, _, _, _, _ = M.match_features(bank[a], bank[b])
        assert qi == m[:, 0].tolist() and ti == m[:, 1].tolist()


def test_enumerate_observations_is_camera_major():
    """Host mirror of the residual-block enumeration (NViewReconstuct.cpp:1187-1211) against the
    oracle's restatement; no GPU needed."""
    import sfm_opencv_b200 as sfm
    from oracle import geometry as G
    rng = np.random.default_rng(3)
    ids = [rng.integers(-1, 20, n) for n in (7, 0, 13, 5)]
    kps = [rng.uniform(0, 1000, (len(i), 2)).astype(np.float32) for i in ids]
    a = sfm.enumerate_observations(ids, kps)
    b = G.enumerate_observations(ids, kps)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (np.diff(a[0]) >= 0).all() and len(a[0]) == sum(int((i >= 0).sum()) for i in ids)


def _rows_worker(rank, world, port, q):
    """One pair sharded by query rows with the CPU oracle as the per-rank
    matcher: pass 1 per shard, MIN over ranks (one float), pass 2 under the reduced min_dist."""
    import torch
    import torch.distributed as dist
    from oracle import matching as M
    from oracle import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    qd, td = synth.sift_like(700, 5), synth.sift_like(900, 6)
    td[:100] = qd[300:400]
    td[500] = qd[3]                                  # distance 0: the global min_dist sits in shard 0
    lo, hi = S.shard_query_rows(len(qd), world)[rank]
    d, idx = M.knn2_int(qd[lo:hi], td)
    _, _, md_local = M.filter_matches(d, idx)
    t = torch.tensor([float(md_local)], dtype=torch.float32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    md = np.float32(t.item())
    # pass 2 under the reduced value (what sfm_match_rows_finish does)
    d0, d1 = d[:, 0], d[:, 1]
    gate = M.GATE_MULT * max(md, M.DIST_FLOOR)
    keep = ~((d0.astype(np.float64) > M.RATIO * d1.astype(np.float64)) | (d0 > gate))
    out = np.zeros(int(keep.sum()), [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    out["queryIdx"] = np.nonzero(keep)[0] + lo
    out["trainIdx"] = idx[keep, 0]
    out["distance"] = d0[keep]
    parts = S.gather_match_lists([out], rank, rank + 1, world)
    if rank == 0:
        q.put((float(md_local), float(md), np.concatenate(parts).tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_pair_world2_gloo_equals_unsharded():
    import torch.multiprocessing as mp
    from oracle import matching as M
    from oracle import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    md_local0, md, got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    qd, td = synth.sift_like(700, 5), synth.sift_like(900, 6)
    td[:100] = qd[300:400]
    td[500] = qd[3]
    m, d0, wmd = M.match_features(qd, td)[:3]
    want = np.zeros(len(m), [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    want["queryIdx"], want["trainIdx"], want["distance"] = m[:, 0], m[:, 1], d0
    assert md == float(wmd) == 0.0 and got == want.tobytes() and len(m) > 50


def test_huber_cost_is_additive_over_observation_ranges():
    """Residual sharding: the cost 0.5 * sum rho(|r|^2) of the whole problem is
    the sum of the costs of contiguous observation ranges (the one scalar the ranks exchange)."""
    from oracle import geometry as G
    from oracle import synth
    sc = synth.scene(3000, 3, seed=4)
    cam, pt = synth.observations_camera_major(3000, 3)
    r = G.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2))
    whole = G.huber_cost(r, 4.0)
    for world in (2, 3, 8):
        parts = [G.huber_cost(r[s:e], 4.0) for s, e in S.shard_range(len(r), world)]
        assert abs(sum(parts) - whole) <= 1e-12 * whole


# ---- staged (overlapped) multi-GPU upload: image regions, per-stage pair shards, indexed gather
@pytest.mark.parametrize("n_img,world,stages", [(200, 8, 3), (200, 2, 3), (10, 4, 3), (7, 2, 1), (5, 8, 2)])
def test_staged_image_ranges_partition_the_images(n_img, world, stages):
    regions = S.staged_image_ranges(n_img, world, stages)
    assert len(regions) == stages and all(len(row) == world for row in regions)
    flat = [rng for row in regions for rng in row]
    nxt = 0
    for first, count in flat:                                  # bank order: region-major, rank-major, contiguous
        assert first == nxt and count >= 0
        nxt += count
    assert nxt == n_img
    per_rank = [sum(regions[k][r][1] for k in range(stages)) for r in range(world)]
    assert sorted(per_rank) == sorted(e - s for s, e in S.shard_range(n_img, world))   # still 1/N per rank
    if n_img % world == 0:
        assert all(len({c for _, c in row}) == 1 for row in regions)                   # equal parts: in-place all-gather
    reg = S.image_regions(n_img, regions)
    assert np.all(np.diff(reg) >= 0) and reg.max() <= stages - 1


def test_shard_pairs_staged_partitions_and_balances_every_stage():
    n_img, world, stages = 40, 4, 3
    sizes = [1000 + 37 * (i % 5) for i in range(n_img)]
    pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
    reg = S.image_regions(n_img, S.staged_image_ranges(n_img, world, stages))
    shards = S.shard_pairs_staged(pairs, sizes, world, reg)
    allidx = np.concatenate(shards)
    assert len(allidx) == len(pairs) and len(np.unique(allidx)) == len(pairs)
    p = np.asarray(pairs)
    stage = np.maximum(reg[p[:, 0]], reg[p[:, 1]])
    cost = S.pair_cost(sizes, pairs)
    for idx in shards:
        assert np.all(np.diff(stage[idx]) >= 0)                # earliest stage first
    for s in range(stages):
        per_rank = [cost[idx[stage[idx] == s]].sum() for idx in shards]
        assert max(per_rank) <= 1.1 * (sum(per_rank) / world) + cost.max()


def _worker_indexed(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_img = 6
    pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
    reg = S.image_regions(n_img, S.staged_image_ranges(n_img, world, 2))
    mine = S.shard_pairs_staged(pairs, [100] * n_img, world, reg)[rank]
    local = [np.full(3, 100 * pairs[i][0] + pairs[i][1], np.int32) for i in mine]    # stands for a match list
    full = S.gather_match_lists_indexed(local, mine, len(pairs))
    if rank == 0:
        q.put([int(x[0]) for x in full])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_indexed_gather_world2_gloo_is_in_pair_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_indexed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [100 * i + j for i in range(6) for j in range(i + 1, 6)]
