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
