"""Worker of test_peer_push_exchange_two_processes: rank `rank` of a 2-rank push exchange, one process
per rank as in bench.py --gpus N, here both on cuda:0 (CUDA IPC works between processes that share a
device).  Handles travel through multiprocessing queues.  Every rank uploads its parts from pinned
host memory, pushes them into the other bank with the copy engine and raises the mailbox flags;
nothing synchronises with the host before match_pairs.  Three steps with growing tags; match lists,
min_dist and kNN rows equal the plain upload bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def match_bytes(c, pairs):
    m, md, knn = c.match_pairs(pairs, want_knn=True)
    return b"".join(x.tobytes() for x in m), md.tobytes(), b"".join(k.tobytes() for k in knn)


def run(rank, inbox, outbox, result):
    import sfm_opencv_b200 as sfm
    from oracle import matching as M
    from oracle import synth
    from sfm_opencv_b200.sharding import image_regions, shard_pairs_staged, staged_image_ranges
    try:
        sizes = [700, 300, 513, 256, 300, 1100, 2, 900, 650, 1500]
        bank = [synth.sift_like(n, 240 + k) for k, n in enumerate(sizes)]
        bank[3][:100] = bank[0][:100]
        bank[8][:200] = bank[1][:200]
        pairs = M.all_pairs(len(bank))
        regions = staged_image_ranges(len(bank), 2, 2)
        shards = shard_pairs_staged(pairs, sizes, 2, image_regions(len(bank), regions))
        mine = [pairs[i] for i in shards[rank]]
        with sfm.Context(0) as whole:
            whole.upload_descriptors(bank)
            want = match_bytes(whole, mine)
        with sfm.Context(0) as c:
            c.bank_layout(sizes)
            host = []
            for row in regions:
                first, count = row[rank]
                part = []
                for i in range(first, first + count):
                    h = c.pinned_empty(bank[i].shape, np.float32, f"img{i}")
                    h[...] = bank[i]
                    part.append(h)
                host.append(part)
            outbox.put(c.peer_export())
            handles = [None, None]
            handles[rank] = c.peer_export()
            handles[1 - rank] = inbox.get(timeout=60)
            c.peer_connect(rank, handles)
            for tag in (1, 2, 3):
                c.bank_layout(sizes, overlap=True)
                c.bank_ready(tag)
                for k, row in enumerate(regions):
                    first, count = row[rank]
                    c.bank_upload_range(first, host[k], overlap=True)
                    c.bank_push_range(first, count, k, tag)
                    c.bank_pull_commit(1 - rank, row[1 - rank][0], row[1 - rank][1], k, tag)
                assert match_bytes(c, mine) == want, f"rank {rank} step {tag}: results differ"
            c.peer_disconnect()
            outbox.put("done")                    # nobody frees a bank the other still maps
            assert inbox.get(timeout=60) == "done"
        result.put((rank, "ok"))
    except BaseException as e:                    # noqa: BLE001 -- reported to the parent
        result.put((rank, repr(e)))
        raise


def run_two_devices(result):
    """One process, one host thread, contexts on cuda:0 and cuda:1 (the reference is a single
    process): peer memory is a plain device pointer with peer access enabled; every enqueue is
    non-blocking (pinned sources), so one thread can drive both ranks of the protocol."""
    import sfm_opencv_b200 as sfm
    from oracle import matching as M
    from oracle import synth
    from sfm_opencv_b200.sharding import image_regions, shard_pairs_staged, staged_image_ranges
    try:
        try:
            c1 = sfm.Context(1)
        except sfm.SfmError:
            result.put("skip")
            return
        sizes = [700, 300, 513, 256, 300, 1100, 2, 900, 650, 1500]
        bank = [synth.sift_like(n, 340 + k) for k, n in enumerate(sizes)]
        bank[3][:100] = bank[0][:100]
        pairs = M.all_pairs(len(bank))
        regions = staged_image_ranges(len(bank), 2, 2)
        shards = shard_pairs_staged(pairs, sizes, 2, image_regions(len(bank), regions))
        mine = [[pairs[i] for i in shards[r]] for r in range(2)]
        with sfm.Context(0) as c0, c1:
            ranks = (c0, c1)
            c0.upload_descriptors(bank)
            want = [match_bytes(c0, mine[r]) for r in range(2)]
            host = []
            for r, c in enumerate(ranks):
                c.bank_layout(sizes)
                parts = []
                for row in regions:
                    first, count = row[r]
                    part = []
                    for i in range(first, first + count):
                        h = c.pinned_empty(bank[i].shape, np.float32, f"img{i}")
                        h[...] = bank[i]
                        part.append(h)
                    parts.append(part)
                host.append(parts)
            handles = [c.peer_export() for c in ranks]
            for r, c in enumerate(ranks):
                c.peer_connect(r, handles)
            for tag in (1, 2):
                for r, c in enumerate(ranks):
                    c.bank_layout(sizes, overlap=True)
                    c.bank_ready(tag)
                    for k, row in enumerate(regions):
                        first, count = row[r]
                        c.bank_upload_range(first, host[r][k], overlap=True)
                        c.bank_push_range(first, count, k, tag)
                        c.bank_pull_commit(1 - r, row[1 - r][0], row[1 - r][1], k, tag)
                for r, c in enumerate(ranks):
                    assert match_bytes(c, mine[r]) == want[r], f"rank {r} step {tag}: results differ"
            for c in ranks:
                c.peer_disconnect()
        result.put("ok")
    except BaseException as e:                    # noqa: BLE001 -- reported to the parent
        result.put(repr(e))
        raise
