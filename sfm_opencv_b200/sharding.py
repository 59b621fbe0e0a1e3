"""Pair / point / observation sharding over the GPUs of one box.

The path has no cross-pair, cross-point or cross-observation state
(OpenCV_SFM/NViewReconstuct.cpp:857-870, :1151-1156, :1187-1211), so every rank works on its
own shard with NO data-path collective; the only exchange is the final gather of results,
which keeps the reference's order (pair order, ascending queryIdx inside a pair).
"""
from __future__ import annotations

import numpy as np


def pair_cost(n_desc, pairs) -> np.ndarray:
    """Work of each pair in multiply-accumulates / 128: Nq * Nt."""
    n = np.asarray(n_desc, np.int64)
    p = np.asarray(pairs, np.int64).reshape(-1, 2)
    return n[p[:, 0]] * n[p[:, 1]]


def shard_pairs(pairs, n_desc, world_size: int):
    """Splits the pair list into `world_size` CONTIGUOUS blocks of near-equal cost.

    Contiguous blocks keep the reference's pair order inside a rank (pairs sharing a query
    image stay together, which keeps that image's tiles warm in L2) and make the gather a
    plain concatenation.  Returns a list of (start, stop) index ranges, one per rank; ranges
    may be empty when there are fewer pairs than ranks.
    """
    p = np.asarray(pairs, np.int64).reshape(-1, 2)
    n_pairs = p.shape[0]
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if n_pairs == 0:
        return [(0, 0)] * world_size
    cost = pair_cost(n_desc, p).astype(np.float64)
    cost = np.maximum(cost, 1.0)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world_size):
        target = cum[-1] * r / world_size
        # first index whose cumulative cost reaches the target, never moving backwards
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the closer of k-1 / k
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n_pairs)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n_pairs))
    bounds.append(n_pairs)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_range(n: int, world_size: int):
    """Contiguous, near-equal ranges of points / observations, one per rank."""
    base, rem = divmod(int(n), world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def shard_query_rows(n_query: int, world_size: int, block: int = 256):
    """One huge pair (BASELINE config 4): query rows are independent, so each rank matches a
    contiguous range of query rows against the whole (replicated) train set -- no collective, the
    k-NN rows / match lists concatenate in rank order.  Ranges are multiples of the kernel's
    256-row query block (except the last) so that no rank pads a block another rank also holds."""
    blocks = (int(n_query) + block - 1) // block
    out = []
    for s, e in shard_range(blocks, world_size):
        out.append((min(s * block, n_query), min(e * block, n_query)))
    return out


def _split_weighted(n: int, weights):
    """n items in len(weights) consecutive parts proportional to the weights (largest remainders)."""
    w = np.asarray(weights, np.float64)
    if len(w) == 0 or (w <= 0).any():
        raise ValueError("weights must be positive")
    ideal = n * w / w.sum()
    parts = np.floor(ideal).astype(np.int64)
    for k in np.argsort(-(ideal - parts), kind="stable")[: n - int(parts.sum())]:
        parts[k] += 1
    return [int(x) for x in parts]


def staged_image_ranges(n_img: int, world_size: int, n_stages: int, weights=None):
    """Bank order for the staged (overlapped) multi-GPU upload: the image list is cut into
    `n_stages` REGIONS; region k holds one contiguous part of every rank, rank-major, so that one
    in-place all-gather moves a whole region.  Rank r uploads regions[k][r] = (first_img, count)
    for every k -- still 1/N of the images per rank, each image crosses PCIe once.  With n_img a
    multiple of world_size the parts of a region are equal (what ncclAllGather needs).
    weights (one per stage, default equal): relative sizes of the regions.  A small first region
    shortens the time before matching can start; matching its pairs then has to cover the arrival
    of the next one (the work of the first k regions grows with the square of their share)."""
    if world_size <= 0 or n_stages <= 0:
        raise ValueError("world_size and n_stages must be positive")
    if weights is None:
        weights = [1.0] * n_stages
    if len(weights) != n_stages:
        raise ValueError("one weight per stage")
    per_rank = [e - s for s, e in shard_range(n_img, world_size)]
    parts = [_split_weighted(c, weights) for c in per_rank]                    # [rank][stage]
    regions, first = [], 0
    for k in range(n_stages):
        row = []
        for r in range(world_size):
            row.append((first, parts[r][k]))
            first += parts[r][k]
        regions.append(row)
    return regions


def image_regions(n_img: int, regions) -> np.ndarray:
    """Region index of every image of a staged_image_ranges layout."""
    out = np.zeros(n_img, np.int32)
    for k, row in enumerate(regions):
        for first, count in row:
            out[first:first + count] = k
    return out


def shard_pairs_staged(pairs, n_desc, world_size: int, img_region):
    """Pair shards for the staged upload: a pair can be matched once the later of its two images'
    regions has arrived (its stage); the pairs of every stage are split into `world_size`
    contiguous cost-balanced blocks, and rank r takes block r of every stage, earliest stage first.
    Every rank therefore has work as soon as region 0 is there.  Returns one int64 index array
    (into `pairs`) per rank; the arrays partition the pair list."""
    p = np.asarray(pairs, np.int64).reshape(-1, 2)
    reg = np.asarray(img_region, np.int64)
    stage = np.maximum(reg[p[:, 0]], reg[p[:, 1]]) if len(p) else np.zeros(0, np.int64)
    out = [[] for _ in range(world_size)]
    for s in np.unique(stage):
        idx = np.nonzero(stage == s)[0]
        for r, (lo, hi) in enumerate(shard_pairs(p[idx], n_desc, world_size)):
            out[r].append(idx[lo:hi])
    return [np.concatenate(o) if o else np.zeros(0, np.int64) for o in out]


def gather_match_lists_indexed(local_matches, pair_index, n_pairs: int, group=None):
    """gather_match_lists for shards that are index lists (shard_pairs_staged) instead of ranges:
    local_matches[k] belongs to pair pair_index[k]; the result is in pair order."""
    import torch.distributed as dist
    idx = [int(i) for i in pair_index]
    payload = (idx, [np.asarray(m) for m in local_matches])
    if not dist.is_available() or not dist.is_initialized():
        gathered = [payload]
    else:
        gathered = [None] * dist.get_world_size(group)
        dist.all_gather_object(gathered, payload, group=group)
    out = [None] * n_pairs
    for ids, ms in gathered:
        assert len(ids) == len(ms)
        for i, m in zip(ids, ms):
            assert out[i] is None, "pair matched by two ranks"
            out[i] = m
    assert all(m is not None for m in out), "pair shards do not cover the pair list"
    return out


def gather_match_lists(local_matches, start: int, stop: int, n_pairs: int, group=None):
    """Final host gather (the only exchange on the path): every rank contributes the match
    arrays of its pair range; rank order == pair order, so the result is the reference's
    `matches_for_all`.  Uses torch.distributed.all_gather_object on the given group
    (gloo or nccl process groups both work: the payload is host data)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        assert start == 0 and stop == n_pairs
        return list(local_matches)
    world = dist.get_world_size(group)
    payload = (int(start), int(stop), [np.asarray(m) for m in local_matches])
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    out = [None] * n_pairs
    for s, e, ms in gathered:
        assert len(ms) == e - s
        for k, m in enumerate(ms):
            out[s + k] = m
    assert all(m is not None for m in out), "pair ranges do not cover the pair list"
    return out
