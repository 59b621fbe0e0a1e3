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
