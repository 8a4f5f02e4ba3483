"""Image-by-image sharding of a batch across the GPUs of one host.

Every image is one independent coder stream, so ranks share nothing: no collective touches the data
path.  `plan_shards` is deterministic and identical on every rank (longest-processing-time greedy on
pixel counts), so ranks agree on the partition without communicating.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence


def plan_shards(pixel_counts: Sequence[int], world: int) -> List[List[int]]:
    """Partition image indices over `world` ranks, balancing total pixels (LPT greedy).
    Returns one ascending index list per rank."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(pixel_counts)), key=lambda i: (-int(pixel_counts[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + int(pixel_counts[i]), r))
    return [sorted(s) for s in shards]


def gather_streams(local_streams: Sequence[bytes], local_indices: Sequence[int], n_total: int, dist=None) -> List[bytes]:
    """Host-side gather of the per-image byte streams of all ranks into batch order.  `dist` is an
    initialised torch.distributed module (any backend) or None for a single rank."""
    out: List[bytes] = [b""] * n_total
    if dist is None or dist.get_world_size() == 1:
        for i, s in zip(local_indices, local_streams):
            out[i] = s
        return out
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (list(local_indices), list(local_streams)))
    for idxs, streams in parts:
        for i, s in zip(idxs, streams):
            out[i] = s
    return out
