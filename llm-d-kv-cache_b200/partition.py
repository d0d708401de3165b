"""Host-side planning for the multi-GPU paths (pure Python, backend-agnostic: tested with gloo on CPU).

The offload path shards by KV partition: one engine per GPU over its own pool, no data-path collective
(the reference runs one engine per worker and separates ranks only by directory, file_mapper.py:60-67).
Migration has one exchange step; torch.distributed carries only its control plane (pool descriptors, block-id
lists, checksums), the bytes move over NVLink by peer stores from the gather kernel."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple:
    """Contiguous [lo, hi) share of n_items for this rank; shares differ by at most one item."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ring_peers(rank: int, world: int) -> tuple:
    """(destination, source) of this rank in the all-pairs ring r -> r+1."""
    return (rank + 1) % world, (rank - 1) % world


def fanout_plan(n_blocks: int, world: int, root: int = 0) -> dict:
    """Prefill -> decode hand-off from `root` to every other GPU: peer p receives its own n_blocks pages
    (different prefixes).  Returns {dst_rank: (src_lo, src_hi)} into the root's list of world-1 prefixes."""
    peers = [r for r in range(world) if r != root]
    return {p: (i * n_blocks, (i + 1) * n_blocks) for i, p in enumerate(peers)}


def exchange_objects(obj, dist=None) -> list:
    """all_gather of one picklable object per rank (pool descriptors, checksums)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def page_checksum_np(tensors: Sequence[np.ndarray], ids) -> int:
    """Order-independent 64-bit checksum of the listed pages (sum of little-endian u64 words mod 2^64)."""
    total = 0
    ids = np.asarray(ids, dtype=np.int64)
    for t in tensors:
        pages = np.ascontiguousarray(t[ids])
        total = (total + int(pages.view(np.uint64).sum(dtype=np.uint64))) & ((1 << 64) - 1)
    return total


def page_checksum_torch(tensors, ids_dev) -> int:
    import torch
    total = 0
    for t in tensors:
        pages = t[ids_dev].contiguous().view(torch.int64)
        total = (total + int(pages.sum().item())) & ((1 << 64) - 1)
    return total
