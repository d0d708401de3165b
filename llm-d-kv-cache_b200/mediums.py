"""Load/store specs used by the handlers.

The reference subclasses vLLM's ``LoadStoreSpec`` types (llmd_fs_backend/mediums.py; vLLM
``vllm.v1.kv_offload.mediums.GPULoadStoreSpec``).  vLLM moved those classes between releases
(0.19: ``kv_offload.mediums``/``abstract``; 0.22: ``kv_offload.base``), so this package carries
structurally identical stand-ins and accepts vLLM's own objects by duck typing
(``.block_ids`` / ``.block_hashes``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional, Sequence

import numpy as np


class GPULoadStoreSpec:
    """Blocks of the GPU KV cache (vLLM GPULoadStoreSpec: ``block_ids`` as int64 array)."""

    def __init__(self, block_ids: Sequence[int], group_sizes=None, block_indices=None):
        self.block_ids = np.asarray(block_ids, dtype=np.int64)
        self.group_sizes = group_sizes
        self.block_indices = block_indices

    @staticmethod
    def medium() -> str:
        return "GPU"


class SharedStorageLoadStoreSpec:
    """Offloaded blocks identified by their hashes (llmd_fs_backend/mediums.py)."""

    def __init__(self, block_hashes: Sequence):
        self.block_hashes = list(block_hashes)

    @staticmethod
    def medium() -> str:
        return "SHARED_STORAGE"


TransferType = tuple  # (src_medium, dst_medium)


class TransferResult(NamedTuple):
    """vLLM ``TransferResult`` (job_id, success, transfer_size, transfer_time, transfer_type)."""

    job_id: int
    success: bool
    transfer_size: Optional[int] = None
    transfer_time: Optional[float] = None
    transfer_type: Optional[tuple] = None
