"""Scheduler-side manager — mirror of kv_connectors/llmd_fs_backend/llmd_fs_backend/manager.py:31-102.
Stateless: a block is "offloaded" iff its file (or host-arena entry) exists; stores are always accepted and never
evict (the tier's own policy does that: the PVC evictor for files, LRU for the arena)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Iterable, Optional

from .file_mapper import FileMapper, hashes_low64
from .mediums import SharedStorageLoadStoreSpec


@dataclass
class PrepareStoreOutput:
    """vLLM ``PrepareStoreOutput`` (block_hashes_to_store, store_spec, block_hashes_evicted)."""

    block_hashes_to_store: list
    store_spec: SharedStorageLoadStoreSpec
    block_hashes_evicted: list = field(default_factory=list)


class SharedStorageOffloadingManager:
    def __init__(self, file_mapper: FileMapper, exists: Optional[Callable[[str], bool]] = None, engine=None):
        """``engine``: a StorageOffloadEngine — ``lookup`` then asks the library ONCE per call
        (``kvb_engine_lookup_prefix``: arena hash-map probes under one lock, or one ``statx`` per file in C) instead of
        one existence probe per block from Python.  Without an engine the reference's loop runs with ``exists``
        (default os.path.exists, manager.py:49-53)."""
        self.file_mapper = file_mapper
        self._engine = engine
        self._exists = exists or os.path.exists

    def lookup(self, block_hashes: Iterable) -> int:
        """How many consecutive blocks from the start are already offloaded (manager.py:43-53)."""
        if self._engine is not None:  # ONE library call; the file names are built in C from 8 bytes per block
            return self._engine.lookup_prefix_hashes(self.file_mapper.base_path, hashes_low64(block_hashes))
        hits = 0
        for h in block_hashes:
            if not self._exists(self.file_mapper.get_file_name(h)):
                break
            hits += 1
        return hits

    def prepare_load(self, block_hashes: Iterable) -> SharedStorageLoadStoreSpec:
        return SharedStorageLoadStoreSpec(block_hashes)  # stateless (manager.py:58-62)

    def touch(self, block_hashes: Iterable) -> None:
        pass  # atime is bumped by the engine when a store finds the file (manager.py:64-70)

    def complete_load(self, block_hashes: Iterable) -> None:
        pass

    def prepare_store(self, block_hashes: Iterable) -> PrepareStoreOutput:
        to_store = list(block_hashes)  # always accepted, nothing evicted (manager.py:79-97)
        return PrepareStoreOutput(to_store, SharedStorageLoadStoreSpec(to_store), [])

    def complete_store(self, block_hashes: Iterable, success: bool = True) -> None:
        pass
