"""SharedStorageOffloadingSpec — mirror of kv_connectors/llmd_fs_backend/llmd_fs_backend/spec.py:38-157, the object
vLLM's OffloadingConnector instantiates from ``kv_connector_extra_config``.

vLLM is not imported here: the reference subclasses ``vllm.v1.kv_offload.spec.OffloadingSpec`` (0.19) /
``vllm.v1.kv_offload.base.OffloadingSpec`` (0.22) only to receive ``vllm_config`` and ``kv_cache_config``; this class
reads the same attributes by duck typing, so a two-line subclass in the deployment binds it to whichever vLLM is
installed."""
from __future__ import annotations

from typing import Iterator

from .file_mapper import FileMapper
from .manager import SharedStorageOffloadingManager
from .mediums import GPULoadStoreSpec, SharedStorageLoadStoreSpec
from .worker import (DEFAULT_MAX_STAGING_MEMORY_GB, DEFAULT_MAX_WRITE_QUEUED_SECONDS,
                     DEFAULT_READ_PREFERRING_WORKERS_RATIO, DEFAULT_THREADS_PER_GPU, StorageOffloadingHandlers)

DEFAULT_STORAGE_BLOCK_SIZE = 256  # spec.py:35


class SharedStorageOffloadingSpec:
    def __init__(self, vllm_config, kv_cache_config=None, *, extra_config=None, gpu_block_size=None):
        self.vllm_config = vllm_config
        self.kv_cache_config = kv_cache_config
        if extra_config is None:
            kt = getattr(vllm_config, "kv_transfer_config", None)
            extra_config = dict(getattr(kt, "kv_connector_extra_config", None) or {})
        self.extra_config = extra_config
        if gpu_block_size is None:
            gpu_block_size = [vllm_config.cache_config.block_size]
        self.gpu_block_size = list(gpu_block_size) if isinstance(gpu_block_size, (list, tuple)) else [gpu_block_size]
        self._manager = None
        self._handlers = None
        # kv_connector_extra_config keys and defaults (reference README.md:107-118, spec.py:50-85)
        knobs = (("threads_per_gpu", int, DEFAULT_THREADS_PER_GPU),
                 ("max_staging_memory_gb", int, DEFAULT_MAX_STAGING_MEMORY_GB),
                 ("read_preferring_ratio", float, DEFAULT_READ_PREFERRING_WORKERS_RATIO),
                 ("max_write_queued_seconds", float, DEFAULT_MAX_WRITE_QUEUED_SECONDS))
        for key, cast, default in knobs:
            setattr(self, key, cast(self.extra_config.get(key, default)))
        self.offloaded_block_size = int(self.extra_config.get("block_size", DEFAULT_STORAGE_BLOCK_SIZE))
        if len(self.gpu_block_size) != 1:
            raise AssertionError(f"Expected exactly one KV cache group, got {len(self.gpu_block_size)}")
        gbs = self.gpu_block_size[0]
        if self.offloaded_block_size % gbs:
            raise AssertionError("offloaded_block_size must be a multiple of gpu_block_size")
        self.gpu_blocks_per_file = self.offloaded_block_size // gbs

        par = vllm_config.parallel_config
        sizes = {"tp_size": par.tensor_parallel_size, "pp_size": par.pipeline_parallel_size,
                 "pcp_size": getattr(par, "prefill_context_parallel_size", 1)}
        if par.world_size != sizes["tp_size"] * sizes["pp_size"] * sizes["pcp_size"]:
            raise AssertionError("world_size != tp * pp * pcp")
        self.file_mapper = FileMapper(
            root_dir=self.extra_config.get("shared_storage_path", "/tmp/shared-kv"),
            model_name=vllm_config.model_config.model, gpu_block_size=gbs,
            gpu_blocks_per_file=self.gpu_blocks_per_file, rank=par.rank,
            dtype=str(vllm_config.cache_config.cache_dtype).replace("torch.", ""), **sizes)

    def get_manager(self) -> SharedStorageOffloadingManager:
        assert self.vllm_config.parallel_config.rank == 0, "Scheduler rank should be 0"
        if self._manager is None:
            self._manager = SharedStorageOffloadingManager(self.file_mapper)
        return self._manager

    def get_handlers(self, kv_caches) -> Iterator[tuple]:
        if self._handlers is None:
            self._handlers = StorageOffloadingHandlers(
                file_mapper=self.file_mapper, gpu_blocks_per_file=self.gpu_blocks_per_file,
                gpu_block_size=self.gpu_block_size[0], kv_caches=kv_caches, threads_per_gpu=self.threads_per_gpu,
                max_staging_memory_gb=self.max_staging_memory_gb, read_preferring_ratio=self.read_preferring_ratio,
                max_write_queued_seconds=self.max_write_queued_seconds, extra_config=self.extra_config)
        yield GPULoadStoreSpec, SharedStorageLoadStoreSpec, self._handlers.gpu_to_storage_handler
        yield SharedStorageLoadStoreSpec, GPULoadStoreSpec, self._handlers.storage_to_gpu_handler
