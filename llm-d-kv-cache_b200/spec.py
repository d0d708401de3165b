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
        cfg = self.extra_config
        self.threads_per_gpu = int(cfg.get("threads_per_gpu", DEFAULT_THREADS_PER_GPU))
        shared_storage_path = cfg.get("shared_storage_path", "/tmp/shared-kv")
        self.max_staging_memory_gb = int(cfg.get("max_staging_memory_gb", DEFAULT_MAX_STAGING_MEMORY_GB))
        self.offloaded_block_size = int(cfg.get("block_size", DEFAULT_STORAGE_BLOCK_SIZE))
        assert len(self.gpu_block_size) == 1, f"Expected exactly one KV cache group, got {len(self.gpu_block_size)}"
        assert self.offloaded_block_size % self.gpu_block_size[0] == 0, \
            "offloaded_block_size must be a multiple of gpu_block_size"
        self.gpu_blocks_per_file = self.offloaded_block_size // self.gpu_block_size[0]
        self.read_preferring_ratio = float(cfg.get("read_preferring_ratio", DEFAULT_READ_PREFERRING_WORKERS_RATIO))
        self.max_write_queued_seconds = float(cfg.get("max_write_queued_seconds", DEFAULT_MAX_WRITE_QUEUED_SECONDS))
        pc = vllm_config.parallel_config
        tp, pp = pc.tensor_parallel_size, pc.pipeline_parallel_size
        pcp = getattr(pc, "prefill_context_parallel_size", 1)
        assert pc.world_size == tp * pp * pcp
        dtype = str(vllm_config.cache_config.cache_dtype).replace("torch.", "")
        self.file_mapper = FileMapper(root_dir=shared_storage_path, model_name=vllm_config.model_config.model,
                                      gpu_block_size=self.gpu_block_size[0], gpu_blocks_per_file=self.gpu_blocks_per_file,
                                      tp_size=tp, pp_size=pp, pcp_size=pcp, rank=pc.rank, dtype=dtype)

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
