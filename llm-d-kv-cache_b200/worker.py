"""Worker-side offloading handlers — mirror of kv_connectors/llmd_fs_backend/llmd_fs_backend/worker.py
(GPUToStorageHandler / StorageToGPUHandler / StorageOffloadingHandlers), i.e. the save_blocks /
load_blocks surface vLLM's OffloadingConnector drives.  The engine underneath is libkvb.so."""
from __future__ import annotations

import math
import os
import time
from typing import Sequence

from .engine import StorageOffloadEngine
from .file_mapper import FileMapper
from .mediums import TransferResult

DEFAULT_MAX_STAGING_MEMORY_GB = 150  # worker.py:58-61
DEFAULT_THREADS_PER_GPU = 64
DEFAULT_READ_PREFERRING_WORKERS_RATIO = 0.75
DEFAULT_MAX_WRITE_QUEUED_SECONDS = 10.0
# This engine packs a chunk of whole files in HBM before the host leg (the reference has no HBM staging at all), and vLLM
# has already sized its KV cache from gpu_memory_utilization when the handlers are built: keep the workers' packed
# chunks inside a fixed HBM budget instead of io_threads x 64 MiB.
DEFAULT_MAX_HBM_STAGING_MB = 1024


def plan_worker_memory(per_block_bytes: int, gpu_blocks_per_file: int, threads_per_gpu: int,
                       max_staging_memory_gb: float, extra_config: dict | None = None):
    """Staging budget clamp (worker.py:303-319) for THIS engine.  What one of its workers allocates is a chunk of whole
    files (>= one file) of HBM plus, in the file tier, the same amount of pinned host memory — not the reference's one
    file-sized host buffer and no HBM at all — so the clamp is on the chunk and covers both memories: the host side by
    ``max_staging_memory_gb`` (the reference's knob), the HBM side by ``max_hbm_staging_mb`` (default 1 GiB).
    Returns (worker threads, chunk_bytes)."""
    extra_config = extra_config or {}
    file_bytes = per_block_bytes * gpu_blocks_per_file
    chunk_bytes = max(int(extra_config.get("chunk_bytes", 0)) or file_bytes, file_bytes)
    chunk_mb = math.ceil(chunk_bytes / (1 << 20))
    budget_mb = min(max_staging_memory_gb * 1024, int(extra_config.get("max_hbm_staging_mb", DEFAULT_MAX_HBM_STAGING_MB)))
    if chunk_mb * threads_per_gpu > budget_mb:
        threads_per_gpu = max(1, min(threads_per_gpu, int(budget_mb / chunk_mb)))
    return threads_per_gpu, chunk_bytes


class BaseStorageOffloadingHandler:
    """Common bookkeeping of both directions (worker.py:64-193)."""

    def __init__(self, gpu_blocks_per_file: int, file_mapper: FileMapper, engine, transfer_type, per_block_bytes: int):
        self.file_mapper = file_mapper
        self.gpu_blocks_per_file = gpu_blocks_per_file
        self.engine = engine
        self.transfer_type = transfer_type
        self.per_block_bytes = per_block_bytes
        # job_id -> (submit_time, bytes, transfer_type); shared by the two handlers.  The type travels with the job: both
        # handlers drain the SAME engine, so whichever vLLM polls first reports the other's jobs too — the reference
        # labels those with the polling handler's type (worker.py:107-122), a metrics quirk not reproduced here
        self._pending_jobs: dict = {}

    def _record_job(self, job_id: int, num_blocks: int) -> None:
        self._pending_jobs[job_id] = (time.monotonic(), num_blocks * self.per_block_bytes, self.transfer_type)

    def get_finished(self) -> list:
        now = time.monotonic()
        results = []
        for job_id, success in self.engine.get_finished():
            info = self._pending_jobs.pop(job_id, None)
            if info is None:  # unknown job: still reported, without metrics (worker.py:139-145)
                results.append(TransferResult(job_id=job_id, success=success))
                continue
            t_submit, size, ttype = info
            results.append(TransferResult(job_id=job_id, success=success, transfer_size=size,
                                          transfer_time=now - t_submit, transfer_type=ttype))
        return results

    def wait(self, job_ids) -> None:
        for job_id in job_ids:
            self.engine.wait_job(job_id)

    def _build_file_block_mapping(self, block_hashes: Sequence, block_ids: Sequence[int]):
        """One file per offloaded-block hash; the FIRST file takes the remainder
        ``len(ids) % bpf or bpf``, all later files ``bpf`` ids (worker.py:158-193)."""
        bpf = self.gpu_blocks_per_file
        total = len(block_ids)
        take = total % bpf or bpf
        files, groups, pos = [], [], 0
        for h in block_hashes:
            files.append(self.file_mapper.get_file_name(h))
            groups.append(block_ids[pos:min(pos + take, total)])
            pos += take
            take = bpf
        return files, groups

    def _transfer(self, submit, job_id: int, hashes, block_ids) -> bool:
        files, groups = self._build_file_block_mapping(hashes, block_ids)
        ok = submit(job_id, files, groups)
        if ok:
            self._record_job(job_id, sum(len(g) for g in groups))
        return ok


class GPUToStorageHandler(BaseStorageOffloadingHandler):
    """save_blocks: GPU -> storage (worker.py:196-227)."""

    def transfer_async(self, job_id: int, spec) -> bool:
        src_spec, dst_spec = spec
        return self._transfer(self.engine.async_store_gpu_blocks, job_id, dst_spec.block_hashes, src_spec.block_ids)


class StorageToGPUHandler(BaseStorageOffloadingHandler):
    """load_blocks: storage -> GPU (worker.py:230-261)."""

    def transfer_async(self, job_id: int, spec) -> bool:
        src_spec, dst_spec = spec
        return self._transfer(self.engine.async_load_gpu_blocks, job_id, src_spec.block_hashes, dst_spec.block_ids)


class StorageOffloadingHandlers:
    """Builds the engine and the two handlers (worker.py:264-407).

    ``kv_caches`` is either vLLM's CanonicalKVCaches (``.tensors[i].tensor``) or a plain list of
    (num_blocks, page_bytes) CUDA tensors."""

    def __init__(self, kv_caches, file_mapper: FileMapper, gpu_block_size: int, gpu_blocks_per_file: int,
                 threads_per_gpu: int, max_staging_memory_gb: int = DEFAULT_MAX_STAGING_MEMORY_GB,
                 read_preferring_ratio: float = DEFAULT_READ_PREFERRING_WORKERS_RATIO,
                 max_write_queued_seconds: float = DEFAULT_MAX_WRITE_QUEUED_SECONDS, extra_config: dict | None = None):
        extra_config = extra_config or {}
        threads_per_gpu = min(threads_per_gpu, int(os.cpu_count() or 1))  # worker.py:280
        raw = getattr(kv_caches, "tensors", kv_caches)
        tensors = [getattr(t, "tensor", t) for t in raw]
        assert tensors
        per_block_bytes = sum(t.stride(0) * t.element_size() for t in tensors)  # worker.py:336
        threads_per_gpu, chunk_bytes = plan_worker_memory(per_block_bytes, gpu_blocks_per_file, threads_per_gpu,
                                                          max_staging_memory_gb, extra_config)
        extra_config = dict(extra_config, chunk_bytes=chunk_bytes)
        read_preferring_workers = max(1, int(threads_per_gpu * read_preferring_ratio))  # worker.py:322
        self.engine = self._create_engine(
            io_threads=threads_per_gpu, gpu_blocks_per_file=gpu_blocks_per_file, tensors=tensors,
            read_preferring_workers=read_preferring_workers, max_write_queued_seconds=max_write_queued_seconds,
            extra_config=extra_config, gds_mode=extra_config.get("gds_mode", "disabled"))
        pending: dict = {}
        self.gpu_to_storage_handler = GPUToStorageHandler(
            gpu_blocks_per_file, file_mapper, self.engine, ("GPU", "SHARED_STORAGE"), per_block_bytes)
        self.storage_to_gpu_handler = StorageToGPUHandler(
            gpu_blocks_per_file, file_mapper, self.engine, ("SHARED_STORAGE", "GPU"), per_block_bytes)
        self.gpu_to_storage_handler._pending_jobs = pending
        self.storage_to_gpu_handler._pending_jobs = pending

    def _create_engine(self, io_threads, gpu_blocks_per_file, tensors, read_preferring_workers,
                       max_write_queued_seconds, extra_config, gds_mode):
        return StorageOffloadEngine(
            io_threads, gpu_blocks_per_file, tensors, read_preferring_workers, gds_mode, max_write_queued_seconds,
            tier=extra_config.get("tier", "file"), host_arena_bytes=int(extra_config.get("host_arena_bytes", 0)),
            chunk_bytes=int(extra_config.get("chunk_bytes", 0)))
