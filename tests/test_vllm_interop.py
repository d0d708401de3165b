"""The save_blocks / load_blocks handlers driven by the INSTALLED vLLM's own OffloadingWorker and spec classes
(vllm.v1.kv_offload.base / .worker.worker in vLLM 0.22; the reference pins 0.19 where they lived in .mediums / .abstract):
registration by medium, routing of (src, dst) specs, the file / block-id mapping of worker.py:158-193 and the shape of the
results vLLM reads back.  CPU only: the engine is a recording stub with the StorageEngine Protocol (worker.py:36-52)."""
import importlib

import numpy as np
import pytest

from oracle import offload_oracle as oo

vllm_base = pytest.importorskip("vllm.v1.kv_offload.base")
vllm_worker = pytest.importorskip("vllm.v1.kv_offload.worker.worker")


class RecordingEngine:
    """StorageEngine Protocol: submit-only store / load, (job_id, ok) draining, wait, shutdown."""

    def __init__(self):
        self.calls, self.done = [], []

    def async_store_gpu_blocks(self, job_id, files, ids):
        self.calls.append(("store", job_id, list(files), [list(map(int, g)) for g in ids]))
        self.done.append((job_id, True))
        return True

    def async_load_gpu_blocks(self, job_id, files, ids):
        self.calls.append(("load", job_id, list(files), [list(map(int, g)) for g in ids]))
        self.done.append((job_id, job_id != 13))
        return True

    def get_finished(self):
        out, self.done = self.done, []
        return out

    def wait_job(self, job_id):
        self.calls.append(("wait", job_id))

    def shutdown(self):
        pass


def test_handlers_under_the_installed_vllm_worker():
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    bpf, block_bytes = 4, 1 << 20
    fm = kvb.file_mapper.FileMapper("/tmp/kvb-vllm", "m", 16, bpf, 1, 1, 1, 0, "torch.float16")
    eng = RecordingEngine()
    store = kvb.worker.GPUToStorageHandler(bpf, fm, eng, ("GPU", "SHARED_STORAGE"), block_bytes)
    load = kvb.worker.StorageToGPUHandler(bpf, fm, eng, ("SHARED_STORAGE", "GPU"), block_bytes)
    load._pending_jobs = store._pending_jobs           # as StorageOffloadingHandlers wires them
    GPU, SHARED = vllm_base.GPULoadStoreSpec, kvb.mediums.SharedStorageLoadStoreSpec
    assert issubclass(GPU, vllm_base.LoadStoreSpec) and GPU.medium() == "GPU" and SHARED.medium() == "SHARED_STORAGE"
    w = vllm_worker.OffloadingWorker()
    w.register_handler(GPU, SHARED, store)             # vLLM's own registry, keyed by (src.medium(), dst.medium())
    w.register_handler(SHARED, GPU, load)
    # 10 GPU blocks -> 3 offloaded blocks of 4: the FIRST file takes the remainder (worker.py:174-191)
    block_ids = [17, 3, 9, 4, 28, 1, 0, 11, 30, 2]
    hashes = [bytes([i]) * 32 for i in (1, 2, 3)]
    gpu_spec = GPU(block_ids, group_sizes=[len(block_ids)], block_indices=[0])      # vLLM 0.22 signature
    assert w.transfer_async(7, (gpu_spec, SHARED(hashes)))
    assert w.transfer_async(13, (SHARED(hashes), gpu_spec))
    want_files, want_groups = oo.build_file_block_mapping([fm.get_file_name(h) for h in hashes], block_ids, bpf)
    assert eng.calls[0] == ("store", 7, want_files, [list(g) for g in want_groups])
    assert eng.calls[1] == ("load", 13, want_files, [list(g) for g in want_groups])
    res = {r.job_id: r for r in w.get_finished()}      # vLLM reads .job_id / .success / .transfer_size / .transfer_type
    assert res[7].success is True and res[13].success is False
    assert res[7].transfer_size == len(block_ids) * block_bytes and res[7].transfer_type == ("GPU", "SHARED_STORAGE")
    assert res[13].transfer_type == ("SHARED_STORAGE", "GPU") and res[7].transfer_time >= 0
    as_vllm = vllm_worker.TransferResult(*res[7])      # field for field the dataclass vLLM declares
    assert (as_vllm.job_id, as_vllm.success, as_vllm.transfer_size) == (7, True, len(block_ids) * block_bytes)
    store.wait({7})
    assert eng.calls[-1] == ("wait", 7)
