"""Host-side mirrors (no GPU): file naming, file grouping, extra-key logic and CBOR suffixes vs the oracle;
handler bookkeeping with a fake engine (the reference tests its handlers the same way with mocks)."""
import random

import numpy as np
import pytest

from oracle import kvblock_oracle as ko
from oracle import offload_oracle as oo


def test_file_mapper_matches_oracle(kvb):
    fm = kvb.file_mapper.FileMapper("/mnt/kv", "meta-llama/Llama-3-8B", 16, 16, 2, 1, 1, 3, "torch.float16")
    base = oo.base_path("/mnt/kv", "meta-llama/Llama-3-8B", 16, 16, 2, 1, 1, 3, "torch.float16")
    assert fm.base_path == base
    assert base.endswith("/block_size_16_blocks_per_file_16/tp_2_pp_size_1_pcp_size_1/rank_3/torch.float16")
    rnd = random.Random(3)
    for _ in range(200):
        h = rnd.getrandbits(rnd.choice([8, 64, 65, 256]))
        assert fm.get_file_name(h) == oo.file_name(base, h)
        b = h.to_bytes(32, "big")
        assert fm.get_file_name(b) == oo.file_name(base, b)
    assert fm.get_file_name(0xABCDEF0123456789).endswith("/abc/de/abcdef0123456789.bin")
    assert fm.get_file_name((1 << 70) | 5).endswith("/000/00/0000000000000005.bin")   # only the low 64 bits
    with pytest.raises((TypeError, AssertionError)):
        fm.get_file_name("nope")


class _FakeEngine:
    def __init__(self):
        self.calls, self.finished = [], []

    def async_store_gpu_blocks(self, job_id, files, ids):
        self.calls.append(("store", job_id, list(files), [list(map(int, g)) for g in ids]))
        self.finished.append((job_id, True))
        return True

    def async_load_gpu_blocks(self, job_id, files, ids):
        self.calls.append(("load", job_id, list(files), [list(map(int, g)) for g in ids]))
        self.finished.append((job_id, job_id != 13))
        return True

    def get_finished(self):
        out, self.finished = self.finished, []
        return out

    def wait_job(self, job_id):
        self.calls.append(("wait", job_id))


@pytest.mark.parametrize("bpf", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("n", [1, 3, 8, 16, 17, 33])
def test_file_block_mapping_matches_oracle(kvb, bpf, n):
    W, M = kvb.worker, kvb.mediums
    fm = kvb.file_mapper.FileMapper("/r", "m", 16, bpf, 1, 1, 1, 0, "torch.float16")
    eng = _FakeEngine()
    h = W.GPUToStorageHandler(bpf, fm, eng, ("GPU", "SHARED_STORAGE"), 1000)
    ids = list(range(100, 100 + n))
    hashes = list(range(1, -(-n // bpf) + 1))
    assert h.transfer_async(5, (M.GPULoadStoreSpec(ids), M.SharedStorageLoadStoreSpec(hashes)))
    _, _, files, groups = eng.calls[0]
    oh, og = oo.build_file_block_mapping(hashes, ids, bpf)
    assert groups == og and files == [oo.file_name(fm.base_path, x) for x in oh]
    assert len(groups[0]) == (n % bpf or bpf) and sum(map(len, groups)) == n
    res = h.get_finished()
    assert len(res) == 1 and res[0].job_id == 5 and res[0].success and res[0].transfer_size == n * 1000
    assert res[0].transfer_time >= 0 and res[0].transfer_type == ("GPU", "SHARED_STORAGE")


def test_handlers_shared_pending_and_unknown_jobs(kvb):
    W, M = kvb.worker, kvb.mediums
    fm = kvb.file_mapper.FileMapper("/r", "m", 16, 4, 1, 1, 1, 0, "x")
    eng = _FakeEngine()
    put = W.GPUToStorageHandler(4, fm, eng, ("GPU", "SHARED_STORAGE"), 10)
    get = W.StorageToGPUHandler(4, fm, eng, ("SHARED_STORAGE", "GPU"), 10)
    get._pending_jobs = put._pending_jobs      # single completion queue (worker.py:346-364)
    assert put.transfer_async(1, (M.GPULoadStoreSpec([1, 2, 3, 4, 5]), M.SharedStorageLoadStoreSpec([7, 8])))
    assert get.transfer_async(13, (M.SharedStorageLoadStoreSpec([7, 8]), M.GPULoadStoreSpec([1, 2, 3, 4, 5])))
    eng.finished.append((99, True))            # a job nobody recorded: still reported, without metrics
    res = {r.job_id: r for r in put.get_finished()}
    assert set(res) == {1, 13, 99} and res[13].success is False and res[99].transfer_size is None
    assert eng.calls[1][0] == "load" and eng.calls[1][3] == [[1], [2, 3, 4, 5]]
    put.wait({1, 13})
    assert ("wait", 1) in eng.calls and ("wait", 13) in eng.calls


def test_extra_features_and_suffix_match_oracle(kvb):
    K = kvb.kvblock
    rnd = random.Random(5)
    for _ in range(200):
        n = rnd.randrange(1, 300)
        bs = rnd.choice([1, 4, 16, 64])
        hashes = {"image": ["%064x" % rnd.getrandbits(256) for _ in range(rnd.randrange(0, 4))],
                  "audio": ["a%d" % rnd.randrange(100) for _ in range(rnd.randrange(0, 3))]}
        mk = lambda cls: {m: [cls(rnd.randrange(0, n), rnd.randrange(0, 50)) for _ in range(rnd.randrange(0, 4))]
                          for m in ("image", "audio") if rnd.random() < 0.8}
        state = rnd.getstate()
        ph_k = mk(K.PlaceholderRange)
        rnd.setstate(state)
        ph_o = mk(ko.PlaceholderRange)
        got = K.compute_block_extra_features(hashes, ph_k, bs, n)
        want = ko.compute_block_extra_features(hashes, ph_o, bs, n)
        assert (got is None) == (want is None)
        if got is None:
            continue
        assert len(got) == len(want) == n // bs
        for g, w in zip(got, want):
            assert (g is None) == (w is None)
            if g is not None:
                assert [m.hash for m in g.mm_hashes] == [m.hash for m in w.mm_hashes]
                assert K.encode_extra(g) == ko.encode_extra_suffix(w)
    assert K.encode_extra(None) == b"\xf6"
    assert K.encode_extra(K.BlockExtraFeatures([])) == b"\x80"
    long_id = "x" * 300
    assert K.encode_extra(K.BlockExtraFeatures([K.MMHash(long_id)])) == ko.cbor_canonical([ko.MMHash(long_id)])
    raw = [None, ["h1"], [["h2", 5]], [7], [], ["a", ["b", 1], 3]]
    got, want = K.parse_raw_extra_keys(raw), ko.parse_raw_extra_keys(raw)
    assert [None if g is None else [m.hash for m in g.mm_hashes] for g in got] == \
           [None if w is None else [m.hash for m in w.mm_hashes] for w in want]
    assert K.parse_raw_extra_keys(None) is None


def test_offload_oracle_layout_properties():
    """The oracle's own layout statement: tail alignment, file size floor, partial-tail reads."""
    rng = np.random.default_rng(0)
    T, N, frag, bpf = 3, 20, 64, 4
    tensors = [rng.integers(0, 256, (N, frag), dtype=np.uint8) for _ in range(T)]
    assert oo.staging_size(T, frag, bpf) == 16 * 1024 * 1024          # 16 MiB floor (thread_pool.cpp:35)
    assert oo.staging_size(64, 32768, 16) == 64 * 32768 * 16
    img = oo.file_image(tensors, [5, 9], bpf)
    off = oo.slot_offset(T, frag, bpf, 2)
    assert off == 2 * T * frag and img[:off].sum() == 0
    assert bytes(img[off:off + frag]) == bytes(tensors[0][5]) and bytes(img[off + frag:off + 2 * frag]) == bytes(tensors[1][5])
    full = oo.file_image(tensors, [1, 2, 3, 4], bpf)
    dst = [np.zeros_like(t) for t in tensors]
    oo.load_from_image(dst, [10, 11], bpf, full)                       # last two slots of the file
    for d, t in zip(dst, tensors):
        assert np.array_equal(d[10], t[3]) and np.array_equal(d[11], t[4]) and d.sum() == t[3].sum() + t[4].sum()
    p = oo.pack_blocks(tensors, [7, 7, 0])
    back = [np.zeros_like(t) for t in tensors]
    oo.unpack_blocks(back, [7, 7, 0], p)
    for b, t in zip(back, tensors):
        assert np.array_equal(b[7], t[7]) and np.array_equal(b[0], t[0])


def test_manager_and_spec(kvb, tmp_path):
    """manager.py:43-102 and spec.py:38-157 behaviours with plain config objects (no vLLM import)."""
    from types import SimpleNamespace as NS
    cfg = NS(cache_config=NS(block_size=16, cache_dtype="torch.float16"),
             parallel_config=NS(tensor_parallel_size=2, pipeline_parallel_size=1, prefill_context_parallel_size=1,
                                world_size=2, rank=0),
             model_config=NS(model="meta-llama/Llama-3-8B"),
             kv_transfer_config=NS(kv_connector_extra_config={"shared_storage_path": str(tmp_path), "block_size": 64,
                                                              "threads_per_gpu": 3}))
    spec = kvb.spec.SharedStorageOffloadingSpec(cfg)
    assert spec.gpu_blocks_per_file == 4 and spec.threads_per_gpu == 3
    assert spec.file_mapper.base_path == oo.base_path(str(tmp_path), "meta-llama/Llama-3-8B", 16, 4, 2, 1, 1, 0, "float16")
    mgr = spec.get_manager()
    hashes = [11, 22, 33]
    assert mgr.lookup(hashes) == 0
    import os
    for h in hashes[:2]:
        p = spec.file_mapper.get_file_name(h)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        open(p, "wb").close()
    assert mgr.lookup(hashes) == 2 and mgr.lookup([33, 11]) == 0           # consecutive hits from the start only
    out = mgr.prepare_store(iter(hashes))
    assert out.block_hashes_to_store == hashes and out.block_hashes_evicted == [] and out.store_spec.block_hashes == hashes
    assert mgr.prepare_load(hashes).medium() == "SHARED_STORAGE"
    with pytest.raises(AssertionError):
        kvb.spec.SharedStorageOffloadingSpec(cfg, extra_config={"block_size": 40})   # not a multiple of gpu block size
    cfg.parallel_config.rank = 1
    with pytest.raises(AssertionError):
        kvb.spec.SharedStorageOffloadingSpec(cfg).get_manager()               # scheduler rank must be 0


def test_worker_memory_plan_counts_the_hbm_chunk(kvb):
    """worker.py:303-319 clamps threads by host staging only; this engine also packs a chunk in HBM per worker, after
    vLLM has sized its KV cache — the plan bounds io_threads x chunk for both memories."""
    plan = kvb.worker.plan_worker_memory
    block = 2 << 20                                             # Llama-3-8B block
    # default: chunk = one file (16 blocks = 32 MiB), 1 GiB of HBM staging -> 32 workers, not 64 x 64 MiB = 4 GiB
    assert plan(block, 16, 64, 150) == (32, 32 << 20)
    assert plan(block, 16, 8, 150) == (8, 32 << 20)             # small pools are left alone
    assert plan(block, 16, 64, 150, {"chunk_bytes": 64 << 20}) == (16, 64 << 20)
    assert plan(block, 16, 64, 150, {"chunk_bytes": 1}) == (32, 32 << 20)          # a chunk always holds a whole file
    assert plan(block, 16, 64, 0.25) == (8, 32 << 20)           # the reference's host budget still binds (256 MiB)
    assert plan(block, 16, 64, 150, {"max_hbm_staging_mb": 4096}) == (64, 32 << 20)
    assert plan(block, 1024, 64, 150)[0] == 1                   # one 2 GiB file per worker: a single worker, never zero
