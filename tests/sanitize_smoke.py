#!/usr/bin/env python
"""Every kernel of libkvb.so at small sizes, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck  python tests/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitize_smoke.py
Results are checked against the oracle as in smoke()."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import kvblock_oracle as ko  # noqa: E402
from oracle import offload_oracle as oo  # noqa: E402

torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
K = kvb.kvblock

# paged copy: bulk + ldg, aligned and ragged fragments, gather / scatter / migrate
for T, N, frag in ((4, 24, 16384), (3, 17, 48), (2, 9, 100), (5, 12, 8208)):
    tensors = [torch.randint(0, 256, (N, frag), dtype=torch.uint8, device="cuda") for _ in range(T)]
    ref = [t.cpu().numpy() for t in tensors]
    pool = kvb.pool.KVPool(tensors)
    ids = np.random.default_rng(1).permutation(N)[: N // 2 + 1].astype(np.int64)
    packed = torch.zeros(ids.size * T * frag, dtype=torch.uint8, device="cuda")
    for variant in (1, 2):
        pool.gather(ids, packed, flags=variant)
        torch.cuda.synchronize()
        assert np.array_equal(packed.cpu().numpy(), oo.pack_blocks(ref, ids))
        pool.scatter(ids, packed, flags=variant)
    dst = [torch.zeros_like(t) for t in tensors]
    kvb.migrate.migrate_blocks(pool, kvb.pool.KVPool(dst), ids, ids)
    torch.cuda.synchronize()
    for d, t in zip(dst, tensors):
        assert torch.equal(d[ids], t[ids])

# engine (host arena), hash (both kernels, several block sizes, multimodal), index apply / lookup / score
tensors = [torch.randint(0, 256, (32, 4096), dtype=torch.uint8, device="cuda") for _ in range(4)]
eng = kvb.engine.StorageOffloadEngine(2, 4, tensors, 1, "disabled", 0.0, tier="host_arena", host_arena_bytes=8 << 20)
eng.async_store_gpu_blocks(1, ["a", "b"], [[1, 2, 3], [4, 5, 6, 7]])
while not eng.get_finished():
    pass
eng.async_load_gpu_blocks(2, ["a", "b"], [[1, 2, 3], [4, 5, 6, 7]])
while not eng.get_finished():
    pass
eng.shutdown()
rng = np.random.default_rng(0)
for family, bs in [(f, b) for f in ("", "lanes", "wpc", "chain", "spec") for b in (4, 16, 17, 64)]:  # warp-per-prompt and lane-per-prompt kernels
    if family:
        os.environ["KVB_HASH_KERNEL"] = family
    else:
        os.environ.pop("KVB_HASH_KERNEL", None)
    tp, otp = K.ChunkedTokenDatabase(bs, "s"), ko.TokenProcessor(bs, "s")
    prompts = [rng.integers(0, 1 << int(rng.choice([5, 8, 16, 17, 32])), int(rng.integers(0, 5 * bs + 3)), dtype=np.uint64).astype(np.uint32)
               for _ in range(40 if family == "spec" else 70)]   # the table kernel takes up to 64 prompts
    keys, off = tp.tokens_to_kv_block_keys_batch(prompts, "m")
    for i, p in enumerate(prompts):
        assert [int(k) for k in keys[off[i]:off[i + 1]]] == (otp.tokens_to_kv_block_keys(0, [int(x) for x in p], "m") or [])
os.environ.pop("KVB_HASH_KERNEL", None)
idx = K.Index(expected_keys=64)
tp, otp = K.ChunkedTokenDatabase(16, ""), ko.TokenProcessor(16, "")
prompts = [rng.integers(0, 128256, int(rng.integers(0, 700))).astype(np.uint32) for _ in range(40)]
for p in prompts:
    ks = otp.tokens_to_kv_block_keys(0, [int(x) for x in p], "m") or []
    if ks:
        d = int(rng.integers(1, len(ks) + 1))
        idx.add(None, ks[:d], [K.PodEntry("p%d" % int(rng.integers(0, 5)), "gpu")])
oidx = ko.InMemoryIndex()
idx2 = K.Index(size=40, pod_cache_size=3, expected_keys=16)          # at capacity: the sequential apply path + order array
o2 = ko.InMemoryIndex(size=40, pod_cache_size=3)
keyspace = [int(x) for x in rng.integers(1, 1 << 62, 120)]
for step in range(60):
    ks = [keyspace[int(i)] for i in rng.integers(0, 120, int(rng.integers(1, 30)))]
    e = ("p%d" % int(rng.integers(0, 6)), "gpu")
    idx2.add(None, ks, [K.PodEntry(*e)])
    o2.add(None, ks, [ko.PodEntry(*e)])
    if step % 7 == 0:
        idx2.evict(ks[0], K.REQUEST_KEY, [K.PodEntry(*e)])
        o2.evict(ks[0], ko.REQUEST_KEY, [ko.PodEntry(*e)])
    probe = keyspace[::5]
    assert set(idx2.lookup(probe)) == set(o2.lookup(probe)), step
big = [int(x) for x in rng.integers(1, 1 << 62, 5000)]               # parallel apply path + rehash
idx3 = K.Index(expected_keys=16)
idx3.add(None, big, [K.PodEntry("p", "gpu"), K.PodEntry("q", "cpu")])
idx3.add(None, big[:2500], [K.PodEntry("r", "gpu")])
assert len(idx3.lookup(big)) == 5000 and len(idx3) == 5000
oix = ko.Indexer(otp, oidx)
for p in prompts:
    ks = otp.tokens_to_kv_block_keys(0, [int(x) for x in p], "m") or []
ix = kvb.indexer.Indexer(tp, idx)
got = ix.score_tokens_batch(prompts, "m")                             # fused tokens -> scores launch (chain kernel + scorer warp)
assert len(got) == 40
one = ix.score_tokens(prompts[3], "m")                                # single-prompt form (arguments in the launch)
assert one == got[3] or (one is None and got[3] is None)
two = idx.score_tokens_flat(16, *tp.prepare_batch(prompts, "m")[:3], flags=kvb._lib.SCORE_TWO_KERNELS)
assert int(two[0].sum()) == sum(len(g or {}) for g in got)
assert idx.lookup([1, 2, 3]) == {}
print("sanitize smoke ok; kernels launched:", kvb.lib.kvb_launch_count())
