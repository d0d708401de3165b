#!/usr/bin/env python
"""Driver for ncu captures of the index kernels: a 1 M-key device-side build (sort + index_apply_par_kernel), then the
fused tokens -> scores launch (chain_kernel<.., SCORE>) and the two-kernel form (hash + index_score_kernel) for 1024 prompts
with pinned buffers, one single-prompt call and a 16-prompt call (the table kernel), then the eviction planner of an index at
capacity (add-only batch, mixed batch)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
K, L = kvb.kvblock, kvb._lib
rng = np.random.default_rng(2)
tp = K.ChunkedTokenDatabase(16, "")
n, ntok = 1024, 1000
idx = K.Index(expected_keys=(1 << 20) + (1 << 16))
for c in range(4):
    idx.add(None, rng.integers(1, 1 << 63, 1 << 18, dtype=np.int64).astype(np.uint64), [K.PodEntry("pod-%d" % c, "gpu")])
idx.flush()
pin_t = kvb.pool.PinnedBuffer(n * ntok * 4)
tok = pin_t.numpy(np.uint32)
tok[:] = rng.integers(0, 128256, n * ntok).astype(np.uint32)
off = np.arange(0, (n + 1) * ntok, ntok, dtype=np.int64)
parents = np.full(n, tp.get_init_hash("m"), dtype=np.uint64)
keys, koff = tp.tokens_to_kv_block_keys_batch([tok[off[i]:off[i + 1]] for i in range(64)], "m")
for i in range(64):
    idx.add(None, keys[koff[i]:koff[i] + 40], [K.PodEntry("pod-%d" % (i % 8), "gpu")])
pin_o = kvb.pool.PinnedBuffer(n * 136 + 1024)
raw = pin_o.numpy(np.uint8)
b1 = (n * 4 + 255) // 256 * 256
b2 = b1 + (n * 26 + 255) // 256 * 256
out = (raw[:n * 4].view(np.int32), raw[b1:b1 + n * 26].view(np.uint16), raw[b2:b2 + n * 104].view(np.float64))
for _ in range(2):
    idx.score_tokens_flat(16, tok, off, parents, out=out, flags=L.SCORE_PINNED_IO)
    idx.score_tokens_flat(16, tok, off, parents, out=out, flags=L.SCORE_PINNED_IO | L.SCORE_TWO_KERNELS | L.SCORE_COPY_TOKENS)
    idx.score_tokens_flat(16, tok[:ntok], off[:2], parents[:1], out=out, flags=L.SCORE_PINNED_IO)
print("done")
# round 2, later: the table ("spec") kernel on 1 and 16 prompts (fused), and the eviction planner on an index at capacity
# (an add-only batch, then one that mixes Add and Evict)
for _ in range(2):
    idx.score_tokens_flat(16, tok[:16 * ntok], off[:17], parents[:16], out=out, flags=L.SCORE_PINNED_IO)
cap = 1 << 18
idx_c = K.Index(size=cap, expected_keys=cap)
fill = rng.integers(1, 1 << 62, cap + 60_000, dtype=np.int64).astype(np.uint64)
ent = [K.PodEntry("pod-0", "gpu")]
idx_c.add(None, fill[:cap], ent)
idx_c.flush()
idx_c.add(None, fill[cap:cap + 30_000], ent)
idx_c.flush()
for c in range(10):
    idx_c.add(None, fill[cap + 30_000 + c * 3000:cap + 30_000 + (c + 1) * 3000], ent)
    for k in fill[cap - 1000 * (c + 1):cap - 1000 * c:10]:
        idx_c.evict(int(k), K.REQUEST_KEY, ent)
idx_c.flush()
st = idx_c.stats()
assert st["flushes_planned"] == 2 and st["plan_fallbacks"] == 0, st
print("planner done", st["lru_evictions"])
