#!/usr/bin/env python
"""Tiny driver for ncu captures of the hash and score kernels (1024 prompts x 1000 tokens)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
K = kvb.kvblock
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(2)
tp = K.ChunkedTokenDatabase(16, "")
tokens = torch.from_numpy(rng.integers(0, 128256, n * 1000).astype(np.uint32)).pin_memory().numpy()
off = np.arange(0, (n + 1) * 1000, 1000, dtype=np.int64)
parents = np.full(n, tp.get_init_hash("m"), dtype=np.uint64)
idx = K.Index(expected_keys=1 << 16)
keys, koff = tp.tokens_to_kv_block_keys_batch([tokens[off[i]:off[i + 1]] for i in range(min(n, 256))], "m")
for i in range(min(n, 256)):
    idx.add(None, keys[koff[i]:koff[i] + 30], [K.PodEntry("p%d" % (i % 64), "gpu")])
for _ in range(4):
    idx.score_tokens_flat(16, tokens, off, parents)
print("done")
