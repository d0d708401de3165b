#!/usr/bin/env python
"""Pageable vs pinned token buffers through the hash-only and the fused tokens -> scores calls (1024 x 1000 tokens).
Pageable memory is staged by the CUDA driver; an own staging pipeline through the library's pinned scratch was tried
and measured slower (0.38-0.43 vs 0.34 ms), so it was removed."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
K = kvb.kvblock
n, ntok, bs = 1024, 1000, 16
rng = np.random.default_rng(2)
tp = K.ChunkedTokenDatabase(bs, "")
tokens = rng.integers(0, 128256, n * ntok).astype(np.uint32)
off = np.arange(0, (n + 1) * ntok, ntok, dtype=np.int64)
parents = np.full(n, tp.get_init_hash("m"), dtype=np.uint64)
keys = np.empty(n * (ntok // bs), np.uint64)
koff = np.empty(n + 1, np.int64)
idx = K.Index(expected_keys=1 << 16)
pin = kvb.pool.PinnedBuffer(tokens.nbytes)
tpin = pin.numpy(np.uint32)
tpin[:] = tokens


def med(fn, it=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(it):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3


def hash_only(t):
    kvb._lib.check(kvb.lib.kvb_hash_token_blocks(0, t.ctypes.data, off.ctypes.data, parents.ctypes.data, n, bs, None, None,
                                                 keys.ctypes.data, koff.ctypes.data, None))


print("pageable: hash %.3f ms  fused %.3f ms" % (med(lambda: hash_only(tokens)), med(lambda: idx.score_tokens_flat(bs, tokens, off, parents))))
print("pinned  : hash %.3f ms  fused %.3f ms" % (med(lambda: hash_only(tpin)), med(lambda: idx.score_tokens_flat(bs, tpin, off, parents))))
