#!/usr/bin/env python
"""Driver for an ncu capture of the paged-copy kernels on the 70B-fp8 block shape (160 tensors x 16 KiB fragments):
one gather and one scatter of N blocks (default 8000 = 21 GB per pass) from a 12288-block pool."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
T, FRAG, POOL = 160, 16384, 12288
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
big = torch.empty((T, POOL, FRAG), dtype=torch.uint8, device="cuda")
big.random_(0, 256)
pool = kvb.pool.KVPool(list(big.unbind(0)))
ids = torch.from_numpy(np.random.default_rng(3).permutation(POOL)[:n].astype(np.int64)).cuda()
packed = torch.empty(n * T * FRAG, dtype=torch.uint8, device="cuda")
pool.gather_dev(ids, packed)
pool.scatter_dev(ids, packed)
torch.cuda.synchronize()
print("done", n, "blocks,", n * T * FRAG, "bytes per pass")
