#!/usr/bin/env python
"""Single-process, two-GPU driver for an ncu capture of the NVLink page->page migration kernel:
source pool on cuda:0, destination pool on cuda:1, peer access enabled in-process (no second rank, so ncu is safe).
Prints GB/s from CUDA events when run without ncu."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

kvb = importlib.import_module("llm-d-kv-cache_b200")
T, N, frag, n = 64, 4096, 32768, 2048
torch.cuda.set_device(0)
kvb.migrate.enable_peer_access(0, 1)
src_t = list(torch.randint(0, 256, (T, N, frag), dtype=torch.uint8, device="cuda:0").unbind(0))
dst_t = list(torch.zeros((T, N, frag), dtype=torch.uint8, device="cuda:1").unbind(0))
src = kvb.pool.KVPool(src_t, 0)
dst = kvb.pool.KVPool(None, 0, ptrs=[t.data_ptr() for t in dst_t], num_blocks=N, frag_bytes=frag, stride_bytes=frag)
kvb._lib.check(kvb.lib.kvb_pool_mark_peer(dst.handle, 1))
rng = np.random.default_rng(0)
s_ids, d_ids = rng.permutation(N)[:n].astype(np.int64), rng.permutation(N)[:n].astype(np.int64)
payload = n * T * frag
for name, flags in (("default(ldg u4 c8)", 0), ("bulk", 2)):
    for _ in range(2):
        kvb.migrate.migrate_blocks(src, dst, s_ids, d_ids, flags=flags)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        kvb.migrate.migrate_blocks(src, dst, s_ids, d_ids, flags=flags)
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {3 * payload / a.elapsed_time(b) / 1e6:.1f} GB/s one-way GPU0->GPU1 ({payload/1e9:.2f} GB per launch)")
torch.cuda.synchronize(1)
for st, dt in zip(src_t[::16], dst_t[::16]):
    assert torch.equal(dt[torch.from_numpy(d_ids).to("cuda:1")].cpu(), st[torch.from_numpy(s_ids).cuda()].cpu())
print("bit-exact")
