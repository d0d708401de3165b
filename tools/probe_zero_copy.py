#!/usr/bin/env python
"""Can the gather / scatter kernels target pinned HOST memory directly (fused gather+D2H / H2D+scatter, no HBM staging)?
Measures GB/s over PCIe for both movers against the staged path (kernel to HBM + one cudaMemcpyAsync)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.set_device(0)
kvb = importlib.import_module("llm-d-kv-cache_b200")
T, N, frag, n = 64, 4096, 32768, 1024
tensors = list(torch.randint(0, 256, (T, N, frag), dtype=torch.uint8, device="cuda").unbind(0))
ref = [t.clone() for t in tensors]
pool = kvb.pool.KVPool(tensors)
ids = torch.from_numpy(np.random.default_rng(1).permutation(N)[:n].astype(np.int64)).cuda()
payload = n * T * frag
host = torch.empty(payload, dtype=torch.uint8).pin_memory()
dev = torch.empty(payload, dtype=torch.uint8, device="cuda")

def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return iters * payload / a.elapsed_time(b) / 1e6

def staged_store():
    pool.gather_dev(ids, dev); host.copy_(dev, non_blocking=True)
def staged_load():
    dev.copy_(host, non_blocking=True); pool.scatter_dev(ids, dev)
print(f"staged  store (gather->HBM, memcpy D2H): {timed(staged_store):6.1f} GB/s   load: {timed(staged_load):6.1f} GB/s")
for name, flags in (("ldg u4 c8", 1 | (4 << 8) | (8 << 12)), ("ldg u8 c4", 1 | (8 << 8) | (4 << 12)), ("bulk default", 2),
                    ("bulk d1 c2 p16K", 2 | (2 << 8) | (2 << 12) | (4 << 20))):
    try:
        s = timed(lambda: pool.gather_dev(ids, host, flags=flags))
        host_copy = host.clone()
        for t in tensors: t[ids] = 0
        l = timed(lambda: pool.scatter_dev(ids, host, flags=flags))
        ok = all(torch.equal(t, r) for t, r in zip(tensors, ref))
        print(f"direct  {name:16s} store {s:6.1f} GB/s   load {l:6.1f} GB/s   bit-exact={ok}")
    except Exception as e:
        print(name, "FAILED", e)
