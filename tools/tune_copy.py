#!/usr/bin/env python
"""Sweep the paged-copy kernel variants on one GPU and print achieved HBM GB/s (2 x payload / time).
Not a bench line: a tuning aid whose output is summarised under profiles/."""
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
kvb = importlib.import_module("llm-d-kv-cache_b200")


def time_kernel(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "8b"
    T, frag = (64, 32768) if shape == "8b" else (160, 16384)
    N, n = (6144, 5000) if shape == "8b" else (4915, 4000)
    torch.cuda.set_device(0)
    big = torch.empty((T, N, frag), dtype=torch.uint8, device="cuda")
    big.random_(0, 256)
    tensors = list(big.unbind(0))
    pool = kvb.pool.KVPool(tensors)
    ids = torch.from_numpy(np.random.default_rng(1).permutation(N)[:n].astype(np.int64)).cuda()
    packed = torch.empty(n * T * frag, dtype=torch.uint8, device="cuda")
    payload = n * T * frag
    out = []
    # plain contiguous copy of the same byte count: the "measured peak" reference on this box
    src = torch.empty(payload, dtype=torch.uint8, device="cuda")
    med, best = time_kernel(lambda: packed.copy_(src))
    print(f"torch copy_ {payload/1e9:.2f} GB: median {2*payload/med/1e6:.0f} GB/s best {2*payload/best/1e6:.0f} GB/s", flush=True)
    out.append({"cfg": "torch_copy", "gbs_median": 2 * payload / med / 1e6})
    del src
    cfgs = []
    for unroll in (2, 4, 8):
        for ctas in (2, 4, 6, 8):
            cfgs.append(("ldg u%d c%d" % (unroll, ctas), 1 | (unroll << 8) | (ctas << 12)))
    for deep in (1, 2, 3):
        for ctas in (1, 2, 3, 4):
            for pc in (3, 4, 5):   # 8, 16, 32 KiB
                cfgs.append(("bulk d%d c%d p%dK" % (deep - 1, ctas, 1 << pc), 2 | (deep << 8) | (ctas << 12) | (pc << 20)))
    for name, flags in cfgs:
        try:
            g_med, g_best = time_kernel(lambda: pool.gather_dev(ids, packed, flags=flags))
            s_med, s_best = time_kernel(lambda: pool.scatter_dev(ids, packed, flags=flags))
        except Exception as e:
            print(name, "FAILED", e, flush=True)
            continue
        rec = {"cfg": name, "flags": flags, "gather_gbs": 2 * payload / g_med / 1e6, "gather_best": 2 * payload / g_best / 1e6,
               "scatter_gbs": 2 * payload / s_med / 1e6, "scatter_best": 2 * payload / s_best / 1e6}
        out.append(rec)
        print(f"{name:18s} gather {rec['gather_gbs']:7.0f} (best {rec['gather_best']:7.0f})  scatter {rec['scatter_gbs']:7.0f} (best {rec['scatter_best']:7.0f}) GB/s", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/tune_copy_{shape}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
