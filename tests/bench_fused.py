#!/usr/bin/env python
"""Fused tokens -> scores call (kvb_index_score_tokens_batch) in its variants, each checked against the oracle's C
restatement before it is timed:  one launch reading pinned tokens in place (default) | one launch after a token copy |
hash kernel + score kernel.  1024 prompts (BASELINE config #5 shape, smaller index) and one prompt (config #1).
    python tests/bench_fused.py [index_keys]      -> one JSON line"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import kvblock_oracle_c as oc  # noqa: E402  (checker)

BS, NTOK, MODEL = 16, 1000, "meta-llama/Llama-3-8B"


def med(fn, iters, warm):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


def main():
    torch.cuda.set_device(0)
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    K, L = kvb.kvblock, kvb._lib
    n_keys = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    rng = np.random.default_rng(2)
    tp = K.ChunkedTokenDatabase(BS, "")
    out = {}
    for n in (1024, 1):
        tokens = rng.integers(0, 128256, n * NTOK).astype(np.uint32)
        off = np.arange(0, (n + 1) * NTOK, NTOK, dtype=np.int64)
        parents = np.full(n, tp.get_init_hash(MODEL), dtype=np.uint64)
        keys_c, koff = oc.hash_batch(tokens, off, parents, BS)
        idx = K.Index(expected_keys=n_keys + (1 << 16))
        cix = oc.load().kvo_index_new(1 << int(np.ceil(np.log2(max(n_keys, 1024) * 2.5))))
        pods = ["10.0.%d.%d" % (i // 8, i % 8) for i in range(64)]
        for p in pods:
            idx.pods.get(p)
        tier = {"gpu": idx._tier_id("gpu"), "cpu": idx._tier_id("cpu")}
        bg = rng.integers(1, 1 << 63, n_keys, dtype=np.int64).astype(np.uint64)
        for c in range(0, n_keys, 50_000):
            ks = np.ascontiguousarray(bg[c:c + 50_000])
            p, t = int(rng.integers(0, 64)), ("gpu" if rng.random() < 0.8 else "cpu")
            idx.add(None, ks, [K.PodEntry(pods[p], t)])
            oc.load().kvo_index_add(cix, ks.ctypes.data, ks.size, p, tier[t])
        nk = NTOK // BS
        for i in range(n):
            d = int(rng.integers(0, nk + 1)) if n > 1 else nk
            for _ in range(int(rng.integers(1, 5))):
                if d == 0:
                    break
                dd = int(rng.integers(1, d + 1))
                chain = np.ascontiguousarray(keys_c[koff[i]:koff[i] + dd])
                p, t = int(rng.integers(0, 64)), ("gpu" if rng.random() < 0.8 else "cpu")
                idx.add(None, chain, [K.PodEntry(pods[p], t)])
                oc.load().kvo_index_add(cix, chain.ctypes.data, dd, p, tier[t])
        w = np.ones(256)
        w[tier["cpu"]] = 0.8
        c_n, c_p, c_s = np.zeros(n, np.int32), np.zeros(n * 13, np.uint16), np.zeros(n * 13, np.float64)
        oc.load().kvo_score_batch(cix, keys_c.ctypes.data, koff.ctypes.data, n, w.ctypes.data, c_n.ctypes.data, c_p.ctypes.data,
                                  c_s.ctypes.data, 0)
        want = [{int(c_p[p * 13 + j]): float(c_s[p * 13 + j]) for j in range(int(c_n[p]))} for p in range(n)]
        pin_t = kvb.pool.PinnedBuffer(tokens.nbytes)
        tok_pin = pin_t.numpy(np.uint32)
        tok_pin[:] = tokens
        pin_o = kvb.pool.PinnedBuffer(n * 136 + 1024)
        raw = pin_o.numpy(np.uint8)
        b1 = (n * 4 + 255) // 256 * 256
        b2 = b1 + (n * 26 + 255) // 256 * 256
        o_pin = (raw[:n * 4].view(np.int32), raw[b1:b1 + n * 26].view(np.uint16), raw[b2:b2 + n * 104].view(np.float64))
        o_pg = (np.zeros(n, np.int32), np.zeros(n * 13, np.uint16), np.zeros(n * 13, np.float64))
        row = {}
        variants = {
            "fused_in_place": (tok_pin, o_pin, 0),
            "fused_token_copy": (tok_pin, o_pin, L.SCORE_COPY_TOKENS),
            "two_kernels": (tok_pin, o_pin, L.SCORE_TWO_KERNELS | L.SCORE_COPY_TOKENS),
            "fused_pageable_buffers": (tokens, o_pg, 0),
        }
        variants["fused_in_place_pinned_io"] = (tok_pin, o_pin, L.SCORE_PINNED_IO)
        for merged in ("auto", "0", "1"):
            if merged == "auto":
                os.environ.pop("KVB_HASH_MERGED", None)
            else:
                os.environ["KVB_HASH_MERGED"] = merged
            for name, (tk, o, fl) in variants.items():
                if merged != "auto" and not name.startswith("fused_in_place"):
                    continue
                for a in o:
                    a[:] = 0
                idx.score_tokens_flat(BS, tk, off, parents, out=o, flags=fl)
                got = [{int(o[1][p * 13 + j]): float(o[2][p * 13 + j]) for j in range(int(o[0][p]))} for p in range(n)]
                assert got == want, (n, name, merged)
                row[f"{name}_merged_{merged}"] = round(med(lambda: idx.score_tokens_flat(BS, tk, off, parents, out=o, flags=fl),
                                                           iters=300 if n == 1 else 60, warm=30 if n == 1 else 10), 2)
        os.environ.pop("KVB_HASH_MERGED", None)
        if n > 1:  # the token buffer on transparent huge pages (fewer IOMMU / ATS translations for 1024 scattered streams?)
            pin_h = kvb.pool.PinnedBuffer(tokens.nbytes, huge_pages=True)
            tok_h = pin_h.numpy(np.uint32)[:tokens.size]
            tok_h[:] = tokens
            o = o_pin
            for a in o:
                a[:] = 0
            idx.score_tokens_flat(BS, tok_h, off, parents, out=o, flags=L.SCORE_PINNED_IO)
            got = [{int(o[1][p * 13 + j]): float(o[2][p * 13 + j]) for j in range(int(o[0][p]))} for p in range(n)]
            assert got == want, (n, "thp")
            row["fused_in_place_tokens_on_huge_pages"] = round(med(lambda: idx.score_tokens_flat(BS, tok_h, off, parents, out=o, flags=L.SCORE_PINNED_IO),
                                                                   iters=60, warm=10), 2)
            del tok_h
            pin_h.free()
        if n > 1:  # token-fetch geometry of the in-place form (chunk tokens x chunks ahead)
            for geo in ("128x2", "256x2", "256x4", "128x6"):
                os.environ["KVB_CHAIN_FETCH"] = geo
                tk, o, fl = variants["fused_in_place_pinned_io"]
                for a in o:
                    a[:] = 0
                idx.score_tokens_flat(BS, tk, off, parents, out=o, flags=fl)
                got = [{int(o[1][p * 13 + j]): float(o[2][p * 13 + j]) for j in range(int(o[0][p]))} for p in range(n)]
                assert got == want, (n, geo)
                row[f"fused_in_place_fetch_{geo}"] = round(med(lambda: idx.score_tokens_flat(BS, tk, off, parents, out=o, flags=fl),
                                                               iters=60, warm=10), 2)
            os.environ.pop("KVB_CHAIN_FETCH", None)
        # the same call as a host-language shim makes it: arguments bound once, no Python wrapper around the C entry point
        fn = kvb.lib.kvb_index_score_tokens_batch
        args = (idx._h, tok_pin.ctypes.data, off.ctypes.data, parents.ctypes.data, n, BS, None, None, None, 0, L.SCORE_PINNED_IO,
                o_pin[0].ctypes.data, o_pin[1].ctypes.data, o_pin[2].ctypes.data)
        for a in o_pin:
            a[:] = 0
        assert fn(*args) == 0
        got = [{int(o_pin[1][p * 13 + j]): float(o_pin[2][p * 13 + j]) for j in range(int(o_pin[0][p]))} for p in range(n)]
        assert got == want, (n, "raw abi")
        row["raw_c_abi_pinned_io"] = round(med(lambda: fn(*args), iters=500 if n == 1 else 80, warm=50 if n == 1 else 10), 2)
        out[str(n)] = row
        idx.close()
        pin_t.free()
        pin_o.free()
    print(json.dumps({"index_keys": n_keys, "us_per_call": out, "bit_exact_vs_oracle": True}))


if __name__ == "__main__":
    main()
