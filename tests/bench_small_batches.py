#!/usr/bin/env python
"""Small batches (1 .. 64 prompts of 1000 tokens) through the fused tokens -> scores call, pinned buffers read and written
in place, raw C-ABI call: the table ("spec") kernel against the chain kernel, each checked against the oracle's C
restatement before it is timed; plus the single-core C restatement of the same work (hash + lookup + score).
    python tests/bench_small_batches.py [index_keys]      -> one JSON line"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import kvblock_oracle_c as oc  # noqa: E402  (checker and CPU baseline)

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
    C = oc.load()
    idx = K.Index(expected_keys=n_keys + (1 << 16))
    cix = C.kvo_index_new(1 << int(np.ceil(np.log2(max(n_keys, 1024) * 2.5))))
    pods = ["10.0.%d.%d" % (i // 8, i % 8) for i in range(64)]
    for p in pods:
        idx.pods.get(p)
    tier = {"gpu": idx._tier_id("gpu"), "cpu": idx._tier_id("cpu")}
    bg = rng.integers(1, 1 << 63, n_keys, dtype=np.int64).astype(np.uint64)
    for c in range(0, n_keys, 50_000):
        ks = np.ascontiguousarray(bg[c:c + 50_000])
        p, t = int(rng.integers(0, 64)), ("gpu" if rng.random() < 0.8 else "cpu")
        idx.add(None, ks, [K.PodEntry(pods[p], t)])
        C.kvo_index_add(cix, ks.ctypes.data, ks.size, p, tier[t])
    n_max = 64
    tokens = rng.integers(0, 128256, n_max * NTOK).astype(np.uint32)
    off_all = np.arange(0, (n_max + 1) * NTOK, NTOK, dtype=np.int64)
    parents_all = np.full(n_max, tp.get_init_hash(MODEL), dtype=np.uint64)
    keys_c, koff = oc.hash_batch(tokens, off_all, parents_all, BS)
    nk = NTOK // BS
    for i in range(n_max):  # every prompt has cached prefixes on a few pods (prompt 0: the whole chain, config #1's shape)
        d = nk if i == 0 else int(rng.integers(1, nk + 1))
        for _ in range(int(rng.integers(1, 5))):
            dd = int(rng.integers(1, d + 1))
            chain = np.ascontiguousarray(keys_c[koff[i]:koff[i] + dd])
            p, t = int(rng.integers(0, 64)), ("gpu" if rng.random() < 0.8 else "cpu")
            idx.add(None, chain, [K.PodEntry(pods[p], t)])
            C.kvo_index_add(cix, chain.ctypes.data, dd, p, tier[t])
    w = np.ones(256)
    w[tier["cpu"]] = 0.8
    pin_t = kvb.pool.PinnedBuffer(tokens.nbytes)
    tok_pin = pin_t.numpy(np.uint32)
    tok_pin[:] = tokens
    pin_o = kvb.pool.PinnedBuffer(n_max * 136 + 1024)
    raw = pin_o.numpy(np.uint8)
    out = {}
    fn = kvb.lib.kvb_index_score_tokens_batch
    for n in (1, 2, 4, 8, 16, 32, 64):
        off = np.ascontiguousarray(off_all[:n + 1])
        parents = np.ascontiguousarray(parents_all[:n])
        c_n, c_p, c_s = np.zeros(n, np.int32), np.zeros(n * 13, np.uint16), np.zeros(n * 13, np.float64)
        ck = np.zeros(n * nk, np.uint64)
        ckoff = np.ascontiguousarray(koff[:n + 1])

        def cpu_call():   # one thread: hash the chains, then probe and walk
            C.kvo_hash_batch(tokens.ctypes.data, off.ctypes.data, parents.ctypes.data, n, BS, None, None, ck.ctypes.data,
                             ckoff.ctypes.data, 1)
            C.kvo_score_batch(cix, ck.ctypes.data, ckoff.ctypes.data, n, w.ctypes.data, c_n.ctypes.data, c_p.ctypes.data,
                              c_s.ctypes.data, 1)
        cpu_call()
        want = [{int(c_p[p * 13 + j]): float(c_s[p * 13 + j]) for j in range(int(c_n[p]))} for p in range(n)]
        assert any(want)
        b1 = (n * 4 + 255) // 256 * 256
        b2 = b1 + (n * 26 + 255) // 256 * 256
        o = (raw[:n * 4].view(np.int32), raw[b1:b1 + n * 26].view(np.uint16), raw[b2:b2 + n * 104].view(np.float64))
        args = (idx._h, tok_pin.ctypes.data, off.ctypes.data, parents.ctypes.data, n, BS, None, None, None, 0, L.SCORE_PINNED_IO,
                o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data)
        row = {"cpu_c_restatement_1_thread": round(med(cpu_call, 200 if n <= 4 else 40, 10), 2)}
        for name, env in (("spec_kernel", {"KVB_HASH_KERNEL": "spec"}), ("chain_kernel", {"KVB_HASH_SPEC": "0"}), ("default", {})):
            for k in ("KVB_HASH_KERNEL", "KVB_HASH_SPEC"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for a in o:
                a[:] = 0
            assert fn(*args) == 0
            got = [{int(o[1][p * 13 + j]): float(o[2][p * 13 + j]) for j in range(int(o[0][p]))} for p in range(n)]
            assert got == want, (n, name)
            row[name] = round(med(lambda: fn(*args), 400, 50), 2)
        for k in ("KVB_HASH_KERNEL", "KVB_HASH_SPEC"):
            os.environ.pop(k, None)
        out[str(n)] = row
    print(json.dumps({"index_keys": n_keys, "tokens_per_prompt": NTOK, "block_size": BS, "us_per_call": out,
                      "bit_exact_vs_oracle": True}))


if __name__ == "__main__":
    main()
