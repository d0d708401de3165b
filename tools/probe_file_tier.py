#!/usr/bin/env python
"""File tier on tmpfs: throughput (store / load GB/s) and one-file latency for a sweep of worker counts and write paths
(pwrite per worker vs shared-mapping copy, KVB_FILE_WRITE is read once per process -> one subprocess per mode).
    python tools/probe_file_tier.py            -> one JSON line"""
import importlib
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T, FRAG, POOL, BPF, N = 64, 32768, 6144, 16, 4096
BLOCK = T * FRAG


def drain(eng, job):
    while True:
        for j, ok in eng.get_finished():
            if j == job:
                assert ok
                return
        time.sleep(0.0003)


def child():
    import torch
    torch.cuda.set_device(0)
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    big = torch.empty((T, POOL, FRAG), dtype=torch.uint8, device="cuda")
    big.random_(0, 256)
    tensors = list(big.unbind(0))
    ids = np.random.default_rng(1).permutation(POOL)[:N].astype(np.int64)
    groups = [ids[i * BPF:(i + 1) * BPF].tolist() for i in range(N // BPF)]
    payload = N * BLOCK
    rows = {}
    for threads in [int(x) for x in os.environ.get("PROBE_THREADS", "8,16,24,32").split(",")]:
        root = f"/dev/shm/kvb_probe_{threads}"
        eng = kvb.engine.StorageOffloadEngine(threads, BPF, tensors, max(1, int(threads * 0.75)), "disabled", 0.0, tier="file",
                                              chunk_bytes=BPF * BLOCK)
        best = None
        for rep in range(3):
            files = [f"{root}/r{rep}/{i:06d}.bin" for i in range(len(groups))]
            t0 = time.perf_counter()
            assert eng.async_store_gpu_blocks(1, files, groups)
            drain(eng, 1)
            t1 = time.perf_counter()
            assert eng.async_load_gpu_blocks(2, files, groups)
            drain(eng, 2)
            t2 = time.perf_counter()
            shutil.rmtree(f"{root}/r{rep}", ignore_errors=True)
            cur = (payload / (t1 - t0) / 1e9, payload / (t2 - t1) / 1e9)
            if rep and (best is None or sum(cur) > sum(best)):
                best = cur
        st, ld = [], []
        for j in range(30):
            f = [f"{root}/lat/{j:04d}.bin"]
            g = [groups[j]]
            t0 = time.perf_counter()
            assert eng.async_store_gpu_blocks(100 + 2 * j, f, g)
            drain(eng, 100 + 2 * j)
            t1 = time.perf_counter()
            assert eng.async_load_gpu_blocks(101 + 2 * j, f, g)
            drain(eng, 101 + 2 * j)
            t2 = time.perf_counter()
            if j >= 5:
                st.append(t1 - t0)
                ld.append(t2 - t1)
        eng.shutdown()
        shutil.rmtree(root, ignore_errors=True)
        rows[str(threads)] = {"store_gbs": round(best[0], 2), "load_gbs": round(best[1], 2),
                              "save_plus_load_gbs": round(2 / (1 / best[0] + 1 / best[1]), 2),
                              "one_file_store_ms": round(float(np.median(st)) * 1e3, 2),
                              "one_file_load_ms": round(float(np.median(ld)) * 1e3, 2)}
    print(json.dumps(rows))


def main():
    out = {}
    modes = {"spread_nodes": {"KVB_FILE_SPREAD": "1"}, "gpu_node_only": {"KVB_FILE_SPREAD": "0"},
             "gpu_node_only_pwrite": {"KVB_FILE_SPREAD": "0", "KVB_FILE_WRITE": "pwrite"},
             "gpu_node_only_mmap": {"KVB_FILE_SPREAD": "0", "KVB_FILE_WRITE": "mmap"}}
    if os.environ.get("PROBE_MODES"):
        modes = {k: v for k, v in modes.items() if k in os.environ["PROBE_MODES"].split(",")}
    for mode, extra in modes.items():
        env = dict(os.environ, PROBE_CHILD="1", **extra)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=900)
        try:
            out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            out[mode] = {"error": (r.stderr or r.stdout)[-400:]}
    print(json.dumps({"blocks": N, "payload_gb": N * BLOCK / 1e9, "tmpfs": "/dev/shm", "modes": out}))


if __name__ == "__main__":
    child() if os.environ.get("PROBE_CHILD") else main()
