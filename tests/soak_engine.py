#!/usr/bin/env python
"""Randomised soak of the offload engine against a model of its contract (run on a GPU box, bounded by --seconds):
many small jobs in flight, both tiers, a host arena small enough to evict, waits that cancel, duplicate stores,
loads of absent files — every block that the engine reports loaded must hold exactly the bytes that were stored
under that file name (checked against the oracle's packing of the pool at store time)."""
import argparse
import importlib
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import offload_oracle as oo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40.0)
    ap.add_argument("--tier", default="host_arena", choices=["host_arena", "file"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    rng = np.random.default_rng(args.seed)
    T, N, frag, bpf = 6, 512, 4096, 4
    block = T * frag
    tensors = [torch.randint(0, 256, (N, frag), dtype=torch.uint8, device="cuda") for _ in range(T)]
    root = "/dev/shm/kvb_soak_%d" % os.getpid()
    shutil.rmtree(root, ignore_errors=True)
    # arena holds ~40 files of 4 blocks: stores beyond that evict the least recently used unpinned entries
    eng = kvb.engine.StorageOffloadEngine(8, bpf, tensors, 6, "disabled", 0.0, tier=args.tier,
                                          host_arena_bytes=40 * bpf * block, chunk_bytes=8 * bpf * block,
                                          strict_load_errors=True)
    stored = {}       # file -> packed bytes (numpy) of what was stored under that name
    job = 0
    stats = dict(stores=0, loads=0, loads_ok=0, loads_missing=0, blocks_checked=0, waits=0, dup_stores=0)
    t_end = time.time() + args.seconds

    def drain(ids):
        want, got = set(ids), {}
        t0 = time.time()
        while want - set(got) and time.time() - t0 < 60:
            for j, ok in eng.get_finished():
                got[j] = ok
        assert not (want - set(got)), "jobs never finished: %s" % (want - set(got))
        return got

    while time.time() < t_end:
        # ---- a burst of store jobs (new files, and sometimes a file that already exists)
        pending = {}
        dup_targets = set()
        for _ in range(int(rng.integers(1, 6))):
            nf = int(rng.integers(1, 5))
            files, groups = [], []
            for _ in range(nf):
                if stored and rng.random() < 0.2:
                    f = list(stored)[int(rng.integers(0, len(stored)))]
                    dup_targets.add(f)
                    stats["dup_stores"] += 1
                else:
                    f = "%s/%03d/%06d.bin" % (root, int(rng.integers(0, 7)), int(rng.integers(0, 10 ** 6)))
                files.append(f)
                groups.append([int(x) for x in rng.permutation(N)[: int(rng.integers(1, bpf + 1))]])
            job += 1
            snap = [t.cpu().numpy() for t in tensors]
            assert eng.async_store_gpu_blocks(job, files, groups)
            pending[job] = (files, groups, snap)
            stats["stores"] += 1
        if rng.random() < 0.3:                      # wait_job cancels what is still queued, then blocks
            j = list(pending)[0]
            eng.wait_job(j)
            stats["waits"] += 1
        res = drain(pending)
        for j, (files, groups, snap) in pending.items():
            assert res[j], "store job %d failed" % j
            for f, g in zip(files, groups):
                if f not in stored and eng.exists(f):  # first writer wins; a cancelled store may not have written
                    stored[f] = oo.pack_blocks(snap, np.asarray(g, dtype=np.int64))
        if args.tier == "host_arena":
            # an arena entry may have been evicted by this very burst before its duplicate store ran, in which case the
            # duplicate was a real store of different bytes: the model cannot know which, so it forgets those names
            for f in dup_targets:
                stored.pop(f, None)
        # ---- mutate the pool, then load a mix of present and absent files into fresh places
        for t in tensors:
            t.random_(0, 256)
        names = list(stored)
        pending = {}
        for _ in range(int(rng.integers(1, 6))):
            nf = int(rng.integers(1, 4))
            files, groups = [], []
            for _ in range(nf):
                if names and rng.random() < 0.85:
                    f = names[int(rng.integers(0, len(names)))]
                    nb = stored[f].size // block
                    take = int(rng.integers(1, nb + 1))   # the LAST `take` blocks of the file (tail-aligned)
                else:
                    f, take = "%s/none/%06d.bin" % (root, int(rng.integers(0, 10 ** 6))), int(rng.integers(1, bpf + 1))
                files.append(f)
                groups.append([int(x) for x in rng.permutation(N)[:take]])
            job += 1
            present = [eng.exists(f) for f in files]
            assert eng.async_load_gpu_blocks(job, files, groups)
            pending[job] = (files, groups, present)
            stats["loads"] += 1
        res = drain(pending)
        torch.cuda.synchronize()
        # jobs run concurrently, so only pages written by exactly one (file, job) are compared
        counts = {}
        for j in pending:
            for f, g, p in zip(*pending[j]):
                if p and f in stored:
                    for b in g:
                        counts[b] = counts.get(b, 0) + 1
        host = [t.cpu().numpy() for t in tensors]
        for j, (files, groups, present) in pending.items():
            all_present = all(p and f in stored for f, p in zip(files, present))
            if all_present:
                assert res[j], "load job %d of present files failed" % j
                stats["loads_ok"] += 1
            elif not res[j]:
                stats["loads_missing"] += 1
            for f, g, p in zip(files, groups, present):
                if not (p and f in stored) or not eng.exists(f):
                    continue                            # absent, or evicted while the burst ran
                nb = stored[f].size // block
                tail = stored[f].reshape(nb, T, frag)[nb - len(g):]
                for k, b in enumerate(g):
                    if counts[b] != 1:
                        continue
                    for ti in range(T):
                        assert np.array_equal(host[ti][b], tail[k, ti]), "block %d of %s differs" % (b, f)
                    stats["blocks_checked"] += 1
        # forget files the arena evicted (file tier never evicts)
        for f in [f for f in stored if not eng.exists(f)]:
            del stored[f]
    eng.shutdown()
    shutil.rmtree(root, ignore_errors=True)
    assert stats["blocks_checked"] > 100, stats
    print("soak ok (%s tier): %s" % (args.tier, stats))


if __name__ == "__main__":
    main()
