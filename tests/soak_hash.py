#!/usr/bin/env python
"""Fuzz of the device hash kernels against the oracle's C restatement (run on a GPU box, bounded by --seconds):
random batch sizes either side of every dispatch boundary, block sizes, token widths (1/2/3/5-byte CBOR forms),
short and long parents, ragged prompts, pre-encoded multimodal extras.  Both kernel families must agree with the
oracle bit for bit."""
import argparse
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import kvblock_oracle_c as oc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=25.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    rounds = keys_checked = with_extra = 0
    while time.time() < t_end:
        bs = int(rng.choice([4, 8, 16, 16, 16, 5, 17, 32, 64, 130, 300]))
        n = int(rng.choice([1, 2, 16, 17, 31, 32, 33, 64, 65, int(rng.integers(1, 300)), 1536, 1537, 2100]))
        if n > 600:
            lens = rng.integers(0, 6 * bs, n)
        else:
            lens = rng.integers(0, 40 * bs, n)
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum(lens)
        total = int(off[-1])
        width = rng.choice([4, 5, 8, 9, 16, 17, 20, 32], total)
        tokens = (rng.integers(0, 1 << 32, total, dtype=np.uint64) & ((np.uint64(1) << width.astype(np.uint64)) - np.uint64(1))).astype(np.uint32)
        parents = rng.integers(0, 1 << 63, n, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rng.integers(0, 2, n).astype(np.uint64)
        small = rng.random(n) < 0.15
        parents[small] = rng.choice([0, 1, 23, 24, 255, 256, 65535, 65536, (1 << 32) - 1, 1 << 32], int(small.sum())).astype(np.uint64)
        koff = np.zeros(n + 1, np.int64)
        koff[1:] = np.cumsum(lens // bs)
        nk = int(koff[-1])
        extra = extra_off = None
        if rng.random() < 0.3 and nk:
            # pre-encoded X(extra): arbitrary bytes per block, empty for most blocks (nil extra -> f6)
            sizes = np.where(rng.random(nk) < 0.3, rng.integers(1, 60, nk), 0)
            extra_off = np.zeros(nk + 1, np.int64)
            extra_off[1:] = np.cumsum(sizes)
            extra = rng.integers(0, 256, max(int(extra_off[-1]), 1), dtype=np.uint8)
            with_extra += 1
        want, woff = oc.hash_batch(tokens, off, parents, bs, extra, extra_off)
        for family in ("", "lanes", "wpc", "chain", "spec"):   # "spec": the table kernel wherever it applies (<= 64 prompts)
            if family:
                os.environ["KVB_HASH_KERNEL"] = family
            else:
                os.environ.pop("KVB_HASH_KERNEL", None)
            got = np.empty(max(nk, 1), np.uint64)
            goff = np.empty(n + 1, np.int64)
            kvb._lib.check(kvb.lib.kvb_hash_token_blocks(
                0, tokens.ctypes.data if total else None, off.ctypes.data, parents.ctypes.data, n, bs,
                None if extra is None else extra.ctypes.data, None if extra_off is None else extra_off.ctypes.data,
                got.ctypes.data, goff.ctypes.data, None))
            assert np.array_equal(goff, woff), (family, bs, n)
            assert np.array_equal(got[:nk], want), (family or "default", bs, n, int(np.argmax(got[:nk] != want)))
        os.environ.pop("KVB_HASH_KERNEL", None)
        rounds += 1
        keys_checked += 5 * nk
    print("hash fuzz ok: %d batches (%d with extras), %d keys compared with the C oracle across the five kernel selections (default, lanes, wpc, chain, spec)"
          % (rounds, with_extra, keys_checked))


if __name__ == "__main__":
    main()
