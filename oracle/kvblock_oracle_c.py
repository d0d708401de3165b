"""ctypes loader of the oracle's C restatement (oracle/kvblock_oracle.c) — CPU-baseline timing only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libkvblock_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = C.CDLL(SO)
        lib.kvo_fnv64a.restype = C.c_uint64
        lib.kvo_fnv64a.argtypes = [C.c_void_p, C.c_size_t]
        lib.kvo_init_hash.restype = C.c_uint64
        lib.kvo_init_hash.argtypes = [C.c_uint64, C.c_char_p, C.c_size_t]
        lib.kvo_hash_batch.restype = None
        lib.kvo_hash_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
        lib.kvo_index_new.restype = C.c_void_p
        lib.kvo_index_new.argtypes = [C.c_uint64]
        lib.kvo_index_free.argtypes = [C.c_void_p]
        lib.kvo_index_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint16, C.c_uint8]
        lib.kvo_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def hash_batch(tokens: np.ndarray, prompt_off: np.ndarray, parents: np.ndarray, block_size: int, extra=None,
               extra_off=None, threads: int = 0):
    lib = load()
    n = len(prompt_off) - 1
    key_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.diff(prompt_off) // block_size, out=key_off[1:])
    keys = np.empty(max(int(key_off[-1]), 1), dtype=np.uint64)
    lib.kvo_hash_batch(tokens.ctypes.data, prompt_off.ctypes.data, parents.ctypes.data, n, block_size,
                       None if extra is None else extra.ctypes.data, None if extra_off is None else extra_off.ctypes.data,
                       keys.ctypes.data, key_off.ctypes.data, threads)
    return keys[:int(key_off[-1])], key_off


def init_hash(seed_hash: int, model: str) -> int:
    raw = model.encode()
    return int(load().kvo_init_hash(seed_hash, raw, len(raw)))
