"""Cross-GPU KV-block migration over NVLink (kvb.h section 5).  No reference counterpart: the
reference has no cross-GPU path; only the paged layout is shared.

One process per GPU.  The destination rank exports its KV tensors over CUDA IPC, the source rank
maps them and ONE kernel reads source pages and writes destination pages (gather fused with the
peer write and the scatter).  torch.distributed carries only the control plane (handles, ids)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import IpcMem, check
from .pool import KVPool, _ids_array, _stream_ptr


def export_pool(pool: KVPool) -> dict:
    """Picklable description of a pool's tensors (CUDA IPC handles), for the peer that will write/read it."""
    lib = _lib.load()
    mems = []
    for t in pool._tensors:
        m = IpcMem()
        check(lib.kvb_ipc_export(pool.device, t.data_ptr(), C.byref(m)))
        mems.append((bytes(m.handle), int(m.offset)))
    return {"device": pool.device, "num_blocks": pool.num_blocks, "frag_bytes": pool.frag_bytes, "mems": mems}


class RemotePool(KVPool):
    """A peer GPU's pool mapped into this process (pointers usable from ``local_device``)."""

    def __init__(self, desc: dict, local_device: int):
        lib = _lib.load()
        self._imported = []
        ptrs = []
        cache: dict = {}
        for handle, offset in desc["mems"]:
            base = cache.get(handle)
            if base is None:
                m = IpcMem()
                C.memmove(m.handle, handle, 64)
                m.offset = 0
                p = C.c_void_p()
                check(lib.kvb_ipc_import(local_device, C.byref(m), C.byref(p)))
                base = cache[handle] = p.value
                self._imported.append(base)
            ptrs.append(base + offset)
        super().__init__(None, local_device, ptrs=ptrs, num_blocks=desc["num_blocks"],
                         frag_bytes=desc["frag_bytes"], stride_bytes=desc["frag_bytes"])
        check(lib.kvb_pool_mark_peer(self.handle, 1))

    def close(self):
        super().close()
        for base in getattr(self, "_imported", []):
            _lib.load().kvb_ipc_close(self.device, base, 0)
        self._imported = []


def migrate_blocks(src: KVPool, dst: KVPool, src_ids: Sequence[int], dst_ids: Sequence[int], stream=None,
                   flags: int = 0) -> None:
    """src pages src_ids[i] -> dst pages dst_ids[i]; ``dst`` is a local pool or a RemotePool."""
    s, d = _ids_array(src_ids), _ids_array(dst_ids)
    if s.size != d.size:
        raise ValueError("src_ids and dst_ids differ in length")
    check(_lib.load().kvb_migrate_blocks(src.handle, dst.handle, s.ctypes.data_as(C.POINTER(C.c_int64)),
                                         d.ctypes.data_as(C.POINTER(C.c_int64)), s.size, _stream_ptr(stream), flags))


def enable_peer_access(device: int, peer: int) -> None:
    check(_lib.load().kvb_enable_peer_access(device, peer))
