#!/usr/bin/env python
"""Sweep the page->page migration kernel over NVLink (torchrun, >= 2 ranks): ring r -> r+1, all ranks at once.
Prints GB/s of egress per GPU for each kernel configuration (tuning aid; summary kept under profiles/)."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    part, mig = kvb.partition, kvb.migrate
    shape = sys.argv[1] if len(sys.argv) > 1 else "8b"
    T, frag = (64, 32768) if shape == "8b" else (160, 16384)
    N, n = 6144, 2048
    big = torch.empty((T, N, frag), dtype=torch.uint8, device="cuda")
    big.random_(0, 256)
    tensors = list(big.unbind(0))
    pool = kvb.pool.KVPool(tensors)
    descs = part.exchange_objects(mig.export_pool(pool), dist)
    dst, src = part.ring_peers(rank, world)
    remote = mig.RemotePool(descs[dst], local)
    src_ids = np.random.default_rng(rank).permutation(N // 2)[:n].astype(np.int64)
    dst_ids = (N // 2 + np.random.default_rng(9 + rank).permutation(N // 2)[:n]).astype(np.int64)
    payload = n * T * frag

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(flags, only_rank0=False, iters=5):
        def go():
            if only_rank0 and rank != 0:
                return
            mig.migrate_blocks(pool, remote, src_ids, dst_ids, flags=flags)
        for _ in range(2):
            go()
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            go()
        b.record()
        sync()
        t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfgs = []
    for unroll in (4, 8):
        for ctas in (2, 4, 8):
            cfgs.append(("ldg u%d c%d" % (unroll, ctas), 1 | (unroll << 8) | (ctas << 12)))
    for deep in (1, 2, 3):
        for ctas in (1, 2, 3, 4):
            for pc in (2, 3, 4, 5):   # 4, 8, 16, 32 KiB
                cfgs.append(("bulk d%d c%d p%dK" % (deep - 1, ctas, 1 << pc), 2 | (deep << 8) | (ctas << 12) | (pc << 20)))
    out = []
    for name, flags in cfgs:
        try:
            ring = timed(flags)
            uni = timed(flags, only_rank0=True)
        except Exception as e:
            if rank == 0:
                print(name, "FAILED", e, flush=True)
            continue
        rec = {"cfg": name, "flags": flags, "ring_gbs_per_gpu": payload / ring / 1e6, "unidir_gbs": payload / uni / 1e6}
        out.append(rec)
        if rank == 0:
            print(f"{name:18s} ring {rec['ring_gbs_per_gpu']:7.1f} GB/s/GPU   one-way {rec['unidir_gbs']:7.1f} GB/s", flush=True)
    # plain peer memcpy of a contiguous buffer: the NVLink ceiling probe on this box
    a_buf = torch.empty(payload, dtype=torch.uint8, device="cuda")
    peer_ptr_t = tensors[0]  # unused; torch peer copy needs a tensor on the peer: use NCCL send/recv instead
    recv = torch.empty(payload, dtype=torch.uint8, device="cuda")

    def nccl_ring():
        ops = [dist.P2POp(dist.isend, a_buf, dst), dist.P2POp(dist.irecv, recv, src)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for _ in range(2):
        nccl_ring()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        nccl_ring()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"nccl send/recv ring of a contiguous {payload/1e9:.2f} GB buffer: {payload / float(t.item()) / 1e6:.1f} GB/s/GPU", flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"world": world, "shape": shape, "rows": out, "nccl_contiguous_ring_gbs": payload / float(t.item()) / 1e6},
                  open(f"gpurun_out/tune_migrate_{shape}_n{world}.json", "w"), indent=1)
    remote.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
