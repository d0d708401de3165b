"""world_size-2 gloo tests (CPU) of the N>1 host logic: partition planning and the control-plane exchange."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = importlib.import_module("llm-d-kv-cache_b200.partition")
        # every rank owns an independent KV partition: block ids and checksums are per-rank, exchanged as objects
        lo, hi = part.shard_range(10001, world, rank)
        rng = np.random.default_rng(100 + rank)
        tensors = [rng.integers(0, 256, (64, 256), dtype=np.uint8) for _ in range(3)]
        ids = rng.permutation(64)[:20]
        mine = {"rank": rank, "range": (lo, hi), "sum": part.page_checksum_np(tensors, ids), "ids": ids.tolist()}
        everyone = part.exchange_objects(mine, dist)
        assert [e["rank"] for e in everyone] == list(range(world))
        assert everyone[0]["range"][0] == 0 and everyone[-1]["range"][1] == 10001
        assert all(everyone[i]["range"][1] == everyone[i + 1]["range"][0] for i in range(world - 1))
        dst, src = part.ring_peers(rank, world)
        # "migration" of the control plane: my ring source's ids must be what that rank published
        rng_src = np.random.default_rng(100 + src)
        t_src = [rng_src.integers(0, 256, (64, 256), dtype=np.uint8) for _ in range(3)]
        ids_src = rng_src.permutation(64)[:20]
        assert everyone[src]["ids"] == ids_src.tolist()
        assert everyone[src]["sum"] == part.page_checksum_np(t_src, ids_src)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_control_plane():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_partition_planning(kvb):
    part = kvb.partition
    for world in (1, 2, 3, 8):
        spans = [part.shard_range(10000, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 10000
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert part.ring_peers(0, 8) == (1, 7) and part.ring_peers(7, 8) == (0, 6) and part.ring_peers(0, 1) == (0, 0)
    plan = part.fanout_plan(2048, 8, root=0)
    assert sorted(plan) == list(range(1, 8)) and plan[1] == (0, 2048) and plan[7] == (6 * 2048, 7 * 2048)
    with pytest.raises(ValueError):
        part.shard_range(10, 2, 2)
    t = [np.arange(64 * 16, dtype=np.uint8).reshape(64, 16)]
    assert part.page_checksum_np(t, [1, 2]) == part.page_checksum_np(t, [2, 1])
