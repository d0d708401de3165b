#!/usr/bin/env python
"""Randomised sweep of the index's mutation logic on the HOST-SIMULATION build (no GPU): index sizes 1 … 2 500, batches of
1 … 500 keys, 1 … 13 pods per key, with and without Evict ops and lookups in between; every read is compared with the
oracle (tests/test_index_sim.py::_random_traffic).  Bounded by the first argument (seconds).
    python tests/soak_index_sim.py 600"""
import ctypes as C
import importlib
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tests.test_index_sim as T  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("build_index_sim", os.path.join(ROOT, "tests", "cpp", "build_index_sim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    kvb = importlib.import_module("llm-d-kv-cache_b200")
    kvb._lib.bind(lib, T.SIM_NAMES)
    lib.kvb_last_error = lambda: b"(host simulation)"
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    t0, n = time.time(), 0
    tot = {k: 0 for k in ("flushes_parallel", "flushes_planned", "plan_fallbacks", "flushes_sequential", "replay_resumes",
                          "order_scans", "lru_evictions", "rehashes")}
    r = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    while time.time() - t0 < budget:
        cfg = dict(seed=r.randrange(1 << 30), size=r.choice([1, 2, 3, 5, 17, 40, 100, 300, 1000, 2500]),
                   max_batch=r.choice([1, 3, 10, 40, 150, 500]), lookup_frac=r.choice([0.1, 0.3, 0.5]),
                   evict_frac=r.choice([0.0, 0.02, 0.2, 0.35]), ppk=r.choice([1, 2, 3, 4, 10, 13]), steps=r.choice([120, 300]))
        cfg["n_keys"] = int(cfg["size"] * r.choice([1.2, 1.7, 3, 6])) + r.randrange(5, 60)
        try:
            st = T._random_traffic(kvb, lib, **cfg)
        except AssertionError as e:
            print("DIVERGENCE from the oracle:", cfg, str(e)[:300])
            sys.exit(1)
        n += 1
        for k in tot:
            tot[k] += st[k]
    print("index sim sweep ok: %d scenarios in %.0f s, every read equal to the oracle; totals %s" % (n, time.time() - t0, tot))


if __name__ == "__main__":
    main()
