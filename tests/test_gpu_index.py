"""Index contract + scorer parity (through the C ABI) against the oracle and the reference's known answers."""
import numpy as np
import pytest

from oracle import kvblock_oracle as o
from tests import scenarios as sc

pytestmark = pytest.mark.gpu


def _pe(kvb, pod, tier, spec=False):
    return kvb.kvblock.PodEntry(pod, tier, spec)


def _as_tuples(d):
    return {int(k): [(e.pod_identifier, e.device_tier, bool(e.speculative)) for e in v] for k, v in d.items()}


@pytest.mark.parametrize("case", sc.SCORER_CASES, ids=lambda c: c[0])
def test_scorer_known_answers(kvb, torch_cuda, case):
    _, weights, keys, hit, want = case
    idx = kvb.kvblock.Index(medium_weights=weights)
    for k, pods in hit.items():
        idx.add([k], [k], [_pe(kvb, p, t) for p, t in pods])
    got = kvb.indexer.LongestPrefixScorer(idx).score(keys)
    assert got == want


@pytest.mark.parametrize("case", sc.INDEXER_CASES, ids=lambda c: c[0])
def test_indexer_known_answers(kvb, torch_cuda, case):
    _, keys, entries, flt, want = case
    idx = kvb.kvblock.Index()
    for k, pods in entries.items():
        idx.add([k], [k], [_pe(kvb, p, t) for p, t in pods])
    got = kvb.indexer.LongestPrefixScorer(idx).score(keys, flt)
    oidx = o.InMemoryIndex()
    for k, pods in entries.items():
        oidx.add([k], [k], [o.PodEntry(p, t) for p, t in pods])
    exact = o.longest_prefix_score(keys, oidx.lookup(keys, flt))
    assert got == exact                                   # bit-exact float64 against the oracle
    assert set(got) == set(want) and all(abs(got[p] - want[p]) < 1e-4 for p in want)


def test_score_tokens_end_to_end(kvb, torch_cuda):
    """Indexer.ScoreTokens: hash -> lookup -> score, vs the oracle Indexer on the same state."""
    K = kvb.kvblock
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, 128256, 1000).astype(np.uint32)           # BASELINE config #1
    model = "meta-llama/Llama-3-8B"
    tp, otp = K.ChunkedTokenDatabase(16, ""), o.TokenProcessor(16, "")
    keys = tp.tokens_to_kv_block_keys(0, tokens, model)
    assert len(keys) == 62
    ix = kvb.indexer.Indexer(tp)
    oix = o.Indexer(otp)
    for i in range(4):
        held = keys[: 62 * (i + 1) // 4]
        ix.index.add(held, held, [_pe(kvb, "pod-%d" % i, "gpu")])
        oix.index.add(held, held, [o.PodEntry("pod-%d" % i, "gpu")])
    ix.index.add(keys[:20], keys[:20], [_pe(kvb, "pod-3", "cpu")])
    oix.index.add(keys[:20], keys[:20], [o.PodEntry("pod-3", "cpu")])
    want = oix.score_tokens([int(t) for t in tokens], model)
    assert ix.score_tokens(tokens, model) == want
    assert want == {"pod-0": 15.0, "pod-1": 31.0, "pod-2": 46.0, "pod-3": 62.0}
    assert ix.score_tokens(tokens, model, ["pod-1", "pod-2"]) == oix.score_tokens([int(t) for t in tokens], model, ["pod-1", "pod-2"])
    assert ix.score_tokens(tokens[:10], model) is None                   # no full block -> nil
    assert ix.score_tokens(tokens, "other-model") == {}


def test_index_contract(kvb, torch_cuda):
    """index_test.go:119-264,589-735; in_memory_test.go:45-236 — same assertions as tests/test_oracle_kvblock.py."""
    K = kvb.kvblock
    P = lambda p, t, s=False: K.PodEntry(p, t, s)
    idx = K.Index()
    with pytest.raises(ValueError):
        idx.lookup([])
    with pytest.raises(ValueError):
        idx.add([1], [], [P("p", "gpu")])
    with pytest.raises(ValueError):
        idx.evict(1, K.ENGINE_KEY, [])
    idx.add([1, 2], [11, 12], [P("p1", "gpu"), P("p2", "gpu")])
    assert idx.lookup([11, 12]) == {11: [P("p1", "gpu"), P("p2", "gpu")], 12: [P("p1", "gpu"), P("p2", "gpu")]}
    idx.add([1], [11], [P("p1", "gpu")])
    assert len(idx.lookup([11])[11]) == 2
    assert idx.lookup([11, 12], {"p1"}) == {11: [P("p1", "gpu")], 12: [P("p1", "gpu")]}
    assert idx.lookup([11], {"nobody"}) == {}
    assert idx.lookup([999, 11]).keys() == {11}
    idx.add([3], [13], [P("p3", "gpu"), P("p3", "cpu")])
    idx.evict(3, K.ENGINE_KEY, [P("p3", "cpu")])
    assert idx.lookup([13]) == {13: [P("p3", "gpu")]}
    idx.add([20, 21, 22, 23], [30], [P("p", "gpu")])
    assert idx.get_request_key(20) == 30 and idx.get_request_key(23) == 30
    idx.evict(21, K.ENGINE_KEY, [P("p", "gpu")])
    assert idx.lookup([30]) == {}
    idx.add([40], [50, 51, 52, 53], [P("p", "gpu")])
    assert idx.get_request_key(40) == 53
    with pytest.raises(KeyError):
        idx.get_request_key(12345)
    idx.evict(777, K.ENGINE_KEY, [P("p", "gpu")])
    idx.add(None, [60], [P("p", "gpu", True)])
    idx.add([61], [60], [P("p", "gpu", False)])
    assert idx.lookup([60])[60] == [P("p", "gpu", True), P("p", "gpu", False)]
    idx.evict(60, K.REQUEST_KEY, [P("p", "gpu", True)])
    assert idx.lookup([60])[60] == [P("p", "gpu", False)]
    small = K.Index(size=2, pod_cache_size=2)
    for k in (1, 2, 3):
        small.add([k], [k], [P("p", "gpu")])
    assert small.lookup([1, 2, 3]).keys() == {2, 3}
    small.add([2], [2], [P("a", "gpu"), P("b", "gpu"), P("c", "gpu")])
    assert small.lookup([2])[2] == [P("b", "gpu"), P("c", "gpu")]
    with pytest.raises(Exception):
        K.Index(pod_cache_size=14)      # more than a device bucket holds: refused, not truncated


def test_random_ops_match_oracle(kvb, torch_cuda):
    """Random Add / Evict / Lookup / Score traffic: every read must equal the oracle's, including LRU effects."""
    K = kvb.kvblock
    rng = np.random.default_rng(123)
    idx, oidx = K.Index(size=300, pod_cache_size=4, expected_keys=64), o.InMemoryIndex(size=300, pod_cache_size=4)
    pods = ["pod-%d" % i for i in range(12)]
    tiers = ["gpu", "cpu", "GPU", "disk"]
    keyspace = [int(x) for x in rng.integers(1, 1 << 63, 500)]
    for step in range(1500):
        op = rng.random()
        if op < 0.45:
            n = int(rng.integers(1, 6))
            rks = [keyspace[int(i)] for i in rng.integers(0, len(keyspace), n)]
            mode = rng.integers(0, 3)
            eks = None if mode == 0 else [int(x) for x in rng.integers(1, 200, n if mode == 1 else 1)]
            ents = [(pods[int(rng.integers(0, 12))], tiers[int(rng.integers(0, 4))], bool(rng.integers(0, 2)))
                    for _ in range(int(rng.integers(1, 4)))]
            idx.add(eks, rks, [K.PodEntry(*e) for e in ents])
            oidx.add(eks, rks, [o.PodEntry(*e) for e in ents])
        elif op < 0.65:
            ents = [(pods[int(rng.integers(0, 12))], tiers[int(rng.integers(0, 4))], bool(rng.integers(0, 2)))
                    for _ in range(int(rng.integers(1, 3)))]
            if rng.random() < 0.5:
                k, kt = int(rng.integers(1, 200)), 0
            else:
                k, kt = keyspace[int(rng.integers(0, len(keyspace)))], 1
            idx.evict(k, kt, [K.PodEntry(*e) for e in ents])
            oidx.evict(k, kt, [o.PodEntry(*e) for e in ents])
        else:
            ks = [keyspace[int(i)] for i in rng.integers(0, len(keyspace), int(rng.integers(1, 40)))]
            flt = [] if rng.random() < 0.5 else [pods[int(i)] for i in rng.integers(0, 12, 3)] + ["unknown-pod"]
            got = _as_tuples(idx.lookup(ks, flt))
            want = _as_tuples(oidx.lookup(ks, flt))
            assert got == want, step
            sgot = kvb.indexer.LongestPrefixScorer(idx).score(ks, flt)
            assert sgot == o.longest_prefix_score(ks, oidx.lookup(ks, flt), {"gpu": 1.0, "cpu": 0.8}), step
        if step % 100 == 0:
            assert len(idx) == len(oidx.data)


def test_batch_scoring_matches_oracle(kvb, torch_cuda):
    """Config #5 in miniature: many prompts x many pods against a larger index, fused tokens->scores."""
    K = kvb.kvblock
    rng = np.random.default_rng(2)
    tp, otp = K.ChunkedTokenDatabase(16, ""), o.TokenProcessor(16, "")
    idx, oidx = K.Index(expected_keys=1 << 16), o.InMemoryIndex()
    model = "m"
    pods = ["10.0.0.%d" % i for i in range(64)]
    prompts = []
    for p in range(96):
        n = int(rng.integers(0, 1100))
        prompts.append(rng.integers(0, 128256, n).astype(np.uint32))
    # index a random prefix of every prompt on random pods / tiers
    for p, toks in enumerate(prompts):
        keys = otp.tokens_to_kv_block_keys(0, [int(t) for t in toks], model) or []
        for _ in range(int(rng.integers(0, 5))):
            depth = int(rng.integers(0, len(keys) + 1))
            if depth == 0:
                continue
            ent = (pods[int(rng.integers(0, 64))], "gpu" if rng.random() < 0.8 else "cpu")
            idx.add(None, keys[:depth], [K.PodEntry(*ent)])
            oidx.add(None, keys[:depth], [o.PodEntry(*ent)])
    ix = kvb.indexer.Indexer(tp, idx)
    got = ix.score_tokens_batch(prompts, model)
    oix = o.Indexer(otp, oidx)
    for p, toks in enumerate(prompts):
        assert got[p] == oix.score_tokens([int(t) for t in toks], model), p
    flt = pods[:7]
    gotf = ix.score_tokens_batch(prompts, model, flt)
    for p, toks in enumerate(prompts):
        assert gotf[p] == oix.score_tokens([int(t) for t in toks], model, flt), p


def test_table_growth_and_tombstones(kvb, torch_cuda):
    K = kvb.kvblock
    idx, oidx = K.Index(expected_keys=16), o.InMemoryIndex()
    keys = [int(x) for x in np.random.default_rng(4).integers(1, 1 << 63, 20000)]
    e, oe = [K.PodEntry("p", "gpu")], [o.PodEntry("p", "gpu")]
    for i in range(0, 20000, 1000):
        idx.add(None, keys[i:i + 1000], e)
        oidx.add(None, keys[i:i + 1000], oe)
        assert set(idx.lookup(keys[i:i + 1000])) == set(keys[i:i + 1000])
    for k in keys[:5000]:
        idx.evict(k, K.REQUEST_KEY, e)
        oidx.evict(k, o.REQUEST_KEY, oe)
    idx.add(None, keys[:2500], e)
    oidx.add(None, keys[:2500], oe)
    got = idx.lookup(keys)
    assert set(got) == set(oidx.lookup(keys)) and len(got) == 17500


def test_concurrent_stress(kvb, torch_cuda):
    """The reference's contract suite hammers every backend from many goroutines (index_test.go:267-585).  Same idea
    with threads: adders, evictors, lookers and scorers run concurrently; the index must stay consistent —
    keys only ever touched by adders end up with exactly the expected entries."""
    import threading
    K = kvb.kvblock
    idx = K.Index(expected_keys=1 << 14)
    n_threads, per = 8, 300
    stable = {t: [int(x) for x in np.random.default_rng(100 + t).integers(1, 1 << 62, per)] for t in range(n_threads)}
    churn = [int(x) for x in np.random.default_rng(7).integers(1, 1 << 62, 500)]
    errors = []

    def worker(t):
        try:
            rng = np.random.default_rng(t)
            ent = [K.PodEntry("pod-%d" % t, "gpu")]
            for i in range(per):
                idx.add([stable[t][i] ^ 0x5555], [stable[t][i]], ent)
                k = churn[int(rng.integers(0, len(churn)))]
                op = rng.random()
                if op < 0.4:
                    idx.add(None, [k], [K.PodEntry("pod-%d" % int(rng.integers(0, 4)), "cpu")])
                elif op < 0.6:
                    idx.evict(k, K.REQUEST_KEY, [K.PodEntry("pod-%d" % int(rng.integers(0, 4)), "cpu")])
                elif op < 0.8:
                    got = idx.lookup([k, stable[t][i]])
                    assert stable[t][i] in got
                else:
                    kvb.indexer.LongestPrefixScorer(idx).score([stable[t][i], k])
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
    for t in range(n_threads):
        got = idx.lookup(stable[t])
        assert len(got) == per
        assert all(v == [K.PodEntry("pod-%d" % t, "gpu")] for v in got.values())
        assert idx.get_request_key(stable[t][0] ^ 0x5555) == stable[t][0]


def test_pinned_token_buffer_scores_like_pageable(kvb, torch_cuda):
    """kvb_host_alloc memory (pinned, NUMA-local) takes the no-staging path of the fused scorer; same results."""
    K = kvb.kvblock
    rng = np.random.default_rng(21)
    tp = K.ChunkedTokenDatabase(16, "")
    idx = K.Index()
    n, ntok = 40, 333
    tokens = rng.integers(0, 128256, n * ntok).astype(np.uint32)
    off = np.arange(0, (n + 1) * ntok, ntok, dtype=np.int64)
    parents = np.full(n, tp.get_init_hash("m"), dtype=np.uint64)
    keys, koff = tp.tokens_to_kv_block_keys_batch([tokens[off[i]:off[i + 1]] for i in range(n)], "m")
    for i in range(0, n, 2):
        idx.add(None, keys[koff[i]:koff[i] + 1 + i % 7], [K.PodEntry("pod-%d" % (i % 5), "gpu")])
    want = idx.score_tokens_flat(16, tokens, off, parents)
    buf = kvb.pool.PinnedBuffer(tokens.nbytes)
    pinned = buf.numpy(np.uint32)
    pinned[:] = tokens
    got = idx.score_tokens_flat(16, pinned, off, parents)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    del pinned
    buf.free()
    buf.free()  # idempotent


@pytest.mark.parametrize("bs", [16, 4])
def test_large_pinned_batch_matches_pageable_and_oracle(kvb, torch_cuda, bs):
    """701 prompts of pinned tokens read in place by ONE fused launch: ragged lengths, empty prompts, a pod filter; against
    the oracle and against the same call on pageable tokens (which are copied first)."""
    K = kvb.kvblock
    rng = np.random.default_rng(77 + bs)
    tp, otp = K.ChunkedTokenDatabase(bs, ""), o.TokenProcessor(bs, "")
    idx, oidx = K.Index(), o.InMemoryIndex()
    n = 701
    lens = rng.integers(0, 40 * bs, n)
    lens[[0, 350, 700]] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tokens = rng.integers(0, 128256, int(off[-1])).astype(np.uint32)
    parents = np.full(n, tp.get_init_hash("m"), dtype=np.uint64)
    prompts = [tokens[off[i]:off[i + 1]] for i in range(n)]
    keys, koff = tp.tokens_to_kv_block_keys_batch(prompts, "m")
    pods = ["pod-%d" % i for i in range(9)]
    for i in range(0, n, 3):
        nk = int(koff[i + 1] - koff[i])
        if nk:
            d = int(rng.integers(1, nk + 1))
            ent = (pods[int(rng.integers(0, 9))], "gpu" if rng.random() < 0.7 else "cpu")
            ks = [int(k) for k in keys[koff[i]:koff[i] + d]]
            idx.add(None, ks, [K.PodEntry(*ent)])
            oidx.add(None, ks, [o.PodEntry(*ent)])
    buf = kvb.pool.PinnedBuffer(max(tokens.nbytes, 64))
    pinned = buf.numpy(np.uint32)[:tokens.size]
    pinned[:] = tokens
    idx.flush()
    launches0 = kvb.lib.kvb_launch_count()
    got = idx.score_tokens_flat(bs, pinned, off, parents)
    assert kvb.lib.kvb_launch_count() - launches0 == 1          # tokens -> scores is one launch
    ref = idx.score_tokens_flat(bs, tokens, off, parents)        # pageable tokens: one launch
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    for i in range(0, n, 7):
        ks = otp.tokens_to_kv_block_keys(0, [int(x) for x in prompts[i]], "m") or []
        want = o.longest_prefix_score(ks, oidx.lookup(ks, []), {"gpu": 1.0, "cpu": 0.8}) if ks else {}
        mine = {idx.pods.names[int(got[1][i * 13 + j])]: float(got[2][i * 13 + j]) for j in range(int(got[0][i]))}
        assert mine == want, i
    flt = ["pod-1", "pod-4"]
    got_f = idx.score_tokens_flat(bs, pinned, off, parents, pod_identifiers=flt)
    ref_f = idx.score_tokens_flat(bs, tokens, off, parents, pod_identifiers=flt)
    for a, b in zip(ref_f, got_f):
        assert np.array_equal(a, b)
    del pinned
    buf.free()


def test_scoring_refreshes_recency_by_default(kvb, torch_cuda):
    """Capacity pressure with SCORING between adds: the reference's ScoreTokens goes through Lookup, whose data.Get
    refreshes every key it finds (in_memory.go:120), so which keys the outer LRU evicts depends on what was scored.
    The batched scoring paths do the same by default (the kernel stamps the slots); `touch_lru=False` opts out."""
    K = kvb.kvblock
    rng = np.random.default_rng(77)
    idx, oidx = K.Index(size=64, pod_cache_size=3, expected_keys=16), o.InMemoryIndex(size=64, pod_cache_size=3)
    keyspace = [int(x) for x in rng.integers(1, 1 << 63, 160)]
    pods = ["pod-%d" % i for i in range(6)]
    for step in range(400):
        r = rng.random()
        if r < 0.5:
            ks = [keyspace[int(i)] for i in rng.integers(0, len(keyspace), int(rng.integers(1, 5)))]
            ent = (pods[int(rng.integers(0, 6))], "gpu" if rng.random() < 0.7 else "cpu")
            idx.add(None, ks, [K.PodEntry(*ent)])
            oidx.add(None, ks, [o.PodEntry(*ent)])
        else:
            # a batch of "prompts" scored in one call: the oracle scores them one after the other
            prompts = [[keyspace[int(i)] for i in rng.integers(0, len(keyspace), int(rng.integers(1, 12)))]
                       for _ in range(int(rng.integers(1, 6)))]
            flat = np.asarray([k for p in prompts for k in p], dtype=np.uint64)
            off = np.cumsum([0] + [len(p) for p in prompts]).astype(np.int64)
            got = idx.score_keys_batch(flat, off)
            for p, g in zip(prompts, got):
                assert g == o.longest_prefix_score(p, oidx.lookup(p), {"gpu": 1.0, "cpu": 0.8}), step
        if step % 50 == 49:
            assert set(idx.lookup(keyspace)) == set(oidx.lookup(keyspace)), step   # same survivors
    assert idx.stats()["lru_evictions"] > 20
    # opt-out: read-only scoring leaves the eviction order alone
    e = [K.PodEntry("p", "gpu")]
    ro = K.Index(size=3, pod_cache_size=2)
    for k in (1, 2, 3):
        ro.add(None, [k], e)
    ro.score_keys_batch(np.asarray([1], dtype=np.uint64), np.asarray([0, 1], dtype=np.int64), touch_lru=False)
    ro.add(None, [4], e)
    assert set(ro.lookup([1, 2, 3, 4])) == {2, 3, 4}          # 1 was scored read-only: still the oldest, evicted


def test_device_side_build_uses_the_parallel_path(kvb, torch_cuda):
    """A bulk build never touches a host copy: ops are queued, sorted and replayed per key on the device."""
    K = kvb.kvblock
    rng = np.random.default_rng(3)
    idx, oidx = K.Index(expected_keys=1 << 12), o.InMemoryIndex()
    keys = [int(x) for x in rng.integers(1, 1 << 63, 50000)]
    for i in range(0, 50000, 5000):
        ents = [("pod-%d" % int(rng.integers(0, 30)), "gpu" if rng.random() < 0.8 else "cpu") for _ in range(int(rng.integers(1, 5)))]
        chunk = keys[i:i + 5000] + keys[max(0, i - 700):i]           # overlaps: the same key gets entries from several calls
        idx.add(None, chunk, [K.PodEntry(*e) for e in ents])
        oidx.add(None, chunk, [o.PodEntry(*e) for e in ents])
    st = idx.stats()
    assert st["live_keys"] == len(oidx.data) == 50000 and st["flushes_parallel"] >= 1 and st["rehashes"] >= 1
    sample = [keys[int(i)] for i in rng.integers(0, 50000, 3000)]
    assert _as_tuples(idx.lookup(sample)) == _as_tuples(oidx.lookup(sample))


def test_at_capacity_add_only_batches_are_planned_on_the_device(kvb, torch_cuda):
    """The scenarios of tests/test_index_sim.py on the real kernels (cub sort / scan, atomics): add-only batches into a
    full index are applied in parallel with the LRU victims planned up front, reads equal the oracle's at every step."""
    from tests.test_index_sim import _random_traffic
    st = _random_traffic(kvb, kvb.lib, seed=21, size=1500, ppk=3, n_keys=5000, steps=400, max_batch=400, lookup_frac=0.4,
                         evict_frac=0.0)
    assert st["flushes_planned"] > 50 and st["plan_fallbacks"] == 0 and st["lru_evictions"] > 10000, st
    st = _random_traffic(kvb, kvb.lib, seed=22, size=700, ppk=2, n_keys=900, steps=500, max_batch=120, lookup_frac=0.5,
                         evict_frac=0.02)
    assert st["flushes_planned"] > 50, st
    st = _random_traffic(kvb, kvb.lib, seed=5, size=3, ppk=2, n_keys=60, steps=300, max_batch=40)   # index smaller than a batch
    assert st["live_keys"] <= 3
    st = _random_traffic(kvb, kvb.lib, seed=57, size=1967, ppk=3, n_keys=3421, steps=300, max_batch=429, lookup_frac=0.25,
                         evict_frac=0.25)                     # a quarter of the ops remove pods: live moves both ways
    assert st["flushes_planned"] > 30 and st["plan_fallbacks"] == 0, st


def test_planned_eviction_walks_several_windows_of_stale_records(kvb, torch_cuda):
    """200 000 keys at capacity; a lookup re-stamps the 150 000 oldest, so the order array starts with 150 000 stale
    records (more than two planning windows) before the first real victim."""
    K = kvb.kvblock
    n = 200_000
    keys = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    idx, oidx = K.Index(size=n, pod_cache_size=2, expected_keys=n), o.InMemoryIndex(size=n, pod_cache_size=2)
    ent, oent = [K.PodEntry("a", "gpu")], [o.PodEntry("a", "gpu")]
    idx.add(None, keys, ent)
    oidx.add(None, [int(k) for k in keys], oent)
    idx.flush()
    idx.add(None, keys[:1] + np.uint64(1), ent)              # at capacity from here: builds the order array
    oidx.add(None, [int(keys[0] + np.uint64(1))], oent)
    assert len(idx.lookup(keys[1:150_001])) == 150_000       # refresh: their order records are stale now
    oidx.lookup([int(k) for k in keys[1:150_001]])
    new = keys[:20_000] + np.uint64(7)
    idx.add(None, new, ent)
    oidx.add(None, [int(k) for k in new], oent)
    st = idx.stats()
    assert st["flushes_planned"] >= 1 and st["plan_fallbacks"] == 0 and st["lru_evictions"] == 20_001, st
    assert st["order_stale_skipped"] >= 150_000
    probe = np.concatenate([keys[:2], keys[149_990:150_010], keys[169_990:170_020], keys[-5:], new[:5]])
    assert _as_tuples(idx.lookup(probe)) == _as_tuples(oidx.lookup([int(k) for k in probe]))
    assert len(idx) == len(oidx.data) == n


@pytest.mark.parametrize("bs", [4, 8, 16, 5, 32])
def test_fused_scoring_every_block_size_and_shape(kvb, torch_cuda, bs):
    """The fused tokens -> scores launch (block sizes 4 / 8 / 16) and the two-kernel form (any other size) against the oracle:
    ragged prompts incl. empty and shorter-than-a-block ones, unaligned prompt starts (odd token offsets, so the 16 B token
    granules begin before the prompt), a pod filter, multimodal extras on some prompts, one prompt alone, and pinned buffers."""
    K, L = kvb.kvblock, kvb._lib
    rng = np.random.default_rng(100 + bs)
    tp, otp = K.ChunkedTokenDatabase(bs, "seed"), o.TokenProcessor(bs, "seed")
    idx, oidx = K.Index(expected_keys=1 << 12), o.InMemoryIndex()
    pods = ["pod-%d" % i for i in range(9)]
    lens = [0, 1, bs - 1, bs, bs + 1, 3 * bs + 2] + [int(x) for x in rng.integers(0, 40 * bs, 40)]
    prompts = [rng.integers(0, 1 << int(rng.choice([7, 16, 17, 20])), n).astype(np.uint32) for n in lens]
    feats = []
    for i, p in enumerate(prompts):                      # every 5th prompt carries an image over its first blocks
        nb = len(p) // bs
        feats.append([K.BlockExtraFeatures([K.MMHash("img-%d" % i)]) if b < 2 else None for b in range(nb)] if (i % 5 == 0 and nb) else None)
    ofeats = [None if f is None else [None if x is None else o.BlockExtraFeatures([o.MMHash(m.hash) for m in x.mm_hashes]) for x in f]
              for f in feats]
    for i, p in enumerate(prompts):
        keys = otp.tokens_to_kv_block_keys(0, [int(t) for t in p], "m", ofeats[i]) or []
        for _ in range(int(rng.integers(0, 4))):
            d = int(rng.integers(0, len(keys) + 1))
            if d:
                ent = (pods[int(rng.integers(0, 9))], "gpu" if rng.random() < 0.7 else "cpu")
                idx.add(None, keys[:d], [K.PodEntry(*ent)])
                oidx.add(None, keys[:d], [o.PodEntry(*ent)])
    ix, oix = kvb.indexer.Indexer(tp, idx), o.Indexer(otp, oidx)
    for flt in (None, pods[:4], ["nobody"]):
        got = ix.score_tokens_batch(prompts, "m", flt, feats)
        for i, p in enumerate(prompts):
            assert got[i] == oix.score_tokens([int(t) for t in p], "m", flt or (), ofeats[i]), (bs, i, flt)
    for i in (3, 5, 17):                                  # one prompt per call: its offsets travel in the kernel arguments
        assert ix.score_tokens(prompts[i], "m", None, feats[i]) == oix.score_tokens([int(t) for t in prompts[i]], "m", (), ofeats[i])
    # pinned buffers read / written in place, text-only prompts packed back to back (odd starts)
    text = [p for p, f in zip(prompts, feats) if f is None]
    off = np.zeros(len(text) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in text], out=off[1:])
    pin = kvb.pool.PinnedBuffer(int(off[-1]) * 4 + 64)
    tok = pin.numpy(np.uint32)[1:1 + int(off[-1])]       # start 4 bytes into the buffer: never 16 B aligned
    tok[:] = np.concatenate(text) if off[-1] else 0
    parents = np.full(len(text), tp.get_init_hash("m"), dtype=np.uint64)
    n = len(text)
    pout = kvb.pool.PinnedBuffer(n * 136 + 1024)
    raw = pout.numpy(np.uint8)
    b1 = (n * 4 + 255) // 256 * 256
    b2 = b1 + (n * 26 + 255) // 256 * 256
    out = (raw[:n * 4].view(np.int32), raw[b1:b1 + n * 26].view(np.uint16), raw[b2:b2 + n * 104].view(np.float64))
    for flags in (L.SCORE_PINNED_IO, L.SCORE_PINNED_IO | L.SCORE_TWO_KERNELS, L.SCORE_COPY_TOKENS, L.SCORE_NO_TOUCH):
        for a in out:
            a[:] = 0
        idx.score_tokens_flat(bs, tok, off, parents, out=out, flags=flags)
        for i, p in enumerate(text):
            g = {idx.pods.names[int(out[1][i * 13 + j])]: float(out[2][i * 13 + j]) for j in range(int(out[0][i]))}
            want = oix.score_tokens([int(t) for t in p], "m") or {}
            assert g == want, (bs, i, flags)
    del tok
    pin.free()
    pout.free()
