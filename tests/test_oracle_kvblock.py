"""Pins the CPU oracle: the reference's golden vectors, its known-answer tests, and CBOR cross-checks."""
import random

import pytest

from oracle import kvblock_oracle as o
from tests import scenarios as sc


def _mm_features(m):
    ph = {k: [o.PlaceholderRange(r["offset"], r["length"]) for r in v] for k, v in m["mm_placeholders"].items()}
    return o.compute_block_extra_features(m["mm_hashes"], ph, m["block_size"], len(m["tokens"]))


def test_golden_text(golden):
    t = golden["text"]
    tp = o.TokenProcessor(t["block_size"], t["hash_seed"])
    assert tp.tokens_to_kv_block_keys(0, t["tokens"], t["model"]) == t["request_keys"]
    # payload bytes quoted in SURVEY §8c
    init = tp.get_init_hash(t["model"])
    assert o.hash_payload(init, t["tokens"][:4], None).hex() == "831b24c143712f37321e841920051901b61901421949eef6"


def test_golden_multimodal(golden):
    m = golden["multimodal"]
    keys = o.TokenProcessor(m["block_size"], m["hash_seed"]).tokens_to_kv_block_keys(
        0, m["tokens"], m["model"], _mm_features(m))
    assert keys == m["request_keys"] and len(keys) == 104
    feats = _mm_features(m)
    assert feats[0] is None and feats[2] is None          # tokens 0..11 are text
    assert feats[3].mm_hashes == [o.MMHash(m["mm_hashes"]["image"][0])]   # placeholder starts at token 15
    assert feats[101].mm_hashes and feats[102] is None    # placeholder ends at token 406


def test_cbor_matches_cbor2():
    cbor2 = pytest.importorskip("cbor2")
    rnd = random.Random(7)
    for _ in range(2000):
        p = rnd.getrandbits(rnd.choice([3, 5, 8, 16, 32, 33, 64]))
        toks = [rnd.getrandbits(rnd.choice([4, 5, 8, 16, 17, 32])) for _ in range(rnd.choice([1, 4, 16, 23, 24, 255, 256, 300]))]
        ex = rnd.choice([None, "model", "", [{"Hash": "ab" * 32}], {"b": 1, "aa": 2, "a": 3}, 5, -3, ["x", 1]])
        assert o.hash_payload(p, toks, ex) == cbor2.dumps([p, toks, ex], canonical=True)
    assert o.hash_payload(1, None, "m") == cbor2.dumps([1, None, "m"], canonical=True)


def test_fnv_known_answers():
    assert o.fnv64a(b"") == 0xCBF29CE484222325
    assert o.fnv64a(b"a") == 0xAF63DC4C8601EC8C
    assert o.fnv64a(b"foobar") == 0x85944171F73967E8


def test_token_processor_properties():
    """token_processor_test.go:30-272,536-628 (behavioural)."""
    with pytest.raises(ValueError):
        o.TokenProcessor(0)
    with pytest.raises(ValueError):
        o.TokenProcessor(-1)
    tp = o.TokenProcessor(16, "")
    assert tp.get_init_hash("m1") == tp.get_init_hash("m1")
    assert tp.get_init_hash("m1") != tp.get_init_hash("m2")
    assert o.TokenProcessor(16, "seed1").get_init_hash("m") != o.TokenProcessor(16, "seed2").get_init_hash("m")
    toks = list(range(1, 257))
    k16 = o.TokenProcessor(16).tokens_to_kv_block_keys(0, toks, "m")
    k64 = o.TokenProcessor(64).tokens_to_kv_block_keys(0, toks, "m")
    assert len(k16) == 16 and len(k64) == 4 and not set(k16) & set(k64)
    assert o.TokenProcessor(16).tokens_to_kv_block_keys(0, toks[:15], "m") is None      # partial block -> no key
    assert len(o.TokenProcessor(16).tokens_to_kv_block_keys(0, toks[:33], "m")) == 2    # tail dropped
    # parentKey continues a chain
    full = tp.tokens_to_kv_block_keys(0, toks[:64], "m")
    cont = tp.tokens_to_kv_block_keys(full[1], toks[32:64], "m")
    assert cont == full[2:]
    with pytest.raises(ValueError):
        tp.tokens_to_kv_block_keys(0, toks[:32], "m", [None])  # extraFeatures length mismatch
    # extra differentiation (token_processor_test.go:374-440)
    h = lambda e: o.block_hash(1, [1, 2], e)
    assert h(None) != h(0) and h(1) != h(2) and h("a") != h("b") and h("1") != h(1)
    assert h({"a": 1}) != h({"a": 2}) and h({"a": 1}) != h({"b": 1}) and h({"a": 1}) != h(None)
    assert h({"a": 1, "b": 2}) == h({"b": 2, "a": 1})   # canonical key order


def test_extra_keys():
    """extra_keys_test.go behaviours."""
    assert o.parse_raw_extra_keys(None) is None
    r = o.parse_raw_extra_keys([None, ["h1"], [["h2", 5]], [7], []])
    assert r[0] is None and r[1].mm_hashes == [o.MMHash("h1")] and r[2].mm_hashes == [o.MMHash("h2")]
    assert r[3] is None and r[4] is None
    assert o.compute_block_extra_features({}, {}, 16, 64) is None
    f = o.compute_block_extra_features({"image": ["A", "B"]},
                                       {"image": [o.PlaceholderRange(20, 4), o.PlaceholderRange(4, 10)]}, 8, 32)
    assert [None if x is None else [m.hash for m in x.mm_hashes] for x in f] == [["B"], ["B"], ["A"], None]


def _index_with(entries):
    idx = o.InMemoryIndex()
    for k, pods in entries.items():
        idx.add([k], [k], [o.PodEntry(p, t) for p, t in pods])
    return idx


@pytest.mark.parametrize("case", sc.SCORER_CASES, ids=lambda c: c[0])
def test_scorer_known_answers(case):
    _, weights, keys, hit, want = case
    got = o.longest_prefix_score(keys, {k: [o.PodEntry(p, t) for p, t in v] for k, v in hit.items()}, weights)
    assert got == want


@pytest.mark.parametrize("case", sc.INDEXER_CASES, ids=lambda c: c[0])
def test_indexer_known_answers(case):
    _, keys, entries, flt, want = case
    idx = _index_with(entries)
    got = o.longest_prefix_score(keys, idx.lookup(keys, flt))
    assert set(got) == set(want)
    for p in want:
        assert abs(got[p] - want[p]) < 1e-4
    assert o.Indexer(o.TokenProcessor(16)).score_tokens([], "m") is None      # "empty tokens" -> nil


def test_index_contract():
    """index_test.go:119-264,589-735 and in_memory_test.go:45-236 (behavioural)."""
    P = o.PodEntry
    idx = o.InMemoryIndex()
    with pytest.raises(ValueError):
        idx.lookup([])
    with pytest.raises(ValueError):
        idx.add([1], [], [P("p", "gpu")])
    with pytest.raises(ValueError):
        idx.evict(1, o.ENGINE_KEY, [])
    idx.add([1, 2], [11, 12], [P("p1", "gpu"), P("p2", "gpu")])
    assert idx.lookup([11, 12]) == {11: [P("p1", "gpu"), P("p2", "gpu")], 12: [P("p1", "gpu"), P("p2", "gpu")]}
    idx.add([1], [11], [P("p1", "gpu")])                       # duplicate pod: no double entry
    assert len(idx.lookup([11])[11]) == 2
    assert idx.lookup([11, 12], {"p1"}) == {11: [P("p1", "gpu")], 12: [P("p1", "gpu")]}
    assert idx.lookup([11], {"nobody"}) == {}                  # filtered-out key omitted
    assert idx.lookup([999, 11]) .keys() == {11}               # absent key skipped, search continues
    idx.add([3], [13], [P("p3", "gpu"), P("p3", "cpu")])
    idx.evict(3, o.ENGINE_KEY, [P("p3", "cpu")])               # exact tier match only
    # engine mapping 3 was dropped by the evict; re-add and check many:1 / 1:many / last request key
    idx.add([20, 21, 22, 23], [30], [P("p", "gpu")])           # many:1
    assert idx.get_request_key(20) == 30 and idx.get_request_key(23) == 30
    idx.evict(21, o.ENGINE_KEY, [P("p", "gpu")])
    assert idx.lookup([30]) == {}                              # shared request key lost the pod
    idx.add([40], [50, 51, 52, 53], [P("p", "gpu")])           # 1:many
    assert idx.get_request_key(40) == 53
    with pytest.raises(KeyError):
        idx.get_request_key(12345)
    idx.evict(777, o.ENGINE_KEY, [P("p", "gpu")])              # unknown engine key: no-op
    # speculative entries coexist with confirmed ones and are evicted separately
    idx.add(None, [60], [P("p", "gpu", True)])
    idx.add([61], [60], [P("p", "gpu", False)])
    assert idx.lookup([60])[60] == [P("p", "gpu", True), P("p", "gpu", False)]
    idx.evict(60, o.REQUEST_KEY, [P("p", "gpu", True)])
    assert idx.lookup([60])[60] == [P("p", "gpu", False)]
    # LRU caps
    small = o.InMemoryIndex(size=2, pod_cache_size=2)
    for k in (1, 2, 3):
        small.add([k], [k], [P("p", "gpu")])
    assert small.lookup([1, 2, 3]).keys() == {2, 3}
    small.add([2], [2], [P("a", "gpu"), P("b", "gpu"), P("c", "gpu")])
    assert small.lookup([2])[2] == [P("b", "gpu"), P("c", "gpu")]
