"""Property tests (hypothesis) of the oracle's invariants — the same invariants the GPU path is held to."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import kvblock_oracle as ko
from oracle import offload_oracle as oo

tokens_st = st.lists(st.integers(min_value=0, max_value=(1 << 32) - 1), min_size=0, max_size=200)


@settings(max_examples=150, deadline=None)
@given(tokens=tokens_st, bs=st.sampled_from([1, 4, 16, 23]), cut=st.integers(min_value=0, max_value=200),
       seed=st.text(max_size=5))
def test_chain_prefix_and_continuation(tokens, bs, cut, seed):
    tp = ko.TokenProcessor(bs, seed)
    full = tp.tokens_to_kv_block_keys(0, tokens, "m") or []
    assert len(full) == len(tokens) // bs
    k = min(cut, len(full))
    # keys of a prefix are a prefix of the keys
    assert (tp.tokens_to_kv_block_keys(0, tokens[:k * bs], "m") or []) == full[:k]
    # continuing from key k-1 over the remaining tokens reproduces the tail
    if 0 < k < len(full):
        assert tp.tokens_to_kv_block_keys(full[k - 1], tokens[k * bs:], "m") == full[k:]
    # a different model or seed changes every key
    other = ko.TokenProcessor(bs, seed + "x").tokens_to_kv_block_keys(0, tokens, "m") or []
    assert all(a != b for a, b in zip(full, other))


@settings(max_examples=100, deadline=None)
@given(n=st.integers(min_value=1, max_value=70), bpf=st.integers(min_value=1, max_value=17))
def test_file_grouping_partitions_the_ids(n, bpf):
    ids = list(range(1000, 1000 + n))
    hashes = list(range(-(-n // bpf)))
    _, groups = oo.build_file_block_mapping(hashes, ids, bpf)
    assert [x for g in groups for x in g] == ids
    assert len(groups[0]) == (n % bpf or bpf) and all(len(g) == bpf for g in groups[1:])


@settings(max_examples=60, deadline=None)
@given(data=st.data())
def test_pack_unpack_roundtrip(data):
    T = data.draw(st.integers(1, 4))
    N = data.draw(st.integers(1, 12))
    frag = data.draw(st.sampled_from([1, 3, 16, 48]))
    rng = np.random.default_rng(data.draw(st.integers(0, 1 << 30)))
    tensors = [rng.integers(0, 256, (N, frag), dtype=np.uint8) for _ in range(T)]
    ids = data.draw(st.lists(st.integers(0, N - 1), min_size=0, max_size=N, unique=True))
    packed = oo.pack_blocks(tensors, ids)
    assert packed.size == len(ids) * T * frag
    dst = [np.zeros_like(t) for t in tensors]
    oo.unpack_blocks(dst, ids, packed)
    for d, t in zip(dst, tensors):
        for b in range(N):
            assert np.array_equal(d[b], t[b] if b in ids else np.zeros(frag, np.uint8))
    bpf = max(len(ids), 1) + data.draw(st.integers(0, 3))
    if ids:
        img = oo.file_image(tensors, ids, bpf)
        dst2 = [np.zeros_like(t) for t in tensors]
        oo.load_from_image(dst2, ids, bpf, img)
        assert all(np.array_equal(a, b) for a, b in zip(dst, dst2))


@settings(max_examples=80, deadline=None)
@given(data=st.data())
def test_scorer_matches_direct_definition(data):
    """score[pod] = sum over the consecutive prefix of the max tier weight."""
    pods = ["a", "b", "c"]
    weights = {"gpu": 1.0, "cpu": 0.8}
    nkeys = data.draw(st.integers(1, 8))
    keys = list(range(1, nkeys + 1))
    k2p = {}
    for k in keys:
        ents = data.draw(st.lists(st.tuples(st.sampled_from(pods), st.sampled_from(["gpu", "cpu", "disk"])), max_size=4))
        if ents:
            k2p[k] = [ko.PodEntry(p, t) for p, t in ents]
    got = ko.longest_prefix_score(keys, k2p, weights)
    want = {}
    for p in pods:
        total, seen = 0.0, False
        for k in keys:
            ws = [weights.get(e.device_tier, 1.0) for e in k2p.get(k, []) if e.pod_identifier == p]
            if not ws:
                break
            total, seen = total + max(ws), True
        if seen:
            want[p] = total
    assert got == want
