"""KV-events write path: adapter + processEventBatch.
CPU: the oracle against the reference's pool_test.go assertions, and the product's adapter / realign against the
oracle.  GPU: EventProcessor (hashing on the device, index through the C ABI) against the oracle on event streams."""
import random

import msgpack
import numpy as np
import pytest

from oracle import kvblock_oracle as ko
from oracle import kvevents_oracle as eo


def _tokens(n):
    return list(range(1, n + 1))            # makeTokens (pool_test.go:38-44)


def _keys(n, base):
    return [base + i for i in range(n)]      # makeEngineKeys (pool_test.go:47-53)


def _pool(bs):
    return ko.InMemoryIndex(), ko.TokenProcessor(bs, "")


# ------------------------------------------------------------------------------------------ oracle vs reference tests
def test_oracle_write_paths():
    """pool_test.go:58-330 (FallbackLegacy, ManyToOne, OneToMany, Eviction_Eager, UnknownEngineKey, PartialBlockDrop)."""
    idx, tp = _pool(16)                                       # 1:1
    eo.process_event_batch(idx, tp, [eo.BlockStored(_keys(4, 500), _tokens(64))], "pod-legacy", "test-model")
    for ek in _keys(4, 500):
        assert idx.get_request_key(ek) != 0
    idx, tp = _pool(64)                                       # many:1 (engine 16 < canonical 64)
    eo.process_event_batch(idx, tp, [eo.BlockStored(_keys(8, 100), _tokens(128))], "pod-a", "test-model")
    ck = tp.tokens_to_kv_block_keys(0, _tokens(128), "test-model")
    assert len(ck) == 2
    for k in ck:
        assert [e.pod_identifier for e in idx.lookup([k])[k]] == ["pod-a"]
    assert idx.get_request_key(100) == ck[0] and idx.get_request_key(104) == ck[1]
    assert idx.lookup([ck[0]])[ck[0]][0].device_tier == "GPU"           # default tier constant is upper-case
    idx, tp = _pool(16)                                       # 1:many (engine 64 > canonical 16)
    eo.process_event_batch(idx, tp, [eo.BlockStored(_keys(2, 200), _tokens(128), device_tier="CPU")], "pod-b", "m")
    ck = tp.tokens_to_kv_block_keys(0, _tokens(128), "m")
    assert len(ck) == 8 and idx.get_request_key(200) == ck[3] and idx.get_request_key(201) == ck[7]
    assert idx.lookup(ck)[ck[0]][0].device_tier == "cpu"                # a given medium is lower-cased
    eo.process_event_batch(idx, tp, [eo.BlockRemoved([200], "CPU")], "pod-b", "m")   # eager eviction of 4 keys
    assert set(idx.lookup(ck)) == set(ck[4:])
    eo.process_event_batch(idx, tp, [eo.BlockRemoved([999])], "pod-b", "m")           # unknown engine key: no-op
    idx, tp = _pool(16)                                       # partial block -> dropped
    eo.process_event_batch(idx, tp, [eo.BlockStored([7], _tokens(10))], "p", "m")
    assert len(idx.data) == 0
    # parent chaining: second event continues from the first one's last engine key
    eo.process_event_batch(idx, tp, [eo.BlockStored(_keys(2, 10), _tokens(32)),
                                     eo.BlockStored(_keys(2, 20), list(range(33, 65)), parent_hash=11)], "p", "m")
    full = tp.tokens_to_kv_block_keys(0, _tokens(64), "m")
    assert idx.get_request_key(21) == full[3]
    eo.process_event_batch(idx, tp, [eo.BlockStored([30], _tokens(16), parent_hash=424242)], "p", "m")  # unknown parent
    with pytest.raises(KeyError):
        idx.get_request_key(30)
    # LoRA name replaces the model name in the chain seed (pool.go:271-274)
    idx2, tp2 = _pool(16)
    eo.process_event_batch(idx2, tp2, [eo.BlockStored([1], _tokens(16), lora_name="adapter-x")], "p", "m")
    assert idx2.get_request_key(1) == tp2.tokens_to_kv_block_keys(0, _tokens(16), "adapter-x")[0]


def test_realign_known_answers(kvb):
    """TestRealignExtraFeatures (pool_test.go:333-396) for the oracle and the product function."""
    for mod, F, M in ((eo, ko.BlockExtraFeatures, ko.MMHash),
                      (kvb.kvevents, kvb.kvblock.BlockExtraFeatures, kvb.kvblock.MMHash)):
        feats = [None] * 4
        assert mod.realign_extra_features(feats, 4) is feats
        assert mod.realign_extra_features([None, None], 8) == [None] * 8
        f0 = F([M("img0")])
        r = mod.realign_extra_features([f0, None], 4)
        assert r[0] is f0 and r[1] is f0 and r[2] is None and r[3] is None
        assert mod.realign_extra_features([None] * 8, 2) == [None, None]
        r = mod.realign_extra_features([F([M("a")]), F([M("b")]), None, F([M("c")])], 2)
        assert [m.hash for m in r[0].mm_hashes] == ["a", "b"] and [m.hash for m in r[1].mm_hashes] == ["c"]


def _rand_payload(rnd, with_extra=True):
    events, plain = [], []
    for _ in range(rnd.randrange(1, 5)):
        kind = rnd.random()
        if kind < 0.6:
            nb = rnd.randrange(1, 4)
            hashes = [rnd.choice([rnd.getrandbits(63), rnd.getrandbits(64), rnd.randbytes(32), rnd.randbytes(5)]) for _ in range(nb)]
            parent = rnd.choice([None, rnd.getrandbits(60), rnd.randbytes(32)])
            toks = [rnd.randrange(0, 200000) for _ in range(16 * nb + rnd.randrange(0, 3))]
            ev = ["BlockStored", hashes, parent, toks, 16]
            tail = [rnd.choice([None, 3]), rnd.choice([None, "GPU", "cpu"]), rnd.choice([None, "lora-a"]),
                    rnd.choice([None, [None, ["mm1"], [["mm2", 4]]][:nb]]) if with_extra else None]
            ev += tail[: rnd.randrange(0, 5)]
            events.append(ev)
        elif kind < 0.9:
            ev = ["BlockRemoved", [rnd.getrandbits(64), rnd.randbytes(32)]]
            if rnd.random() < 0.5:
                ev.append(rnd.choice([None, "CPU"]))
            events.append(ev)
        else:
            events.append(["AllBlocksCleared"])
    batch = [rnd.random() * 1e9, events]
    if rnd.random() < 0.5:
        batch.append(rnd.choice([None, 0, 3]))
    return msgpack.packb(batch, use_bin_type=True)


def test_adapter_matches_oracle(kvb):
    rnd = random.Random(1)
    ad = kvb.kvevents.VLLMAdapter()
    for _ in range(300):
        payload = _rand_payload(rnd)
        topic = rnd.choice(["kv@10.0.0.%d@meta-llama/Llama-3-8B" % rnd.randrange(9), "weird-topic"])
        pod, model, batch = ad.parse_message(topic, payload)
        opod, omodel, ots, oevents = eo.parse_vllm_message(topic, payload)
        assert (pod, model, batch.timestamp) == (opod, omodel, ots) and len(batch.events) == len(oevents)
        for a, b in zip(batch.events, oevents):
            assert type(a).__name__.replace("Event", "") == type(b).__name__
            assert {k: v for k, v in vars(a).items()} == {k: v for k, v in vars(b).items()}
    assert ad.parse_topic("kv@pod@model") == ("pod", "model") and ad.parse_topic("x@y") == ("x@y", "")
    assert ad.sharding_key("kv@pod@model") == "pod"
    for bad in (msgpack.packb([1.0, [["Nope"]]]), msgpack.packb([1.0, [[5]]]), msgpack.packb([1.0, [["BlockStored", [1]]]]),
                msgpack.packb([1.0, [["BlockStored", 5, None, [1], 16]]]), msgpack.packb("str"), b"\xc1"):
        with pytest.raises(ValueError):
            ad.parse_message("kv@p@m", bad)
        with pytest.raises(Exception):
            eo.parse_vllm_message("kv@p@m", bad)


# ------------------------------------------------------------------------------------------ GPU parity
def _stream(rnd, n_events, key_base, with_parent=True, with_extra=True):
    """A plausible per-pod stream: stores (sometimes chained to an earlier engine key), removals of earlier keys."""
    events, oevents, known = [], [], []
    next_key = key_base
    for _ in range(n_events):
        r = rnd.random()
        if r < 0.7 or not known:
            eng_bs = rnd.choice([8, 16, 32])
            nb = rnd.randrange(1, 5)
            toks = [rnd.randrange(0, 128256) for _ in range(eng_bs * nb + rnd.choice([0, 0, 3]))]
            hashes = list(range(next_key, next_key + nb))
            next_key += nb
            parent = rnd.choice(known) if (with_parent and known and rnd.random() < 0.5) else 0
            if rnd.random() < 0.1:
                parent = 987654321              # unknown parent -> skipped
            tier = rnd.choice(["", "GPU", "CPU", "disk"])
            lora = rnd.choice([None, None, "lora-1"])
            extra = None
            if with_extra and rnd.random() < 0.3:
                extra = [rnd.choice([None, ["img-%d" % rnd.randrange(4)], [["aud-%d" % rnd.randrange(3), 7]], []])
                         for _ in range(nb)]
            kw = dict(block_hashes=hashes, tokens=toks, parent_hash=parent, device_tier=tier, lora_name=lora, extra_keys=extra)
            events.append(("S", kw))
            known.extend(hashes)
        elif r < 0.95:
            hs = [rnd.choice(known) for _ in range(rnd.randrange(1, 3))] + ([555] if rnd.random() < 0.2 else [])
            events.append(("R", dict(block_hashes=hs, device_tier=rnd.choice(["", "GPU", "cpu"]))))
        else:
            events.append(("C", {}))
    return events


def _mk(kind, kw, E, oracle):
    if kind == "S":
        return (eo.BlockStored if oracle else E.BlockStoredEvent)(**kw)
    if kind == "R":
        return (eo.BlockRemoved if oracle else E.BlockRemovedEvent)(**kw)
    return (eo.AllBlocksCleared if oracle else E.AllBlocksClearedEvent)()


def _assert_same_index(kvb, idx, oidx):
    assert len(idx) == len(oidx.data)
    for rk, pods in oidx.data.d.items():
        got = idx.host_peek(rk)
        want = [(p.pod_identifier, p.device_tier, p.speculative) for p in pods.keys()]
        assert [(g.pod_identifier, g.device_tier, g.speculative) for g in got] == want


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [16, 64])
def test_event_processor_matches_oracle(kvb, torch_cuda, bs):
    E, K = kvb.kvevents, kvb.kvblock
    rnd = random.Random(bs)
    idx, tp = K.Index(), K.ChunkedTokenDatabase(bs, "")
    oidx, otp = ko.InMemoryIndex(), ko.TokenProcessor(bs, "")
    proc = E.EventProcessor(idx, tp)
    for pod in range(6):
        spec = _stream(rnd, 40, 1000 * (pod % 3))          # pods 0/3, 1/4, 2/5 share engine-key ranges (same content)
        proc.process_event_batch([_mk(k, kw, E, False) for k, kw in spec], "pod-%d" % pod, "base-model")
        eo.process_event_batch(oidx, otp, [_mk(k, kw, E, True) for k, kw in spec], "pod-%d" % pod, "base-model")
    _assert_same_index(kvb, idx, oidx)
    keys = list(oidx.data.d.keys())[:200]
    assert {k: [(e.pod_identifier, e.device_tier) for e in v] for k, v in idx.lookup(keys).items()} == \
           {k: [(e.pod_identifier, e.device_tier) for e in v] for k, v in oidx.lookup(keys).items()}
    for ek in range(0, 3000, 37):
        try:
            want = oidx.get_request_key(ek)
        except KeyError:
            with pytest.raises(KeyError):
                idx.get_request_key(ek)
        else:
            assert idx.get_request_key(ek) == want


@pytest.mark.gpu
def test_process_many_equals_per_pod_processing(kvb, torch_cuda):
    """Batched rounds (one device hash call per round for all pods) give the same index as pod-by-pod processing
    when pods do not share engine keys (the only cross-pod coupling of the write path)."""
    E, K = kvb.kvevents, kvb.kvblock
    rnd = random.Random(9)
    specs = [("pod-%d" % p, "m", _stream(rnd, 30, 100000 * (p + 1))) for p in range(12)]
    idx_a, idx_b, tp = K.Index(), K.Index(), K.ChunkedTokenDatabase(16, "")
    oidx, otp = ko.InMemoryIndex(), ko.TokenProcessor(16, "")
    pa, pb = E.EventProcessor(idx_a, tp), E.EventProcessor(idx_b, tp)
    launches0 = kvb.lib.kvb_launch_count()
    pa.process_many([(pod, m, [_mk(k, kw, E, False) for k, kw in spec]) for pod, m, spec in specs])
    launches_many = kvb.lib.kvb_launch_count() - launches0
    launches0 = kvb.lib.kvb_launch_count()
    for pod, m, spec in specs:
        pb.process_event_batch([_mk(k, kw, E, False) for k, kw in spec], pod, m)
        eo.process_event_batch(oidx, otp, [_mk(k, kw, E, True) for k, kw in spec], pod, m)
    launches_seq = kvb.lib.kvb_launch_count() - launches0
    _assert_same_index(kvb, idx_a, oidx)
    _assert_same_index(kvb, idx_b, oidx)
    assert launches_many * 4 < launches_seq          # ~12x fewer hash launches: one per round instead of one per event


@pytest.mark.gpu
def test_native_ingest_equals_the_oracle(kvb, torch_cuda):
    """kvb_index_ingest_events: a decoded batch of many pods applied inside the library (parents through the engine-key map,
    one hash launch per round, BlockRemoved by engine key, unknown parents and short events skipped) gives the same index
    as the oracle's processEventBatch pod by pod — text-only streams; a batch with multimodal extras falls back."""
    E, K = kvb.kvevents, kvb.kvblock
    rnd = random.Random(31)
    specs = [("pod-%d" % p, "m", _stream(rnd, 40, 100000 * (p + 1), with_extra=False)) for p in range(16)]
    idx, tp = K.Index(), K.ChunkedTokenDatabase(16, "")
    oidx, otp = ko.InMemoryIndex(), ko.TokenProcessor(16, "")
    proc = E.EventProcessor(idx, tp)
    launches0 = kvb.lib.kvb_launch_count()
    skipped = proc.process_many_native([(pod, m, [_mk(k, kw, E, False) for k, kw in spec]) for pod, m, spec in specs])
    launches = kvb.lib.kvb_launch_count() - launches0
    for pod, m, spec in specs:
        eo.process_event_batch(oidx, otp, [_mk(k, kw, E, True) for k, kw in spec], pod, m)
    _assert_same_index(kvb, idx, oidx)
    assert skipped > 0                                   # the streams contain unknown parents
    assert launches < 40 * 8                             # hash launches: one per ROUND (40), plus the index's own kernels
    for ek in range(100000, 100400, 7):
        try:
            want = oidx.get_request_key(ek)
        except KeyError:
            with pytest.raises(KeyError):
                idx.get_request_key(ek)
        else:
            assert idx.get_request_key(ek) == want
    # a batch with extras takes the Python path and still matches
    specs2 = [("pod-x%d" % p, "m", _stream(rnd, 20, 900000 + 10000 * p)) for p in range(3)]
    proc.process_many_native([(pod, m, [_mk(k, kw, E, False) for k, kw in spec]) for pod, m, spec in specs2])
    for pod, m, spec in specs2:
        eo.process_event_batch(oidx, otp, [_mk(k, kw, E, True) for k, kw in spec], pod, m)
    _assert_same_index(kvb, idx, oidx)


@pytest.mark.gpu
def test_raw_messages_end_to_end(kvb, torch_cuda):
    E, K = kvb.kvevents, kvb.kvblock
    idx, tp = K.Index(), K.ChunkedTokenDatabase(16, "")
    proc, ad = E.EventProcessor(idx, tp), E.VLLMAdapter()
    toks = list(range(1, 49))
    payload = msgpack.packb([1.5, [["BlockStored", [b"\x00" * 24 + (77).to_bytes(8, "big"), 78, 79], None, toks, 16, None, "GPU"]]],
                            use_bin_type=True)
    proc.process_raw_message(ad, "kv@10.1.2.3:8000@meta-llama/Llama-3-8B", payload)
    keys = tp.tokens_to_kv_block_keys(0, toks, "meta-llama/Llama-3-8B")
    assert idx.get_request_key(77) == keys[0] and idx.get_request_key(79) == keys[2]
    assert idx.lookup(keys)[keys[1]] == [K.PodEntry("10.1.2.3:8000", "gpu")]
    proc.process_raw_message(ad, "kv@10.1.2.3:8000@meta-llama/Llama-3-8B", msgpack.packb([2.0, [["BlockRemoved", [78], "GPU"]]]))
    assert set(idx.lookup(keys)) == {keys[0], keys[2]}
    proc.process_raw_message(ad, "kv@p@m", b"garbage")          # logged and dropped
    assert proc.skipped == 1


def test_adapter_known_answers(kvb):
    """common_test.go:28-91 and vllm_adapter_test.go:29-440 (values transcribed), for the product adapter and the oracle."""
    E = kvb.kvevents
    ad = E.VLLMAdapter()
    for parse in (ad.parse_topic, eo.parse_topic):
        assert parse("kv@pod-123@llama-2-7b") == ("pod-123", "llama-2-7b")
        assert parse("pod-123@llama-2-7b") == ("pod-123@llama-2-7b", "")
        assert parse("fallback") == ("fallback", "")
    for h in (E._hash_u64, eo.hash_as_u64):
        assert h(42) == 42 and h((12345).to_bytes(8, "big")) == 12345
        assert h(b"\x01" * 24 + (7).to_bytes(8, "big")) == 7 and h(b"\x01\x02") == 0x0102   # last 8 bytes / left-padded
        assert h(-1) == (1 << 64) - 1                                                          # int64 reinterpretation
        for bad in (b"", "not a hash"):
            with pytest.raises(ValueError):
                h(bad)
    stored = ["BlockStored", [100, 101], 99, [1, 2, 3], 16, None, "gpu", None, None]
    pod, model, batch = ad.parse_message("kv@pod-1@llama-2-7b", msgpack.packb([1234567890.0, [stored], None]))
    assert (pod, model, batch.timestamp, len(batch.events)) == ("pod-1", "llama-2-7b", 1234567890.0, 1)
    ev = batch.events[0]
    assert ev.block_hashes == [100, 101] and ev.parent_hash == 99 and ev.tokens == [1, 2, 3] and ev.device_tier == "gpu"
    # forward compatibility: unknown trailing fields are ignored, present ones parsed
    ev = ad.decode_event(["BlockStored", [400, 401], 399, [10, 11, 12], 16, None, "gpu", "my-lora", [["extra", "keys"]],
                          "completely-unknown-field"])
    assert (ev.lora_id, ev.lora_name, ev.extra_keys) == (None, "my-lora", [["extra", "keys"]])
    # backward compatibility: missing trailing fields
    for fields, want in ((["BlockStored", [1], None, [5], 16, 7, "cpu"], (7, "cpu", None)),
                         (["BlockStored", [1], None, [5], 16, 7], (7, "", None)),
                         (["BlockStored", [1], None, [5], 16], (None, "", None))):
        ev = ad.decode_event(fields)
        assert (ev.lora_id, ev.device_tier, ev.lora_name) == want and ev.parent_hash == 0 and ev.extra_keys is None
    rm = ad.decode_event(["BlockRemoved", [5, 6], "cpu", "future-field"])
    assert rm.block_hashes == [5, 6] and rm.device_tier == "cpu"
    assert ad.decode_event(["BlockRemoved", [5]]).device_tier == ""
    assert isinstance(ad.decode_event(["AllBlocksCleared"]), E.AllBlocksClearedEvent)
    for bad in (["BlockStored", [1], None, [5], 16, None, None, None, ["not-a-list"]],   # invalid extra_keys entry
                ["Unknown"], [], [17]):
        with pytest.raises(ValueError):
            ad.decode_event(bad)
    for payload in (b"", b"\x93\x01"):                                                    # empty / malformed
        with pytest.raises(ValueError):
            ad.parse_message("kv@p@m", payload)
    assert isinstance(E.SGLangAdapter().decode_event(["BlockStored", [1], None, [5], 16]), E.BlockStoredEvent)
