"""CPU oracle for the KV-events write path — TEST INFRASTRUCTURE ONLY.

Restates (paths relative to /root/reference/pkg/kvevents):
  engineadapter/common.go:34-130      parseTopic, getHashAsUint64, convertBlockHashes, convertExtraKeys
  engineadapter/vllm_adapter.go:64-260  ParseMessage / decodeVLLMEvent / convertBlockStored|Removed|Cleared
  pool.go:206-249                     realignExtraFeatures
  pool.go:253-398                     processEventBatch
over the oracle index / token processor (oracle/kvblock_oracle.py).  Wire format: msgpack (third-party codec
vmihailenco/msgpack v5 in the reference; the `msgpack` Python package here — the format is the MessagePack spec).
Parity status: no literal golden bytes exist in the reference for this path; behaviour is pinned by transcribing
pool_test.go's assertions (tests/test_kvevents.py)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import msgpack

from . import kvblock_oracle as ko

DEFAULT_TIER = "GPU"  # pool.go:32 (upper-case when the event carries no medium)


@dataclass
class BlockStored:  # events.go:71-79
    block_hashes: list
    tokens: list
    parent_hash: int = 0
    device_tier: str = ""
    lora_id: Optional[int] = None
    lora_name: Optional[str] = None
    extra_keys: Optional[list] = None


@dataclass
class BlockRemoved:  # events.go:87-90
    block_hashes: list
    device_tier: str = ""


@dataclass
class AllBlocksCleared:  # events.go:98-100
    device_tier: str = ""


def parse_topic(topic: str):
    """common.go:34-45: "kv@<pod>@<model>" -> (pod, model); anything else -> (topic, "")."""
    parts = topic.split("@")
    return (parts[1], parts[2]) if len(parts) == 3 else (topic, "")


def hash_as_u64(raw) -> int:
    """common.go:50-71: ints as-is (two's complement for negatives), bytes -> last 8 bytes big-endian."""
    if isinstance(raw, bool):
        raise ValueError("unsupported hash type: bool")
    if isinstance(raw, int):
        return raw & ((1 << 64) - 1)
    if isinstance(raw, (bytes, bytearray)):
        if len(raw) == 0:
            raise ValueError("hash byte slice is empty")
        return int.from_bytes(bytes(raw[-8:]), "big")
    raise ValueError(f"unsupported hash type: {type(raw)}")


def _field(fields, i):
    return fields[i] if i < len(fields) else None


def decode_vllm_event(fields):
    """vllm_adapter.go:105-260 (single decoded []any per event)."""
    if len(fields) < 1:
        raise ValueError("malformed tagged union: no tag")
    tag = fields[0]
    if not isinstance(tag, str):
        raise ValueError(f"event tag is not a string: {type(tag)}")
    if tag == "BlockStored":
        if len(fields) < 5:
            raise ValueError(f"BlockStored: need at least 5 fields, got {len(fields)}")
        if not isinstance(fields[1], (list, tuple)):
            raise ValueError("BlockStored: block_hashes is not an array")
        hashes = [hash_as_u64(h) for h in fields[1]]
        parent = hash_as_u64(fields[2]) if fields[2] is not None else 0
        if not isinstance(fields[3], (list, tuple)):
            raise ValueError("token_ids is not an array")
        tokens = [int(t) & 0xFFFFFFFF for t in fields[3]]
        lora_id = _field(fields, 5)
        medium = _field(fields, 6)
        if medium is not None and not isinstance(medium, str):
            raise ValueError("BlockStored: medium is not a string")
        lora_name = _field(fields, 7)
        if lora_name is not None and not isinstance(lora_name, str):
            raise ValueError("BlockStored: lora_name is not a string")
        extra = _field(fields, 8)
        if extra is not None:
            if not isinstance(extra, (list, tuple)):
                raise ValueError("BlockStored: extra_keys is not an array")
            for i, k in enumerate(extra):
                if k is not None and not isinstance(k, (list, tuple)):
                    raise ValueError(f"extra_keys[{i}] has invalid type")
            extra = [None if k is None else list(k) for k in extra]
        return BlockStored(hashes, tokens, parent, medium or "", lora_id, lora_name, extra)
    if tag == "BlockRemoved":
        if len(fields) < 2:
            raise ValueError(f"BlockRemoved: need at least 2 fields, got {len(fields)}")
        if not isinstance(fields[1], (list, tuple)):
            raise ValueError("BlockRemoved: block_hashes is not an array")
        medium = _field(fields, 2)
        if medium is not None and not isinstance(medium, str):
            raise ValueError("BlockRemoved: medium is not a string")
        return BlockRemoved([hash_as_u64(h) for h in fields[1]], medium or "")
    if tag == "AllBlocksCleared":
        medium = _field(fields, 1)
        return AllBlocksCleared(medium if isinstance(medium, str) else "")
    raise ValueError(f"unknown vLLM event tag: {tag}")


def parse_vllm_message(topic: str, payload: bytes):
    """VLLMAdapter.ParseMessage (vllm_adapter.go:64-88): batch = [ts, [event...], dp_rank?]."""
    pod, model = parse_topic(topic)
    batch = msgpack.unpackb(payload, raw=False, strict_map_key=False)
    if not isinstance(batch, (list, tuple)) or len(batch) < 2:
        raise ValueError("failed to decode vLLM event batch")
    events = [decode_vllm_event(list(e)) for e in batch[1]]
    return pod, model, float(batch[0]), events


def realign_extra_features(engine_features, canonical_count: int):
    """pool.go:206-249."""
    n = len(engine_features)
    if n == canonical_count:
        return engine_features
    if n == 0 or canonical_count == 0:  # Go would panic on the empty slice; nothing can be produced anyway
        return [None] * canonical_count
    out = [None] * canonical_count
    if n < canonical_count:
        for i in range(canonical_count):
            out[i] = engine_features[i * n // canonical_count]
    else:
        for i, ef in enumerate(engine_features):
            ci = i * canonical_count // n
            if ef is None:
                continue
            if out[ci] is None:
                out[ci] = ko.BlockExtraFeatures(None)  # &BlockExtraFeatures{}: MMHashes is a nil slice (pool.go:240-242)
            if ef.mm_hashes:  # append(nil, <no elements>...) stays nil -> CBOR null, not an empty array
                out[ci].mm_hashes = list(out[ci].mm_hashes or []) + list(ef.mm_hashes)
    return out


def process_event_batch(index: ko.InMemoryIndex, tp: ko.TokenProcessor, events, pod: str, model: str):
    """Pool.processEventBatch (pool.go:253-398).  Errors are logged-and-skipped in the reference: `continue`."""
    for ev in events:
        if isinstance(ev, BlockStored):
            tier = ev.device_tier.lower() if ev.device_tier else DEFAULT_TIER
            eff_model = ev.lora_name if ev.lora_name else model
            entries = [ko.PodEntry(pod, tier)]
            parent_rk = 0
            if ev.parent_hash != 0:
                try:
                    parent_rk = index.get_request_key(ev.parent_hash)
                except KeyError:
                    continue
            feats = None
            if ev.extra_keys is not None:
                feats = ko.parse_raw_extra_keys(ev.extra_keys)
            if feats is not None:
                cnt = len(ev.tokens) // tp.block_size()
                if len(feats) != cnt:
                    feats = realign_extra_features(feats, cnt)
            try:
                rks = tp.tokens_to_kv_block_keys(parent_rk, ev.tokens, eff_model, feats)
            except ValueError:
                continue
            if not rks:
                continue
            try:
                index.add(list(ev.block_hashes), rks, entries)
            except (ValueError, ZeroDivisionError):
                continue
        elif isinstance(ev, BlockRemoved):
            tier = ev.device_tier.lower() if ev.device_tier else DEFAULT_TIER
            for h in ev.block_hashes:
                index.evict(h, ko.ENGINE_KEY, [ko.PodEntry(pod, tier)])
        # AllBlocksCleared: log only (pool.go:388-392)
