"""CPU oracle for the kvblock read path: block-key hashing, index, scorer, Indexer.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line (relative to /root/reference) whose behaviour it restates.

Third-party arithmetic restated here (absent from /root/reference, named in go.mod):
  * github.com/fxamacker/cbor/v2 v2.7.0 (go.mod:11), ``CanonicalEncOptions`` — RFC 8949
    §4.2.1 core deterministic encoding: shortest-form integer heads, definite
    lengths, map keys sorted bytewise by their encoding (length-first for text).
    Go specifics: nil slice -> null (0xf6); struct -> map keyed by field name.
  * Go stdlib hash/fnv New64a: offset 0xcbf29ce484222325, prime 0x100000001b3.
  * github.com/hashicorp/golang-lru/v2 v2.0.7 (go.mod:15): Add/Get move to front,
    Keys() oldest->newest, ContainsOrAdd/Peek/Contains/Len/Keys do not touch recency.
Parity is PINNED by the reference's golden vectors (tests/golden/kvblock_golden.json).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

MASK64 = (1 << 64) - 1
FNV64_OFFSET = 0xCBF29CE484222325
FNV64_PRIME = 0x100000001B3


# --------------------------------------------------------------------------- FNV / CBOR
def fnv64a(data: bytes, h: int = FNV64_OFFSET) -> int:
    """Go hash/fnv New64a().Write(data).Sum64()  (token_processor.go:92-94,132-134)."""
    for b in data:
        h = ((h ^ b) * FNV64_PRIME) & MASK64
    return h


def _head(major: int, n: int) -> bytes:
    """CBOR initial byte + shortest-form argument (RFC 8949 §3, §4.2.1)."""
    m = major << 5
    if n < 24:
        return bytes([m | n])
    if n < 1 << 8:
        return bytes([m | 24, n])
    if n < 1 << 16:
        return bytes([m | 25]) + n.to_bytes(2, "big")
    if n < 1 << 32:
        return bytes([m | 26]) + n.to_bytes(4, "big")
    if n < 1 << 64:
        return bytes([m | 27]) + n.to_bytes(8, "big")
    raise OverflowError("CBOR integer out of range")


def cbor_canonical(obj) -> bytes:
    """Canonical CBOR of the value kinds ``hash()`` can be handed (token_processor.go:113-126)."""
    if obj is None:
        return b"\xf6"
    if obj is True:
        return b"\xf5"
    if obj is False:
        return b"\xf4"
    if isinstance(obj, int):
        return _head(0, obj) if obj >= 0 else _head(1, -1 - obj)
    if isinstance(obj, str):
        raw = obj.encode("utf-8")
        return _head(3, len(raw)) + raw
    if isinstance(obj, (bytes, bytearray)):
        return _head(2, len(obj)) + bytes(obj)
    if isinstance(obj, MMHash):  # struct MMHash{Hash string} -> {"Hash": s}  (extra_keys.go:26-28)
        return cbor_canonical({"Hash": obj.hash})
    if isinstance(obj, dict):
        items = sorted((cbor_canonical(k), cbor_canonical(v)) for k, v in obj.items())
        return _head(5, len(items)) + b"".join(k + v for k, v in items)
    if isinstance(obj, (list, tuple)):
        return _head(4, len(obj)) + b"".join(cbor_canonical(x) for x in obj)
    try:  # numpy integers
        return cbor_canonical(int(obj))
    except Exception as e:  # pragma: no cover
        raise TypeError(f"unsupported CBOR type {type(obj)}") from e


def hash_payload(parent: int, tokens: Optional[Sequence[int]], extra) -> bytes:
    """Bytes that get FNV-folded: CBOR of the 3-array [parent, tokens, extra] (token_processor.go:124-126)."""
    toks = None if tokens is None else [int(t) for t in tokens]
    return b"\x83" + cbor_canonical(int(parent)) + cbor_canonical(toks) + cbor_canonical(extra)


def block_hash(parent: int, tokens: Optional[Sequence[int]], extra) -> int:
    """chunkedTokenDatabase.hash (token_processor.go:123-135)."""
    return fnv64a(hash_payload(parent, tokens, extra))


# --------------------------------------------------------------------------- extra features
@dataclass(frozen=True)
class MMHash:
    """extra_keys.go:26-28."""

    hash: str


@dataclass
class BlockExtraFeatures:
    """extra_keys.go:32-34; ``None`` in a list of these means a pure-text block."""

    mm_hashes: Optional[list] = None  # list[MMHash]


@dataclass(frozen=True)
class PlaceholderRange:
    """extra_keys.go:38-41."""

    offset: int
    length: int


def parse_raw_extra_keys(raw):
    """ParseRawExtraKeys (extra_keys.go:49-85)."""
    if raw is None:
        return None
    out = [None] * len(raw)
    for i, keys in enumerate(raw):
        if keys is None:
            continue
        hs = []
        for e in keys:
            if isinstance(e, str):
                hs.append(MMHash(e))
            elif isinstance(e, (list, tuple)):
                if len(e) >= 1 and isinstance(e[0], str):
                    hs.append(MMHash(e[0]))
        if hs:
            out[i] = BlockExtraFeatures(hs)
    return out


def compute_block_extra_features(mm_hashes, mm_placeholders, block_size: int, num_tokens: int):
    """ComputeBlockExtraFeatures (extra_keys.go:100-163)."""
    if not mm_hashes or block_size <= 0 or num_tokens <= 0:
        return None
    items = []
    for modality, hashes in mm_hashes.items():
        ranges = (mm_placeholders or {}).get(modality)
        if ranges is None:
            continue
        for h, r in zip(hashes, ranges):  # n = min(len(hashes), len(ranges))
            items.append((r.offset, r.offset + r.length, h))
    if not items:
        return None
    items.sort(key=lambda it: it[0])  # sort.Slice by start (not stable in Go; ties only reorder equal starts)
    nblocks = num_tokens // block_size
    res = [None] * nblocks
    for b in range(nblocks):
        bs, be = b * block_size, (b + 1) * block_size
        hs = []
        for s, e, h in items:
            if e <= bs:
                continue
            if s >= be:
                break
            hs.append(MMHash(h))
        if hs:
            res[b] = BlockExtraFeatures(hs)
    return res


# --------------------------------------------------------------------------- token processor
class TokenProcessor:
    """chunkedTokenDatabase (token_processor.go:72-205)."""

    def __init__(self, block_size: int = 16, hash_seed: str = ""):
        if block_size <= 0:  # token_processor.go:86-88
            raise ValueError(f"blockSize must be greater than 0, got {block_size}")
        self._block_size = block_size
        self.hash_seed = hash_seed
        self.init_hash = fnv64a(hash_seed.encode("utf-8"))  # :90-95

    def block_size(self) -> int:
        return self._block_size

    def get_init_hash(self, model_name: str) -> int:
        """getInitHash (:109-111): tokens nil -> CBOR null."""
        return block_hash(self.init_hash, None, model_name)

    def tokens_to_kv_block_keys(self, parent_key: int, tokens: Sequence[int], model_name: str,
                                extra_features=None) -> Optional[list]:
        """TokensToKVBlockKeys (:177-205); returns None for "nil, nil" (no full block)."""
        parent = parent_key if parent_key != 0 else self.get_init_hash(model_name)
        bs = self._block_size
        nchunks = len(tokens) // bs  # chunkTokens drops the tail (:161-174)
        if nchunks == 0:
            return None
        if extra_features is None:
            extra_features = [None] * nchunks
        elif len(extra_features) != nchunks:
            raise ValueError(
                f"extraFeatures length {len(extra_features)} does not match token chunk count {nchunks} "
                f"(blockSize={bs}, tokens={len(tokens)})")
        keys = []
        for i in range(nchunks):  # prefixHashes (:139-153)
            ef = extra_features[i]
            extra = None if ef is None else ef.mm_hashes
            parent = block_hash(parent, tokens[i * bs:(i + 1) * bs], extra)
            keys.append(parent)
        return keys


def encode_extra_suffix(ef: Optional[BlockExtraFeatures]) -> bytes:
    """The trailing CBOR item X(extra) of one block's payload."""
    return cbor_canonical(None if ef is None else ef.mm_hashes)


# --------------------------------------------------------------------------- LRU (golang-lru v2 semantics)
class LRU:
    """hashicorp/golang-lru/v2 Cache subset: newest at the end of the OrderedDict."""

    def __init__(self, size: int):
        if size <= 0:
            raise ValueError("must provide a positive size")
        self.size = size
        self.d: OrderedDict = OrderedDict()

    def add(self, k, v) -> bool:
        if k in self.d:
            self.d[k] = v
            self.d.move_to_end(k)
            return False
        self.d[k] = v
        if len(self.d) > self.size:
            self.d.popitem(last=False)
            return True
        return False

    def get(self, k):
        if k in self.d:
            self.d.move_to_end(k)
            return self.d[k], True
        return None, False

    def contains_or_add(self, k, v):
        if k in self.d:
            return True, False
        return False, self.add(k, v)

    def remove(self, k) -> bool:
        if k in self.d:
            del self.d[k]
            return True
        return False

    def keys(self) -> list:
        return list(self.d.keys())  # oldest -> newest

    def __len__(self):
        return len(self.d)


# --------------------------------------------------------------------------- index
@dataclass(frozen=True)
class PodEntry:
    """index.go:176-183."""

    pod_identifier: str
    device_tier: str
    speculative: bool = False


ENGINE_KEY = 0  # index.go:155-161
REQUEST_KEY = 1


class InMemoryIndex:
    """InMemoryIndex (in_memory.go:57-304).  Single-threaded restatement."""

    def __init__(self, size: int = int(1e8), pod_cache_size: int = 10):
        self.data = LRU(size)  # requestKey -> LRU[PodEntry]
        self.engine_to_request = LRU(size)  # engineKey -> [requestKey]
        self.pod_cache_size = pod_cache_size
        if pod_cache_size <= 0:
            # lru.New fails lazily inside Add (in_memory.go:190-193); surface it at Add time too
            pass

    def lookup(self, request_keys: Sequence[int], pod_filter: Iterable[str] = ()) -> dict:
        """Lookup (in_memory.go:107-148)."""
        if len(request_keys) == 0:
            raise ValueError("no requestKeys provided for lookup")
        flt = set(pod_filter or ())
        out: dict = {}
        for rk in request_keys:
            pods, found = self.data.get(rk)
            if not found:
                continue  # absent: keep going (:143-145)
            if pods is None or len(pods) == 0:
                return out  # present but empty: cut (:120-124)
            if not flt:
                out[rk] = pods.keys()
            else:
                for p in pods.keys():
                    if p.pod_identifier in flt:
                        out.setdefault(rk, []).append(p)
        return out

    def add(self, engine_keys: Optional[Sequence[int]], request_keys: Sequence[int], entries: Sequence[PodEntry]):
        """Add (in_memory.go:154-224)."""
        if len(request_keys) == 0 or len(entries) == 0:
            raise ValueError("no keys or entries provided for adding to index")
        if engine_keys is not None:
            # Go: an empty non-nil slice would divide by zero; callers never do that.
            new_map: dict = {}
            n = max(len(engine_keys), len(request_keys))
            for i in range(n):
                ek = engine_keys[i * len(engine_keys) // n]
                rk = request_keys[i * len(request_keys) // n]
                new_map.setdefault(ek, []).append(rk)
            for ek, rks in new_map.items():  # Go iterates the map in random order; we use first-seen order
                self.engine_to_request.add(ek, rks)
        for rk in request_keys:
            pods, found = self.data.get(rk)
            if not found:
                pods = LRU(self.pod_cache_size)
                self.data.contains_or_add(rk, pods)
            for e in entries:
                pods.add(e, None)

    def evict(self, key: int, key_type: int, entries: Sequence[PodEntry]):
        """Evict (in_memory.go:229-255)."""
        if len(entries) == 0:
            raise ValueError("no entries provided for eviction from index")
        if key_type == ENGINE_KEY:
            rks, found = self.engine_to_request.get(key)
            if not found:
                return
            for rk in rks:
                self._evict_from_request_key(rk, entries)
            self.engine_to_request.remove(key)
        elif key_type == REQUEST_KEY:
            self._evict_from_request_key(key, entries)
        else:
            raise ValueError(f"unknown key type: {key_type}")

    def _evict_from_request_key(self, rk: int, entries):
        """evictPodsFromRequestKey (in_memory.go:259-293)."""
        pods, found = self.data.get(rk)
        if not found or pods is None:
            return
        for e in entries:
            pods.remove(e)
        if len(pods) != 0:
            return
        cur, still = self.data.get(rk)
        if still and cur is not None and len(cur) == 0:
            self.data.remove(rk)

    def get_request_key(self, engine_key: int) -> int:
        """GetRequestKey (in_memory.go:298-304): LAST mapped request key."""
        rks, found = self.engine_to_request.get(engine_key)
        if not found or len(rks) == 0:
            raise KeyError(f"engine key not found: {engine_key}")
        return rks[-1]


# --------------------------------------------------------------------------- scorer
DEFAULT_MEDIUM_WEIGHTS = {"gpu": 1.0, "cpu": 0.8}  # backend.go:26-31


def _fill_max_weights(entries, weights) -> dict:
    """fillMaxWeights (kvblock_scorer.go:91-103)."""
    dst: dict = {}
    for e in entries:
        w = 1.0
        if weights is not None and e.device_tier in weights:
            w = weights[e.device_tier]
        if e.pod_identifier not in dst or w > dst[e.pod_identifier]:
            dst[e.pod_identifier] = w
    return dst


def longest_prefix_score(keys: Sequence[int], key_to_pods: dict, weights=DEFAULT_MEDIUM_WEIGHTS) -> dict:
    """LongestPrefixScorer.Score (kvblock_scorer.go:106-154); float64 adds in key order."""
    if len(keys) == 0:
        return {}
    cur = _fill_max_weights(key_to_pods.get(keys[0], ()), weights)
    scores = dict(cur)
    active = set(cur)
    for i in range(1, len(keys)):
        if not active:
            break
        cur = _fill_max_weights(key_to_pods.get(keys[i], ()), weights)
        for pod in list(active):
            if pod in cur:
                scores[pod] = scores[pod] + cur[pod]
            else:
                active.discard(pod)
    return scores


# --------------------------------------------------------------------------- indexer
class Indexer:
    """kvcache.Indexer.ScoreTokens (indexer.go:239-304) over the in-memory index."""

    def __init__(self, token_processor: TokenProcessor, index: Optional[InMemoryIndex] = None,
                 weights=DEFAULT_MEDIUM_WEIGHTS):
        self.token_processor = token_processor
        self.index = index if index is not None else InMemoryIndex()
        self.weights = weights

    def score_tokens(self, tokens, model_name: str, pod_identifiers=(), extra_features=None):
        keys = self.token_processor.tokens_to_kv_block_keys(0, tokens, model_name, extra_features)
        if not keys:
            return None  # "nil, nil" (indexer.go:266-270)
        key_to_pods = self.index.lookup(keys, set(pod_identifiers or ()))
        return longest_prefix_score(keys, key_to_pods, self.weights)
