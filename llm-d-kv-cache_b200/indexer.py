"""kvcache.Indexer / LongestPrefixScorer — mirror of pkg/kvcache/{indexer.go,kvblock_scorer.go,backend.go}."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .kvblock import ChunkedTokenDatabase, Index, PodEntry

LONGEST_PREFIX_MATCH = "LongestPrefix"  # kvblock_scorer.go:29-32


@dataclass
class KVCacheBackendConfig:  # backend.go:19-24
    name: str
    weight: float


def default_kv_cache_backend_config():  # backend.go:26-31
    return [KVCacheBackendConfig("gpu", 1.0), KVCacheBackendConfig("cpu", 0.8)]


class LongestPrefixScorer:
    """KVBlockScorer (kvblock_scorer.go:50-57,75-154).  The reference scorer consumes the map Lookup
    returned; here the same prefix walk runs on the GPU against the index the map came from, so
    ``score`` takes the keys and the Index."""

    def __init__(self, index: Index, medium_weights: Optional[dict] = None):
        self.index = index
        if medium_weights is not None:
            index.set_medium_weights(medium_weights)

    def strategy(self) -> str:
        return LONGEST_PREFIX_MATCH

    def score(self, keys: Sequence[int], pod_identifiers=None, touch_lru: bool = True) -> dict:
        if len(keys) == 0:
            return {}  # kvblock_scorer.go:111-113
        k = np.asarray(keys, dtype=np.uint64)
        return self.index.score_keys_batch(k, np.array([0, k.size], dtype=np.int64), pod_identifiers, touch_lru)[0]


class Indexer:
    """kvcache.Indexer (indexer.go:65-304): tokens -> block keys -> Lookup -> Score."""

    def __init__(self, token_processor: ChunkedTokenDatabase, index: Optional[Index] = None,
                 backend_configs=None, device: int = 0):
        if token_processor is None:
            raise ValueError("tokenProcessor cannot be nil")  # indexer.go:82-84
        cfgs = default_kv_cache_backend_config() if backend_configs is None else backend_configs
        weights = {c.name: c.weight for c in cfgs}
        self.token_processor = token_processor
        self.index = index if index is not None else Index(device=device)
        self.scorer = LongestPrefixScorer(self.index, weights)

    def kv_block_index(self) -> Index:
        return self.index

    def compute_block_keys_from_tokens(self, tokens, model_name: str, extra_features=None):
        return self.token_processor.tokens_to_kv_block_keys(0, tokens, model_name, extra_features)

    def score_tokens(self, tokens, model_name: str, pod_identifiers=None, extra_features=None):
        """ScoreTokens (indexer.go:239-304): None when the prompt has no full block ("nil, nil")."""
        if len(tokens) // self.token_processor.block_size() == 0:
            return None
        res, _ = self.index.score_tokens_batch(
            self.token_processor, [tokens], [model_name], pod_identifiers,
            None if extra_features is None else [extra_features], touch_lru=True)
        return res[0]

    def score_tokens_batch(self, prompts, model_names, pod_identifiers=None, extra_features=None):
        """Data-parallel ScoreTokens: one result per prompt (None for prompts without a full block)."""
        res, nblk = self.index.score_tokens_batch(self.token_processor, prompts, model_names, pod_identifiers,
                                                  extra_features)
        return [r if int(nblk[i]) > 0 else None for i, r in enumerate(res)]
