"""Multimodal taint — the reference's extra_keys_test.go:127-324 scenarios, run against the oracle (CPU) and the device
hash path (GPU)."""
import pytest

from oracle import kvblock_oracle as ko

TOK16 = list(range(1, 17))
TOK64 = list(range(1, 65))


def _impls(kvb=None):
    """(name, token processor, features class, mmhash class, compute fn, parse fn)"""
    out = [("oracle", ko.TokenProcessor(16, "test"), ko.BlockExtraFeatures, ko.MMHash, ko.compute_block_extra_features,
            ko.parse_raw_extra_keys, ko.PlaceholderRange)]
    if kvb is not None:
        K = kvb.kvblock
        out.append(("gpu", K.ChunkedTokenDatabase(16, "test"), K.BlockExtraFeatures, K.MMHash, K.compute_block_extra_features,
                    K.parse_raw_extra_keys, K.PlaceholderRange))
    return out


def _run(impl):
    name, tp, F, M, compute, parse, PR = impl
    # TestComputeBlockExtraFeatures_NoOverlap / _SingleImage / _TextOnlyBlocksBetweenImages
    assert compute(None, None, 16, 64) is None
    r = compute({"image": ["hash_A"]}, {"image": [PR(0, 48)]}, 16, 64)
    assert len(r) == 4 and [None if x is None else x.mm_hashes[0].hash for x in r] == ["hash_A"] * 3 + [None]
    r = compute({"image": ["hashA", "hashB"]}, {"image": [PR(0, 32), PR(48, 32)]}, 16, 80)
    assert [None if x is None else x.mm_hashes[0].hash for x in r] == ["hashA", "hashA", None, "hashB", "hashB"]
    # TestMMFeatures_DifferentImagesProduceDifferentHashes
    ka = tp.tokens_to_kv_block_keys(0, TOK16, "model", [F([M("image_hash_A")])])
    kb = tp.tokens_to_kv_block_keys(0, TOK16, "model", [F([M("image_hash_B")])])
    assert ka[0] != kb[0]
    # TestMMFeatures_NilFeaturesSameAsTextOnly
    assert tp.tokens_to_kv_block_keys(0, TOK16, "model", None) == tp.tokens_to_kv_block_keys(0, TOK16, "model", [None])
    # TestMMFeatures_OnlyAffectOverlappingBlocks
    text = tp.tokens_to_kv_block_keys(0, TOK64, "model", None)
    img = tp.tokens_to_kv_block_keys(0, TOK64, "model", [None, None, F([M("image_X")]), None])
    assert len(text) == len(img) == 4
    assert text[0] == img[0] and text[1] == img[1] and text[2] != img[2] and text[3] != img[3]
    # TestMMFeatures_MismatchedLengthReturnsError
    with pytest.raises(ValueError, match="does not match token chunk count"):
        tp.tokens_to_kv_block_keys(0, list(range(32)), "model", [None])
    # TestParseAndComputeProduceSameFeatures
    parsed = parse([["img_hash"], ["img_hash"], ["img_hash"], None])
    computed = compute({"image": ["img_hash"]}, {"image": [PR(0, 48)]}, 16, 64)
    assert [None if p is None else [m.hash for m in p.mm_hashes] for p in parsed] == \
           [None if c is None else [m.hash for m in c.mm_hashes] for c in computed]
    return ka, kb, text, img


def test_mm_scenarios_oracle():
    _run(_impls()[0])


@pytest.mark.gpu
def test_mm_scenarios_gpu_match_oracle(kvb, torch_cuda):
    oracle, gpu = _impls(kvb)
    assert _run(gpu) == _run(oracle)          # same scenarios, identical keys on both sides
