#!/usr/bin/env python
"""Extract the reference's golden vectors for the kvblock hash path into JSON fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Source of every value (reference @ 82d31d1):
  tests/e2e/uds_tokenizer/uds_e2e_test.go:337-348         goldenTokenIDs / goldenRequestKeys
  tests/e2e/uds_tokenizer/uds_e2e_mm_test.go:386-458      goldenMM{TokenIDs,Hashes,Placeholders,RequestKeys}
  tests/e2e/uds_tokenizer/uds_e2e_suite_test.go           block size / model names / hash seed
Only DATA (token ids, expected keys, config constants) is extracted; no reference code.
"""
import json
import os
import re
import sys

REF = os.environ.get("KVB_REFERENCE_ROOT", "/root/reference")
E2E = os.path.join(REF, "tests/e2e/uds_tokenizer")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kvblock_golden.json")


def _block(src: str, name: str) -> str:
    m = re.search(re.escape(name) + r"\s*=\s*\[\][\w.]+\{(.*?)\n?\s*\}", src, re.S)
    if not m:
        raise SystemExit(f"cannot find {name}")
    return m.group(1)


def _ints(txt: str):
    txt = re.sub(r"//.*", "", txt)
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", txt)]


def main():
    text = open(os.path.join(E2E, "uds_e2e_test.go")).read()
    mm = open(os.path.join(E2E, "uds_e2e_mm_test.go")).read()
    suite = open(os.path.join(E2E, "uds_e2e_suite_test.go")).read()

    model_text = re.search(r'defaultModelName\s*=\s*"([^"]+)"', suite)
    model_text = model_text.group(1) if model_text else re.search(r'"(ibm-granite/[^"]+)"', suite).group(1)
    model_mm = re.search(r'"(Qwen/Qwen2-VL-2B-Instruct)"', mm).group(1)
    bs = re.search(r"BlockSize:\s*(\d+)", suite)
    block_size = int(bs.group(1)) if bs else 4
    seed = re.search(r'HashSeed:\s*"([^"]*)"', suite)
    hash_seed = seed.group(1) if seed else ""

    ph = re.search(r"goldenMMPlaceholders\s*=\s*\[\]kvblock\.PlaceholderRange\{\{Offset:\s*(\d+),\s*Length:\s*(\d+)\}\}", mm)
    hashes = re.findall(r'"([0-9a-f]{64})"', _block(mm, "goldenMMHashes"))

    out = {
        "_source": "llm-d/llm-d-kv-cache@82d31d1 tests/e2e/uds_tokenizer/{uds_e2e_test.go:337-348,uds_e2e_mm_test.go:386-458}",
        "text": {
            "model": model_text, "block_size": block_size, "hash_seed": hash_seed,
            "tokens": _ints(_block(text, "goldenTokenIDs")),
            "request_keys": _ints(_block(text, "goldenRequestKeys")),
        },
        "multimodal": {
            "model": model_mm, "block_size": block_size, "hash_seed": hash_seed,
            "tokens": _ints(_block(mm, "goldenMMTokenIDs")),
            "mm_hashes": {"image": hashes},
            "mm_placeholders": {"image": [{"offset": int(ph.group(1)), "length": int(ph.group(2))}]},
            "request_keys": _ints(_block(mm, "goldenMMRequestKeys")),
        },
    }
    assert len(out["text"]["tokens"]) == 7 and len(out["text"]["request_keys"]) == 1
    assert len(out["multimodal"]["tokens"]) == 418 and len(out["multimodal"]["request_keys"]) == 104
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, model_text, model_mm, block_size, repr(hash_seed))


if __name__ == "__main__":
    sys.exit(main())
