"""The C-ABI library loads and exports every symbol include/kvb.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kvb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kvb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(kvb):
    names = _declared()
    assert len(names) >= 40
    lib = ctypes.CDLL(os.path.join(ROOT, "llm-d-kv-cache_b200", "libkvb.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in kvb.h but not exported: {missing}"


def test_binding_covers_header(kvb):
    from importlib import import_module
    sig = import_module("llm-d-kv-cache_b200._lib").SIGNATURES
    assert sorted(sig) == _declared()


def test_abi_version_and_no_device_is_loud(kvb):
    assert kvb.lib.kvb_abi_version() == kvb._lib.ABI_VERSION == 5
    if kvb.lib.kvb_device_count() == 0:
        # no CPU fallback: compute entry points must fail, not silently compute on the host
        tp = kvb.kvblock.ChunkedTokenDatabase(16, "")
        import pytest
        with pytest.raises(Exception):
            tp.tokens_to_kv_block_keys(0, list(range(32)), "m")
        with pytest.raises(Exception):
            kvb.kvblock.Index(size=16)


def test_header_cites_reference():
    src = open(os.path.join(ROOT, "include", "kvb.h")).read()
    for needle in ("tensor_copier.cu", "storage_offload.cpp", "token_processor.go", "in_memory.go",
                   "kvblock_scorer.go", "indexer.go"):
        assert needle in src


def test_oracle_is_only_imported_by_the_checkers():
    """The product package and tools/ must not import, link or execute anything under oracle/."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|oracle[/.](kvblock|offload|kvevents)_oracle|oracle/_ref", re.M)
    offenders = []
    for sub in ("llm-d-kv-cache_b200", "tools", "include", "go"):
        for dp, _, fns in os.walk(os.path.join(root, sub)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".h", ".hpp", ".go", ".cpp")):
                    if pat.search(open(os.path.join(dp, fn), errors="ignore").read()):
                        offenders.append(os.path.relpath(os.path.join(dp, fn), root))
    assert offenders == []
