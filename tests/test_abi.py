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


def _prototypes():
    """name -> number of parameters, parsed from include/kvb.h (comments stripped)."""
    src = open(os.path.join(ROOT, "include", "kvb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(kvb_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_binding_argument_counts_match_the_header(kvb):
    """The ctypes table must take exactly as many arguments as the prototype in include/kvb.h declares — a prototype that
    grows an argument (kvb_hash_token_blocks_dev did in round 2) shows up here, on a box without a GPU."""
    from importlib import import_module
    sig = import_module("llm-d-kv-cache_b200._lib").SIGNATURES
    protos = _prototypes()
    wrong = {n: (len(a), protos.get(n)) for n, (_, a) in sig.items() if protos.get(n) != len(a)}
    assert not wrong, wrong


def test_python_callers_pass_the_declared_number_of_arguments():
    """Every direct `lib.kvb_*(...)` call in the benches, tools and tests passes as many arguments as the header declares
    (a static check: the calls that only run on a GPU box are exactly the ones a CPU-only run cannot execute)."""
    protos = _prototypes()
    call = re.compile(r"\b(kvb_[a-z0-9_]+)\s*\(")
    offenders = []
    files = [os.path.join(ROOT, f) for f in ("bench.py", "bench_extras.py", "__graft_entry__.py")]
    for sub in ("tests", "tools", "llm-d-kv-cache_b200"):
        for dp, _, fns in os.walk(os.path.join(ROOT, sub)):
            files += [os.path.join(dp, fn) for fn in fns if fn.endswith(".py")]
    for path in files:
        src = open(path).read()
        for m in call.finditer(src):
            name = m.group(1)
            if name not in protos or src[max(0, m.start() - 4):m.start()].endswith("def "):
                continue
            if not re.search(r"(lib|_lib|L)\.\s*$", src[max(0, m.start() - 8):m.start()]):
                continue
            depth, i, n_args, seen = 1, m.end(), 0, False
            while i < len(src) and depth:
                c = src[i]
                if c in "([{":
                    depth += 1
                elif c in ")]}":
                    depth -= 1
                elif c == "," and depth == 1:
                    n_args += 1
                elif not c.isspace() and depth >= 1:
                    seen = True
                i += 1
            n_args = n_args + 1 if seen else 0
            if "*" in src[m.end():i]:          # star-args: cannot be counted statically
                continue
            if n_args != protos[name]:
                offenders.append((os.path.relpath(path, ROOT), name, n_args, protos[name]))
    assert not offenders, offenders
