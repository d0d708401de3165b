"""The C++ host-side mirror of the Go interfaces (include/kvb_kvblock.hpp): compiles everywhere (CPU), runs its parity
program on the GPU box (golden vectors, Index contract, ScoreTokens known answers, Go error texts)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_kvblock.cpp")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
PKG = os.path.join(ROOT, "llm-d-kv-cache_b200")


def _write_golden(golden, path):
    def block(g, hashes=(), ranges=()):
        return "\n".join([
            "model " + g["model"], "block_size %d" % g["block_size"],
            "tokens " + " ".join(map(str, g["tokens"])), "keys " + " ".join(map(str, g["request_keys"])),
            "mm_hashes " + " ".join(hashes), "mm_ranges " + " ".join("%d %d" % r for r in ranges)])
    m = golden["multimodal"]
    with open(path, "w") as f:
        f.write(block(golden["text"]) + "\n")
        f.write(block(m, m["mm_hashes"]["image"], [(r["offset"], r["length"]) for r in m["mm_placeholders"]["image"]]) + "\n")


def test_header_compiles(kvb, tmp_path):
    """Syntax/type check of the header and the test program without a GPU (compile only)."""
    r = subprocess.run([CXX, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", SRC,
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_mirror_parity(kvb, torch_cuda, golden, tmp_path):
    exe, gold = str(tmp_path / "test_kvblock"), str(tmp_path / "golden.txt")
    _write_golden(golden, gold)
    r = subprocess.run([CXX, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", PKG, "-lkvb",
                        "-Wl,-rpath," + PKG, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, gold], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK ")


def test_engine_header_compiles(kvb, tmp_path):
    r = subprocess.run([CXX, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", "/usr/local/cuda/include", "-c", os.path.join(ROOT, "tests", "cpp", "test_engine.cpp"),
                        "-o", str(tmp_path / "e.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_engine_roundtrip(kvb, torch_cuda, tmp_path):
    exe = str(tmp_path / "test_engine")
    r = subprocess.run([CXX, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "tests", "cpp", "test_engine.cpp"), "-o", exe, "-L", PKG, "-lkvb",
                        "-Wl,-rpath," + PKG, "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64",
                        "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(tmp_path / "files")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK engine round trips" in r.stdout
