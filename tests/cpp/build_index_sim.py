"""Build tests/_build/libkvb_indexsim.so: csrc/index.cu compiled by g++ with -DKVB_HOST_SIM (tests/cpp/sim_cuda.h stands in
for the device).  TEST INFRASTRUCTURE: it exists so that the index's mutation logic can be checked against the oracle on a
box without a GPU; the product (libkvb.so) is built by nvcc from the same source and has no host path."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "llm-d-kv-cache_b200", "csrc", "index.cu")
OUT = os.path.join(ROOT, "tests", "_build", "libkvb_indexsim.so")


def build() -> str:
    deps = [SRC, os.path.join(ROOT, "tests", "cpp", "sim_cuda.h"), os.path.join(ROOT, "include", "kvb.h"), __file__]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DKVB_HOST_SIM", "-x", "c++", SRC, "-o", OUT,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"), "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function",
           # libkvb.so may already be loaded RTLD_GLOBAL in the same process and defines the same names: bind this
           # library's own references to its own definitions
           "-Wl,-Bsymbolic", "-fno-semantic-interposition"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
