"""Builds libplb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "plb200_api.cu")
OUT = os.path.join(HERE, "plb200", "libplb200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("plb200_api.cu", "gram.cuh", "syrk.cuh", "prelude.cuh", "ptx.cuh", "hoststage.h", "multi.inl", "eigh.cuh")] + [
    os.path.join(HERE, "..", "include", "plb200.h")
]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-pthread", "-shared", "-ldl", "-lpthread"]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
