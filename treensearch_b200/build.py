"""In-tree build of libtnsb.so with nvcc for sm_100a (no JIT cache: the .so must travel with the repository snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "tnsb.cu")
OUT = os.path.join(_HERE, "libtnsb.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # the distance test must round exactly like the reference ("parity")
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcdir = os.path.join(_HERE, "csrc")
    deps = [os.path.join(srcdir, f) for f in os.listdir(srcdir)] + [os.path.join(_HERE, "..", "include", "tnsb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.run(cmd, check=True)
    return OUT


BENCH_CPP_SRC = os.path.join(_HERE, "..", "tools", "bench_cpp.cpp")
BENCH_CPP = os.path.join(_HERE, "bench_cpp")


def build_bench_cpp(force: bool = False) -> str:
    """tools/bench_cpp.cpp against include/TreeNSearch + libtnsb.so: the drop-in C++ caller bench.py times end to end."""
    build_library()
    if not force and os.path.exists(BENCH_CPP) and os.path.getmtime(BENCH_CPP) > max(os.path.getmtime(BENCH_CPP_SRC), os.path.getmtime(OUT)):
        return BENCH_CPP
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-I", os.path.join(_HERE, "..", "include"), BENCH_CPP_SRC, "-o", BENCH_CPP,
           "-L", _HERE, "-ltnsb", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return BENCH_CPP


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
