"""Build csrc/libjfnk.so for sm_100a, in-tree (nvcc cross-compiles without a GPU).

    python iterative-solvers-summer-2020_b200/csrc/build.py [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libjfnk.so")
SOURCES = ["abi.cpp", "engine.cpp", "cuda_ops.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh", ".cu", ".cpp"))]
    deps.append(os.path.join(ROOT, "include", "jfnk.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
           "-Xcompiler", "-Wno-unused-function", "-Xcompiler", "-Wno-unknown-pragmas", "-cudart", "static"] + ARCH
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-I", HERE, "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES] + ["-ldl", "-Xlinker", "-Bsymbolic"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
