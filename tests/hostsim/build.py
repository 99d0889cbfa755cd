"""Build the CPU test double of the device backend (TEST INFRASTRUCTURE, see host_ops.cpp).

Links the product's host-side solver logic (csrc/abi.cpp, csrc/engine.cpp) against
tests/hostsim/host_ops.cpp instead of the CUDA backend, into tests/hostsim/_build/.
The product package never loads this library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "iterative-solvers-summer-2020_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libjfnk_hostsim.so")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, "abi.cpp"), os.path.join(CSRC, "engine.cpp"), os.path.join(HERE, "host_ops.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(ROOT, "include", "jfnk.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-ffp-contract=off", "-Wl,-Bsymbolic",
           "-I", CSRC, "-o", LIB] + srcs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
