"""The C ABI used from C++ with no Python in the loop: examples/cpp_sh_driver/main.cpp is the reference's C++
Swift-Hohenberg driver (cpp_work/.../Project1/main.cpp: N = 5, d = 2, nonlin_solve(residual, Uo, 6e-6, ...)) against
libjfnk.so.  CPU: it must compile and link.  GPU: its per-step output must match the SciPy oracle."""
import os
import subprocess

import numpy as np
import pytest

from jfnk_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "cpp_sh_driver", "main.cpp")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    exe = str(tmp_path / "sh_driver")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["g++", "-O2", "-std=c++17", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", libdir, "-l:libjfnk.so", "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}",
           f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def xorshift_state(nn):
    s = 88172645463325252
    out = np.empty(nn)
    mask = (1 << 64) - 1
    for i in range(nn):
        s ^= (s << 13) & mask
        s ^= s >> 7
        s ^= (s << 17) & mask
        out[i] = 2.0 * ((s >> 11) / 9007199254740992.0) - 1.0
    return out


def test_cpp_driver_compiles_and_links(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("N,d", [(5, 2.0), (64, 40.0)])
def test_cpp_driver_matches_oracle(tmp_path, N, d):
    from oracle.sh import SHOracle

    steps = 5
    exe = build(tmp_path)
    out = subprocess.run([exe, str(N), str(d), str(steps)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == steps
    o = SHOracle(N=N, d=d)
    U = xorshift_state(N * N)
    for row in rows:
        hist = []
        U = o.step(U, history=hist, f_tol=6e-6)
        assert int(row[1]) == len(hist[0]["iters"])
        assert abs(float(row[3]) - U.sum()) <= 1e-8 * np.abs(U).sum()
        assert abs(float(row[4]) - np.linalg.norm(U)) <= 1e-8 * np.linalg.norm(U)
