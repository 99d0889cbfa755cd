"""N > 1 path on CPU: world_size-2 (and 3) `gloo` runs of the slab-decomposed solver.

The engine's host logic (csrc/engine.cpp: where halos are exchanged, which scalars are all-reduced, that every
rank takes identical control-flow decisions) runs for real; the device backend is the CPU test double with its
collectives routed through torch.distributed gloo (tests/hostsim/sim.py).  On GPUs the same call sites use NCCL
send/recv + allreduce (csrc/cuda_ops.cu)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import jfnk_b200 as jf
from oracle.sh import SHOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=int(sys.argv[2]))
    import jfnk_b200 as jf
    from tests.hostsim.sim import SimBuffers
    N, nsteps = {N}, {nsteps}
    comm = jf.SlabComm()
    F = jf.SHResidual(N=N, d=0.625 * N, comm=comm, buffers=SimBuffers())
    row0, nrows = F.rows
    U = jf.seeded_slab_state(N, row0, nrows, seed=1234)
    y = F.spmv_L(U)
    hist = []
    U = F.steps(U, nsteps, history=hist)
    full = comm.gather_rows(U, N, N)
    fully = comm.gather_rows(y, N, N)
    if comm.rank == 0:
        np.savez({out!r}, U=full, Lu=fully, nit=[h["nit"] for h in hist], nfev=[h["nfev"] for h in hist])
    dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, N, nsteps, tmp_path):
    out = str(tmp_path / f"out_{world}.npz")
    script = tmp_path / f"worker_{world}.py"
    script.write_text(WORKER.format(root=ROOT, port=_free_port(), N=N, nsteps=nsteps, out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return np.load(out)


def test_slab_rows_partition_and_seeded_state():
    for ny, size in [(64, 2), (61, 3), (16384, 8), (10, 4)]:
        rows = [jf.slab_rows(ny, r, size) for r in range(size)]
        assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
        for (a, n), (b, _) in zip(rows, rows[1:]):
            assert a + n == b
        assert max(n for _, n in rows) - min(n for _, n in rows) <= 1
    whole = jf.seeded_slab_state(32, 0, 32)
    parts = np.concatenate([jf.seeded_slab_state(32, *jf.slab_rows(32, r, 3)) for r in range(3)])
    assert np.array_equal(whole, parts)


@pytest.mark.parametrize("world", [2, 3])
def test_two_and_three_rank_solver_matches_single_rank_and_oracle(world, tmp_path):
    N, nsteps = 48, 3
    multi = _run(world, N, nsteps, tmp_path)
    U0 = jf.seeded_slab_state(N, 0, N, seed=1234)
    o = SHOracle(N=N, d=0.625 * N)
    assert np.abs(multi["Lu"] - o.L @ U0).max() / np.abs(o.L @ U0).max() < 1e-14  # halo exchange incl. periodic wrap
    href = []
    Uref = o.run(U0, nsteps, history=href)
    assert np.linalg.norm(multi["U"] - Uref) / np.linalg.norm(Uref) < 1e-8
    assert list(multi["nit"]) == [len(h["iters"]) for h in href]
    from tests.hostsim.sim import SimBuffers

    single = jf.SHResidual(N=N, d=0.625 * N, buffers=SimBuffers()).steps(U0, nsteps)
    # reduction order differs between 1 and P ranks; the FD-JVP amplifies it to ~1e-10
    assert np.linalg.norm(multi["U"] - single) / np.linalg.norm(single) < 1e-8
