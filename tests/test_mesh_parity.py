"""PMA2 and droplet residual / Newton-Krylov parity: the engine's matrix-free moving-mesh kernels (CUDA under
`-m gpu`, the CPU test double otherwise) against the golden vectors of the reference's own modules and against
the oracle (oracle/mesh.py, itself pinned bit-for-bit to those goldens in tests/test_oracle.py).

Tolerances: operator applications 1e-11 relative to the field's max (matrix-free rows vs the reference's
COO/CSR products differ only by summation order; the metric terms involve divisions by J and 1/h^4 factors);
fields after Newton-Krylov steps 1e-8 relative L2 (north_star)."""
import os

import numpy as np
import pytest

import jfnk_b200 as jf
from oracle.mesh import DropletOracle, PMA2Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relmax(a, b):
    return np.abs(np.ravel(a) - np.ravel(b)).max() / np.abs(np.ravel(b)).max()


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def test_pma2_operators_and_residual_vs_reference_golden(buffers):
    g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    F = jf.PMA2Residual(N=51, buffers=buffers)
    F.set_mesh(g["op_Q"])
    vxx, vyy = F.laplace(g["op_u"])
    assert relmax(vxx, g["op_vxx"]) < 1e-11 and relmax(vyy, g["op_vyy"]) < 1e-11
    F.set_prev(g["op_Uval"])
    assert relmax(F(g["op_u"]), g["op_residual"]) < 1e-10
    U = jf.newton_krylov(F, g["op_Uval"], verbose=0)
    assert rel(U, g["op_Unew"]) < 1e-8
    assert F.last_history["nit"] == len(g["op_hist"])


def test_pma2_time_loop_vs_reference_golden(buffers):
    """PMA2_nk.py main() loop (:80-106): device Newton-Krylov per step; the mesh update (DCT PMA solve) is the
    'next' component of SURVEY.md section 8f and runs on the host oracle, feeding Q to the engine each step."""
    g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    o = PMA2Oracle(N=51)
    F = jf.PMA2Residual(N=51, buffers=buffers)
    U = np.zeros(51 * 51)
    Uo = U.copy()
    for s in range(3):
        F.set_mesh(o.Q)
        F.set_prev(U)
        Unew = jf.newton_krylov(F, U, verbose=0)
        Uo = o.step(Uo)  # advances o.Q exactly as the script does
        assert rel(Unew, g[f"run_U{s}"]) < 1e-8, s
        U = Unew


def test_droplet_operators_and_residual_vs_reference_golden(buffers):
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    F = jf.DropletResidual(buffers=buffers)
    F.set_mesh(g["state_Q"])
    vxx, vyy = F.laplace(g["op_u"])
    assert relmax(vxx, g["op_vxx"]) < 1e-11 and relmax(vyy, g["op_vyy"]) < 1e-11
    F.set_prev(g["state_U"], 1e-4)
    assert relmax(F(g["op_u"]), g["op_residual"]) < 1e-10


def test_droplet_coalescence_steps_vs_reference_golden(buffers):
    """config 2 (first steps): initdrop_coal_1_91-61 state, newton_krylov(..., maxiter=20, f_tol=1e-7) per step
    (droplet.py:383); loop_pma(3e-9, 400) on the host oracle between steps."""
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    o = DropletOracle()
    o.Q = g["state_Q"].copy()
    F = jf.DropletResidual(buffers=buffers)
    U = g["state_U"].copy()
    Uo = U.copy()
    for s in range(2):
        F.set_mesh(o.Q)
        F.set_prev(U, 1e-4)
        Unew = jf.newton_krylov(F, U, verbose=0, maxiter=20, f_tol=1e-7)
        assert F.last_history["f_max"][-1] <= 1e-7
        Uo = o.step(Uo, 1e-4)
        assert rel(Unew, g[f"run_U{s}"]) < 1e-8, s
        assert F.last_history["nit"] == len(g[f"run_hist{s}"])
        U = Unew


def test_mesh_relaxation_matches_loop_pma_and_solve_pma(buffers):
    """loop_pma / solve_PMA / compute_and_smooth_monitor on the engine (dense
    orthonormal DCT-II products instead of scipy.fft) against the oracle's restatement of the reference loop."""
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    o = DropletOracle()
    o.set_mesh(g["state_Q"])
    o.set_prev(g["state_U"], 1e-4)
    o.loop_pma(3e-9, 25)
    F = jf.DropletResidual(buffers=buffers)
    Q = F.relax_mesh(g["state_Q"], g["state_U"], 3e-9, loops=25)
    dq_ref = o.Q - g["state_Q"]
    assert rel(Q - g["state_Q"], dq_ref) < 1e-11  # the increment itself, not just Q
    # the golden of the reference's own two steps: Q after step 0 is loop_pma(3e-9, 400) applied to the state
    Q400 = F.relax_mesh(g["state_Q"], g["state_U"], 3e-9, loops=400)
    assert rel(Q400 - g["state_Q"], g["run_Q0"] - g["state_Q"]) < 1e-10
    # PMA2: one solve_PMA + explicit update (PMA2_nk.py:94,103) with the epsilon = 0 monitor 1/(1+u)^6
    p = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    po = PMA2Oracle(N=51)
    po.set_mesh(p["op_Q"])
    po.Uval = p["op_Uval"].copy()
    Qdt = po.ops.solve_pma(po.monitor(), po.met["J"], po.alpha, po.gamma)
    P = jf.PMA2Residual(N=51, buffers=buffers)
    Q2 = P.relax_mesh(p["op_Q"], p["op_Uval"], 1e-4, loops=1)
    assert rel(Q2 - p["op_Q"], 1e-4 * Qdt) < 1e-12
    with pytest.raises(ValueError):
        P(p["op_u"])  # metric fields are stale after a relaxation: set_mesh must be called again


def test_pma2_full_loop_on_engine_vs_reference_golden(buffers):
    """PMA2_nk.py main() loop (:80-106) entirely on the engine: metrics, CN term, Newton-Krylov, mesh update."""
    g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    N, k = 51, 1e-4
    F = jf.PMA2Residual(N=N, buffers=buffers)
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    Q = np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2, N * N)
    U = np.zeros(N * N)
    for s in range(3):
        F.set_mesh(Q)
        F.set_prev(U)
        dt = min((1 + U) ** 3) * k  # compute_g() * k (:91, :446-450)
        Unew = jf.newton_krylov(F, U, verbose=0)
        Q = F.relax_mesh(Q, U, dt, loops=1)  # solve_PMA() with the OLD solution; Q.val += dt * Q.dt (:94,:103)
        assert rel(Unew, g[f"run_U{s}"]) < 1e-8, s
        assert rel(Q, g[f"run_Q{s}"]) < 1e-12, s
        U = Unew


def test_droplet_state_file_round_trip(tmp_path):
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    p = tmp_path / "initdrop_test.txt"
    jf.save_droplet_state(str(p), g["state_U"], g["state_Q"])
    U, Q = jf.load_droplet_state(str(p))
    assert np.array_equal(U, g["state_U"]) and np.array_equal(Q, g["state_Q"])


def test_mesh_argument_errors(buffers):
    F = jf.PMA2Residual(N=51, buffers=buffers)
    with pytest.raises(ValueError):
        F.set_prev(np.zeros(51 * 51))  # mesh potential not set
    with pytest.raises(ValueError):
        jf.PMA2Residual(N=5, buffers=buffers).context()
