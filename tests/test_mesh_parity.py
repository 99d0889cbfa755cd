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


def test_mesh_jvp_matches_krylov_jacobian_matvec(buffers):
    """KrylovJacobian.matvec (scipy/optimize/_nonlin.py:1557-1565) on the moving-mesh residual:
    (F(x0 + sc v) - F(x0))/sc with sc = omega/||v||, omega = rdiff max(1,|x0|_inf)/max(1,|F(x0)|_inf), at the script's
    initial state (U = 0 on the uniform mesh, PMA2_nk.py:68-71) where F = O(1) and omega ~ 1e-8.  Both sides are FD
    quotients: they agree to rounding(F)/sc, ~1e-9 relative here (tolerance 1e-6)."""
    N = 51
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    Q = (0.5 * X ** 2 + 0.5 * Y ** 2).reshape(-1)
    x0 = np.zeros(N * N)
    o = PMA2Oracle(N=N)
    o.set_mesh(Q)
    o.set_prev(x0)
    F = jf.PMA2Residual(N=N, buffers=buffers)
    F.set_mesh(Q)
    F.set_prev(x0)
    v = (np.sin(2 * X) * np.cos(Y) * (1 - X ** 2) * (1 - Y ** 2)).reshape(-1)
    f0 = o.residual(x0)
    omega = np.sqrt(np.finfo(float).eps) * max(1.0, np.abs(x0).max()) / max(1.0, np.abs(f0).max())
    sc = omega / np.linalg.norm(v)
    ref = (o.residual(x0 + sc * v) - f0) / sc
    F.linearize(x0)
    assert relmax(F.jvp(v), ref) < 1e-6


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


# ---- marching Laplace kernel (mesh_march.cuh): the large-grid path of config 3 ---------------------------------
def _pma2_state(N, seed=3):
    """smooth non-trivial mesh potential and solution on the N x N grid (metric J stays positive)"""
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    # (a small perturbation: a larger one drives Q_etaeta -> 0 on the top edge, where A22 = Q_ksiksi/Q_etaeta then
    #  amplifies the rounding of the 1/h^2 boundary closure by 1e4 and no two fp64 evaluations agree to 1e-9)
    Q = 0.5 * X ** 2 + 0.5 * Y ** 2 + 0.002 * np.cos(2.1 * X + 0.3) * np.sin(1.7 * Y - 0.2)
    rng = np.random.default_rng(seed)
    u = 0.05 * np.sin(3 * X + 1) * np.cos(2 * Y) + 1e-3 * rng.standard_normal((N, N))
    u0 = 0.04 * np.cos(2 * X) * np.cos(3 * Y + 0.5)
    return Q.reshape(-1), u.reshape(-1), u0.reshape(-1)


@pytest.mark.gpu
def test_pma2_marching_kernel_small_grid_vs_reference_golden(cuda_buffers):
    """kernel_variant=2 forces the marching kernel on the reference's own 51 x 51 grid: same goldens as above."""
    g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    F = jf.PMA2Residual(N=51, buffers=cuda_buffers, kernel_variant=2)
    F.set_mesh(g["op_Q"])
    F.set_prev(g["op_Uval"])
    assert relmax(F(g["op_u"]), g["op_residual"]) < 1e-10
    U = jf.newton_krylov(F, g["op_Uval"], verbose=0)
    assert rel(U, g["op_Unew"]) < 1e-8
    assert F.last_history["nit"] == len(g["op_hist"])


@pytest.mark.gpu
@pytest.mark.parametrize("N", [300, 523])
def test_pma2_marching_kernel_vs_point_kernel_and_oracle(cuda_buffers, N):
    """Default rule (marching kernel from 256^2 points up: 2 and 3 strips of 250 columns, several row chunks) against
    the one-thread-per-point kernels (kernel_variant=1) and against the oracle's sparse-matrix residual."""
    Q, u, u0 = _pma2_state(N)
    dt = 1e-4 * ((2.0 / (N - 1)) / 0.04) ** 4
    Fm = jf.PMA2Residual(N=N, dt=dt, buffers=cuda_buffers)
    Fp = jf.PMA2Residual(N=N, dt=dt, buffers=cuda_buffers, kernel_variant=1)
    for F in (Fm, Fp):
        F.set_mesh(Q)
        F.set_prev(u0)
    rm, rp = Fm(u), Fp(u)
    assert relmax(rm, rp) < 1e-13  # same formulas, same order: only FMA contraction may differ
    o = PMA2Oracle(N=N, k=dt)
    o.set_mesh(Q)
    o.set_prev(u0)
    # summation-order / FMA differences are amplified by the two 1/h^2 Laplacians: measured 7e-11 against the
    # reference's goldens at N = 51, growing like (N/51)^4
    assert relmax(rm, o.residual(u)) < 2e-10 * (N / 51.0) ** 4
    # FD-JVP at the same linearisation point.  Linearised at U = 0 (the script's initial state), where F = O(1) and
    # omega ~ 1e-8: at the rough state above |F|_inf ~ 1e9 drives omega to ~1e-17 and the quotient is rounding noise
    # in any implementation.  Both kernels then agree to rounding(F)/sc.
    zero = np.zeros(N * N)
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    v = (np.sin(2 * X) * np.cos(Y) + 1e-6 * np.random.default_rng(5).standard_normal((N, N))).reshape(-1)
    for F in (Fm, Fp):
        F.set_prev(zero)
        F.linearize(zero)
    jm, jp = Fm.jvp(v), Fp.jvp(v)
    assert relmax(jm, jp) < 1e-6
    # line-search form t = x + s dx with the trial iterate kept, through Newton-Krylov: same converged field
    Um = jf.newton_krylov(Fm, zero, verbose=0)
    Up = jf.newton_krylov(Fp, zero, verbose=0)
    assert rel(Um, Up) < 1e-8
    assert Fm.last_history["nit"] == Fp.last_history["nit"]


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1100, 1101])
def test_pma2_marching_kernel_long_row_chunks(cuda_buffers, N):
    """Row chunks of ~22 rows per CTA (the grids above give every CTA exactly one period of the 8x unrolled row loop): the
    operand ring wraps several times per chunk, so the compile-time stage / mbarrier-phase indices of the bulk-TMA path
    (N even) and the cp.async groups (N odd) are exercised across periods.  Residual and FD-JVP against the
    one-thread-per-point kernels, tolerances as in the test above."""
    Q, u, u0 = _pma2_state(N)
    dt = 1e-4 * ((2.0 / (N - 1)) / 0.04) ** 4
    Fm = jf.PMA2Residual(N=N, dt=dt, buffers=cuda_buffers)
    Fp = jf.PMA2Residual(N=N, dt=dt, buffers=cuda_buffers, kernel_variant=1)
    for F in (Fm, Fp):
        F.set_mesh(Q)
        F.set_prev(u0)
    assert relmax(Fm(u), Fp(u)) < 1e-13
    zero = np.zeros(N * N)
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    v = (np.sin(2 * X) * np.cos(Y) + 1e-6 * np.random.default_rng(5).standard_normal((N, N))).reshape(-1)
    for F in (Fm, Fp):
        F.set_prev(zero)
        F.linearize(zero)
    assert relmax(Fm.jvp(v), Fp.jvp(v)) < 1e-6


@pytest.mark.gpu
def test_droplet_marching_laplace_rectangular_grid(cuda_buffers):
    """MARCH_LAP with deriv_bc=1 on the 91 x 61 droplet grid (set_prev and the mesh relaxation use the summed
    Laplace_operator): goldens of the reference's own modules."""
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    F = jf.DropletResidual(buffers=cuda_buffers, kernel_variant=2)
    F.set_mesh(g["state_Q"])
    F.set_prev(g["state_U"], 1e-4)
    assert relmax(F(g["op_u"]), g["op_residual"]) < 1e-10
    Q400 = F.relax_mesh(g["state_Q"], g["state_U"], 3e-9, loops=400)
    assert rel(Q400 - g["state_Q"], g["run_Q0"] - g["state_Q"]) < 1e-10


def test_droplet_initialisers_vs_reference_golden(buffers):
    """SURVEY.md section 8f rank 3 on the engine: compute_U2 on the device + mesh relaxation, driven by the host loop of
    droplet.py:132-248 / :316-358, against the outputs of the reference's own functions.  Tolerances: U 1e-11 relative
    to max (exp / log of the smoothed contact line are library functions on both sides), mesh increment 1e-9."""
    g = np.load(os.path.join(GOLD, "droplet_init_91x61.npz"))
    F = jf.DropletResidual(buffers=buffers)
    Q0 = F.uniform_mesh_potential()
    U, Q = F.initialise_coalescing_droplets(4, g["coal_info"].tolist(), 5e-9, 20)
    assert relmax(U, g["coal_U"]) < 1e-11
    assert rel(Q - Q0, g["coal_Q"] - Q0) < 1e-9
    U, Q = F.initialise_droplet(3, 5e-9, 20)
    assert relmax(U, g["rect_U"]) < 1e-11
    assert rel(Q - Q0, g["rect_Q"] - Q0) < 1e-9
    pmaloops, Rfinal, tol, dtR, dtmesh = g["evolveR_args"]
    U, Q, R, _ = F.evolve_R_explicit(U, Q, int(pmaloops), Rfinal, tol, dtR=dtR, dtmesh=dtmesh)
    assert abs(R - float(g["evolveR_R"])) < 1e-14
    assert relmax(U, g["evolveR_U"]) < 1e-11
    assert rel(Q - Q0, g["evolveR_Q"] - Q0) < 1e-9
    with pytest.raises(ValueError):
        F.compute_U2(Q0, [[0.0, 0.0, -1.0, 1.0]])  # non-positive radius


@pytest.mark.parametrize("key,Nx,Ny,endr,eps", [("rect61", 61, 61, 3.0, 0.01), ("coal81", 81, 61, 5.0, 0.005)])
def test_droplet_step_on_other_shipped_states(buffers, key, Nx, Ny, endr, eps):
    """The other grid sizes the reference ships states for (61 x 61 on [-3,3]^2, 81 x 61 on [-3,5] x [-3,3]; other precursor
    film heights): one pass of evolve_with_PDE's loop body (droplet.py:371-384) -- Newton-Krylov with maxiter=20,
    f_tol=1e-7, then loop_pma(3e-9, 20) -- against the oracle (pinned bit for bit to the reference at 91 x 61)."""
    g = np.load(os.path.join(GOLD, "droplet_states.npz"))
    U0, Q0 = g[key + "_U"], g[key + "_Q"]
    o = DropletOracle(Nx=Nx, Ny=Ny, endl=-3, endr=endr, endb=-3, endt=3, epsilon=eps)
    o.Q = Q0.copy()
    hist = []
    Uref = o.step(U0, 1e-4, dtmesh=3e-9, pmaloops=20, history=hist)
    F = jf.DropletResidual(Nx=Nx, Ny=Ny, endl=-3.0, endr=endr, endb=-3.0, endt=3.0, epsilon=eps, buffers=buffers)
    F.set_mesh(Q0)
    F.set_prev(U0, 1e-4)
    U = jf.newton_krylov(F, U0, verbose=0, maxiter=20, f_tol=1e-7)
    assert F.last_history["f_max"][-1] <= 1e-7
    assert rel(U, Uref) < 1e-8
    assert F.last_history["nit"] == len(hist[0]["iters"])
    Q = F.relax_mesh(Q0, U0, 3e-9, loops=20)
    assert rel(Q - Q0, o.Q - Q0) < 1e-10


def _rect_state(Nx, Ny, endl, endr, endb, endt, eps=0.01):
    """A smooth droplet-like state on a mildly deformed mesh of a grid whose two spacings DIFFER."""
    ksi, eta = np.meshgrid(np.linspace(endl, endr, Nx), np.linspace(endb, endt, Ny))
    Q = 0.5 * ksi ** 2 + 0.5 * eta ** 2 + 0.05 * np.cos(np.pi * (ksi - endl) / (endr - endl)) * np.cos(np.pi * (eta - endb) / (endt - endb))
    U = eps + 0.8 * np.exp(-((ksi - 1.0) ** 2 + eta ** 2) / 0.8) + 0.5 * np.exp(-((ksi - 3.0) ** 2 + (eta - 0.3) ** 2) / 0.5)
    return U.reshape(-1), Q.reshape(-1)


def test_rectangular_spacing_dksi_differs_from_deta(buffers):
    """Every shipped grid has dksi == deta == 0.1, so nothing above tells dksi^2, deta^2 and the dksi*deta divisor of
    M.Leig (droplet.py:831-833, reference quirk 3) apart.  91 x 41 on [-3,6] x [-3,3] has dksi = 0.1, deta = 0.15:
    Laplace_operator, the droplet residual, a Newton-Krylov step and the DCT mesh relaxation against the oracle."""
    Nx, Ny, box = 91, 41, (-3.0, 6.0, -3.0, 3.0)
    U, Q = _rect_state(Nx, Ny, *box)
    o = DropletOracle(Nx=Nx, Ny=Ny, endl=box[0], endr=box[1], endb=box[2], endt=box[3])
    assert abs(o.ops.dksi - 0.1) < 1e-15 and abs(o.ops.deta - 0.15) < 1e-15
    F = jf.DropletResidual(Nx=Nx, Ny=Ny, endl=box[0], endr=box[1], endb=box[2], endt=box[3], buffers=buffers)
    o.set_mesh(Q)
    F.set_mesh(Q)
    vxx, vyy = F.laplace(U)
    rxx, ryy = o.ops.laplace_of(U, o.met)
    assert relmax(vxx, rxx) < 1e-11 and relmax(vyy, ryy) < 1e-11
    o.set_prev(U, 1e-4)
    F.set_prev(U, 1e-4)
    u = U * (1 + 1e-3 * np.sin(np.arange(Nx * Ny)))
    assert relmax(F(u), o.residual(u)) < 1e-10
    # mesh relaxation: the spectral divide uses Leig = lam/(dksi*deta); with dksi^2 or deta^2 the increment is off by 50 %
    o.loop_pma(3e-9, 5)
    Qn = F.relax_mesh(Q, U, 3e-9, loops=5)
    assert rel(Qn - Q, o.Q - Q) < 1e-10
    wrong = DropletOracle(Nx=Nx, Ny=Ny, endl=box[0], endr=box[1], endb=box[2], endt=box[3])
    wrong.ops.Leig = wrong.ops.Leig * (wrong.ops.dksi * wrong.ops.deta) / wrong.ops.dksi ** 2
    wrong.set_mesh(Q)
    wrong.set_prev(U, 1e-4)
    wrong.loop_pma(3e-9, 5)
    assert rel(Qn - Q, wrong.Q - Q) > 1e-3  # the test does discriminate the divisor


@pytest.mark.parametrize("model", ["pma2", "droplet"])
def test_lgmres_restarts_on_mesh_problems(buffers, model):
    """scipy.sparse.linalg.lgmres with maxiter > 1 (outer restarts: r = b - A x needs ||x|| kept across an operator
    application whose unfused mesh residual kernels park their norms in the scratch scalars) on the FD Jacobian of the
    moving-mesh residuals, against SciPy on the oracle's KrylovJacobian."""
    from scipy.optimize._nonlin import KrylovJacobian
    from scipy.sparse.linalg import lgmres

    if model == "pma2":
        g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
        o, F = PMA2Oracle(N=51), jf.PMA2Residual(N=51, buffers=buffers, inner_m=10, outer_k=2)
        o.set_mesh(g["op_Q"]); F.set_mesh(g["op_Q"])
        o.set_prev(g["op_Uval"]); F.set_prev(g["op_Uval"])
        x0 = g["op_u"]
    else:
        g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
        o, F = DropletOracle(), jf.DropletResidual(buffers=buffers, inner_m=10, outer_k=2)
        o.set_mesh(g["state_Q"]); F.set_mesh(g["state_Q"])
        o.set_prev(g["state_U"], 1e-4); F.set_prev(g["state_U"], 1e-4)
        x0 = g["state_U"]
    jac = KrylovJacobian()
    b = o.residual(x0)
    jac.setup(x0.copy(), b, o.residual)
    F.linearize(x0)
    # a short Arnoldi length forces several outer cycles (the FD noise floor of these operators is ~6e-5 relative)
    xs, info_s = lgmres(jac.op, b, rtol=1e-4, atol=0, maxiter=40, inner_m=10, outer_k=2)
    xg, info_g = F.lgmres(b, rtol=1e-4, maxiter=40)
    assert info_s == 0 and info_g == 0
    assert F.last_lgmres["inner"] > 12  # more than one cycle was needed
    r = b - jac.matvec(xg)
    assert np.linalg.norm(r) <= 2e-4 * np.linalg.norm(b)
    assert rel(xg, xs) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("N", [128, 256])
def test_fft_dct_mesh_update_matches_scipy_fft(monkeypatch, N):
    """solve_PMA (PMA2_nk.py:393-403) transforms with scipy.fft.dct / idct(norm="ortho").  On power-of-two grids the
    engine uses its own FFT-based DCT (csrc/dct_fft.cuh: Makhoul permutation + shared-memory radix-2 FFT per row,
    transposes for the column direction) instead of the dense DCT-matrix products; both against the oracle's scipy.fft
    path, on the increment of one mesh update."""
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    Q = np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2 + 0.01 * np.cos(np.pi * X) * np.cos(np.pi * Y), N * N)
    U = np.reshape(-0.3 * np.exp(-4 * (X ** 2 + Y ** 2)), N * N)
    po = PMA2Oracle(N=N)
    po.set_mesh(Q)
    po.Uval = U.copy()
    Qdt = po.ops.solve_pma(po.monitor(), po.met["J"], po.alpha, po.gamma)
    monkeypatch.setenv("JFNK_RELAX_FUSED", "0")  # the per-stage path (what a 2048^2 grid runs)
    got = {}
    for fft in ("1", "0"):
        monkeypatch.setenv("JFNK_DCT_FFT", fft)
        P = jf.PMA2Residual(N=N)
        got[fft] = P.relax_mesh(Q, U, 1e-4, loops=1) - Q
        assert rel(got[fft], 1e-4 * Qdt) < 1e-11, fft
    assert rel(got["1"], got["0"]) < 1e-12
    # several passes (ping-pong buffers, CUDA-graph replay of the recorded pass)
    monkeypatch.setenv("JFNK_DCT_FFT", "1")
    P = jf.PMA2Residual(N=N)
    Q5 = P.relax_mesh(Q, U, 1e-5, loops=6)
    monkeypatch.setenv("JFNK_DCT_FFT", "0")
    P0 = jf.PMA2Residual(N=N)
    assert rel(Q5 - Q, P0.relax_mesh(Q, U, 1e-5, loops=6) - Q) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("N", [300])
def test_dense_dct_mesh_update_tensor_core_gemm(monkeypatch, N):
    """Grids that are neither a power of two nor small enough for the cluster kernel transform with dense DCT-matrix
    products; those run on the fp64 tensor cores (dmma_gemm_kernel: 32 x 32 tiles, edges zero-filled -- 300 = 9 x 32 + 12).
    Against the oracle's scipy.fft path and against the CUDA-core kernel it replaced."""
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    Q = np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2 + 0.01 * np.cos(np.pi * X) * np.cos(np.pi * Y), N * N)
    U = np.reshape(-0.3 * np.exp(-4 * (X ** 2 + Y ** 2)), N * N)
    po = PMA2Oracle(N=N)
    po.set_mesh(Q)
    po.Uval = U.copy()
    Qdt = po.ops.solve_pma(po.monitor(), po.met["J"], po.alpha, po.gamma)
    got = {}
    for dmma in ("1", "0"):
        monkeypatch.setenv("JFNK_DMMA_GEMM", dmma)
        P = jf.PMA2Residual(N=N)
        got[dmma] = P.relax_mesh(Q, U, 1e-4, loops=1) - Q
        assert rel(got[dmma], 1e-4 * Qdt) < 1e-11, dmma
    assert rel(got["1"], got["0"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("grid", ["91x61", "61x61", "81x61", "91x41"])
@pytest.mark.parametrize("gs,tau", [("cgs-ifneeded", 0.25), ("cgs2", 0.25)])
def test_one_launch_droplet_cycle_matches_streaming_path(monkeypatch, grid, gs, tau):
    """The droplet grids run a whole LGMRES cycle plus the first line-search trial as one launch of a 16-CTA cluster with the
    basis and the metric fields in shared memory and the three stencil stages of the Jacobian-vector product chained over
    DSMEM (csrc/mesh_cycle.cuh).  Same point functions and decisions as the streaming kernels (JFNK_CYCLE_FUSED=0): equal
    Newton counts, fields far inside the 1e-8 criterion, and an order of magnitude fewer launches."""
    if grid == "91x61":
        g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
        U0, Q0, kw = g["state_U"], g["state_Q"], {}
    elif grid == "91x41":
        kw = dict(Nx=91, Ny=41, endl=-3.0, endr=6.0, endb=-3.0, endt=3.0)
        U0, Q0 = _rect_state(91, 41, -3.0, 6.0, -3.0, 3.0)
    else:
        g = np.load(os.path.join(GOLD, "droplet_states.npz"))
        key, kw = {"61x61": ("rect61", dict(Nx=61, Ny=61, endl=-3.0, endr=3.0, endb=-3.0, endt=3.0, epsilon=0.01)),
                   "81x61": ("coal81", dict(Nx=81, Ny=61, endl=-3.0, endr=5.0, endb=-3.0, endt=3.0, epsilon=0.005))}[grid]
        U0, Q0 = g[key + "_U"], g[key + "_Q"]
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JFNK_CYCLE_FUSED", fused)
        F = jf.DropletResidual(gs=gs, gs_tau=tau, **kw)
        U, Q, scale = U0.copy(), Q0.copy(), 1.0
        nits, l0 = [], None
        for s in range(4):
            F.set_mesh(Q)
            F.set_prev(U, 1e-4 * scale)
            if l0 is None:
                l0 = F.context().launches()
            Un = jf.newton_krylov(F, U, verbose=0, maxiter=20, f_tol=1e-7)
            nits.append((F.last_history["nit"], F.last_history["nfev"]))
            scale += np.exp(-10 * np.linalg.norm(Un - U))
            U = Un
        out[fused] = (U, nits, F.context().launches() - l0)
    assert [n[0] for n in out["1"][1]] == [n[0] for n in out["0"][1]]
    assert all(abs(a[1] - b[1]) <= 3 for a, b in zip(out["1"][1], out["0"][1]))
    # (the two paths sum their dots in different orders; the FD-JVP noise turns that into ~1e-9 after four steps on the
    #  stiffest shipped state -- measured 1.3e-9 on the 81 x 61 one, 1e-10 on the others)
    assert rel(out["1"][0], out["0"][0]) < 3e-9
    assert out["1"][2] < out["0"][2] / 8


@pytest.mark.gpu
@pytest.mark.parametrize("gs,tau", [("cgs-ifneeded", 0.25), ("cgs-ifneeded", 0.97)])
def test_one_launch_pma2_cycle_matches_streaming_path(monkeypatch, gs, tau):
    """PMA2_nk.py's own 51 x 51 grid through the two-stage operator of csrc/mesh_cycle.cuh (Lap, pull, Lap + pointwise terms)
    against the streaming kernels, three steps of the script's loop (mesh update included)."""
    N, k = 51, 1e-4
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JFNK_CYCLE_FUSED", fused)
        F = jf.PMA2Residual(N=N, gs=gs, gs_tau=tau)
        Q = np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2, N * N)
        U = np.zeros(N * N)
        nits, l0 = [], None
        for s in range(3):
            F.set_mesh(Q)
            F.set_prev(U)
            if l0 is None:
                l0 = F.context().launches()
            Unew = jf.newton_krylov(F, U, verbose=0)
            nits.append((F.last_history["nit"], F.last_history["nfev"], F.last_history["reorth"]))
            Q = F.relax_mesh(Q, U, min((1 + U) ** 3) * k, loops=1)
            U = Unew
        out[fused] = (U, nits, F.context().launches() - l0)
    assert [n[0] for n in out["1"][1]] == [n[0] for n in out["0"][1]]
    assert all(abs(a[1] - b[1]) <= 3 for a, b in zip(out["1"][1], out["0"][1]))
    assert rel(out["1"][0], out["0"][0]) < 3e-9
    assert out["1"][2] < out["0"][2] / 5
    if tau > 0.9:
        assert sum(n[2] for n in out["1"][1]) > 5  # the second Gram-Schmidt pass of the cycle kernel was exercised


def test_time_loop_drivers_match_the_oracle_loops(buffers):
    """``DropletResidual.evolve_with_PDE`` (droplet.py:360-411) and ``PMA2Residual.run`` (PMA2_nk.py:80-106): the scripts' own
    time loops as one call, against the oracle stepping the same loops."""
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    F = jf.DropletResidual(buffers=buffers)
    hist = []
    U, Q, t, scale = F.evolve_with_PDE(g["state_U"], g["state_Q"], dt=1e-4, iterMax=3, dtmesh=3e-9, pmaloops=20, history=hist)
    o = DropletOracle()
    o.Q = g["state_Q"].copy()
    Ur, sc, tr = g["state_U"].copy(), 1.0, 0.0
    for _ in range(2):
        Un = o.step(Ur, 1e-4 * sc, dtmesh=3e-9, pmaloops=20)
        tr += 1e-4 * sc
        sc += np.exp(-10 * np.linalg.norm(Un - Ur))
        Ur = Un
    assert len(hist) == 2 and abs(t - tr) < 1e-12 and abs(scale - sc) < 1e-7  # (scale feeds on |dU|, known to 1e-8)
    assert rel(U, Ur) < 1e-8 and rel(Q, o.Q) < 1e-10
    N = 51
    P = jf.PMA2Residual(N=N, buffers=buffers)
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    U2, Q2, t2 = P.run(np.zeros(N * N), np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2, N * N), 2)
    po = PMA2Oracle(N=N)
    Up = np.zeros(N * N)
    for _ in range(2):
        Up = po.step(Up)
    assert rel(U2, Up) < 1e-8 and rel(Q2, po.Q) < 1e-12
