"""Swift-Hohenberg parity: the engine (CUDA kernels under `-m gpu`, the CPU test double of the device
backend otherwise) against the reference's SciPy path restated in oracle/sh.py.

Tolerances (BASELINE.json north_star; SURVEY.md section 8a "parity note"):
  * single operator applications (SpMV, F(u), one JVP at fixed omega): 1e-13 relative (rounding only);
  * the field after N time steps: 1e-8 relative L2;
  * per-iteration Newton residual norms: the FD-JVP amplifies rounding differences by 1/omega, so two
    *correct* fp64 implementations differ by 1e-5..3e-2 relative after the first update (measured for the
    oracle against an algebraically identical stencil in SURVEY.md); they are compared at 5e-2 with the
    first-evaluation norm ||F(u0)|| at 1e-12, and the iteration counts must agree.
"""
import numpy as np
import pytest

import jfnk_b200 as jf
from oracle.sh import SHLinearisedOracle, SHOracle, seeded_state


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def relmax(a, b):
    return np.abs(np.ravel(a) - np.ravel(b)).max() / np.abs(np.ravel(b)).max()


CASES = [(61, 40.0), (64, 40.0)]  # config 1: the droplet-sized 61 and the script's N = 64


@pytest.mark.parametrize("N,d", CASES + [(96, 60.0), (5, 2.0)])  # (5, 2.0): the C++ driver's grid (main.cpp:3-4)
def test_spmv_matches_reference_matrices(buffers, N, d):
    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    u = seeded_state(N, 0)
    assert relmax(F.spmv_lap(u), o.Lap @ u) < 1e-14
    assert relmax(F.spmv_L(u), o.L @ u) < 1e-14


@pytest.mark.parametrize("N,d", CASES)
def test_residual_matches_reference(buffers, N, d):
    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    Uo = seeded_state(N)
    u = Uo + 0.05 * seeded_state(N, 7)
    o.set_prev(Uo)
    F.set_prev(Uo)
    assert relmax(F(u), o.residual(u)) < 1e-13
    assert relmax(F(Uo), o.residual(Uo)) < 1e-13


@pytest.mark.parametrize("N,d", CASES)
def test_jvp_matches_krylov_jacobian(buffers, N, d):
    """KrylovJacobian.matvec at a fixed linearisation point (scipy/optimize/_nonlin.py:1557-1565)."""
    from scipy.optimize._nonlin import KrylovJacobian

    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    Uo = seeded_state(N)
    x0 = Uo + 0.01 * seeded_state(N, 3)
    o.set_prev(Uo)
    F.set_prev(Uo)
    jac = KrylovJacobian()
    jac.setup(x0.copy(), o.residual(x0), o.residual)
    F.linearize(x0)
    for seed in (11, 12):
        v = seeded_state(N, seed)
        ref = jac.matvec(v)
        got = F.jvp(v)
        # both are FD quotients with step omega/||v|| ~ 1e-10: rounding noise of F divided by the step
        noise = 64 * np.finfo(float).eps * np.abs(o.residual(x0)).max() / (jac.omega / np.linalg.norm(v))
        assert np.abs(got - ref).max() < max(noise, 1e-9 * np.abs(ref).max())
    assert np.all(F.jvp(np.zeros(N * N)) == 0.0)


@pytest.mark.parametrize("N,d", CASES)
def test_lgmres_matches_scipy_on_linear_operator(buffers, N, d):
    """scipy.sparse.linalg.lgmres on the same (FD-Jacobian) operator, one outer cycle as newton_krylov uses it."""
    from scipy.optimize._nonlin import KrylovJacobian
    from scipy.sparse.linalg import lgmres

    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    Uo = seeded_state(N)
    o.set_prev(Uo)
    F.set_prev(Uo)
    x0 = Uo + 0.01 * seeded_state(N, 3)
    jac = KrylovJacobian()
    b = o.residual(x0)
    jac.setup(x0.copy(), b, o.residual)
    F.linearize(x0)
    rtol = 1e-3
    outer_v = []
    xs, _ = lgmres(jac.op, b, rtol=rtol, atol=0, maxiter=1, outer_k=10, outer_v=outer_v, prepend_outer_v=True,
                   store_outer_Av=False)
    xg, info = F.lgmres(b, rtol=rtol, maxiter=1)
    # both satisfy the same residual bound; they agree to the inner tolerance times a modest factor
    assert rel(xg, xs) < 2e-3
    r = b - jac.matvec(xg)
    assert np.linalg.norm(r) <= 1.5 * rtol * np.linalg.norm(b)
    assert F.last_lgmres["inner"] >= 1


def test_lgmres_exact_linear_operator_with_augmentation(buffers):
    """LGMRES against scipy.sparse.linalg.lgmres on an exactly linear operator (the linearised-SH matrix
    I + D - L k/2, no finite-difference noise): same inner iteration counts, same solutions, and the
    augmentation vectors (outer_v, prepend order, capped at outer_k) carried across calls."""
    from scipy.sparse import diags
    from scipy.sparse.linalg import lgmres

    N = 48
    o = SHLinearisedOracle(N=N, d=30.0)
    S = jf.SHLinearised(N=N, d=30.0, buffers=buffers, outer_k=2, gs="cgs2")
    U, Uo = seeded_state(N, 4), seeded_state(N, 5)
    b = S.prepare(U, Uo)
    D = diags(np.multiply(5 * U - Uo, 5 * U - Uo) * o.k / 16 - o.g * o.k * U, 0)
    A = (o.I + D - o.L * o.k / 2).tocsr()
    assert relmax(b, (o.I + o.L * o.k / 2) @ U) < 1e-14
    v = seeded_state(N, 6)
    assert relmax(S.jvp(v), A @ v) < 1e-14
    outer_v = []
    count = [0]

    def mv(x):
        count[0] += 1
        return A @ x

    from scipy.sparse.linalg import LinearOperator

    op = LinearOperator(A.shape, matvec=mv)
    for call, rtol in enumerate((1e-2, 1e-4, 1e-6, 1e-8)):
        count[0] = 0
        xs, _ = lgmres(op, b, rtol=rtol, atol=0, maxiter=1, inner_m=30, outer_k=2, outer_v=outer_v,
                       prepend_outer_v=True, store_outer_Av=False)
        xg, _ = S.lgmres(b, rtol=rtol, maxiter=1, reset=(call == 0))
        assert S.last_lgmres["inner"] == count[0] - 1  # SciPy spends one matvec on r = A @ 0 - b (lgmres.py:150)
        assert rel(xg, xs) < 50 * rtol * 1e-2 + 1e-9
    assert len(outer_v) == 2
    # full solve with restarts
    xs, info_s = lgmres(A, b, rtol=1e-12, atol=0, maxiter=50, inner_m=30, outer_k=2)
    xg, info_g = S.lgmres(b, rtol=1e-12, maxiter=50)
    assert info_s == 0 and info_g == 0
    assert rel(xg, xs) < 1e-10


@pytest.mark.parametrize("N,d", CASES)
def test_newton_history_and_field_after_steps(buffers, N, d):
    """Config 1: seeded IC, default script parameters, the script's loop (sh_scipy_nk.py:53-61)."""
    nsteps = 12
    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    U0 = seeded_state(N)
    href = []
    Uref = o.run(U0, nsteps, history=href)
    hist = []
    U = F.steps(U0, nsteps, history=hist)
    assert rel(U, Uref) < 1e-8
    for s, (hr, h) in enumerate(zip(href, hist)):
        # first evaluation of the first step sees identical inputs; later steps start from fields that
        # already differ at the 1e-10 level
        ftol = 1e-12 if s == 0 else 1e-7
        assert abs(h["f0_max"] - hr["f0_max"]) <= ftol * hr["f0_max"]
        assert abs(h["f0_l2"] - hr["f0_l2"]) <= ftol * hr["f0_l2"]
        assert h["nit"] == len(hr["iters"])
        assert abs(h["nfev"] - hr["nfev"]) <= 2
        for i, (fmax, fl2) in enumerate(hr["iters"]):
            assert abs(h["f_l2"][i] - fl2) <= 5e-2 * fl2 + 1e-9
            assert abs(h["f_max"][i] - fmax) <= 1e-1 * fmax + 1e-9


def test_newton_krylov_front_end_matches_scipy_signature(buffers, capsys):
    """newton_krylov(F, Uo, verbose=1) as sh_scipy_nk.py:61 calls it; output, trace format, exceptions."""
    from scipy.optimize import newton_krylov as scipy_nk

    N = 64
    o = SHOracle(N=N)
    F = jf.SHResidual(N=N, buffers=buffers)
    Uo = seeded_state(N)
    o.set_prev(Uo)
    F.set_prev(Uo)
    ref = scipy_nk(o.residual, Uo, verbose=1)
    ref_out = capsys.readouterr().out.strip().splitlines()
    got = jf.newton_krylov(F, Uo, verbose=1)
    got_out = capsys.readouterr().out.strip().splitlines()
    assert got.shape == Uo.shape and got.dtype == np.float64
    assert rel(got, ref) < 1e-9
    assert len(got_out) == len(ref_out)
    for a, b in zip(got_out, ref_out):
        assert a.split(":")[0] == b.split(":")[0] and "|F(x)| =" in a and a.endswith("step 1")
    # 2-D input keeps its shape (SciPy's _array_like)
    got2 = jf.newton_krylov(F, Uo.reshape(N, N))
    assert got2.shape == (N, N)
    # callback(x, Fx) once per Newton iteration
    seen = []
    jf.newton_krylov(F, Uo, callback=lambda x, f: seen.append((x.copy(), np.abs(f).max())))
    assert len(seen) == F.last_history["nit"]
    assert rel(seen[-1][0], got) < 1e-12
    # NoConvergence carries the last iterate (scipy/optimize/_nonlin.py:258-260)
    with pytest.raises(jf.NoConvergence) as ei:
        jf.newton_krylov(F, Uo, maxiter=2)
    assert ei.value.args[0].shape == Uo.shape
    with pytest.raises(Exception) as es:
        scipy_nk(o.residual, Uo, maxiter=2)
    assert type(es.value).__name__ == "NoConvergence"
    # un-converged early iterates carry the FD-JVP noise of the first Newton steps (||F|| ~ 650 => omega ~ 1e-10)
    assert rel(ei.value.args[0], es.value.args[0]) < 2e-3
    # f_tol / iter / line_search=None honoured
    x_it = jf.newton_krylov(F, Uo, iter=2)
    assert F.last_history["nit"] == 2
    x_ref = scipy_nk(o.residual, Uo, iter=2)
    assert rel(x_it, x_ref) < 2e-3
    jf.newton_krylov(F, Uo, f_tol=1e-9, line_search=None)
    assert F.last_history["f_max"][-1] <= 1e-9
    with pytest.raises(ValueError):
        jf.newton_krylov(F, Uo, line_search="bogus")
    with pytest.raises(NotImplementedError):
        jf.newton_krylov(F, Uo, method="gmres")
    with pytest.raises(TypeError):
        jf.newton_krylov(o.residual, Uo)


def test_armijo_backtracking_path(buffers):
    """A large time step makes the full Newton step fail the Armijo test: the quadratic / cubic backtracking of
    scalar_search_armijo (scipy/optimize/_linesearch.py:698-753) must take the same steps as SciPy."""
    N = 32
    kw = dict(N=N, d=20.0, k=50.0, r=0.3, g=1.5)
    o = SHOracle(**kw)
    F = jf.SHResidual(buffers=buffers, **kw)
    U0 = 2.0 * seeded_state(N, 5)
    steps_ref = []
    import scipy.optimize._nonlin as nl

    orig = nl._nonlin_line_search

    def spy(*a, **k):
        out = orig(*a, **k)
        steps_ref.append(out[0])
        return out

    nl._nonlin_line_search = spy
    try:
        o.set_prev(U0)
        from scipy.optimize import newton_krylov as scipy_nk

        ref = scipy_nk(o.residual, U0, maxiter=40)
    finally:
        nl._nonlin_line_search = orig
    F.set_prev(U0)
    got = jf.newton_krylov(F, U0, maxiter=40)
    h = F.last_history
    assert any(s < 1.0 for s in steps_ref), "test problem must exercise backtracking"
    assert h["nit"] == len(steps_ref)
    k = int(np.argmax(np.array(steps_ref) < 1.0))  # first backtracking iteration: same step length
    np.testing.assert_allclose(h["step"][: k + 1], steps_ref[: k + 1], rtol=5e-3)
    assert rel(got, ref) < 1e-5


def test_linearised_step_matches_spsolve(buffers):
    """sh_linearised.py:51-57: the reference factorises with SuperLU; device LGMRES to rtol 1e-13."""
    N = 64
    o = SHLinearisedOracle(N=N)
    S = jf.SHLinearised(N=N, buffers=buffers)
    U0 = seeded_state(N, 2)
    ref = o.run(U0, 5)
    U, Uo = U0, U0
    for _ in range(5):
        U, Uo = S.steps(U, Uo, nsteps=1)
    assert S.last_info["info"] == 0
    assert rel(U, ref) < 1e-9
    U5, _ = S.steps(U0, U0, nsteps=5)
    assert rel(U5, ref) < 1e-9


def test_argument_errors(buffers):
    F = jf.SHResidual(N=64, buffers=buffers)
    with pytest.raises(ValueError):
        F(np.zeros(64 * 64))  # set_prev not called yet
    F.set_prev(np.zeros(64 * 64))
    with pytest.raises(ValueError):
        F(np.zeros(10))
    with pytest.raises(ValueError):
        jf.SHResidual(N=64, buffers=buffers, inner_m=60).context()
    with pytest.raises(ValueError):
        jf.SHResidual(N=3, buffers=buffers).context()
    with pytest.raises(ValueError):
        F.jvp(np.ones(64 * 64))  # linearize not called


def test_nonfinite_residual_raises_like_scipy(buffers):
    """KrylovJacobian.matvec raises ValueError('Function returned non-finite results') (_nonlin.py:1563-1564)."""
    N = 32
    F = jf.SHResidual(N=N, d=20.0, buffers=buffers)
    U0 = 1e120 * seeded_state(N, 9)  # u^3 overflows to inf
    F.set_prev(U0)
    with pytest.raises(ValueError):
        jf.newton_krylov(F, U0, maxiter=5)
    # the context stays usable afterwards
    U1 = seeded_state(N, 9)
    F.set_prev(U1)
    out = jf.newton_krylov(F, U1)
    assert np.isfinite(out).all()


def test_preconditioned_newton_krylov_matches_scipy_inner_M(buffers):
    """newton_krylov(..., inner_M=M).  The same Fourier preconditioner
    M = (I/k - L/2)^-1 is given to SciPy (as a LinearOperator on NumPy vectors) and to the engine (acting on device
    vectors); LGMRES applies it from the left in both.  On the script's fixed domain d = 40 at N = 128 the plain solve
    needs several hundred residual evaluations per step; the preconditioned one a few dozen, with the same Newton
    iterates: iteration counts equal, F-evaluation counts within 2, field to 1e-8."""
    from scipy.optimize import newton_krylov
    from scipy.sparse.linalg import LinearOperator

    N, d = 128, 40.0
    o = SHOracle(N=N, d=d)
    F = jf.SHResidual(N=N, d=d, buffers=buffers)
    M = jf.SHFourierPreconditioner.for_residual(F)
    # the symbol inverts the constant-coefficient part of the Jacobian exactly
    v = seeded_state(N, 5)
    Jv = v / o.k - (o.L @ v) / 2
    assert relmax(M.matvec(Jv), v) < 1e-10
    U0 = 0.1 * seeded_state(N)
    o.set_prev(U0)
    nit_ref, nfev0 = [], o.nfev
    Uref = newton_krylov(o.residual, U0, inner_M=LinearOperator((N * N, N * N), matvec=M.matvec, dtype=np.float64),
                         callback=lambda x, f: nit_ref.append(np.abs(f).max()))
    nfev_ref = o.nfev - nfev0
    F.set_prev(U0)
    U = jf.newton_krylov(F, U0, inner_M=M)
    h = F.last_history
    assert rel(U, Uref) < 1e-8
    assert h["nit"] == len(nit_ref)
    assert abs(h["nfev"] - nfev_ref) <= 2
    # and it is what makes this stiff configuration cheap: the plain solve needs an order of magnitude more
    F.set_prev(U0)
    jf.newton_krylov(F, U0)
    assert F.last_history["nfev"] > 4 * h["nfev"]
    # a failing preconditioner surfaces as the Python exception it raised, not as a crash across the C ABI
    def bad(v):
        raise ZeroDivisionError("boom")
    F.set_prev(U0)
    with pytest.raises(ZeroDivisionError):
        jf.newton_krylov(F, U0, inner_M=bad)


def test_engine_vs_reference_script_golden(buffers):
    """The engine against the outputs of the reference scripts themselves (tests/golden/sh_n64.npz: sh_scipy_nk.py and
    sh_linearised.py executed unmodified for three time steps from a seeded np.random.randn state): fields to 1e-8
    relative L2, equal Newton iteration counts, first-iteration norm to 5e-2 (FD-JVP noise, see the module docstring)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sh_n64.npz"))
    F = jf.SHResidual(N=64, buffers=buffers)
    U = g["nk_U0"]
    for s in (1, 2, 3):
        F.set_prev(U)
        U = jf.newton_krylov(F, U)
        assert rel(U, g[f"nk_U{s}"]) < 1e-8, s
        h = F.last_history
        assert h["nit"] == len(g[f"nk_hist{s}"]), s
        assert abs(h["f_max"][0] - g[f"nk_hist{s}"][0]) <= 5e-2 * g[f"nk_hist{s}"][0], s
    FL = jf.SHLinearised(N=64, buffers=buffers)
    U = g["lin_U0"]
    Uo = U.copy()
    for s in (1, 2, 3):
        U, Uo = FL.steps(U, Uo, nsteps=1)
        assert rel(U, g[f"lin_U{s}"]) < 1e-10, s


@pytest.mark.parametrize("kw", [dict(x_tol=1e-6), dict(x_rtol=1e-7), dict(f_rtol=1e-4), dict(f_tol=1e-3, x_tol=1e-5),
                                dict(f_tol=1e-9, f_rtol=1e-6, x_rtol=1e-9)])
def test_termination_conditions_match_scipy(buffers, kw):
    """TerminationCondition (scipy/optimize/_nonlin.py:328-385) beyond the default f_tol: x_tol / x_rtol compare
    |dx|_inf with |x|_inf, f_rtol compares |F|_inf with the first |F|_inf; unset entries are inf.  Same number of
    Newton iterations and the same field as SciPy with the same keywords."""
    from scipy.optimize import newton_krylov as scipy_nk

    N = 64
    o = SHOracle(N=N)
    F = jf.SHResidual(N=N, buffers=buffers)
    Uo = seeded_state(N)
    o.set_prev(Uo)
    F.set_prev(Uo)
    hist = []
    ref = scipy_nk(o.residual, Uo, callback=lambda x, f: hist.append(np.abs(f).max()), **kw)
    got = jf.newton_krylov(F, Uo, **kw)
    assert F.last_history["nit"] == len(hist)
    assert rel(got, ref) < 1e-8


def test_linearised_step_reports_an_unconverged_solve(buffers):
    """sh_linearised.py:57 solves exactly (spsolve); the matrix-free LGMRES replacement must not hand back an
    unconverged field silently."""
    N = 48
    S = jf.SHLinearised(N=N, d=30.0, buffers=buffers)
    U = seeded_state(N, 4)
    with pytest.raises(jf.NoConvergence):
        S.steps(U, nsteps=1, rtol=1e-13, maxiter=1)
    Un, _ = S.steps(U, nsteps=1)
    assert S.last_info["info"] == 0


def test_callback_exception_propagates_like_scipy(buffers):
    """SciPy lets an exception raised inside `callback` leave newton_krylov (_nonlin.py:240-243); ctypes would print and
    swallow it.  The host layer parks it, stops the solve and re-raises."""
    N = 64
    F = jf.SHResidual(N=N, buffers=buffers)
    Uo = seeded_state(N)
    F.set_prev(Uo)
    calls = []

    class Boom(RuntimeError):
        pass

    def cb(x, f):
        calls.append(1)
        raise Boom("from the callback")

    with pytest.raises(Boom):
        jf.newton_krylov(F, Uo, callback=cb)
    assert len(calls) == 1  # the solve stopped at once
    # the context is still usable afterwards
    assert rel(jf.newton_krylov(F, Uo), jf.newton_krylov(F, Uo)) == 0.0


@pytest.mark.parametrize("gs,tau", [("cgs-ifneeded", 0.25), ("cgs-ifneeded", 0.97), ("cgs2", 0.25), ("cgs", 0.25)])
def test_speculative_arnoldi_loop_is_bitwise_the_synchronous_one(buffers, monkeypatch, gs, tau):
    """The engine enqueues Arnoldi step j+1 before it has read the outcome of step j; the device drops it (JS_STOP) when
    step j converged, broke down or asked for a second Gram-Schmidt pass.  Same arithmetic in the same order as the loop
    that reads back after every step (JFNK_SPECULATE=0): identical bits, identical counts of residual evaluations, inner
    iterations and second passes (tau = 0.97 forces a second pass on most steps)."""
    N = 64
    U0 = seeded_state(N)
    out = {}
    monkeypatch.setenv("JFNK_CYCLE_FUSED", "0")  # (at this size the default is the one-launch cycle kernel)
    for spec in ("1", "0"):
        monkeypatch.setenv("JFNK_SPECULATE", spec)
        F = jf.SHResidual(N=N, buffers=buffers, gs=gs, gs_tau=tau)
        hist = []
        U = F.steps(U0, 3, history=hist)
        out[spec] = (U, [(h["nit"], h["nfev"], h["inner_iters"], h["reorth"]) for h in hist])
    assert np.array_equal(out["1"][0], out["0"][0])
    assert out["1"][1] == out["0"][1]
    if gs == "cgs-ifneeded" and tau > 0.9:
        assert sum(c[3] for c in out["1"][1]) > 10  # the second-pass path was exercised
    # the linear solver entry point (several outer cycles) too
    S = jf.SHLinearised(N=N, buffers=buffers, gs=gs, gs_tau=tau)
    lin = {}
    for spec in ("1", "0"):
        monkeypatch.setenv("JFNK_SPECULATE", spec)
        lin[spec] = S.steps(0.1 * U0, nsteps=2)[0]
    assert np.array_equal(lin["1"], lin["0"])


@pytest.mark.gpu
@pytest.mark.parametrize("N", [61, 64, 40, 24])
@pytest.mark.parametrize("gs,tau", [("cgs-ifneeded", 0.25), ("cgs-ifneeded", 0.97), ("cgs2", 0.25), ("cgs", 0.25)])
def test_one_launch_cycle_kernel_matches_streaming_path(monkeypatch, N, gs, tau):
    """Reference-sized grids run a whole LGMRES cycle (_gcrotmk.py:16-183 + lgmres.py:188-208) as one launch of a single
    thread-block cluster with the basis in shared memory (csrc/sh_cycle.cuh).  Same point formulas, Givens / least-squares
    functions and decisions as the streaming kernels (JFNK_CYCLE_FUSED=0); only the summation order of the dots differs, so
    the Newton / inner iteration counts agree and the fields agree far inside the 1e-8 field criterion."""
    U0 = seeded_state(N)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JFNK_CYCLE_FUSED", fused)
        F = jf.SHResidual(N=N, gs=gs, gs_tau=tau)
        hist = []
        U = F.steps(U0, 6, history=hist)
        out[fused] = (U, [h["nit"] for h in hist], [h["inner_iters"] for h in hist], sum(h["reorth"] for h in hist),
                      F.context().launches())
    assert out["1"][1] == out["0"][1]
    assert all(abs(a - b) <= 2 for a, b in zip(out["1"][2], out["0"][2]))
    assert rel(out["1"][0], out["0"][0]) < 1e-9
    assert out["1"][4] < out["0"][4] / 4  # one launch per cycle instead of three per Arnoldi step
    if gs == "cgs2" or tau > 0.9:
        assert out["1"][3] > 10  # the second-pass path of the cycle kernel was exercised
    # against the SciPy oracle: the same bar as the streaming path
    o = SHOracle(N=N)
    href = []
    Ur = o.run(U0, 6, history=href)
    assert rel(out["1"][0], Ur) < 1e-8
    assert out["1"][1] == [len(h["iters"]) for h in href]
    # the linearly-implicit operator through the same kernel (several outer cycles per solve, rtol 1e-13)
    lin = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JFNK_CYCLE_FUSED", fused)
        S = jf.SHLinearised(N=N, gs=gs, gs_tau=tau)
        lin[fused] = S.steps(0.1 * U0, nsteps=3)[0]
    assert rel(lin["1"], lin["0"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("inner_m,outer_k", [(8, 3), (30, 0), (37, 10)])
def test_one_launch_cycle_kernel_other_krylov_sizes(monkeypatch, inner_m, outer_k):
    """The cycle kernel sizes its shared-memory basis from inner_m + outer_k: short restarts (many cycles per Newton step, the
    augmentation ring wrapping), no augmentation at all, and the largest basis the arena allows (inner_m + outer_k + 1 = 48)."""
    N = 64
    U0 = seeded_state(N)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JFNK_CYCLE_FUSED", fused)
        F = jf.SHResidual(N=N, inner_m=inner_m, outer_k=outer_k)
        hist = []
        U = F.steps(U0, 5, history=hist)
        out[fused] = (U, [h["nit"] for h in hist], [h["inner_iters"] for h in hist])
    assert out["1"][1] == out["0"][1]
    assert all(abs(a - b) <= 3 for a, b in zip(out["1"][2], out["0"][2]))
    assert rel(out["1"][0], out["0"][0]) < 3e-9
