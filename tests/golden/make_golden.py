"""Generate the golden vectors of tests/golden/ by running the REFERENCE's own Python modules.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference modules are imported unmodified from /root/reference/python_work behind a MagicMock
matplotlib (not installed here; the scripts only plot with it).  The shipped droplet state is loaded with
np.loadtxt because the reference's read_from_file hard-codes a Windows path separator (droplet.py:568).
Outputs (small, committed):
    pma2_n51.npz      operators, Laplace_operator, CN term, residual and 3 time steps of PMA2_nk.py (N_ = 51)
    droplet_91x61.npz the same for droplet.py from initdrop_coal_1_91-61_100_0.01_0.01_0.1_0.15.txt
    sh_n64.npz        the first 3 time steps of sh_scipy_nk.py EXECUTED UNMODIFIED (it has no import guard and would run
                      2500 steps: its newton_krylov is wrapped to record the result and stop after 3 calls; np.random is
                      seeded so that its `np.random.randn(N**2)` initial state is reproducible) and of sh_linearised.main()
                      (spsolve wrapped the same way)
    droplet_states.npz  two more of the states the reference ships (INPUTS only): initdrop_rect_1_61-61_100_0.01_... on
                      [-3,3]^2 and initdrop_coal_1_81-61_100_0.005_... on [-3,5]x[-3,3] (domains read off Q's corner values)
    droplet_init_91x61.npz  droplet.py's initialisers: initialise_coalescing_droplets (4 volume steps), initialise_droplet
                      (3 steps) followed by evolve_R_explicit (until R = 1.07), each with loop_pma relaxations
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference/python_work"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference(name):
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.animation",
                "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(mod, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return __import__(name)


def make_pma2():
    from scipy.optimize import newton_krylov

    P = _import_reference("PMA2_nk")
    N, NN = P.N_, P.NN_
    out = {}
    # --- operator-level vectors on a deformed mesh and a non-trivial state --------------------------------
    P.make_Ibdy()
    P.make_M()
    xi, eta = P.ksiksi.reshape(NN), P.etaeta.reshape(NN)
    P.Q.val = 0.5 * xi ** 2 + 0.5 * eta ** 2 + 0.02 * np.cos(np.pi * xi) * np.cos(np.pi * eta)
    P.U.val = -0.3 * np.exp(-6 * (xi ** 2 + eta ** 2)) * (1 + 0.3 * xi)
    P.U.new = P.U.val.copy()
    P.compute_Q_spatial_ders()
    P.J = P.Q.d2ksi * P.Q.d2eta - P.Q.dksideta ** 2
    P.compute_u_spatial_ders()
    P.CN_term = P.compute_rhs_pde()
    rng = np.random.default_rng(7)
    u_test = P.U.val + 1e-3 * rng.standard_normal(NN)
    vxx, vyy = P.Laplace_operator(u_test.reshape(N, N), P.M.dksiCentre.dot(u_test), P.M.detaCentre.dot(u_test))
    out.update(op_Q=P.Q.val.copy(), op_Uval=P.U.val.copy(), op_u=u_test, op_d2ksi=P.Q.d2ksi.copy(), op_d2eta=P.Q.d2eta.copy(),
               op_dksideta=P.Q.dksideta.copy(), op_J=P.J.copy(), op_vxx=vxx, op_vyy=vyy, op_CN=P.CN_term.copy(),
               op_residual=P.residual(u_test), op_Uxx=P.U.xx.copy(), op_Uyy=P.U.yy.copy())
    hist = []
    Unew = newton_krylov(P.residual, P.U.val, verbose=0, callback=lambda x, f: hist.append(np.abs(f).max()))
    out.update(op_Unew=Unew, op_hist=np.array(hist))
    # --- the script's own run: U = 0, Q = (xi^2+eta^2)/2, three passes of main()'s loop body (:83-106) ------
    P.Q.val = np.reshape(0.5 * P.ksiksi ** 2 + 0.5 * P.etaeta ** 2, NN)
    P.U.new = np.zeros(NN, dtype=float)
    for s in range(3):
        P.U.val = P.U.new.copy()
        P.compute_Q_spatial_ders()
        P.J = P.Q.d2ksi * P.Q.d2eta - P.Q.dksideta ** 2
        P.compute_u_spatial_ders()
        dt = P.compute_g() * P.k
        P.solve_PMA()
        P.CN_term = P.compute_rhs_pde()
        P.U.new = newton_krylov(P.residual, P.U.val, verbose=0)
        P.Q.val += dt * P.Q.dt
        out[f"run_U{s}"] = P.U.new.copy()
        out[f"run_Q{s}"] = P.Q.val.copy()
    np.savez_compressed(os.path.join(HERE, "pma2_n51.npz"), **out)
    print("pma2_n51.npz", {k: v.shape for k, v in out.items() if k.startswith("op_")})


def make_droplet():
    from scipy.optimize import newton_krylov

    D = _import_reference("droplet")
    NN, Nx, Ny = D.NN_, D.Nx_, D.Ny_
    state = np.loadtxt(os.path.join(REF, "initdrop_coal_1_91-61_100_0.01_0.01_0.1_0.15.txt"))
    out = {"state_U": state[:, 0].copy(), "state_Q": state[:, 1].copy()}
    D.make_Ibdy()
    D.make_M()
    D.U.new = state[:, 0].copy()
    D.U.val = D.U.new.copy()
    D.Q.val = state[:, 1].copy()
    dt = 1e-4
    U_hist, Q_hist, F_hist = [], [], []
    for s in range(2):
        # droplet.py:371-384
        D.U.val = D.U.new.copy()
        D.compute_Q_spatial_ders()
        D.J = D.Q.d2ksi * D.Q.d2eta - D.Q.dksideta ** 2
        D.compute_u_spatial_ders()
        D.P.val = D.pressure(D.U.val, D.U.xx, D.U.yy)
        D.compute_P_spatial_ders()
        F = D.pde_rhs(D.U.val, D.U.xx, D.U.yy)
        if s == 0:
            rng = np.random.default_rng(11)
            u_test = D.U.val * (1 + 1e-3 * rng.standard_normal(NN))
            vxx, vyy = D.Laplace_operator(u_test.reshape(Ny, Nx), D.M.dksiCentre.dot(u_test), D.M.detaCentre.dot(u_test))
            out.update(op_u=u_test, op_d2ksi=D.Q.d2ksi.copy(), op_d2eta=D.Q.d2eta.copy(), op_dksideta=D.Q.dksideta.copy(),
                       op_J=D.J.copy(), op_vxx=vxx, op_vyy=vyy, op_Uxx=D.U.xx.copy(), op_Uyy=D.U.yy.copy(),
                       op_Pval=D.P.val.copy(), op_F=F.copy(), op_residual=D.residual(u_test, F, dt),
                       op_PI=D.PI(D.U.val))
        hist = []
        D.U.new = newton_krylov(lambda u: D.residual(u, F, dt), D.U.val, verbose=0, maxiter=20, f_tol=1e-7,
                                callback=lambda x, f: hist.append(np.abs(f).max()))
        D.loop_pma(3e-9, 400)
        out[f"run_U{s}"] = D.U.new.copy()
        out[f"run_Q{s}"] = D.Q.val.copy()
        out[f"run_hist{s}"] = np.array(hist)
    np.savez_compressed(os.path.join(HERE, "droplet_91x61.npz"), **out)
    print("droplet_91x61.npz", {k: v.shape for k, v in out.items() if k.startswith("op_")})


def make_droplet_init():
    import contextlib
    import io

    D = _import_reference("droplet")
    NN = D.NN_
    D.make_Ibdy()
    D.make_M()
    out = {}
    quiet = io.StringIO()
    # initialise_coalescing_droplets(Vsteps, info, dtmesh, loops, fromfile, tofile) from main()'s start state (:102-105)
    D.Q.val = np.reshape(0.5 * D.ksiksi ** 2 + 0.5 * D.etaeta ** 2, NN)
    D.U.new = np.full(NN, D.epsilon_)
    info = [[0, 0, 1, 1], [3, 0, 1, 1]]
    with contextlib.redirect_stdout(quiet):
        D.initialise_coalescing_droplets(4, info, 5e-9, 20, False, False)
    out.update(coal_info=np.array(info, dtype=float), coal_U=D.U.val.copy(), coal_Q=D.Q.val.copy())
    # initialise_droplet(Vsteps, dtmesh, loops, fromfile, tofile) then evolve_R_explicit(pmaloops, Rfinal, tol)
    D.Q.val = np.reshape(0.5 * D.ksiksi ** 2 + 0.5 * D.etaeta ** 2, NN)
    D.U.new = np.full(NN, D.epsilon_)
    D.V_, D.R_ = 0, 1
    with contextlib.redirect_stdout(quiet):
        D.initialise_droplet(3, 5e-9, 20, False, False)
    out.update(rect_U=D.U.val.copy(), rect_Q=D.Q.val.copy())
    D.U.new = D.U.val.copy()
    with contextlib.redirect_stdout(quiet):
        D.evolve_R_explicit(5, 1.07, 1e-2)
    out.update(evolveR_U=D.U.val.copy(), evolveR_Q=D.Q.val.copy(), evolveR_R=np.array(float(D.R_)),
               evolveR_args=np.array([5, 1.07, 1e-2, D.dtR_, D.dtmesh_]))
    D.R_ = 1
    np.savez_compressed(os.path.join(HERE, "droplet_init_91x61.npz"), **out)
    print("droplet_init_91x61.npz", {k: np.shape(v) for k, v in out.items()}, "R after evolve_R_explicit:", float(out["evolveR_R"]))


def make_droplet_states():
    out = {}
    for key, name in (("rect61", "initdrop_rect_1_61-61_100_0.01_0.01_0.1_0.15.txt"),
                      ("coal81", "initdrop_coal_1_81-61_100_0.005_0.01_0.1_0.15.txt")):
        a = np.loadtxt(os.path.join(REF, name))
        out[key + "_U"], out[key + "_Q"] = a[:, 0].copy(), a[:, 1].copy()
    np.savez_compressed(os.path.join(HERE, "droplet_states.npz"), **out)
    print("droplet_states.npz", {k: v.shape for k, v in out.items()})


class _Stop(Exception):
    pass


def make_sh():
    import contextlib
    import io
    import runpy

    import scipy.optimize
    import scipy.sparse.linalg

    _import_reference("math")  # installs the matplotlib mocks and the reference path
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]  # `import matplotlib.pyplot as plt` binds the attribute
    sys.modules["matplotlib.pyplot"].subplots.return_value = (MagicMock(), MagicMock())  # fig, ax = plt.subplots(...)
    out = {}
    # --- sh_scipy_nk.py: module-level time loop, U = newton_krylov(residual, Uo, verbose=1) (:53-61) ----------------
    real_nk = scipy.optimize.newton_krylov
    rec = {"U": [], "hist": []}

    def nk(F, x0, **kw):
        if not rec["U"]:
            rec["U0"] = np.array(x0, copy=True)
        hist = []
        U = real_nk(F, x0, callback=lambda x, f: hist.append(np.abs(f).max()), **kw)
        rec["U"].append(np.array(U, copy=True))
        rec["hist"].append(np.array(hist))
        if len(rec["U"]) == 3:
            raise _Stop()
        return U

    scipy.optimize.newton_krylov = nk
    np.random.seed(1234)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(os.path.join(REF, "sh_scipy_nk.py"))
    except _Stop:
        pass
    finally:
        scipy.optimize.newton_krylov = real_nk
    np.random.seed(1234)
    assert np.array_equal(rec["U0"], np.random.randn(64 ** 2))
    out.update(nk_U0=rec["U0"], nk_U1=rec["U"][0], nk_U2=rec["U"][1], nk_U3=rec["U"][2],
               nk_hist1=rec["hist"][0], nk_hist2=rec["hist"][1], nk_hist3=rec["hist"][2])
    # --- sh_linearised.main(): U = spsolve(I + D - L k/2, (I + L k/2) Uo) (:51-57) ------------------------------------
    real_sp = scipy.sparse.linalg.spsolve
    lin = []

    def sp(A, b, *a, **kw):
        U = real_sp(A, b, *a, **kw)
        lin.append(np.array(U, copy=True))
        if len(lin) == 3:
            raise _Stop()
        return U

    scipy.sparse.linalg.spsolve = sp
    np.random.seed(4321)
    try:
        SL = _import_reference("sh_linearised")
        SL.main()
    except _Stop:
        pass
    finally:
        scipy.sparse.linalg.spsolve = real_sp
    np.random.seed(4321)
    out.update(lin_U0=np.random.randn(64 ** 2), lin_U1=lin[0], lin_U2=lin[1], lin_U3=lin[2])
    np.savez_compressed(os.path.join(HERE, "sh_n64.npz"), **out)
    print("sh_n64.npz", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    make_sh()
    make_pma2()
    make_droplet()
    make_droplet_init()
    make_droplet_states()
