"""Moving-mesh CPU oracle for the PMA2 (MEMS) and droplet (thin-film) models -- TEST INFRASTRUCTURE,
never imported by the product.

NumPy / scipy.sparse restatement of the reference's operators, class-based instead of module globals so it can
run on the GPU box where /root/reference does not exist.  Each piece cites the reference lines it follows; the
restatement is pinned by tests/golden/*.npz, generated in the build container by importing the reference's own
modules (tests/golden/make_golden.py) and compared in tests/test_oracle.py.

  MeshOps            make_Ibdy / make_M          PMA2_nk.py:165-233,  droplet.py:762-833
  MeshOps.q_ders     compute_Q_spatial_ders + J  PMA2_nk.py:235-248,87  droplet.py:696-711,376
  MeshOps.laplace    Laplace_operator            PMA2_nk.py:263-343,  droplet.py:601-681
  MeshOps.monitor / solve_pma                    PMA2_nk.py:345-403,  droplet.py:578-588,729-760
  PMA2Oracle         residual / rhs / time loop  PMA2_nk.py:80-106,121-159,405-419
  DropletOracle      residual / pressure / loop  droplet.py:360-411,435-473,590-599,713-727

The nonlinear solver is SciPy's own newton_krylov (the reference's dependency), called as the scripts call it.
"""
from __future__ import annotations

import numpy as np
from scipy.fft import dct, idct
from scipy.optimize import newton_krylov
from scipy.sparse import csc_matrix, diags, kron


def _d2_matrix(n, h2):
    """1-D second derivative, 4th order, one-sided closures (droplet.py:784-790 / PMA2_nk.py:187-192)."""
    t = diags([-1, 16, -30, 16, -1], [-2, -1, 0, 1, 2], shape=(n, n), format="lil")
    t[0, :5] = [-415 / 6, 96, -36, 32 / 3, -1.5]
    t[1, :6] = [10, -15, -4, 14, -6, 1]
    t[-1, -5:] = [-1.5, 32 / 3, -36, 96, -415 / 6]
    t[-2, -6:] = [1, -6, 14, -4, -15, 10]
    return csc_matrix(t / (12 * h2))


def _d1_matrix(n, h):
    """1-D first derivative, 4th order, one-sided closures (droplet.py:800-804 / PMA2_nk.py:197-200)."""
    t = diags([1, -8, 8, -1], [-2, -1, 1, 2], shape=(n, n), format="lil")
    t[:2, :5] = [[-25, 48, -36, 16, -3], [-3, -10, 18, -6, 1]]
    t[-2:, -5:] = [[-1, 6, -18, 10, 3], [3, -16, 36, -48, 25]]
    return csc_matrix(t / (12 * h))


def _cons_diff_rows(A, v, h2):
    """(A v_s)_s along the LAST axis with the conservative 4th-order interior formula and the special formulas
    next to the boundary; first / last entries stay 0 (droplet.py:619-668 for the ksi direction; the eta
    direction is the same formula applied to the transposed arrays)."""
    out = np.zeros_like(v)
    out[:, 3:-3] = (4 * np.multiply(A[:, 2:-4], (v[:, :-6] - 8 * v[:, 1:-5] + 8 * v[:, 3:-3] - v[:, 4:-2]))
                    - np.multiply((-A[:, 1:-5] + 9 * A[:, 2:-4] + 9 * A[:, 3:-3] - A[:, 4:-2]),
                                  (v[:, 1:-5] - 27 * v[:, 2:-4] + 27 * v[:, 3:-3] - v[:, 4:-2]))
                    + np.multiply((-A[:, 2:-4] + 9 * A[:, 3:-3] + 9 * A[:, 4:-2] - A[:, 5:-1]),
                                  (v[:, 2:-4] - 27 * v[:, 3:-3] + 27 * v[:, 4:-2] - v[:, 5:-1]))
                    - 4 * np.multiply(A[:, 4:-2], (v[:, 2:-4] - 8 * v[:, 3:-3] + 8 * v[:, 5:-1] - v[:, 6:]))) / (288 * h2)
    out[:, 1] = np.multiply(A[:, 1], (10 * v[:, 0] - 15 * v[:, 1] - 4 * v[:, 2] + 14 * v[:, 3] - 6 * v[:, 4] + v[:, 5])) / (12 * h2) \
        + np.multiply((-3 * v[:, 0] - 10 * v[:, 1] + 18 * v[:, 2] - 6 * v[:, 3] + v[:, 4]),
                      (-3 * A[:, 0] - 10 * A[:, 1] + 18 * A[:, 2] - 6 * A[:, 3] + A[:, 4])) / (144 * h2)
    out[:, -2] = np.multiply(A[:, -2], (10 * v[:, -1] - 15 * v[:, -2] - 4 * v[:, -3] + 14 * v[:, -4] - 6 * v[:, -5] + v[:, -6])) / (12 * h2) \
        + np.multiply((3 * v[:, -1] + 10 * v[:, -2] - 18 * v[:, -3] + 6 * v[:, -4] - v[:, -5]),
                      (3 * A[:, -1] + 10 * A[:, -2] - 18 * A[:, -3] + 6 * A[:, -4] - A[:, -5])) / (144 * h2)
    out[:, 2] = np.multiply(A[:, 2], (-v[:, 0] + 16 * v[:, 1] - 30 * v[:, 2] + 16 * v[:, 3] - v[:, 4])) / (12 * h2) \
        + np.multiply((v[:, 0] - 8 * v[:, 1] + 8 * v[:, 3] - v[:, 4]), (A[:, 0] - 8 * A[:, 1] + 8 * A[:, 3] - A[:, 4])) / (144 * h2)
    out[:, -3] = np.multiply(A[:, -3], (-v[:, -1] + 16 * v[:, -2] - 30 * v[:, -3] + 16 * v[:, -4] - v[:, -5])) / (12 * h2) \
        + np.multiply((v[:, -5] - 8 * v[:, -4] + 8 * v[:, -2] - v[:, -1]), (A[:, -5] - 8 * A[:, -4] + 8 * A[:, -2] - A[:, -1])) / (144 * h2)
    return out


class MeshOps:
    """Derivative matrices, index sets and the mesh-dependent operators on an Nx x Ny computational grid."""

    def __init__(self, Nx, Ny, endl, endr, endb, endt, leig_scale=None):
        self.Nx, self.Ny, self.NN = Nx, Ny, Nx * Ny
        self.endl, self.endr, self.endb, self.endt = endl, endr, endb, endt
        self.dksi = (endr - endl) / (Nx - 1)
        self.deta = (endt - endb) / (Ny - 1)
        self.dksi2, self.deta2 = self.dksi * self.dksi, self.deta * self.deta
        self.ksiksi, self.etaeta = np.meshgrid(np.linspace(endl, endr, Nx), np.linspace(endb, endt, Ny))
        # index sets (make_Ibdy)
        X, Y = self.ksiksi.reshape(-1), self.etaeta.reshape(-1)
        self.Left, self.Right = np.nonzero(X == endl)[0], np.nonzero(X == endr)[0]
        self.Bottom, self.Top = np.nonzero(Y == endb)[0], np.nonzero(Y == endt)[0]
        self.Boundary = np.nonzero((X == endr) | (X == endl) | (Y == endt) | (Y == endb))[0]
        # derivative matrices (make_M)
        eyeX, eyeY = diags([1], shape=(Nx, Nx)), diags([1], shape=(Ny, Ny))
        d1x, d1y = _d1_matrix(Nx, self.dksi), _d1_matrix(Ny, self.deta)
        self.M_d2ksi = kron(eyeY, _d2_matrix(Nx, self.dksi2))
        self.M_d2eta = kron(_d2_matrix(Ny, self.deta2), eyeX)
        self.M_dksi = kron(eyeY, d1x)
        self.M_deta = kron(d1y, eyeX)
        self.M_dksideta = kron(d1y, d1x)
        # eigenvalues used by the DCT solve; the divisor is dksi*deta (droplet.py:831-833; = dksi^2 for PMA2_nk.py:231-233)
        lam = (2 * np.cos(np.pi * np.arange(0, Ny) / (Ny - 1)) - 2).reshape((Ny, 1)) * np.ones(Nx) \
            + np.ones((Ny, 1)) * (2 * np.cos(np.pi * np.arange(0, Nx) / (Nx - 1)) - 2)
        self.Leig = lam / (self.dksi * self.deta)

    # compute_Q_spatial_ders + J
    def q_ders(self, Q):
        m = {}
        m["dksi"] = self.M_dksi.dot(Q)
        m["deta"] = self.M_deta.dot(Q)
        m["dksi"][self.Left] = self.endl; m["dksi"][self.Right] = self.endr
        m["deta"][self.Bottom] = self.endb; m["deta"][self.Top] = self.endt
        t = np.zeros(self.NN); t[self.Left] = 25 / (6 * self.dksi) * abs(self.endl); t[self.Right] = 25 / (6 * self.dksi) * abs(self.endr)
        m["d2ksi"] = self.M_d2ksi.dot(Q) + t
        t = np.zeros(self.NN); t[self.Top] = 25 / (6 * self.deta) * abs(self.endt); t[self.Bottom] = 25 / (6 * self.deta) * abs(self.endb)
        m["d2eta"] = self.M_d2eta.dot(Q) + t
        m["dksideta"] = self.M_dksideta.dot(Q); m["dksideta"][self.Boundary] = 0
        m["J"] = m["d2ksi"] * m["d2eta"] - m["dksideta"] ** 2
        return m

    # Laplace_operator(v, v_dksi, v_deta)
    def laplace(self, v, v_dksi, v_deta, m):
        Ny, Nx = self.Ny, self.Nx
        J = m["J"]
        A11 = np.reshape(np.divide(m["dksideta"] ** 2 + m["d2eta"] ** 2, J), (Ny, Nx))
        A22 = np.reshape(np.divide(m["dksideta"] ** 2 + m["d2ksi"] ** 2, J), (Ny, Nx))
        A12 = -np.divide(np.multiply(m["dksideta"], m["d2ksi"] + m["d2eta"]), J)
        v2 = np.reshape(v, (Ny, Nx))
        v_xx = _cons_diff_rows(A11, v2, self.dksi2)
        v_yy = _cons_diff_rows(np.ascontiguousarray(A22.T), np.ascontiguousarray(v2.T), self.deta2).T
        t = self.M_dksi.dot(np.multiply(A12, v_deta)); t[self.Left] = 0; t[self.Right] = 0
        v_xx = np.reshape(v_xx, self.NN) + t
        t = self.M_deta.dot(np.multiply(A12, v_dksi)); t[self.Top] = 0; t[self.Bottom] = 0
        v_yy = np.reshape(np.ascontiguousarray(v_yy), self.NN) + t
        return np.divide(v_xx, J), np.divide(v_yy, J)

    def laplace_of(self, v, m):
        """Laplace_operator fed with the plain centred derivatives, as both residuals call it."""
        return self.laplace(v, self.M_dksi.dot(v), self.M_deta.dot(v), m)

    # fourth-order smoothing filter of the monitor function (droplet.py:738-754 / PMA2_nk.py:371-386)
    def smooth(self, temp, iters):
        Ny, Nx = self.Ny, self.Nx
        mon = np.zeros((Ny, Nx))
        for _ in range(iters):
            mon[1:-1, 1:-1] = temp[1:-1, 1:-1] + (temp[:-2, 1:-1] + temp[2:, 1:-1] + temp[1:-1, :-2] + temp[1:-1, 2:]) / 8 \
                + (temp[:-2, :-2] + temp[:-2, 2:] + temp[2:, :-2] + temp[2:, 2:]) / 16
            mon[1:-1, Nx - 1] = (4 * temp[1:-1, Nx - 1] + 2 * temp[:-2, Nx - 1] + 2 * temp[2:, Nx - 1] + 2 * temp[1:-1, Nx - 2] + temp[2:, Nx - 2] + temp[:-2, Nx - 2]) / 12
            mon[1:-1, 0] = (4 * temp[1:-1, 0] + 2 * temp[:-2, 0] + 2 * temp[2:, 0] + 2 * temp[1:-1, 1] + temp[2:, 1] + temp[:-2, 1]) / 12
            mon[Ny - 1, 1:-1] = (4 * temp[Ny - 1, 1:-1] + 2 * temp[Ny - 1, :-2] + 2 * temp[Ny - 1, 2:] + 2 * temp[Ny - 2, 1:-1] + temp[Ny - 2, 2:] + temp[Ny - 2, :-2]) / 12
            mon[0, 1:-1] = (4 * temp[0, 1:-1] + 2 * temp[0, :-2] + 2 * temp[0, 2:] + 2 * temp[1, 1:-1] + temp[1, 2:] + temp[1, :-2]) / 12
            mon[0, 0] = (4 * temp[0, 0] + 2 * temp[0, 1] + 2 * temp[1, 0] + temp[1, 1]) / 9
            mon[0, Nx - 1] = (4 * temp[0, Nx - 1] + 2 * temp[0, Nx - 2] + 2 * temp[1, Nx - 1] + temp[1, Nx - 2]) / 9
            mon[Ny - 1, 0] = (4 * temp[Ny - 1, 0] + 2 * temp[Ny - 1, 1] + 2 * temp[Ny - 2, 0] + temp[Ny - 2, 1]) / 9
            mon[Ny - 1, Nx - 1] = (4 * temp[Ny - 1, Nx - 1] + 2 * temp[Ny - 1, Nx - 2] + 2 * temp[Ny - 2, Nx - 1] + temp[Ny - 2, Nx - 2]) / 9
            temp = mon.copy()
        return np.reshape(mon, self.NN)

    # solve_PMA: dQ/dt = (alpha (I - gamma Lap_ksi))^-1 sqrt(M |J|) by a 2-D DCT-II (droplet.py:578-588)
    def solve_pma(self, monitor, J, alpha, gamma):
        q_rhs = np.sqrt(np.multiply(monitor, np.abs(J))) / alpha
        temp = dct(dct(q_rhs.reshape(self.Ny, self.Nx).T, norm="ortho").T, norm="ortho")
        dQdt = idct(idct(np.divide(temp, (1 - gamma * self.Leig)).T, norm="ortho").T, norm="ortho")
        return dQdt.reshape(self.NN)


class PMA2Oracle:
    """PMA2_nk.py with p_ = 2 (the only working branch, :134-139) and its module-level constants :23-40."""

    def __init__(self, N=51, lambd=1.0, beta=0.15, epsilon=0.0, m=3, k=1e-4, alpha=0.1, gamma=0.1, smoothing_iters=4):
        self.N = N
        self.ops = MeshOps(N, N, -1, 1, -1, 1)
        self.lambd, self.beta, self.epsilon, self.m, self.k = lambd, beta, epsilon, m, k
        self.alpha, self.gamma, self.smoothing_iters = alpha, gamma, smoothing_iters
        self.dt = k  # the script's residual always sees the module-level dt = k (PMA2_nk.py:51,66,91,159)
        self.Q = np.reshape(0.5 * self.ops.ksiksi ** 2 + 0.5 * self.ops.etaeta ** 2, N * N)  # :68
        self.met = None
        self.Uval = None
        self.CN = None
        self.nfev = 0

    def set_mesh(self, Q):
        self.Q = np.array(Q, dtype=float, copy=True)
        self.met = self.ops.q_ders(self.Q)

    def _simple_rhs(self, u):
        return -self.lambd / ((1 + u) ** 2) + self.lambd * (self.epsilon ** (self.m - 2)) / ((1 + u) ** self.m)

    def rhs(self, u):
        """compute_rhs_pde (:405-419) == the new_rhs part of residual (:131-157)."""
        o = self.ops
        out = self._simple_rhs(u)
        u_xx, u_yy = o.laplace_of(u, self.met)
        v = u_xx + u_yy
        v_xx, v_yy = o.laplace_of(v, self.met)
        out -= self.beta * self.beta * (v_xx + v_yy)
        out[o.Boundary] = 0
        return out

    def set_prev(self, Uval):
        self.Uval = np.array(Uval, dtype=float, copy=True)
        self.CN = self.rhs(self.Uval)

    def residual(self, u):
        self.nfev += 1
        return (u - self.Uval) / self.dt - (self.rhs(u) + self.CN) / 2

    def monitor(self):
        """compute_and_smooth_monitor for epsilon == 0 (:361-362) or the p = 2 arc-length-like monitor (:367)."""
        o = self.ops
        if self.epsilon == 0:
            temp = (1 / (1 + self.Uval) ** 6).reshape((self.N, self.N))
        else:
            u_xx, u_yy = o.laplace_of(self.Uval, self.met)
            temp = (np.abs(u_xx + u_yy) ** 2).reshape((self.N, self.N))
        mon = o.smooth(temp, self.smoothing_iters)
        mon += np.sum(mon * np.abs(self.met["J"])) * o.dksi2
        return mon

    def step(self, U, history=None, **kw):
        """one pass of the while-loop body (:83-106); returns U.new and advances the mesh"""
        self.set_mesh(self.Q)
        self.Uval = np.array(U, dtype=float, copy=True)
        g = min((1 + self.Uval) ** 3) if self.epsilon == 0 else 1
        dt_mesh = g * self.k  # the *local* dt of :91
        Qdt = self.ops.solve_pma(self.monitor(), self.met["J"], self.alpha, self.gamma)
        self.CN = self.rhs(self.Uval)
        hist = []
        n0 = self.nfev
        Unew = newton_krylov(self.residual, self.Uval, verbose=0,
                             callback=lambda x, f: hist.append((float(np.abs(f).max()), float(np.linalg.norm(f)))), **kw)
        if history is not None:
            history.append({"iters": hist, "nfev": self.nfev - n0})
        self.Q = self.Q + dt_mesh * Qdt
        return Unew


class DropletOracle:
    """droplet.py's evolve_with_PDE path (:360-411) with the module constants of :23-53."""

    def __init__(self, Nx=91, Ny=61, endl=-3, endr=6, endb=-3, endt=3, epsilon=1e-2, n=6, m=3, Bo=0.01, alpha2=0.0,
                 alpha=0.01, gamma=0.1, C=0.15, smoothing_iters=4):
        self.ops = MeshOps(Nx, Ny, endl, endr, endb, endt)
        self.Nx, self.Ny = Nx, Ny
        self.epsilon, self.n, self.m, self.Bo, self.alpha2 = epsilon, n, m, Bo, alpha2
        self.epsilon2 = 1 / (endt - endb)
        self.alpha, self.gamma, self.C, self.smoothing_iters = alpha, gamma, C, smoothing_iters
        self.Q = None
        self.met = None
        self.Uval = None
        self.Uxx = self.Uyy = None
        self.F = None
        self.dt = None
        self.nfev = 0

    def PI(self, h):
        e, n, m = self.epsilon, self.n, self.m
        return (n - 1) * (m - 1) * ((np.divide(e, h) ** m) - (np.divide(e, h) ** n)) / (2 * e * (n - m))

    def pressure(self, h, hxx, hyy):
        return -(hxx + hyy) + self.PI(h) + self.Bo * np.cos(self.alpha2) * h

    def set_mesh(self, Q):
        self.Q = np.array(Q, dtype=float, copy=True)
        self.met = self.ops.q_ders(self.Q)

    def _u_ders(self):
        """compute_u_spatial_ders (:713-727) including its `U_dksi[Ibdy.Bottom] = 0` line (:722)."""
        o = self.ops
        U_dksi = o.M_dksi.dot(self.Uval)
        U_deta = o.M_deta.dot(self.Uval)
        U_dksi[o.Left] = 0; U_dksi[o.Right] = 0
        U_deta[o.Top] = 0; U_dksi[o.Bottom] = 0
        self.Uxx, self.Uyy = o.laplace(self.Uval, U_dksi, U_deta, self.met)

    def _flux_div(self, p, h):
        """(P_x, P_y) with zero normal derivative at the walls, the flux (A, B) and its divergence
        (compute_P_spatial_ders :683-694 + pde_rhs :452-460; identical arithmetic in residual :438-449)."""
        o, m = self.ops, self.met
        P_dksi = o.M_dksi.dot(p); P_deta = o.M_deta.dot(p)
        P_dksi[o.Left] = 0; P_dksi[o.Right] = 0
        P_deta[o.Top] = 0; P_deta[o.Bottom] = 0
        pdx = np.divide(np.multiply(m["d2eta"], P_dksi) - np.multiply(m["dksideta"], P_deta), m["J"])
        pdy = np.divide(-np.multiply(m["dksideta"], P_dksi) + np.multiply(m["d2ksi"], P_deta), m["J"])
        A = (pdx - self.Bo * np.sin(self.alpha2) / self.epsilon2) * (h ** 3) / 3
        B = pdy * (h ** 3) / 3
        return np.divide(m["d2eta"] * o.M_dksi.dot(A) - m["dksideta"] * o.M_deta.dot(A)
                         - m["dksideta"] * o.M_dksi.dot(B) + m["d2ksi"] * o.M_deta.dot(B), m["J"])

    def set_prev(self, Uval, dt):
        """:373-381: U.val, derivatives, P.val, F = pde_rhs"""
        self.Uval = np.array(Uval, dtype=float, copy=True)
        self.dt = dt
        self._u_ders()
        Pval = self.pressure(self.Uval, self.Uxx, self.Uyy)
        self.F = self._flux_div(Pval, self.Uval)

    def residual(self, u):
        """:435-450"""
        self.nfev += 1
        u_xx, u_yy = self.ops.laplace_of(u, self.met)
        pnew = self.pressure(u, u_xx, u_yy)
        F2 = self._flux_div(pnew, u)
        return (u - self.Uval) - self.dt * (F2 + self.F) / 2

    def monitor(self):
        """compute_and_smooth_monitor (:729-760)"""
        o = self.ops
        temp = (np.abs(self.Uxx + self.Uyy) ** 2).reshape((self.Ny, self.Nx))
        mon = o.smooth(temp, self.smoothing_iters)
        mon_integral = np.sum(mon * np.abs(self.met["J"])) * o.dksi * o.deta
        mon += self.C * mon_integral
        return mon

    def loop_pma(self, dt, loops):
        """:590-599 (Uval is the *old* solution during the mesh relaxation, as in the script)"""
        o = self.ops
        self.Q = self.Q + dt * o.solve_pma(self.monitor(), self.met["J"], self.alpha, self.gamma)
        for _ in range(1, loops):
            self.met = o.q_ders(self.Q)
            self._u_ders()
            self.Q = self.Q + dt * o.solve_pma(self.monitor(), self.met["J"], self.alpha, self.gamma)

    # ---- initial states (droplet.py:132-248, :413-429, :544-554) ---------------------------------------------
    def compute_U2(self, info, a=100):
        """:413-423 with G2 (:425-426) and H2 (:428-429) on the current mesh (self.met)"""
        e = self.epsilon
        ret = np.full(self.ops.NN, e, dtype=float)
        for x, y, R, V in info:
            xx = np.sqrt((self.met["dksi"] - x) ** 2 + (self.met["deta"] - y) ** 2)
            psi = R + np.log((1 + np.exp(-2 * a * (xx + R))) / (1 + np.exp(-2 * a * (xx - R)))) / (2 * a)
            ret += (1 - e) * (4 * V * (1 - psi * psi / (R * R)) / (R * R))
        return ret

    def initialise_coalescing_droplets(self, Vsteps, info, dtmesh, loops, a=100):
        """:132-189 without the file / plot branches: inflate the droplets in Vsteps volume increments, relaxing the
        mesh (loop_pma) after each.  Starts from self.Q and U = epsilon (main(), :102-105).  Returns U."""
        Unew = np.full(self.ops.NN, self.epsilon, dtype=float)
        for i in range(1, Vsteps + 1):
            self.Uval = Unew.copy()
            self.met = self.ops.q_ders(self.Q)
            self._u_ders()
            arg = [[d[0], d[1], d[2], d[3] * i / Vsteps] for d in info]
            Unew = self.compute_U2(arg, a)
            self.loop_pma(dtmesh, loops)
        return Unew

    def evolve_R_explicit(self, U, pmaloops, Rfinal, tol, R=1.0, V=1.0, dtR=5e-2, dtmesh=1e-7, a=100, max_iters=None):
        """:316-358: explicit evolution of the droplet radius R' = Rdot() (:553-554) with U = compute_U() (:544-551);
        loop_pma's first pass still sees the Laplacian of the previous U.val (the script refreshes U.val after
        compute_u_spatial_ders).  Returns (U, R, time)."""
        self.Uval = np.array(U, dtype=float, copy=True)
        time, it = 0.0, 0
        while abs(Rfinal - R) > tol and (max_iters is None or it < max_iters):
            self.met = self.ops.q_ders(self.Q)
            self._u_ders()
            dt = dtR * (R ** 2)
            R += dt * (8 * V / R ** 3 - 1) / (3 * np.log(1 / self.epsilon))
            self.Uval = self.compute_U2([[0.0, 0.0, R, V]], a)
            self.loop_pma(dtmesh, pmaloops)
            it += 1
            time += dt
        return self.Uval, R, time

    def step(self, U, dt_n, dtmesh=3e-9, pmaloops=400, history=None, **kw):
        """one pass of evolve_with_PDE's loop body (:371-384); returns U.new"""
        self.set_mesh(self.Q)
        self.set_prev(U, dt_n)
        hist = []
        n0 = self.nfev
        kw.setdefault("maxiter", 20)
        kw.setdefault("f_tol", 1e-7)
        Unew = newton_krylov(self.residual, self.Uval, verbose=0,
                             callback=lambda x, f: hist.append((float(np.abs(f).max()), float(np.linalg.norm(f)))), **kw)
        if history is not None:
            history.append({"iters": hist, "nfev": self.nfev - n0})
        if pmaloops:
            self.loop_pma(dtmesh, pmaloops)
        return Unew
