"""Swift-Hohenberg CPU oracle (TEST INFRASTRUCTURE -- never imported by the product).

Restates the reference's SciPy path for the Swift-Hohenberg scripts:

* operator build ``Lap``, ``L``        -- python_work/sh_scipy_nk.py:31-39
  (same in sh_vscode_nk.py:31-39, sh_linearised.py:31-38, cpp main.cpp:39-81)
* Crank-Nicolson ``residual(u)``       -- python_work/sh_scipy_nk.py:47-49
* implicit time loop                   -- python_work/sh_scipy_nk.py:53-61
* linearly-implicit step               -- python_work/sh_linearised.py:51-57

The nonlinear / linear solvers are NOT restated: the oracle calls the reference's own
third-party dependency, SciPy 1.18.1 (``scipy.optimize.newton_krylov`` ->
``scipy/optimize/_nonlin.py:134`` nonlin_solve, ``:1375`` KrylovJacobian,
``scipy/sparse/linalg/_isolve/lgmres.py:17``, ``_gcrotmk.py:16`` _fgmres).

The sparse matrices are built with the *same scipy.sparse constructors in the same
order* as the reference so that CSR storage order -- and therefore the floating-point
summation order of ``L @ u`` -- is the reference's.

Pinned: tests/golden/sh_n64.npz holds the first three time steps of sh_scipy_nk.py and sh_linearised.py EXECUTED
UNMODIFIED in the build container (tests/golden/make_golden.py: legacy RNG seeded, the solver call wrapped to record
and stop); tests/test_oracle.py shows this restatement reproduces their fields and per-iteration Newton norms bit
for bit.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import diags, block_diag
from scipy.sparse import linalg as sla
from scipy.optimize import newton_krylov
from scipy.optimize import NoConvergence  # noqa: F401  (re-exported for tests)


def build_lap(N: int, h: float):
    """Periodic 5-point Laplacian on an N x N grid (sh_scipy_nk.py:32-35)."""
    e = 1 / h**2
    nn = N**2
    A = np.tile(diags([e, e, -4 * e, e, e], [1 - N, -1, 0, 1, N - 1], shape=(N, N), format="csc"), N)
    Lap = block_diag(A[:]) + diags([e, e, e, e], [N - nn, -N, N, nn - N], shape=(nn, nn), format="csc")
    return Lap


def build_L(N: int, h: float, r: float):
    """L = -Lap*Lap - 2*Lap + (r-1)*I  (sh_scipy_nk.py:38-39)."""
    nn = N**2
    Lap = build_lap(N, h)
    I = diags([1], [0], shape=(nn, nn), format="csc")
    L = -Lap * Lap - 2 * Lap + (r - 1) * I
    return Lap, L


def sh13_coefficients(h: float, r: float):
    """Analytic 13-point coefficients of L: centre, axial+-1, diagonal, axial+-2."""
    e = 1 / h**2
    return (-20 * e * e + 8 * e + (r - 1), 8 * e * e - 2 * e, -2 * e * e, -e * e)


class SHOracle:
    """The reference's Swift-Hohenberg CN/JFNK path with a seeded, injectable state.

    Parameters follow sh_scipy_nk.py:15-29 (d, N, k, r, g); ``h = d/N``.
    """

    def __init__(self, N=64, d=40.0, k=0.2, r=0.01, g=1.0):
        self.N, self.d, self.k, self.r, self.g = N, d, k, r, g
        self.h = d / N
        self.Lap, self.L = build_L(N, self.h, r)
        self.Uo = self.UoUo = self.UoUoUo = None
        self.nfev = 0

    # -- sh_scipy_nk.py:56-58
    def set_prev(self, U):
        self.Uo = np.array(U, dtype=np.float64, copy=True)
        self.UoUo = np.multiply(self.Uo, self.Uo)
        self.UoUoUo = np.multiply(self.Uo, self.UoUo)

    # -- sh_scipy_nk.py:47-49 (evaluation order kept left-to-right as NumPy does it)
    def residual(self, u):
        self.nfev += 1
        L, g, k = self.L, self.g, self.k
        uu = np.multiply(u, u)
        return (u - self.Uo) / k - (L @ u + g * uu - np.multiply(u, uu) + L @ self.Uo + g * self.UoUo - self.UoUoUo) / 2

    # -- sh_scipy_nk.py:61 with the Newton trace recorded through SciPy's own callback hook
    def step(self, U, history=None, **kw):
        self.set_prev(U)
        hist = []

        def cb(x, Fx):
            hist.append((float(np.abs(Fx).max()), float(np.linalg.norm(Fx))))

        nfev0 = self.nfev
        F0 = self.residual(self.Uo)
        self.nfev -= 1
        Unew = newton_krylov(self.residual, self.Uo, callback=cb, **kw)
        if history is not None:
            history.append({
                "f0_max": float(np.abs(F0).max()), "f0_l2": float(np.linalg.norm(F0)),
                "iters": hist, "nfev": self.nfev - nfev0})
        return Unew

    def run(self, U0, nsteps, history=None, **kw):
        U = np.array(U0, dtype=np.float64, copy=True)
        for _ in range(nsteps):
            U = self.step(U, history=history, **kw)
        return U


def seeded_state(N: int, seed: int = 1234):
    """Config-1/4 initial condition: default_rng(seed).standard_normal(N*N)."""
    return np.random.default_rng(seed).standard_normal(N * N)


class SHLinearisedOracle:
    """Linearly-implicit Swift-Hohenberg step (sh_linearised.py:16-57), SuperLU direct solve."""

    def __init__(self, N=64, d=40.0, k=0.2, r=0.2, g=0.0):
        self.N, self.d, self.k, self.r, self.g = N, d, k, r, g
        self.h = d / N
        self.nn = N * N
        self.Lap, self.L = build_L(N, self.h, r)
        self.I = diags([1], [0], shape=(self.nn, self.nn), format="csc")

    # -- sh_linearised.py:51-57.  U is U[s], Uo is U[s-1]; returns (U[s+1], U[s]).
    def step(self, U, Uo):
        k, g, nn = self.k, self.g, self.nn
        D = diags(np.multiply(5 * U - Uo, 5 * U - Uo) * k / 16 - g * k * U, 0, shape=(nn, nn), format="csc")
        Uo = U
        Unew = sla.spsolve((self.I + D - self.L * k / 2).tocsc(), np.transpose((self.I + self.L * k / 2) @ Uo))
        return Unew, Uo

    def run(self, U0, nsteps):
        U = np.array(U0, dtype=np.float64, copy=True)
        Uo = U.copy()  # sh_linearised.py:28
        for _ in range(nsteps):
            U, Uo = self.step(U, Uo)
        return U


def apply_lap_roll(u, N, h):
    """Periodic 5-point Laplacian by np.roll -- the same operator as build_lap without forming a matrix; used
    where the CSR build is too large (checked against the matrices in tests/test_oracle.py)."""
    e = 1 / h**2
    U = np.asarray(u).reshape(N, N)
    out = e * (np.roll(U, 1, 0) + np.roll(U, -1, 0) + np.roll(U, 1, 1) + np.roll(U, -1, 1)) - 4 * e * U
    return out.reshape(-1)


def apply_L_roll(u, N, h, r):
    """L u = -Lap(Lap u) - 2 Lap u + (r-1) u by two roll-Laplacians (sh_scipy_nk.py:39 without the matrix)."""
    lu = apply_lap_roll(u, N, h)
    return -apply_lap_roll(lu, N, h) - 2 * lu + (r - 1) * np.asarray(u).reshape(-1)


def apply_L_roll_rows(U2d, a, b, h, r):
    """Rows [a, b) of L u for an N x N periodic field given as a 2-D array, by the same two roll-Laplacians as
    :func:`apply_L_roll` (bit-identical to its rows a..b-1, tests/test_oracle.py) but touching only rows a-2 .. b+1:
    lets bench.py check the engine at 16384^2, where the full-field rolls would need ~30 GB of temporaries."""
    N = U2d.shape[0]
    e = 1 / h**2
    blk = U2d[np.arange(a - 2, b + 2) % N]

    def lap_rows(B):  # 5-point Laplacian of the interior rows of B (periodic in columns, rows from the block itself)
        # same operand order as apply_lap_roll: e*(up + down + left + right) - 4 e u
        return e * (B[:-2] + B[2:] + np.roll(B[1:-1], 1, 1) + np.roll(B[1:-1], -1, 1)) - 4 * e * B[1:-1]

    l1 = lap_rows(blk)          # rows a-1 .. b
    l2 = lap_rows(l1)           # rows a .. b-1
    return -l2 - 2 * l1[1:-1] + (r - 1) * blk[2:-2]


def residual_roll_rows(U2d, Uo2d, a, b, h, r, g, k):
    """Rows [a, b) of the Crank-Nicolson residual of sh_scipy_nk.py:47-49 with L applied by :func:`apply_L_roll_rows`."""
    u, uo = U2d[a:b], Uo2d[a:b]
    uu, uouo = np.multiply(u, u), np.multiply(uo, uo)
    return (u - uo) / k - (apply_L_roll_rows(U2d, a, b, h, r) + g * uu - np.multiply(u, uu)
                           + apply_L_roll_rows(Uo2d, a, b, h, r) + g * uouo - np.multiply(uo, uouo)) / 2
