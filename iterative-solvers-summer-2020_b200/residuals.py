"""Device residual handles: the objects passed as ``F`` to :func:`newton_krylov`.

Each mirrors one of the reference's ``residual`` functions together with the module-level
state it closes over (``Uo``, ``k``, ``g``, ``L`` ... / ``U.val``, ``dt``, ``CN_term``, ``Q.*``, ``J``):

* :class:`SHResidual`        -- python_work/sh_scipy_nk.py:15-49 (and sh_vscode_nk.py, cpp main.cpp:19-32)
* :class:`SHLinearised`      -- python_work/sh_linearised.py:16-57
* :class:`PMA2Residual`      -- python_work/PMA2_nk.py:121-159 (+ Laplace_operator :263-343)
* :class:`DropletResidual`   -- python_work/droplet.py:435-450 (+ Laplace_operator :601-681)

``F(u)`` evaluates the residual on the GPU; ``F.jvp(v)`` is KrylovJacobian.matvec.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .context import Context, HistoryBuffer


class _DeviceResidual:
    """Common machinery: lazily (re)creates the engine context for the requested Krylov sizes."""

    problem = None

    def __init__(self, nx, ny, *, comm=None, inner_m=30, outer_k=10, gs="cgs-ifneeded", gs_tau=0.25,
                 kernel_variant=0, buffers=None):
        # gs: Gram-Schmidt variant of the Arnoldi process.  SciPy uses modified GS (j dependent reductions per
        # step); the device path batches all dots of a step into one fused reduction (classical GS) and takes a
        # second pass only when ||w_after|| < gs_tau ||w_before|| (cancellation of more than 1/gs_tau), decided from
        # the scalars the host reads back once per Arnoldi step.  "cgs2" always re-orthogonalises, "cgs" never does.
        self.nx, self.ny = int(nx), int(ny)
        self.comm = comm
        self._krylov = (int(inner_m), int(outer_k))
        self._gs = (gs, float(gs_tau))
        self._variant = int(kernel_variant)
        self._buffers = buffers
        self._ctx = None
        self.last_history = None

    # slab owned by this rank
    @property
    def rows(self):
        if self.comm is None:
            return 0, self.ny
        return self.comm.slab(self.ny)

    @property
    def n(self):
        r0, nr = self.rows
        return nr * self.nx

    def context(self, inner_m=None, outer_k=None) -> Context:
        want = (self._krylov[0] if inner_m is None else int(inner_m), self._krylov[1] if outer_k is None else int(outer_k))
        if self._ctx is not None and want == self._krylov:
            return self._ctx
        if self._ctx is not None:
            self._ctx.close()
        self._krylov = want
        r0, nr = self.rows
        rank, nranks = (0, 1) if self.comm is None else (self.comm.rank, self.comm.size)
        self._ctx = Context(self.problem, self.nx, self.ny, row0=r0, nrows=nr, rank=rank, nranks=nranks,
                            inner_m=want[0], outer_k=want[1], gs=self._gs[0], gs_tau=self._gs[1],
                            kernel_variant=self._variant, buffers=self._buffers)
        if self.comm is not None and nranks > 1:
            self.comm.attach(self._ctx)
        self._configure(self._ctx)
        return self._ctx

    def _configure(self, ctx):  # push parameters / per-step state into a fresh context
        raise NotImplementedError

    def profile(self, on=True):
        """switch the per-kernel-class CUDA-event timing of the engine on / off"""
        ctx = self.context()
        ctx.check(ctx.lib.jfnk_profile_enable(ctx.handle, 1 if on else 0))

    def profile_read(self):
        """{class: {"launches", "ms", "bytes"}} since the last read (synchronises the stream)"""
        ctx = self.context()
        arr = (_capi.KernelStat * 32)()
        cnt = C.c_int(0)
        ctx.check(ctx.lib.jfnk_profile_read(ctx.handle, arr, 32, C.byref(cnt)))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].ms), "bytes": float(arr[i].bytes)}
                for i in range(cnt.value)}

    # -- operator level ---------------------------------------------------------------------------
    def __call__(self, u):
        ctx = self.context()
        du = ctx.vec(u, "u")
        dF = ctx.buf.alloc(ctx.n)
        ctx.check(ctx.lib.jfnk_residual(ctx.handle, ctx.buf.ptr(du), ctx.buf.ptr(dF)))
        return ctx.buf.to_user(dF, u)

    def linearize(self, x0, rdiff=None):
        """Fix the Jacobian's linearisation point (KrylovJacobian.setup, _nonlin.py:1579-1598)."""
        ctx = self.context()
        dx = ctx.vec(x0, "x0")
        ctx.check(ctx.lib.jfnk_linearize(ctx.handle, ctx.buf.ptr(dx), -1.0 if rdiff is None else float(rdiff)))
        return self

    def jvp(self, v):
        """KrylovJacobian.matvec(v) at the linearisation point (_nonlin.py:1557-1565)."""
        ctx = self.context()
        dv = ctx.vec(v, "v")
        out = ctx.buf.alloc(ctx.n)
        ctx.check(ctx.lib.jfnk_jvp(ctx.handle, ctx.buf.ptr(dv), ctx.buf.ptr(out)))
        return ctx.buf.to_user(out, v)

    def lgmres(self, b, rtol=1e-5, maxiter=1, reset=True):
        """scipy.sparse.linalg.lgmres(J, b, rtol=rtol, atol=0, maxiter=maxiter, inner_m, outer_k,
        outer_v=<carried list>, prepend_outer_v=True, store_outer_Av=False) -> (x, info)."""
        ctx = self.context()
        if reset:
            ctx.check(ctx.lib.jfnk_lgmres_reset(ctx.handle))
        db = ctx.vec(b, "b")
        dx = ctx.buf.alloc(ctx.n)
        info, inner, res = C.c_int(0), C.c_int(0), C.c_double(0)
        ctx.check(ctx.lib.jfnk_lgmres(ctx.handle, ctx.buf.ptr(db), ctx.buf.ptr(dx), float(rtol), int(maxiter),
                                      C.byref(info), C.byref(res), C.byref(inner)))
        self.last_lgmres = {"info": info.value, "res": res.value, "inner": inner.value}
        return ctx.buf.to_user(dx, b), info.value


class SHResidual(_DeviceResidual):
    """Crank-Nicolson residual of Swift-Hohenberg on an N x N periodic grid (sh_scipy_nk.py:15-49).

    ``h = d/N``; ``set_prev(Uo)`` plays the role of the script's ``Uo = U.copy(); UoUo = ...; UoUoUo = ...``.
    """

    problem = _capi.PROBLEM_SH

    def __init__(self, N=64, d=40.0, k=0.2, r=0.01, g=1.0, *, h=None, **kw):
        super().__init__(N, N, **kw)
        self.N, self.d, self.k, self.r, self.g = int(N), float(d), float(k), float(r), float(g)
        self.h = float(d) / N if h is None else float(h)
        self._prev = None

    def _configure(self, ctx):
        ctx.check(ctx.lib.jfnk_sh_setup(ctx.handle, self.h, self.r, self.g, self.k))
        if self._prev is not None:
            ctx.check(ctx.lib.jfnk_set_prev(ctx.handle, ctx.buf.ptr(self._prev)))

    def set_prev(self, Uo):
        ctx = self.context()
        self._prev = ctx.vec(Uo, "Uo")
        ctx.check(ctx.lib.jfnk_set_prev(ctx.handle, ctx.buf.ptr(self._prev)))
        return self

    def spmv_lap(self, x):
        """y = Lap @ x, the periodic 5-point Laplacian (sh_scipy_nk.py:34-35)."""
        return self._spmv("jfnk_spmv_lap", x)

    def spmv_L(self, x):
        """y = L @ x, L = -Lap*Lap - 2 Lap + (r-1) I (sh_scipy_nk.py:39)."""
        return self._spmv("jfnk_spmv_sh", x)

    def _spmv(self, fn, x):
        ctx = self.context()
        dx = ctx.vec(x, "x")
        dy = ctx.buf.alloc(ctx.n)
        ctx.check(getattr(ctx.lib, fn)(ctx.handle, ctx.buf.ptr(dx), ctx.buf.ptr(dy)))
        return ctx.buf.to_user(dy, x)

    def steps(self, U, nsteps=1, history=None, inplace=False, **opts):
        """``nsteps`` implicit time steps of the script's loop (sh_scipy_nk.py:53-61), entirely on the device:
        ``Uo = U; U = newton_krylov(residual, Uo, **opts)``.  Returns the new field like ``U``.
        ``inplace=True`` advances a flat fp64 device buffer without any copy (and returns it)."""
        ctx = self.context()
        if inplace:
            du = U
            if du.shape[0] != ctx.n:
                raise ValueError(f"U has {du.shape[0]} elements, expected {ctx.n}")
        else:
            du = ctx.vec(U, "U")
        o = Context.make_opts(**opts)
        hists = [HistoryBuffer() for _ in range(nsteps)]
        arr = (_capi.History * nsteps)(*[h.c for h in hists])
        rc = ctx.lib.jfnk_sh_step(ctx.handle, ctx.buf.ptr(du), int(nsteps), C.byref(o), arr)
        for h, c in zip(hists, arr):
            h.c = c
        if history is not None:
            history.extend(h.as_dict() for h in hists)
        self._prev = None  # the engine's per-step constant now belongs to the last step
        if rc != _capi.OK:
            ctx.check(rc, du if inplace else ctx.buf.to_user(du, U))
        return du if inplace else ctx.buf.to_user(du, U)


class SHLinearised(_DeviceResidual):
    """Linearly-implicit Swift-Hohenberg stepper (sh_linearised.py:16-57): per step solve
    ``(I + D - L k/2) U+ = (I + L k/2) U`` with ``D = diag((5U - Uo)^2 k/16 - g k U)``.  The reference
    factorises with SuperLU (``spsolve``); here the system is solved matrix-free by device LGMRES to ``rtol``."""

    problem = _capi.PROBLEM_SH_LINEAR

    def __init__(self, N=64, d=40.0, k=0.2, r=0.2, g=0.0, *, h=None, **kw):
        super().__init__(N, N, **kw)
        self.N, self.d, self.k, self.r, self.g = int(N), float(d), float(k), float(r), float(g)
        self.h = float(d) / N if h is None else float(h)

    def _configure(self, ctx):
        ctx.check(ctx.lib.jfnk_sh_setup(ctx.handle, self.h, self.r, self.g, self.k))

    def steps(self, U, Uo=None, nsteps=1, rtol=1e-13, maxiter=200):
        """Advance ``nsteps``; returns ``(U_new, Uo_new)``.  ``Uo`` defaults to ``U`` (sh_linearised.py:28)."""
        ctx = self.context()
        dU = ctx.vec(U, "U")
        dUo = ctx.vec(U if Uo is None else Uo, "Uo")
        info, mv = C.c_int(0), C.c_int64(0)
        ctx.check(ctx.lib.jfnk_shlin_step(ctx.handle, ctx.buf.ptr(dU), ctx.buf.ptr(dUo), int(nsteps), float(rtol),
                                          int(maxiter), C.byref(info), C.byref(mv)))
        self.last_info = {"info": info.value, "matvecs": mv.value}
        return ctx.buf.to_user(dU, U), ctx.buf.to_user(dUo, U)

    def prepare(self, U, Uo):
        """Returns b = (I + L k/2) U and fixes D for :meth:`jvp` / :meth:`lgmres` (operator-level parity tests)."""
        ctx = self.context()
        dU, dUo = ctx.vec(U, "U"), ctx.vec(Uo, "Uo")
        db = ctx.buf.alloc(ctx.n)
        ctx.check(ctx.lib.jfnk_shlin_prepare(ctx.handle, ctx.buf.ptr(dU), ctx.buf.ptr(dUo), ctx.buf.ptr(db)))
        return ctx.buf.to_user(db, U)


class _MeshResidual(_DeviceResidual):
    """Shared part of the moving-mesh problems: geometry and the mesh potential Q."""

    def __init__(self, nx, ny, dksi, deta, bounds, **kw):
        super().__init__(nx, ny, **kw)
        self.dksi, self.deta = float(dksi), float(deta)
        self.bounds = tuple(float(b) for b in bounds)  # (left, right, bottom, top)
        self._Q = None
        self._prev = None

    def _configure_mesh(self, ctx):
        ctx.check(ctx.lib.jfnk_mesh_setup(ctx.handle, self.dksi, self.deta, *self.bounds))
        if self._Q is not None:
            ctx.check(ctx.lib.jfnk_mesh_set_potential(ctx.handle, ctx.buf.ptr(self._Q)))

    def set_mesh(self, Q):
        """Mesh potential for this step: compute_Q_spatial_ders + J (+ the A_ij of Laplace_operator)."""
        ctx = self.context()
        self._Q = ctx.vec(Q, "Q")
        ctx.check(ctx.lib.jfnk_mesh_set_potential(ctx.handle, ctx.buf.ptr(self._Q)))
        return self

    # defaults of the PMA mesh relaxation, overridden per model
    _pma_defaults = dict(alpha=0.1, gamma=0.1, cnorm=1.0, smoothing_iters=4, monitor_mode=0)

    def relax_mesh(self, Q, Uval, dt, loops=1, **kw):
        """Moving-mesh relaxation on the device: ``loops`` passes of the body of ``loop_pma`` (droplet.py:590-599),
        i.e. ``solve_PMA(); Q.val += dt*Q.dt`` with the derivatives refreshed every pass; ``loops=1`` is the single
        mesh update of PMA2_nk.py:94,103.  ``Uval`` is the OLD solution.  Returns the new ``Q`` (like ``Q``)."""
        p = dict(self._pma_defaults)
        p.update(kw)
        ctx = self.context()
        dQ = ctx.vec(Q, "Q")
        dU = ctx.vec(Uval, "U.val")
        ctx.check(ctx.lib.jfnk_mesh_relax(ctx.handle, ctx.buf.ptr(dQ), ctx.buf.ptr(dU), float(dt), int(loops),
                                          float(p["alpha"]), float(p["gamma"]), float(p["cnorm"]),
                                          int(p["smoothing_iters"]), int(p["monitor_mode"])))
        self._Q = None  # the engine's metric fields are stale until set_mesh(Q) is called again
        self._prev = None
        return ctx.buf.to_user(dQ, Q)

    def laplace(self, v):
        """(v_xx, v_yy) = Laplace_operator(v, D_ksi v, D_eta v)."""
        ctx = self.context()
        dv = ctx.vec(v, "v")
        a, b = ctx.buf.alloc(ctx.n), ctx.buf.alloc(ctx.n)
        ctx.check(ctx.lib.jfnk_mesh_laplace(ctx.handle, ctx.buf.ptr(dv), ctx.buf.ptr(a), ctx.buf.ptr(b)))
        return ctx.buf.to_user(a, v), ctx.buf.to_user(b, v)


class PMA2Residual(_MeshResidual):
    """Residual of PMA2_nk.py:121-159 (p = 2) on an N x N computational grid over [-1, 1]^2.

    The script's ``residual`` always uses the module-level ``dt = k`` (its time loop assigns a *local* ``dt``,
    PMA2_nk.py:66,91), which is what ``dt`` means here.
    """

    problem = _capi.PROBLEM_PMA2
    # PMA2_nk.py:27-31 (alpha_, gamma_, smoothing_iters_); epsilon_ == 0 -> monitor 1/(1+u)**6 (:361-362);
    # mon += mon_integral (:389-390)
    _pma_defaults = dict(alpha=0.1, gamma=0.1, cnorm=1.0, smoothing_iters=4, monitor_mode=1)

    def __init__(self, N=51, *, lambd=1.0, beta=0.15, epsilon=0.0, m=3, dt=1e-4, endl=-1.0, endr=1.0, **kw):
        dksi = (endr - endl) / (N - 1)
        super().__init__(N, N, dksi, dksi, (endl, endr, endl, endr), **kw)
        self.lambd, self.beta, self.epsilon, self.m, self.dt = float(lambd), float(beta), float(epsilon), int(m), float(dt)

    def _configure(self, ctx):
        self._configure_mesh(ctx)
        ctx.check(ctx.lib.jfnk_pma2_setup(ctx.handle, self.lambd, self.beta, self.epsilon, self.m, self.dt))
        if self._prev is not None:
            ctx.check(ctx.lib.jfnk_pma2_set_prev(ctx.handle, ctx.buf.ptr(self._prev)))

    def set_prev(self, Uval):
        """``U.val = U.new.copy()`` and ``CN_term = compute_rhs_pde()`` (PMA2_nk.py:83,97)."""
        ctx = self.context()
        self._prev = ctx.vec(Uval, "U.val")
        ctx.check(ctx.lib.jfnk_pma2_set_prev(ctx.handle, ctx.buf.ptr(self._prev)))
        return self


def _pma2_run(self, U, Q, nsteps, k=None, history=None, **nk):
    """The while-loop body of PMA2_nk.py:80-106, ``nsteps`` times: metrics, Crank-Nicolson term, ``newton_krylov``,
    ``solve_PMA()`` with the old solution and ``Q.val += dt*Q.dt`` with the script's local ``dt = compute_g()*k``.
    Returns ``(U, Q, t)``."""
    from .nonlin import newton_krylov

    k = self.dt if k is None else float(k)
    t = 0.0
    for _ in range(int(nsteps)):
        self.set_mesh(Q)
        self.set_prev(U)
        g = float(((1 + U) ** 3).min()) if self.epsilon == 0 else 1.0  # compute_g (:446-450)
        Unew = newton_krylov(self, U, verbose=0, **nk)
        Q = self.relax_mesh(Q, U, g * k, loops=1)
        U = Unew
        t += g * k
        if history is not None:
            history.append(dict(self.last_history, t=t))
    return U, Q, t


PMA2Residual.run = _pma2_run


class DropletResidual(_MeshResidual):
    """Residual of droplet.py:435-450 on the Nx x Ny moving mesh (defaults: droplet.py:23-53)."""

    problem = _capi.PROBLEM_DROPLET
    # droplet.py:41-44 (alpha_, gamma_, C_), :31 (smoothing_iters_); monitor abs(u_xx + u_yy)**2 (:736)
    _pma_defaults = dict(alpha=0.01, gamma=0.1, cnorm=0.15, smoothing_iters=4, monitor_mode=0)

    def __init__(self, Nx=91, Ny=61, *, endl=-3.0, endr=6.0, endb=-3.0, endt=3.0, epsilon=1e-2, n=6, m=3, Bo=0.01,
                 alpha2=0.0, epsilon2=None, **kw):
        dksi = (endr - endl) / (Nx - 1)
        deta = (endt - endb) / (Ny - 1)
        super().__init__(Nx, Ny, dksi, deta, (endl, endr, endb, endt), **kw)
        self.epsilon, self.n_exp, self.m_exp, self.Bo, self.alpha2 = float(epsilon), int(n), int(m), float(Bo), float(alpha2)
        self.epsilon2 = 1.0 / (endt - endb) if epsilon2 is None else float(epsilon2)
        self._dt = None

    def _configure(self, ctx):
        self._configure_mesh(ctx)
        ctx.check(ctx.lib.jfnk_droplet_setup(ctx.handle, self.epsilon, self.n_exp, self.m_exp, self.Bo, self.alpha2,
                                             self.epsilon2))
        if self._prev is not None:
            ctx.check(ctx.lib.jfnk_droplet_set_prev(ctx.handle, ctx.buf.ptr(self._prev), self._dt))

    def set_prev(self, Uval, dt):
        """Per-step precompute of droplet.py:373-381: ``U.val``, ``P.val``, ``F = pde_rhs(...)`` and ``dt_n``."""
        ctx = self.context()
        self._prev = ctx.vec(Uval, "U.val")
        self._dt = float(dt)
        ctx.check(ctx.lib.jfnk_droplet_set_prev(ctx.handle, ctx.buf.ptr(self._prev), self._dt))
        return self


    def evolve_with_PDE(self, U, Q, dt=1e-4, iterMax=101, dtmesh=3e-9, pmaloops=400, scale=1.0, history=None, **nk):
        """The time loop of droplet.py:360-411 (``evolve_with_PDE(dt, iterMax, dtU, dtmesh, pmaloops)``): ``iterMax - 1``
        steps of ``dt_n = dt*scale`` -- metrics of the mesh, per-step precompute, ``newton_krylov(..., maxiter=20,
        f_tol=1e-7)``, ``loop_pma(dtmesh, pmaloops)`` with the OLD solution, ``scale += exp(-10 |dU|)``.  Returns
        ``(U, Q, t, scale)``; ``history`` (a list) receives one dict per step."""
        from .nonlin import newton_krylov

        nk.setdefault("maxiter", 20)
        nk.setdefault("f_tol", 1e-7)
        t = 0.0
        for _ in range(1, int(iterMax)):
            dt_n = dt * scale
            self.set_mesh(Q)
            self.set_prev(U, dt_n)
            Unew = newton_krylov(self, U, verbose=0, **nk)
            if pmaloops:
                Q = self.relax_mesh(Q, U, dtmesh, loops=pmaloops)
            diff = Unew - U
            nrm = float(diff.norm()) if hasattr(diff, "norm") else float(np.linalg.norm(diff))
            scale += float(np.exp(-10.0 * nrm))
            U = Unew
            t += dt_n
            if history is not None:
                history.append(dict(self.last_history, t=t, dt=dt_n, scale=scale))
        return U, Q, t, scale

    # ---- initial states (droplet.py:132-248, :316-358, :413-429, :544-554) --------------------------------------
    def uniform_mesh_potential(self):
        """Q of the undeformed mesh, 0.5 ksi^2 + 0.5 eta^2 (droplet.py:102)."""
        el, er, eb, et = self.bounds
        ksi, eta = np.meshgrid(np.linspace(el, er, self.nx), np.linspace(eb, et, self.ny))
        return np.reshape(0.5 * ksi ** 2 + 0.5 * eta ** 2, self.nx * self.ny)

    def compute_U2(self, Q, info, a=100.0):
        """``compute_U2(info)`` (droplet.py:413-423) on the mesh with potential ``Q``: the precursor film plus one smoothed
        parabolic cap per droplet, ``info = [[x, y, R, V], ...]``; evaluated on the device at the physical positions
        ``(Q_ksi, Q_eta)`` of the mesh points."""
        ctx = self.context()
        dQ = ctx.vec(Q, "Q")
        flat = np.ascontiguousarray(np.asarray(info, dtype=np.float64).reshape(-1, 4))
        out = ctx.buf.alloc(ctx.n)
        ctx.check(ctx.lib.jfnk_droplet_shape(ctx.handle, ctx.buf.ptr(dQ), int(flat.shape[0]),
                                             flat.ctypes.data_as(C.POINTER(C.c_double)), float(a), ctx.buf.ptr(out)))
        return ctx.buf.to_user(out, Q)

    def compute_U(self, Q, R=1.0, V=1.0, a=100.0):
        """``compute_U()`` (droplet.py:544-551): one droplet of radius ``R`` and volume ``V`` centred at the origin."""
        return self.compute_U2(Q, [[0.0, 0.0, R, V]], a)

    def initialise_coalescing_droplets(self, Vsteps, info, dtmesh, loops, Q=None, a=100.0, tofile=None):
        """``initialise_coalescing_droplets(Vsteps, info, dtmesh, loops, fromfile=False, tofile)`` (droplet.py:132-189):
        inflate the droplets of ``info`` in ``Vsteps`` volume increments, relaxing the mesh with ``loop_pma(dtmesh, loops)``
        after each.  Returns ``(U, Q)``; ``tofile`` writes the state in the reference's format (``fromfile`` is
        :func:`load_droplet_state`)."""
        Q = self.uniform_mesh_potential() if Q is None else Q
        Unew = np.full(self.nx * self.ny, self.epsilon)
        for i in range(1, int(Vsteps) + 1):
            Uval = Unew
            Unew = self.compute_U2(Q, [[d[0], d[1], d[2], d[3] * i / Vsteps] for d in info], a)
            Q = self.relax_mesh(Q, Uval, dtmesh, loops=loops)
        if tofile:
            save_droplet_state(tofile, Unew, Q)
        return Unew, Q

    def initialise_droplet(self, Vsteps, dtmesh, loops, R=1.0, Vf=1.0, Q=None, a=100.0, tofile=None):
        """``initialise_droplet`` (droplet.py:192-248): the single-droplet version (``compute_U`` with V = Vf i/Vsteps)."""
        return self.initialise_coalescing_droplets(Vsteps, [[0.0, 0.0, R, Vf]], dtmesh, loops, Q=Q, a=a, tofile=tofile)

    def evolve_R_explicit(self, U, Q, pmaloops, Rfinal, tol, R=1.0, V=1.0, dtR=5e-2, dtmesh=1e-7, a=100.0, max_iters=None):
        """``evolve_R_explicit`` (droplet.py:316-358): R' = (8V/R^3 - 1)/(3 log(1/eps)) (Rdot, :553-554) by explicit Euler with
        ``dt = dtR R^2``, ``U = compute_U()`` and a mesh relaxation per step.  As in the script the first relaxation pass
        still sees the previous ``U`` (its derivatives are refreshed before ``U.val`` is replaced).  Returns (U, Q, R, t)."""
        time, it = 0.0, 0
        while abs(Rfinal - R) > tol and (max_iters is None or it < max_iters):
            dt = dtR * (R ** 2)
            R += dt * (8 * V / R ** 3 - 1) / (3 * np.log(1 / self.epsilon))
            Unew = self.compute_U(Q, R, V, a)
            Q = self.relax_mesh(Q, U, dtmesh, loops=1)
            if pmaloops > 1:
                Q = self.relax_mesh(Q, Unew, dtmesh, loops=pmaloops - 1)
            U = Unew
            it += 1
            time += dt
        return U, Q, R, time


def load_droplet_state(path):
    """Read an ``initdrop_*.txt`` state: one ``U.val[i] Q.val[i]`` pair per line (written by droplet.py:556-562).
    The reference's own reader (droplet.py:564-576) hard-codes a Windows path separator."""
    a = np.loadtxt(path)
    return a[:, 0].copy(), a[:, 1].copy()


def save_droplet_state(path, U, Q):
    """Write a state in the reference's format (droplet.py:556-562)."""
    with open(path, "w") as f:
        for u, q in zip(np.asarray(U).ravel(), np.asarray(Q).ravel()):
            f.write(str(float(u)) + " " + str(float(q)) + "\n")
