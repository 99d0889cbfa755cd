"""``newton_krylov`` with SciPy's signature, running on the B200 engine.

Drop-in for the reference's call sites
    newton_krylov(residual, Uo, verbose=1)                                   sh_scipy_nk.py:61
    newton_krylov(residual, U.val, verbose=0)                                PMA2_nk.py:100
    newton_krylov(lambda u: residual(u, F, dt_n), U.val, verbose=1, maxiter=20, f_tol=1e-7)   droplet.py:383
where ``F`` is a device residual handle (:mod:`residuals`) instead of a NumPy closure.  Return value,
exceptions (``NoConvergence(x)``, ``ValueError``) and the verbose trace format follow
scipy/optimize/_nonlin.py:134-277.
"""
from __future__ import annotations

import ctypes as C
import sys

from . import _capi
from .context import Context, HistoryBuffer, NoConvergence  # noqa: F401
from .residuals import _DeviceResidual


def newton_krylov(F, xin, iter=None, rdiff=None, method="lgmres", inner_maxiter=20, inner_M=None, outer_k=10,
                  verbose=False, maxiter=None, f_tol=None, f_rtol=None, x_tol=None, x_rtol=None, tol_norm=None,
                  line_search="armijo", callback=None, full_output=False, raise_exception=True, **kw):
    """Find a root of ``F`` with the Jacobian-free Newton-Krylov method (LGMRES inner solver).

    Same parameters as :func:`scipy.optimize.newton_krylov`; the differences are
      * ``F`` is a device residual handle (``SHResidual``, ``PMA2Residual``, ``DropletResidual``),
      * ``method`` must be ``'lgmres'`` (the only inner solver the reference uses), ``tol_norm`` must be None (max-norm),
      * ``inner_M`` (never set by the reference) is a preconditioner acting on DEVICE vectors: a callable or an object
        with ``matvec`` that maps a flat fp64 device tensor (NumPy array with the CPU test double) to ``M @ v``; as in
        SciPy it is applied from the left inside LGMRES, and its optional ``update(x, f)`` method is called after every
        Newton step (KrylovJacobian.update, _nonlin.py:1574-1577).  See :mod:`precond` for a Fourier preconditioner
        of the Swift-Hohenberg Crank-Nicolson operator,
      * ``inner_maxiter`` is accepted and -- exactly as in SciPy for LGMRES -- has no effect
        (KrylovJacobian overrides ``maxiter=1`` and LGMRES' ``inner_m`` stays 30, _nonlin.py:1503-1506);
        pass ``inner_inner_m=...`` to change the Arnoldi length as SciPy's ``inner_*`` forwarding allows.
    """
    if not isinstance(F, _DeviceResidual):
        raise TypeError("F must be a device residual handle (SHResidual, PMA2Residual, DropletResidual); "
                        "there is no CPU path for arbitrary Python callables")
    if method != "lgmres":
        raise NotImplementedError("only method='lgmres' (the reference's inner solver) is implemented")
    if tol_norm is not None:
        raise NotImplementedError("tol_norm: only the default max-norm is implemented")
    inner_m = 30
    for key, value in kw.items():
        if not key.startswith("inner_"):
            raise ValueError(f"Unknown parameter {key}")
        if key == "inner_inner_m":
            inner_m = int(value)
        elif key == "inner_outer_k":
            outer_k = int(value)
        else:
            raise NotImplementedError(f"inner option {key!r} is not supported by the device LGMRES")

    ctx = F.context(inner_m=inner_m, outer_k=outer_k)
    opts = Context.make_opts(f_tol=f_tol, f_rtol=f_rtol, x_tol=x_tol, x_rtol=x_rtol, rdiff=rdiff, maxiter=maxiter,
                             iter=iter, line_search=line_search)
    du = ctx.vec(xin, "xin")
    hist = HistoryBuffer()

    n = ctx.n
    if callback is not None or inner_M is not None:
        ctx.check_stream("newton_krylov(callback= / inner_M=)")
    ps_c = None
    m_update = getattr(inner_M, "update", None) if inner_M is not None else None
    if inner_M is not None:
        apply_M = inner_M.matvec if hasattr(inner_M, "matvec") else inner_M
        if not callable(apply_M):
            raise TypeError("inner_M must be callable or provide matvec(v)")
        ps_err = []

        def _ps(_user, pin, pout):
            try:
                out = apply_M(ctx.buf.raw_view(pin, n))
                dst = ctx.buf.raw_view(pout, n)
                if hasattr(dst, "copy_"):
                    dst.copy_(out.reshape(-1))
                else:
                    dst[:] = out.reshape(-1)
                return 0
            except Exception as e:  # no exception may cross the C ABI
                ps_err.append(e)
                return 1

        ps_c = _capi.PSOLVE(_ps)
        if hasattr(inner_M, "setup"):  # KrylovJacobian.setup (_nonlin.py:1594-1598): preconditioner.setup(x, f, func)
            inner_M.setup(du, F(du), F)
        ctx.check(ctx.lib.jfnk_set_preconditioner(ctx.handle, ps_c, None))
    cb_c = None
    cb_err = []
    if callback is not None or m_update is not None:

        def _cb(_user, _it, px, pF, _fmax, _fl2):
            # SciPy lets an exception raised by `callback` / `inner_M.update` propagate out of newton_krylov
            # (_nonlin.py:240-243).  Nothing may cross the C ABI, so it is parked here, the remaining callbacks of the
            # solve are skipped, and it is re-raised as soon as jfnk_newton returns.
            if cb_err:
                return
            try:
                if m_update is not None:
                    m_update(ctx.buf.raw_view(px, n), ctx.buf.raw_view(pF, n))
                if callback is not None:
                    callback(ctx.buf.view_for_callback(px, n, xin), ctx.buf.view_for_callback(pF, n, xin))
            except BaseException as e:  # noqa: BLE001 -- re-raised below
                cb_err.append(e)
                ctx.lib.jfnk_request_stop(ctx.handle)

        cb_c = _capi.CALLBACK(_cb)
        ctx.check(ctx.lib.jfnk_set_callback(ctx.handle, cb_c, None))
    try:
        rc = ctx.lib.jfnk_newton(ctx.handle, ctx.buf.ptr(du), C.byref(opts), C.byref(hist.c))
    finally:
        if cb_c is not None:
            ctx.lib.jfnk_set_callback(ctx.handle, _capi.CALLBACK(), None)
        if ps_c is not None:
            ctx.lib.jfnk_set_preconditioner(ctx.handle, _capi.PSOLVE(), None)
    if cb_err:
        raise cb_err[0]
    if inner_M is not None and ps_err:
        raise ps_err[0]
    h = hist.as_dict()
    F.last_history = h
    if verbose:
        for i in range(len(h["f_max"])):
            sys.stdout.write(f"{i}:  |F(x)| = {h['f_max'][i]:g}; step {h['step'][i]:g}\n")
        sys.stdout.flush()
    x = ctx.buf.to_user(du, xin)
    status = 1
    if rc == _capi.NO_CONVERGENCE:
        if raise_exception:
            raise NoConvergence(x)
        status = 2
    else:
        ctx.check(rc)
    if full_output:
        info = {"nit": h["nit"] + (1 if status == 1 else 0), "status": status, "success": status == 1,
                "message": {1: "A solution was found at the specified tolerance.",
                            2: "The maximum number of iterations allowed has been reached."}[status],
                "history": h}
        return x, info
    return x
