"""Preconditioners for ``newton_krylov(..., inner_M=...)`` acting on device vectors.

The reference never passes ``inner_M``; without one, the script's fixed domain d = 40 makes the
unpreconditioned Krylov solve need 1234 (N = 256) to 5826 (N = 512) residual evaluations per time step.  The
Crank-Nicolson Jacobian of Swift-Hohenberg is ``J = I/k - (L + 2 g u - 3 u^2)/2`` (sh_scipy_nk.py:47-49); its constant-
coefficient part ``I/k - L/2`` is diagonal in Fourier space on the periodic grid, which gives an exact inverse in
O(n log n): ``M = (I/k - L/2)^-1``.
"""
from __future__ import annotations

import numpy as np


class SHFourierPreconditioner:
    """``M v = ifft2( fft2(v) / (1/k - Lhat/2 + shift) )`` with ``Lhat = -lam^2 - 2 lam + (r - 1)`` and
    ``lam(i, j) = e (2 cos(2 pi i/N) - 2) + e (2 cos(2 pi j/N) - 2)``, the symbol of the script's 5-point periodic
    Laplacian (sh_scipy_nk.py:32-39), ``e = 1/h^2``.  Works on flat fp64 torch device tensors (cuFFT through
    ``torch.fft``) and on NumPy arrays (the oracle / CPU test double); single rank (the FFT is global).

    ``adaptive=True``: ``update(x, f)`` -- called by ``newton_krylov`` after every Newton step, as SciPy calls
    ``inner_M.update`` (KrylovJacobian.update, _nonlin.py:1574-1577) -- sets ``shift`` to the grid mean of the dropped
    diagonal ``(3 x^2 - 2 g x)/2`` of the Jacobian, which matters while ``|u|`` is large (early time steps).
    """

    def __init__(self, N, h, k, r, g=0.0, adaptive=False, shift=0.0):
        self.N, self.k, self.g, self.adaptive = int(N), float(k), float(g), bool(adaptive)
        e = 1.0 / (h * h)
        c = 2.0 * np.cos(2.0 * np.pi * np.arange(self.N) / self.N) - 2.0
        lam = e * (c[:, None] + c[None, :])
        self.base = 1.0 / k - 0.5 * (-lam * lam - 2.0 * lam + (r - 1.0))  # symbol of I/k - L/2
        self._base_dev = None
        self.shift = float(shift)

    @classmethod
    def for_residual(cls, F, **kw):
        """from an :class:`SHResidual` handle"""
        return cls(F.N, F.h, F.k, F.r, g=F.g, **kw)

    @property
    def symbol(self):
        return 1.0 / (self.base + self.shift)

    def update(self, x, f):
        if self.adaptive:
            self.shift = max(0.0, float((1.5 * x * x - self.g * x).mean()))

    def setup(self, x, f, func):
        self.update(x, f)

    def matvec(self, v):
        N = self.N
        if isinstance(v, np.ndarray):
            return np.real(np.fft.ifft2(np.fft.fft2(v.reshape(N, N)) / (self.base + self.shift))).reshape(-1)
        import torch

        if self._base_dev is None or self._base_dev.device != v.device:
            self._base_dev = torch.from_numpy(self.base).to(v.device)
        return torch.fft.ifft2(torch.fft.fft2(v.reshape(N, N)) / (self._base_dev + self.shift)).real.reshape(-1)

    __call__ = matvec
