"""GPU-only checks of the sm_100a kernels through the C ABI: the marching shared-memory stencil kernel
against the one-thread-per-point kernel and the oracle, the fused BLAS-1 kernels against NumPy, and
size-independent properties at bandwidth-relevant sizes."""
import numpy as np
import pytest

import jfnk_b200 as jf
from oracle.sh import SHOracle, apply_L_roll, apply_lap_roll, seeded_state

pytestmark = pytest.mark.gpu


def relmax(a, b):
    return np.abs(np.ravel(a) - np.ravel(b)).max() / np.abs(np.ravel(b)).max()


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.mark.parametrize("N", [64, 130, 256, 520, 1030])
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_stencil_kernels_both_variants_vs_oracle(cuda_buffers, N, variant):
    """variant 1: the per-point kernel; 2: the marching kernel with 1-D bulk-copy staging (even N); 3: the marching kernel
    with tensor-map TMA boxes (even N >= 256; 520 and 1030 have a partial last strip that the TMA unit zero-fills, 256 a
    single-box strip; below 256 variant 3 is the per-point kernel)."""
    h = 0.625
    o = SHOracle(N=N, d=h * N)
    F = jf.SHResidual(N=N, d=h * N, buffers=cuda_buffers, kernel_variant=variant)
    u = seeded_state(N, 0)
    assert relmax(F.spmv_lap(u), o.Lap @ u) < 1e-14
    assert relmax(F.spmv_L(u), o.L @ u) < 1e-14
    Uo = seeded_state(N)
    o.set_prev(Uo)
    F.set_prev(Uo)
    x = Uo + 0.02 * seeded_state(N, 3)
    assert relmax(F(x), o.residual(x)) < 1e-13
    from scipy.optimize._nonlin import KrylovJacobian

    jac = KrylovJacobian()
    jac.setup(x.copy(), o.residual(x), o.residual)
    F.linearize(x)
    v = seeded_state(N, 11)
    ref = jac.matvec(v)
    noise = 64 * np.finfo(float).eps * np.abs(o.residual(x)).max() / (jac.omega / np.linalg.norm(v))
    assert np.abs(F.jvp(v) - ref).max() < max(noise, 1e-9 * np.abs(ref).max())


def test_marching_and_point_kernels_agree_bitwise(cuda_buffers):
    """same arithmetic core, same grouping of the sums: the two kernels must give identical bits."""
    N = 256
    u = seeded_state(N, 0)
    outs = []
    for variant in (1, 2, 3):
        F = jf.SHResidual(N=N, d=0.625 * N, buffers=cuda_buffers, kernel_variant=variant)
        F.set_prev(u)
        outs.append((F.spmv_L(u), F.spmv_lap(u), F(u + 0.5)))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("N", [512, 776])
def test_box_kernel_solver_paths(cuda_buffers, N):
    """every operation of the tensor-map marching kernel inside the solvers: implicit steps (set_prev, residual with and
    without the line-search operand, FD-JVP) and the linearised stepper (prepare + matvec), against the per-point kernel."""
    h = 0.625
    U0 = seeded_state(N)
    res = {}
    for variant in (1, 3):
        F = jf.SHResidual(N=N, d=h * N, buffers=cuda_buffers, kernel_variant=variant)
        hist = []
        U = F.steps(U0, 2, history=hist)
        S = jf.SHLinearised(N=N, d=h * N, buffers=cuda_buffers, kernel_variant=variant)
        V, _ = S.steps(0.1 * U0, nsteps=2)
        res[variant] = (U, [h_["nit"] for h_ in hist], V)
    assert rel(res[3][0], res[1][0]) < 1e-9 and res[3][1] == res[1][1]
    assert rel(res[3][2], res[1][2]) < 1e-11


@pytest.mark.parametrize("ny,nx,nv", [(8, 512, 0), (7, 585, 1), (5, 2001, 3), (16, 4096, 8), (7, 9363, 9),
                                      (8, 25000, 24), (7, 28573, 25), (10, 5000, 41)])
def test_fused_multi_dot_and_axpy(cuda_buffers, ny, nx, nv):
    """even (128-bit path) and odd (scalar path + tail) vector lengths, every accumulator-count instantiation"""
    import torch

    n = nx * ny
    rng = np.random.default_rng(n + nv)
    ctx = jf.Context(0, nx, ny, buffers=cuda_buffers)
    stride = (n + 31) // 32 * 32
    V = rng.standard_normal((max(nv, 1), stride))
    w = rng.standard_normal(n)
    dV = torch.from_numpy(V).cuda()
    dw = torch.from_numpy(w).cuda()
    got = ctx.multi_dot(dV, nv, stride, dw)
    ref = np.array([V[i, :n] @ w for i in range(nv)] + [w @ w])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-9)
    coef = rng.standard_normal(nv)
    n2 = ctx.multi_axpy(dV, nv, stride, coef, dw)
    wref = w - (coef[:, None] * V[:nv, :n]).sum(0) if nv else w
    np.testing.assert_allclose(dw.cpu().numpy(), wref, rtol=1e-12, atol=1e-12)
    assert abs(n2 - wref @ wref) <= 1e-12 * (wref @ wref)
    # run-to-run determinism of the two-stage reduction
    assert np.array_equal(ctx.multi_dot(dV, nv, stride, dw), ctx.multi_dot(dV, nv, stride, dw))
    # an 8-byte-aligned (not 16-byte) operand takes the scalar-load kernels
    box = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    dw2 = box[1:]
    dw2.copy_(torch.from_numpy(w))
    assert dw2.data_ptr() % 16 == 8
    np.testing.assert_allclose(ctx.multi_dot(dV, nv, stride, dw2), ref, rtol=1e-12, atol=1e-9)
    ctx.multi_axpy(dV, nv, stride, coef, dw2)
    np.testing.assert_allclose(dw2.cpu().numpy(), wref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N", [512])
def test_time_steps_at_marching_size(cuda_buffers, N):
    """two implicit steps at a size where the solver runs on the marching kernels (h fixed at 0.625)."""
    h = 0.625
    o = SHOracle(N=N, d=h * N)
    F = jf.SHResidual(N=N, d=h * N, buffers=cuda_buffers)
    U0 = seeded_state(N)
    href, hist = [], []
    Uref = o.run(U0, 2, history=href)
    U = F.steps(U0, 2, history=hist)
    assert rel(U, Uref) < 1e-8
    assert [h["nit"] for h in hist] == [len(h["iters"]) for h in href]


@pytest.mark.parametrize("N", [4096])
def test_large_grid_properties(cuda_buffers, N):
    """BASELINE-size checks that need no matrix: exact stencil vs np.roll, constants, linearity, zero column sums."""
    import torch

    h, r = 0.625, 0.01
    F = jf.SHResidual(N=N, d=h * N, r=r, buffers=cuda_buffers)
    x = seeded_state(N, 0)
    y = seeded_state(N, 1)
    Lx = F.spmv_L(x)
    assert relmax(Lx, apply_L_roll(x, N, h, r)) < 1e-13
    lapx = F.spmv_lap(x)
    assert relmax(lapx, apply_lap_roll(x, N, h)) < 1e-14
    c = np.full(N * N, 3.25)
    assert np.abs(F.spmv_lap(c)).max() < 1e-10
    np.testing.assert_allclose(F.spmv_L(c), (r - 1) * c, rtol=1e-10)
    Ly = F.spmv_L(y)
    assert relmax(F.spmv_L(2.5 * x - 0.75 * y), 2.5 * Lx - 0.75 * Ly) < 1e-13
    assert abs(lapx.sum()) < 1e-6 * np.abs(lapx).sum()  # periodic Laplacian: zero column sums
    # F(u) - F(w) for the CN residual is independent of the per-step constant: G(u) - G(w)
    F.set_prev(y)
    d1 = F(x) - F(y)
    F.set_prev(x)
    d2 = F(x) - F(y)
    assert relmax(d1, d2) < 1e-11
