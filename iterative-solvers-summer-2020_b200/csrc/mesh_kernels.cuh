// Matrix-free kernels of the moving-mesh residuals (PMA2_nk.py:121-159, droplet.py:435-450): the
// reference's COO/CSR derivative matrices (make_M) and the slice arithmetic of Laplace_operator become
// one-thread-per-point stencil kernels with a straight-line interior fast path and the one-sided boundary
// closures selected per point (arithmetic in mesh_math.h).  These kernels serve the reference's 51^2 ... 91 x 61
// grids (launch-latency bound) and the 2048^2 synthetic case, where they are instruction-issue bound
// (~420 instructions per point, DRAM traffic at the ideal 6 fields): 1.2-1.6 TB/s of algorithmic traffic.
#pragma once
#include "cuda_common.cuh"
#include "mesh_math.h"

namespace jfnk {

struct MetricPtrs {
  double* m[7];
};
struct MetricCPtrs {
  const double* m[7];
};

// Stencil kernels walk the grid in 32 x 8 tiles (one tile per 256-thread CTA pass): the vertical neighbours of a
// point are then loaded by the same CTA and hit in L1 (86 % L1 hit rate in ncu), instead of every stencil leg going
// to L2 (a 1-D mapping puts adjacent rows on different SMs).  Rows stay contiguous, so each warp still reads whole
// 256-byte segments.  `body(r, c, e)` is called once per grid point with e = r*nx + c.
template <typename Body>
__device__ __forceinline__ void for_each_point_tiled(const MeshGeom& gm, Body body) {
  const int tiles_x = (gm.nx + 31) >> 5, tiles_y = (gm.ny + 7) >> 3;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int r = (tile / tiles_x) * 8 + ty, c = (tile % tiles_x) * 32 + tx;
    if (r < gm.ny && c < gm.nx) body(r, c, (size_t)r * gm.nx + c);
  }
}

__global__ void __launch_bounds__(256) mesh_metrics_kernel(MeshGeom gm, const double* Q, MetricPtrs M) {
  for_each_point_tiled(gm, [&](int r, int c, size_t) { mesh_metrics_point(gm, Q, r, c, M.m); });
}

__global__ void __launch_bounds__(256) mesh_laplace_kernel(MeshGeom gm, MetricCPtrs M, const double* v, double* vxx,
                                                           double* vyy, int sum_only, int deriv_bc) {
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    double xx, yy;
    mesh_laplace_point(gm, M.m, v, r, c, deriv_bc, xx, yy);
    if (sum_only) vxx[e] = xx + yy;
    else { vxx[e] = xx; vyy[e] = yy; }
  });
}

__global__ void __launch_bounds__(256) pma2_rhs_kernel(MeshGeom gm, Pma2Params pp, const double* u, const double* lap2,
                                                       double* out) {
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    bool bdy = (r == 0 || c == 0 || r == gm.ny - 1 || c == gm.nx - 1);
    out[e] = bdy ? 0.0 : pma2_rhs_point(pp, u[e], lap2[e]);
  });
}

__global__ void __launch_bounds__(256) pma2_combine_kernel(size_t n, Pma2Params pp, const double* u, const double* uval,
                                                           const double* rhs, const double* cn, double* F, double* S,
                                                           int norm_off, ReduceWs ws) {
  double val[3] = {0.0, 0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    double f = pma2_combine_point(pp, u[e], uval[e], rhs[e], cn[e]);
    F[e] = f;
    val[0] = fma(f, f, val[0]); val[1] = fmax(val[1], fabs(f)); val[2] = fmax(val[2], fabs(u[e]));
  }
  grid_reduce<3>(val, 0x6u, ws, S + norm_off);
}

__global__ void __launch_bounds__(256) droplet_pressure_kernel(size_t n, DropletParams dp, const double* h,
                                                               const double* lap, double* p) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
    p[e] = droplet_pressure_point(dp, h[e], lap[e]);
}

__global__ void __launch_bounds__(256) droplet_flux_kernel(MeshGeom gm, DropletParams dp, MetricCPtrs M, const double* p,
                                                           const double* h, double* A, double* B, const double* S,
                                                           int skippable) {
  if (skippable && S[JS_STOP] != 0.0) return;
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    double a, b;
    droplet_flux_point(gm, dp, M.m, p, h, r, c, a, b);
    A[e] = a; B[e] = b;
  });
}

__global__ void __launch_bounds__(256) droplet_div_kernel(MeshGeom gm, MetricCPtrs M, const double* A, const double* B,
                                                          double* out) {
  for_each_point_tiled(gm, [&](int r, int c, size_t e) { out[e] = droplet_div_point(gm, M.m, A, B, r, c); });
}

// ---- fused evaluation chains of the reference-sized grids ---------------------------------------------------------------
// At 51^2 ... 91 x 61 points every launch is pure latency, so the residual / FD-JVP chains are cut to the launches the
// stencil dependencies force: (1) Laplace_operator of t = x + a v formed on the fly (never stored unless the caller wants
// it) with the pointwise pressure in the epilogue, (2) the flux, (3) the divergence with the Crank-Nicolson combination,
// the FD quotient (F - f0)/div or the norms in the epilogue.  `skippable`: part of an Arnoldi step the host enqueued
// ahead of the previous step's outcome (Engine::cycle): return at once while JS_STOP is set.
//   EPI 0: out = t_xx + t_yy (PMA2 pass 1) ; EPI 1: out = p(t, t_xx + t_yy) (droplet.py:468-473)
template <int EPI, bool HAS_V>
__global__ void __launch_bounds__(256) mesh_lap_fused_kernel(MeshGeom gm, MetricCPtrs M, const double* x, const double* v,
                                                             ScalarRef a, DropletParams dp, double* tt, double* out,
                                                             const double* S, int skippable) {
  if (skippable && S[JS_STOP] != 0.0) return;
  const double av = HAS_V ? eval_sref(S, a) : 0.0;
  const size_t nx = gm.nx;
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    double xx, yy, tc;
    if (HAS_V) {
      auto f = [=](int rr, int cc) { const size_t i = (size_t)rr * nx + cc; return combine(x[i], av, v[i]); };
      mesh_laplace_general_g(gm, M.m, f, r, c, 0, xx, yy);
      tc = f(r, c);
      if (tt) tt[e] = tc;
    } else {
      mesh_laplace_point(gm, M.m, x, r, c, 0, xx, yy);
      tc = x[e];
    }
    out[e] = (EPI == 0) ? xx + yy : droplet_pressure_point(dp, tc, xx + yy);
  });
}

// PMA2 pass 2 (PMA2_nk.py:131-159): lap2 = Laplace(lap1) ; rhs(u, lap2) with the boundary zeroed ; F = (u - uval)/dt -
// (rhs + cn)/2 ; QUOT: out = (F - f0)/div, else out = F and the norms {sum F^2, max|F|, max|u|} at S[norm_off..+2]
template <bool QUOT>
__global__ void __launch_bounds__(256) pma2_lap_fused_kernel(MeshGeom gm, MetricCPtrs M, Pma2Params pp, const double* lap1,
                                                             const double* u, const double* uval, const double* cn,
                                                             const double* f0, ScalarRef div, double* out, double* S,
                                                             int norm_off, ReduceWs ws, int skippable) {
  if (skippable && S[JS_STOP] != 0.0) return;
  const double dv = QUOT ? eval_sref(S, div) : 1.0;
  double val[3] = {0.0, 0.0, 0.0};
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    double xx, yy;
    mesh_laplace_point(gm, M.m, lap1, r, c, 0, xx, yy);
    const bool bdy = (r == 0 || c == 0 || r == gm.ny - 1 || c == gm.nx - 1);
    const double rhs = bdy ? 0.0 : pma2_rhs_point(pp, u[e], xx + yy);
    const double f = pma2_combine_point(pp, u[e], uval[e], rhs, cn[e]);
    if (QUOT) out[e] = (f - f0[e]) / dv;
    else {
      out[e] = f;
      val[0] = fma(f, f, val[0]); val[1] = fmax(val[1], fabs(f)); val[2] = fmax(val[2], fabs(u[e]));
    }
  });
  if (!QUOT) grid_reduce<3>(val, 0x6u, ws, S + norm_off);
}

// droplet.py:448-450: F2 = div(A, B) ; F = (u - uval) - dt (F2 + Fprev)/2 ; QUOT / norms as above
template <bool QUOT>
__global__ void __launch_bounds__(256) droplet_div_fused_kernel(MeshGeom gm, MetricCPtrs M, DropletParams dp, const double* A,
                                                                const double* B, const double* u, const double* uval,
                                                                const double* Fprev, const double* f0, ScalarRef div,
                                                                double* out, double* S, int norm_off, ReduceWs ws,
                                                                int skippable) {
  if (skippable && S[JS_STOP] != 0.0) return;
  const double dv = QUOT ? eval_sref(S, div) : 1.0;
  double val[3] = {0.0, 0.0, 0.0};
  for_each_point_tiled(gm, [&](int r, int c, size_t e) {
    const double F2 = droplet_div_point(gm, M.m, A, B, r, c);
    const double f = droplet_combine_point(dp, u[e], uval[e], F2, Fprev[e]);
    if (QUOT) out[e] = (f - f0[e]) / dv;
    else {
      out[e] = f;
      val[0] = fma(f, f, val[0]); val[1] = fmax(val[1], fabs(f)); val[2] = fmax(val[2], fabs(u[e]));
    }
  });
  if (!QUOT) grid_reduce<3>(val, 0x6u, ws, S + norm_off);
}

__global__ void __launch_bounds__(256) droplet_shape_kernel(MeshGeom gm, const double* Q, DropList dl, double a, double eps,
                                                            double* out) {
  for_each_point_tiled(gm, [&](int r, int c, size_t e) { out[e] = droplet_shape_point(gm, Q, r, c, dl, a, eps); });
}

__global__ void __launch_bounds__(256) droplet_combine_kernel(size_t n, DropletParams dp, const double* u,
                                                              const double* uval, const double* F2, const double* Fprev,
                                                              double* F, double* S, int norm_off, ReduceWs ws) {
  double val[3] = {0.0, 0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    double f = droplet_combine_point(dp, u[e], uval[e], F2[e], Fprev[e]);
    F[e] = f;
    val[0] = fma(f, f, val[0]); val[1] = fmax(val[1], fabs(f)); val[2] = fmax(val[2], fabs(u[e]));
  }
  grid_reduce<3>(val, 0x6u, ws, S + norm_off);
}

} // namespace jfnk
