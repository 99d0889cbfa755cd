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
                                                           const double* h, double* A, double* B) {
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
