// The whole moving-mesh relaxation loop (loop_pma, droplet.py:590-599; solve_PMA + explicit Euler update of
// PMA2_nk.py:94,103) as ONE persistent kernel run by ONE thread-block cluster, for the reference's grid sizes
// (51^2 ... 91 x 61).
//
// A pass is a dozen dependent stages over ~5000 points.  As separate launches (even replayed from a CUDA graph) each
// stage costs ~3.7 us of launch/drain latency: 52 us per pass, 21 ms for the 400 passes droplet.py runs per time step;
// a cooperative grid of 22 CTAs with grid-wide barriers measured the same (a grid barrier through L2 costs as much as
// a launch).  Here the stages are grid-stride loops of a single cluster of 16 CTAs (8 where 16 cannot be scheduled),
// separated by the hardware cluster barrier (barrier.cluster, release/acquire at cluster scope, ~0.3 us), and all
// `loops` passes run inside the same launch.
// The arithmetic is the same point functions (mesh_math.h) the per-stage kernels call, in the same order; the dense
// DCT-II products accumulate over k in the same order as small_gemm_kernel; the monitor's weighted sum is reduced per CTA
// and then in CTA order (deterministic).
// The four dense DCT products would otherwise be chains of 61-91 L2 round trips per output: every CTA keeps both DCT
// matrices in shared memory for the whole kernel and copies the (44 KB) operand field of a product into shared memory
// with one coalesced sweep before computing its outputs from shared memory only.
#pragma once
#include <cooperative_groups.h>
#include "cuda_common.cuh"
#include "mesh_kernels.cuh"
#include "pma_kernels.cuh"

namespace jfnk {

struct RelaxArgs {
  MeshGeom gm;
  PmaParams pp;
  double dt, cell;        // mesh time step ; dksi * deta
  int loops, deriv_bc;
  double* Q;
  const double* Uval;
  MetricPtrs M;           // metric fields of the current Q (rewritten every pass)
  double *a, *b, *t, *spec;
  const double *dctx, *dcty; // orthonormal DCT-II matrices (nx x nx, ny x ny)
  double* partials;       // one slot per CTA
  long long* prof;        // optional (JFNK_CYCLE_PROF=1): clock cycles per phase of pma_relax_band_kernel, CTA 0
};

constexpr int kRelaxThreads = 512;

__global__ void __launch_bounds__(kRelaxThreads) pma_relax_kernel(const __grid_constant__ RelaxArgs A) {
  namespace cg = cooperative_groups;
  cg::cluster_group grid = cg::this_cluster(); // the whole grid is one cluster
  __shared__ double red[kRelaxThreads / 32];
  __shared__ double total_s;
  extern __shared__ __align__(16) double relax_smem[]; // Cy [ny*ny] | Cx [nx*nx] | operand field [nx*ny]
  const MeshGeom& g = A.gm;
  const int nx = g.nx, ny = g.ny;
  const size_t n = (size_t)nx * ny;
  const size_t tid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const double* const* Mc = const_cast<const double* const*>(A.M.m);
  double* a = A.a;
  double* b = A.b;
  double* Cys = relax_smem;
  double* Cxs = Cys + (size_t)ny * ny;
  double* fld = Cxs + (size_t)nx * nx;
  for (int i = threadIdx.x; i < ny * ny; i += blockDim.x) Cys[i] = A.dcty[i];
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) Cxs[i] = A.dctx[i];
  auto stage_field = [&](const double* src) { // the CTA's private copy of a field other CTAs have just written
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) fld[i] = __ldcg(src + i);
    __syncthreads();
  };
  __syncthreads();

  for (int it = 0; it < A.loops; ++it) {
    // 1. metric fields of the current potential (compute_Q_spatial_ders + J + A_ij)
    for (size_t e = tid0; e < n; e += stride) {
      const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
      mesh_metrics_point(g, A.Q, r, c, A.M.m);
    }
    grid.sync();
    // 2. monitor function of the (old) solution: |u_xx + u_yy|^2 or 1/(1+u)^6
    for (size_t e = tid0; e < n; e += stride) {
      double lap = 0.0;
      if (A.pp.monitor_mode == 0) {
        const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
        double xx, yy;
        mesh_laplace_point(g, Mc, A.Uval, r, c, A.deriv_bc, xx, yy);
        lap = xx + yy;
      }
      a[e] = pma_monitor_point(A.pp.monitor_mode, A.Uval[e], lap);
    }
    grid.sync();
    // 3. smoothing sweeps (ping-pong)
    for (int s = 0; s < A.pp.smoothing_iters; ++s) {
      for (size_t e = tid0; e < n; e += stride) {
        const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
        b[e] = pma_smooth_point(g, a, r, c);
      }
      grid.sync();
      double* tmp = a; a = b; b = tmp;
    }
    // 4. Mackenzie regularisation: sum mon |J| over the grid, then sqrt((mon + C sum |J| dksi deta) |J|) / alpha
    {
      double acc = 0.0;
      for (size_t e = tid0; e < n; e += stride) acc = fma(a[e], fabs(A.M.m[3][e]), acc);
      acc = warp_sum(acc);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        A.partials[blockIdx.x] = v;
      }
    }
    grid.sync();
    if (threadIdx.x == 0) {
      double v = 0.0;
      for (unsigned k = 0; k < gridDim.x; ++k) v += __ldcg(A.partials + k);
      total_s = v;
    }
    __syncthreads();
    {
      const double add = A.pp.cnorm * A.cell * total_s;
      for (size_t e = tid0; e < n; e += stride) b[e] = sqrt((a[e] + add) * fabs(A.M.m[3][e])) / A.pp.alpha;
    }
    grid.sync();
    // 5. forward DCT-II: t = Cy . b ; spec = (t . Cx^T) / (1 - gamma Leig)
    stage_field(b);
    for (size_t e = tid0; e < n; e += stride) {
      const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
      double acc = 0.0;
      for (int k = 0; k < ny; ++k) acc = fma(Cys[(size_t)r * ny + k], fld[(size_t)k * nx + c], acc);
      A.t[e] = acc;
    }
    grid.sync();
    stage_field(A.t);
    for (size_t e = tid0; e < n; e += stride) {
      const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
      double acc = 0.0;
      for (int k = 0; k < nx; ++k) acc = fma(fld[(size_t)r * nx + k], Cxs[(size_t)c * nx + k], acc);
      A.spec[e] = acc / (1.0 - A.pp.gamma * pma_leig(g, r, c));
    }
    grid.sync();
    // 6. inverse: t = Cy^T . spec ; Q += dt (t . Cx)
    stage_field(A.spec);
    for (size_t e = tid0; e < n; e += stride) {
      const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
      double acc = 0.0;
      for (int k = 0; k < ny; ++k) acc = fma(Cys[(size_t)k * ny + r], fld[(size_t)k * nx + c], acc);
      A.t[e] = acc;
    }
    grid.sync();
    stage_field(A.t);
    for (size_t e = tid0; e < n; e += stride) {
      const int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
      double acc = 0.0;
      for (int k = 0; k < nx; ++k) acc = fma(fld[(size_t)r * nx + k], Cxs[(size_t)k * nx + c], acc);
      // lincomb(Q, 1, Q, dt, Q_t): product rounded, then sum (as the per-stage path and NumPy's Q.val += dt*Q.dt)
      A.Q[e] = combine(1.0 * A.Q[e], A.dt, acc);
    }
    grid.sync();
  }
}

} // namespace jfnk
