// Kernels of the moving-mesh relaxation: loop_pma / solve_PMA /
// compute_and_smooth_monitor of droplet.py:578-599,729-760 and PMA2_nk.py:345-403 on the device.
//   monitor -> 4 passes of the 9-point filter -> Mackenzie regularisation (weighted sum) -> sqrt(M |J|)/alpha
//   -> 2-D orthonormal DCT-II -> divide by (1 - gamma Leig) -> inverse DCT -> Q += dt Q_t
// The reference transforms with scipy.fft; here the DCT is applied as two small dense products with the
// orthonormal DCT-II matrices (the reference grids are 51^2 ... 91 x 61, sizes without a fast factorisation).
#pragma once
#include "cuda_common.cuh"
#include "mesh_math.h"

namespace jfnk {

__global__ void __launch_bounds__(256) pma_monitor_kernel(size_t n, int mode, const double* u, const double* lap, double* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
    out[e] = pma_monitor_point(mode, u[e], mode == 1 ? 0.0 : lap[e]);
}

__global__ void __launch_bounds__(256) pma_smooth_kernel(MeshGeom gm, const double* in, double* out) {
  const size_t n = (size_t)gm.nx * gm.ny, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    int r = (int)(e / gm.nx), c = (int)(e - (size_t)r * gm.nx);
    out[e] = pma_smooth_point(gm, in, r, c);
  }
}

__global__ void __launch_bounds__(256) pma_wsum_kernel(size_t n, const double* mon, const double* J, double* S, int out_off,
                                                       ReduceWs ws) {
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) acc[0] = fma(mon[e], fabs(J[e]), acc[0]);
  grid_reduce<1>(acc, 0u, ws, S + out_off);
}

__global__ void __launch_bounds__(256) pma_rhs_kernel(size_t n, const double* mon, const double* J, ScalarRef add, double alpha,
                                                      const double* S, double* out) {
  const double av = eval_sref(S, add);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
    out[e] = sqrt((mon[e] + av) * fabs(J[e])) / alpha;
}

__global__ void __launch_bounds__(256) pma_divide_kernel(MeshGeom gm, double gamma, double* Y) {
  const size_t n = (size_t)gm.nx * gm.ny, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    int r = (int)(e / gm.nx), c = (int)(e - (size_t)r * gm.nx);
    Y[e] = Y[e] / (1.0 - gamma * pma_leig(gm, r, c));
  }
}

__global__ void __launch_bounds__(256) dct_matrix_kernel(int N, double* C) {
  const size_t n = (size_t)N * N, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    int k = (int)(e / N), j = (int)(e - (size_t)k * N);
    C[e] = dct2_entry(k, j, N);
  }
}

// C[M x N] = op(A) . op(B), row-major; op = transpose when the flag is set.  16 x 16 shared-memory tiles.
__global__ void __launch_bounds__(256) small_gemm_kernel(int M, int N, int K, const double* A, int ta, const double* B, int tb,
                                                         double* C) {
  __shared__ double sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    int ak = k0 + tx, bk = k0 + ty;
    sa[ty][tx] = (row < M && ak < K) ? (ta ? A[(size_t)ak * M + row] : A[(size_t)row * K + ak]) : 0.0;
    sb[ty][tx] = (bk < K && col < N) ? (tb ? B[(size_t)col * K + bk] : B[(size_t)bk * N + col]) : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fma(sa[ty][k], sb[k][tx], acc);
    __syncthreads();
  }
  if (row < M && col < N) C[(size_t)row * N + col] = acc;
}

} // namespace jfnk
