// BLAS-1 kernels of the Arnoldi process: fused multi-dot (all Gram-Schmidt dots of one Arnoldi step in a
// single pass over the basis), fused multi-axpy (the whole GS update / the dx assembly in a single pass),
// and small helpers.  All are HBM-bound streaming kernels: 128-bit coalesced loads, many loads in flight
// per thread, warp-shuffle + last-block reductions.  Algorithmic traffic per grid point (fp64):
//   mdot   (nv vectors) : (nv + 1) * 8 B          gs_update / maxpy_sub : (nv + 2) * 8 B      maxpy : (nz + 1) * 8 B
#pragma once
#include "cuda_common.cuh"

namespace jfnk {

// out[i] = V_i . w (i < nv), out[nv] = w . w
template <int NV>
__global__ void __launch_bounds__(256) mdot_kernel(PtrList V, int nv, const double* __restrict__ w, size_t n,
                                                   double* S, int out_off, int guard, double tau2, ReduceWs ws) {
  if (guard && !gs_second_pass_taken(S, nv, tau2)) return; // grid-uniform
  double acc[NV + 1];
#pragma unroll
  for (int k = 0; k <= NV; ++k) acc[k] = 0.0;
  const size_t n2 = n >> 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 wv = ldg2(w + 2 * i);
    // groups of 8 basis vectors: 8 independent 128-bit loads in flight per thread, then 16 FMAs
#pragma unroll
    for (int kb = 0; kb < NV; kb += 8) {
      double2 vv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (kb + q < NV && kb + q < nv) vv[q] = ldg2(V.p[kb + q] + 2 * i);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (kb + q < NV && kb + q < nv) acc[kb + q] = fma(vv[q].y, wv.y, fma(vv[q].x, wv.x, acc[kb + q]));
    }
    acc[NV] = fma(wv.y, wv.y, fma(wv.x, wv.x, acc[NV]));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { // odd tail element
    double we = w[n - 1];
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < nv) acc[k] = fma(V.p[k][n - 1], we, acc[k]);
    acc[NV] = fma(we, we, acc[NV]);
  }
  // acc[0..nv) are the dots, acc[NV] is w.w; the reduction writes them contiguously: out[0..nv-1], out[nv]
  grid_reduce_sums<NV + 1>(acc, nv, ws, S + out_off);
}

// scalar-load variant for vectors that are not 16-byte aligned / odd leading dimension
template <int NV>
__global__ void __launch_bounds__(256) mdot_scalar_kernel(PtrList V, int nv, const double* __restrict__ w, size_t n,
                                                          double* S, int out_off, int guard, double tau2, ReduceWs ws) {
  if (guard && !gs_second_pass_taken(S, nv, tau2)) return;
  double acc[NV + 1];
#pragma unroll
  for (int k = 0; k <= NV; ++k) acc[k] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double wv = w[i];
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < nv) acc[k] = fma(V.p[k][i], wv, acc[k]);
    acc[NV] = fma(wv, wv, acc[NV]);
  }
  grid_reduce_sums<NV + 1>(acc, nv, ws, S + out_off);
}

// MODE 0: w -= sum_i (S[c_off+i]/S[JS_VN2+i]) V_i   (Gram-Schmidt update with unnormalised basis vectors)
// MODE 1: w -= sum_i S[c_off+i] V_i
// MODE 2: w  = sum_i S[c_off+i] V_i                  (dx assembly; w is write-only)
// S[n2_off] = ||w||^2 of the result.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) maxpy_kernel(PtrList V, int nv, double* __restrict__ w, size_t n, double* S,
                                                    int c_off, int n2_off, int guard, double tau2, ReduceWs ws) {
  if (guard && !gs_second_pass_taken(S, nv, tau2)) return;
  __shared__ double c[JF_MAXV];
  __shared__ const double* vp[JF_MAXV];
  if (threadIdx.x < nv) {
    double cv = S[c_off + threadIdx.x];
    if (MODE == 0) cv /= S[JS_VN2 + threadIdx.x];
    c[threadIdx.x] = cv;
    vp[threadIdx.x] = V.p[threadIdx.x];
  }
  __syncthreads();
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    const size_t n2 = n >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 t;
      int k0 = 0;
      if (MODE == 2) {
        double2 v = ldg2(vp[0] + 2 * i);
        t.x = c[0] * v.x; t.y = c[0] * v.y;
        k0 = 1;
      } else {
        t = *reinterpret_cast<const double2*>(w + 2 * i);
      }
#pragma unroll 8
      for (int k = k0; k < nv; ++k) {
        double2 v = ldg2(vp[k] + 2 * i);
        if (MODE == 2) { t.x = fma(c[k], v.x, t.x); t.y = fma(c[k], v.y, t.y); }
        else { t.x = fma(-c[k], v.x, t.x); t.y = fma(-c[k], v.y, t.y); }
      }
      stg2(w + 2 * i, t);
      acc[0] = fma(t.y, t.y, fma(t.x, t.x, acc[0]));
    }
  }
  // scalar path: everything when !VEC, the odd tail element when VEC
  {
    size_t begin = VEC ? (n & ~(size_t)1) : 0;
    for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double t;
      int k0 = 0;
      if (MODE == 2) { t = c[0] * vp[0][i]; k0 = 1; }
      else t = w[i];
      for (int k = k0; k < nv; ++k) t = (MODE == 2) ? fma(c[k], vp[k][i], t) : fma(-c[k], vp[k][i], t);
      w[i] = t;
      acc[0] = fma(t, t, acc[0]);
    }
  }
  grid_reduce<1>(acc, 0u, ws, S + n2_off);
}

// out = a x + b y (y may be null), optional ||out||^2
template <bool VEC>
__global__ void __launch_bounds__(256) lincomb_kernel(double* out, ScalarRef a, const double* x, ScalarRef b,
                                                      const double* y, size_t n, double* S, int n2_off, ReduceWs ws) {
  const double av = eval_sref(S, a), bv = eval_sref(S, b);
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    const size_t n2 = n >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 xv = *reinterpret_cast<const double2*>(x + 2 * i);
      double2 t;
      t.x = av * xv.x; t.y = av * xv.y;
      if (y) {
        double2 yv = *reinterpret_cast<const double2*>(y + 2 * i);
        // x + a*v with separate rounding of the product, as NumPy evaluates x0 + sc*v
        t.x = __dadd_rn(t.x, __dmul_rn(bv, yv.x));
        t.y = __dadd_rn(t.y, __dmul_rn(bv, yv.y));
      }
      stg2(out + 2 * i, t);
      acc[0] = fma(t.y, t.y, fma(t.x, t.x, acc[0]));
    }
  }
  size_t begin = VEC ? (n & ~(size_t)1) : 0;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double t = av * x[i];
    if (y) t = __dadd_rn(t, __dmul_rn(bv, y[i]));
    out[i] = t;
    acc[0] = fma(t, t, acc[0]);
  }
  if (n2_off >= 0) grid_reduce<1>(acc, 0u, ws, S + n2_off);
}

// out = (x - y) / div
__global__ void __launch_bounds__(256) diff_scale_kernel(double* out, const double* x, const double* y, ScalarRef div,
                                                         size_t n, const double* S) {
  const double dv = eval_sref(S, div);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (x[i] - y[i]) / dv;
}

__global__ void __launch_bounds__(256) maxabs_kernel(const double* v, size_t n, double* S, int out_off, ReduceWs ws) {
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc[0] = fmax(acc[0], fabs(v[i]));
  grid_reduce<1>(acc, 1u, ws, S + out_off);
}

__global__ void givens_kernel(double* S, int j, int pass2, double tau2) { hess_givens_step(S, j, pass2, tau2); }

struct IdxList {
  int v[JF_MAXV];
};
__global__ void lsq_kernel(double* S, int nit, IdxList zn2, int scale_n2_idx) { lsq_solve(S, nit, zn2.v, scale_n2_idx); }

} // namespace jfnk
