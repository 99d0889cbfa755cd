// BLAS-1 kernels of the Arnoldi process: fused multi-dot (all Gram-Schmidt dots of one Arnoldi step in a
// single pass over the basis), fused multi-axpy (the whole GS update / the dx assembly in a single pass),
// and small helpers.  All are HBM-bound streaming kernels: 128-bit coalesced loads, many loads in flight
// per thread, warp-shuffle + last-block reductions.  Algorithmic traffic per grid point (fp64):
//   mdot   (nv vectors) : (nv + 1) * 8 B          gs_update / maxpy_sub : (nv + 2) * 8 B      maxpy : (nz + 1) * 8 B
#pragma once
#include "cuda_common.cuh"
#include "p2p_kernels.cuh"

namespace jfnk {

// out[i] = V_i . w (i < nv), out[nv] = w . w
// NV = accumulators compiled in (nv <= NV at run time), U = double2 elements per thread per sweep: few basis
// vectors -> several elements per thread so that every thread still keeps >= 8 independent 128-bit loads in flight.
// (two 256-thread CTAs per SM: <= 128 registers, checked with -Xptxas -v; one CTA/SM starves the memory pipe)
// nvb > 1: the vector is summed as nvb virtual blocks of n / nvb elements (rank-count-independent sums, cuda_common.cuh)
template <int NV, int U>
__global__ void __launch_bounds__(256, 2) mdot_kernel(PtrList V, int nv, const double* __restrict__ w, size_t n,
                                                      double* S, int out_off, ReduceWs ws, P2PReduceArgs R, int nvb) {
  if (S[JS_STOP] != 0.0) return; // speculatively enqueued Arnoldi step after the process stopped
  double acc[NV + 1];
  const size_t nb2 = (n / (size_t)nvb) >> 1; // double2 elements per virtual block (the launcher makes it whole when nvb > 1)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int vb = 0; vb < nvb; ++vb) {
  const size_t n2 = (size_t)(vb + 1) * nb2; // this block is the index range [vb nb2, n2)
#pragma unroll
  for (int k = 0; k <= NV; ++k) acc[k] = 0.0;
  for (size_t i0 = (size_t)vb * nb2 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n2; i0 += stride * U) {
    double2 wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      wv[u] = (i < n2) ? ldg2(w + 2 * i) : make_double2(0.0, 0.0);
    }
    // groups of 8 basis vectors: 8*U independent 128-bit loads in flight per thread, then the FMAs
#pragma unroll
    for (int kb = 0; kb < NV; kb += 8) {
      double2 vv[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = i0 + (size_t)u * stride;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (kb + q < NV && kb + q < nv && i < n2) vv[u][q] = ldg2(V.p[kb + q] + 2 * i);
          else vv[u][q] = make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (kb + q < NV && kb + q < nv) acc[kb + q] = fma(vv[u][q].y, wv[u].y, fma(vv[u][q].x, wv[u].x, acc[kb + q]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc[NV] = fma(wv[u].y, wv[u].y, fma(wv[u].x, wv[u].x, acc[NV]));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { // odd tail element (only with nvb == 1)
    double we = w[n - 1];
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < nv) acc[k] = fma(V.p[k][n - 1], we, acc[k]);
    acc[NV] = fma(we, we, acc[NV]);
  }
  // acc[0..nv) are the dots, acc[NV] is w.w; the reduction writes them contiguously: out[0..nv-1], out[nv]
  seg_store<NV + 1>(acc, nv, vb, ws);
  } // virtual blocks
  const bool last = seg_finalize(nv + 1, nvb, ws, S + out_off);
  // slab ranks: the finalising CTA all-reduces the nv + 1 sums over peer memory (no separate collective launch)
  if (last && R.nranks > 1) p2p_allreduce_block(R);
}

// scalar-load variant for vectors that are not 16-byte aligned / odd leading dimension
template <int NV>
__global__ void __launch_bounds__(256) mdot_scalar_kernel(PtrList V, int nv, const double* __restrict__ w, size_t n,
                                                          double* S, int out_off, ReduceWs ws, P2PReduceArgs R) {
  if (S[JS_STOP] != 0.0) return;
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
  const bool last = grid_reduce_sums<NV + 1>(acc, nv, ws, S + out_off);
  if (last && R.nranks > 1) p2p_allreduce_block(R);
}

// MODE 0: w -= sum_i (S[c_off+i]/S[JS_VN2+i]) V_i   (Gram-Schmidt update with unnormalised basis vectors)
// MODE 1: w -= sum_i S[c_off+i] V_i
// MODE 2: w  = sum_i S[c_off+i] V_i                  (dx assembly; w is write-only)
// S[n2_off] = ||w||^2 of the result.
// NV = vectors compiled in (nv <= NV at run time; the loop over vectors is fully unrolled so that the vector
// base pointers come straight from the kernel-parameter constant bank and the loads of a group of 8 vectors are
// issued back to back), U = double2 elements per thread per sweep.
template <int MODE, int NV, int U>
__global__ void __launch_bounds__(256, 2) maxpy_kernel(PtrList V, int nv, double* __restrict__ w, size_t n, double* S,
                                                       int c_off, int n2_off, int fuse_j, ReduceWs ws, P2PReduceArgs R,
                                                       int nvb) {
  if (S[JS_STOP] != 0.0) return; // (the host clears JS_STOP before it assembles dx with MODE 2)
  __shared__ double c[JF_MAXV];
  if (threadIdx.x < JF_MAXV) {
    double cv = 0.0;
    if (threadIdx.x < nv) {
      cv = S[c_off + threadIdx.x];
      if (MODE == 0) cv /= S[JS_VN2 + threadIdx.x];
    }
    c[threadIdx.x] = (MODE == 2) ? cv : -cv;
  }
  __syncthreads();
  double acc[1] = {0.0};
  const size_t nb2 = (n / (size_t)nvb) >> 1; // double2 elements per virtual block (rank-count-independent norm, cuda_common.cuh)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int vb = 0; vb < nvb; ++vb) {
  const size_t n2 = (size_t)(vb + 1) * nb2; // this block is the index range [vb nb2, n2)
  acc[0] = 0.0;
  for (size_t i0 = (size_t)vb * nb2 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n2; i0 += stride * U) {
    double2 t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      t[u] = make_double2(0.0, 0.0);
      if (MODE != 2 && i < n2) t[u] = *reinterpret_cast<const double2*>(w + 2 * i);
    }
#pragma unroll
    for (int kb = 0; kb < NV; kb += 8) {
      if (kb < nv) {
        double2 v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = i0 + (size_t)u * stride;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[u][q] = make_double2(0.0, 0.0);
            if (kb + q < NV && kb + q < nv && i < n2) v[u][q] = ldg2(V.p[kb + q] + 2 * i);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (kb + q < NV) {
            const double ck = c[kb + q]; // 0 beyond nv
#pragma unroll
            for (int u = 0; u < U; ++u) {
              t[u].x = fma(ck, v[u][q].x, t[u].x);
              t[u].y = fma(ck, v[u][q].y, t[u].y);
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < n2) {
        stg2(w + 2 * i, t[u]);
        acc[0] = fma(t[u].y, t[u].y, fma(t[u].x, t[u].x, acc[0]));
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { // odd tail element (only with nvb == 1)
    double tt = (MODE == 2) ? 0.0 : w[n - 1];
    for (int k = 0; k < nv; ++k) tt = fma(c[k], V.p[k][n - 1], tt);
    w[n - 1] = tt;
    acc[0] = fma(tt, tt, acc[0]);
  }
  seg_store<1>(acc, 0, vb, ws);
  } // virtual blocks
  const bool last = seg_finalize(1, nvb, ws, S + n2_off);
  // The finalising CTA goes on: with slab ranks it all-reduces the norm over peer memory and then (R.givens_j >= 0) runs
  // the Givens step of the column on the reduced value; on one GPU its thread 0 (which just wrote the norm) runs the
  // Givens step of column fuse_j directly.  Either way an Arnoldi step is operator + multi-dot + this kernel.
  if (last) {
    if (R.nranks > 1) p2p_allreduce_block(R);
    else if (MODE == 0 && fuse_j >= 0 && threadIdx.x == 0) hess_givens_step(S, fuse_j, R.givens_taken, R.givens_rerun);
  }
}

// scalar-load variant for operands that are not 16-byte aligned
template <int MODE>
__global__ void __launch_bounds__(256) maxpy_scalar_kernel(PtrList V, int nv, double* __restrict__ w, size_t n, double* S,
                                                           int c_off, int n2_off, int fuse_j, ReduceWs ws, P2PReduceArgs R) {
  if (S[JS_STOP] != 0.0) return;
  __shared__ double c[JF_MAXV];
  __shared__ const double* vp[JF_MAXV];
  if (threadIdx.x < nv) {
    double cv = S[c_off + threadIdx.x];
    if (MODE == 0) cv /= S[JS_VN2 + threadIdx.x];
    c[threadIdx.x] = (MODE == 2) ? cv : -cv;
    vp[threadIdx.x] = V.p[threadIdx.x];
  }
  __syncthreads();
  double acc[1] = {0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double t = (MODE == 2) ? 0.0 : w[i];
    for (int k = 0; k < nv; ++k) t = fma(c[k], vp[k][i], t);
    w[i] = t;
    acc[0] = fma(t, t, acc[0]);
  }
  const bool last = grid_reduce<1>(acc, 0u, ws, S + n2_off);
  if (last) {
    if (R.nranks > 1) p2p_allreduce_block(R);
    else if (MODE == 0 && fuse_j >= 0 && threadIdx.x == 0) hess_givens_step(S, fuse_j, R.givens_taken, R.givens_rerun);
  }
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

__global__ void givens_kernel(double* S, int j, int taken, int rerun, int skippable) {
  if (skippable && S[JS_STOP] != 0.0) return;
  hess_givens_step(S, j, taken, rerun);
}

struct IdxList {
  int v[JF_MAXV];
};
__global__ void lsq_kernel(double* S, int nit, IdxList zn2, int scale_n2_idx) { lsq_solve(S, nit, zn2.v, scale_n2_idx); }

} // namespace jfnk
