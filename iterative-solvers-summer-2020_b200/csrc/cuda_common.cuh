// Common device helpers: streaming 128-bit loads, warp/block reductions and the deterministic
// "last block finalises" cross-block reduction used by every norm / dot kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_ops.h"

namespace jfnk {

constexpr int kMaxBlocks = 148 * 8; // persistent-style grids: a multiple of the B200's 148 SMs
constexpr int kPartialStride = JF_MAXV + 4;

struct PtrList {
  const double* p[JF_MAXV];
};

constexpr int kMaxVirtualBlocks = 8; // = the largest slab-rank count: see "rank-count-independent sums" below

struct ReduceWs {
  double* partials;   // [kMaxVirtualBlocks][kMaxBlocks * kPartialStride]
  unsigned* ticket;   // zero between launches
};

// fixed pairwise tree over n = 2^k values (n <= 8): ((v0+v1)+(v2+v3)) + ((v4+v5)+(v6+v7)).  The order the slab ranks'
// partial sums are combined in -- by the all-reduce over P ranks and, on fewer ranks, over the virtual blocks of a rank.
__device__ __forceinline__ double tree_sum(const double* v, int n) {
  if (n == 8) return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  if (n == 4) return (v[0] + v[1]) + (v[2] + v[3]);
  if (n == 2) return v[0] + v[1];
  double s = v[0];
  for (int i = 1; i < n; ++i) s += v[i];
  return s;
}

// read-only streaming loads (data is touched once per kernel; keep it out of L1)
__device__ __forceinline__ double2 ldg2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ldg1(const double* p) { return __ldg(p); }
__device__ __forceinline__ void stg2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// streaming store: the line is written once and not read again by this kernel (evict-first in L2)
__device__ __forceinline__ void stg2_cs(double* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// x + a*v exactly as NumPy evaluates `x0 + sc*v` (product rounded, then sum rounded; never contracted to an FMA)
__device__ __forceinline__ double combine(double x, double a, double v) { return __dadd_rn(x, __dmul_rn(a, v)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-reduce K per-thread values (op: 0 sum, 1 max per entry, given by the bit mask `maxmask`), write this
// block's partials, and let the LAST block to arrive reduce all partials in a fixed order and store the K
// results to out[0..K).  Deterministic for a fixed grid.  Must be called by all threads of all blocks.
template <int K>
__device__ __forceinline__ bool grid_reduce(double (&val)[K], unsigned maxmask, const ReduceWs& ws, double* out) {
  __shared__ double sm[32][K];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = ((maxmask >> k) & 1u) ? warp_max(val[k]) : warp_sum(val[k]);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      bool mx = (maxmask >> k) & 1u;
      double v = (lane < nwarp) ? sm[lane][k] : 0.0; // max entries are non-negative (abs values)
      v = mx ? warp_max(v) : warp_sum(v);
      if (lane == 0) ws.partials[(size_t)blockIdx.x * kPartialStride + k] = v;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int k = warp; k < K; k += nwarp) {
      bool mx = (maxmask >> k) & 1u;
      double v = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) {
        double pv = __ldcg(&ws.partials[(size_t)b * kPartialStride + k]);
        v = mx ? fmax(v, pv) : v + pv;
      }
      v = mx ? warp_max(v) : warp_sum(v);
      if (lane == 0) out[k] = v;
    }
    if (threadIdx.x == 0) *ws.ticket = 0u;
  }
  return is_last; // true in every thread of the CTA that finalised (its thread 0 wrote out[0])
}

// ---- rank-count-independent sums -------------------------------------------------------------------------------------------
// With row slabs the order in which a global sum is formed depends on the number of ranks P, and the finite-difference
// Jacobian amplifies that rounding-level difference into ~1e-8 of the field after a few time steps: runs on 1, 2, 4 and 8
// GPUs are then "equal up to noise" instead of equal.  The streaming reductions therefore always sum the GLOBAL vector as
// 8 VIRTUAL BLOCKS of rows (the slabs of an 8-rank run): a rank that owns nvb = 8/P of them sweeps them one after the other
// with the grid and the element -> thread mapping an 8-rank run would use for that block, keeps one set of per-CTA partials
// per block, and its finalising CTA reduces every block over the CTAs in the fixed order and then combines the blocks --
// and the all-reduce combines the ranks -- in one fixed pairwise tree (tree_sum).  The sums are bit-identical for
// P = 1, 2, 4, 8; everything else on the path is pointwise, so the fields are too.
// seg_store: this CTA's partial sums of virtual block vb ; seg_finalize: ticket, then the last CTA does the two-level sum.
template <int KMAX>
__device__ __forceinline__ void seg_store(double (&val)[KMAX], int nv, int vb, const ReduceWs& ws) {
  __shared__ double sm[8][KMAX];
  const int K = nv + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5; // nwarp <= 8
#pragma unroll
  for (int k = 0; k < KMAX - 1; ++k) {
    if (k < nv) {
      double v = warp_sum(val[k]);
      if (lane == 0) sm[warp][k] = v;
    }
  }
  {
    double v = warp_sum(val[KMAX - 1]);
    if (lane == 0) sm[warp][nv] = v;
  }
  __syncthreads();
  double* dst = ws.partials + ((size_t)vb * kMaxBlocks + blockIdx.x) * kPartialStride;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double v = 0.0;
    for (int w = 0; w < nwarp; ++w) v += sm[w][k];
    dst[k] = v;
  }
  __syncthreads(); // (sm is reused by the next virtual block)
}
__device__ __forceinline__ bool seg_finalize(int K, int nvb, const ReduceWs& ws, double* out) {
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int k = warp; k < K; k += nwarp) {
      double vs[kMaxVirtualBlocks];
#pragma unroll
      for (int vb = 0; vb < kMaxVirtualBlocks; ++vb) {
        vs[vb] = 0.0;
        if (vb < nvb) {
          double v = 0.0;
          for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(&ws.partials[((size_t)vb * kMaxBlocks + b) * kPartialStride + k]);
          vs[vb] = warp_sum(v);
        }
      }
      if (lane == 0) out[k] = tree_sum(vs, nvb);
    }
    if (threadIdx.x == 0) *ws.ticket = 0u;
  }
  return is_last;
}

// multi-dot variant: val[0..nv) are dot accumulators (nv <= KMAX-1 at run time), val[KMAX-1] is w.w.
// Writes out[0..nv) = dots, out[nv] = w.w.  All array indices are compile-time so val[] stays in registers.
template <int KMAX>
__device__ __forceinline__ bool grid_reduce_sums(double (&val)[KMAX], int nv, const ReduceWs& ws, double* out) {
  __shared__ double sm[8][KMAX];
  __shared__ bool is_last;
  const int K = nv + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5; // nwarp <= 8
#pragma unroll
  for (int k = 0; k < KMAX - 1; ++k) {
    if (k < nv) {
      double v = warp_sum(val[k]);
      if (lane == 0) sm[warp][k] = v;
    }
  }
  {
    double v = warp_sum(val[KMAX - 1]);
    if (lane == 0) sm[warp][nv] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double v = 0.0;
    for (int w = 0; w < nwarp; ++w) v += sm[w][k];
    ws.partials[(size_t)blockIdx.x * kPartialStride + k] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int k = warp; k < K; k += nwarp) {
      double v = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(&ws.partials[(size_t)b * kPartialStride + k]);
      v = warp_sum(v);
      if (lane == 0) out[k] = v;
    }
    if (threadIdx.x == 0) *ws.ticket = 0u;
  }
  return is_last; // true in every thread of the CTA that finalised
}

} // namespace jfnk
