// Matrix-free Swift-Hohenberg stencil kernels (periodic 5-point Laplacian and the 13-point
// L = -Lap^2 - 2 Lap + (r-1) I), with the nonlinear terms, the Crank-Nicolson combination, the
// finite-difference Jacobian-vector product and the norms fused into the same pass.
//
// Two kernels share one arithmetic core (sh_value):
//  * sh_march_kernel -- the product path for large even grids.  Each warp owns a 64-column strip
//    (one double2 per lane) and marches down a chunk of rows.  A row is read from HBM exactly once per
//    chunk with coalesced 128-bit loads; the horizontal neighbours travel through a double-buffered
//    shared-memory row (one __syncwarp per row, no block barrier), the vertical neighbours stay in a
//    register window (5 rows of u, the horizontal pair sums of 4 rows).  The input combination
//    t = x + a v of the JVP / line search is formed at load time, so F(x0 + sc z) never exists in memory.
//  * sh_point_kernel -- one thread per grid point straight from global memory; used for small / odd
//    grids (61 x 61, 91 x 61 ...) where launch latency, not bandwidth, is the bound, and as an independent
//    cross-check of the marching kernel in the parity tests (kernel_variant = 1).
//
// Algorithmic bytes per grid point (fp64): spmv 16, set_prev 16, residual 24 (+8 when x+s dx is stored),
// JVP 40 (x0, z, d, f0 in; w out), linearised matvec 24.
#pragma once
#include "cuda_common.cuh"

namespace jfnk {

enum ShOp { OP_LAP = 0, OP_L = 1, OP_SETPREV = 2, OP_RESID = 3, OP_JVP = 4, OP_LINPREP = 5, OP_LINMV = 6 };

struct ShArgs {
  const double *x, *xtop, *xbot; // primary field and its 2-row halos (top = rows -2,-1; bot = rows nrows, nrows+1)
  const double *v, *vtop, *vbot; // optional second field: t = x + a v
  ScalarRef a;                   // combination coefficient
  ScalarRef div;                 // JVP divisor ; LINMV output scale
  const double* d;               // RESID/JVP: per-step constant d ; LINMV: diagonal D
  const double* f0;              // JVP: f0 ; LINPREP: Uo
  double* out;                   // result field
  double* out2;                  // RESID: x + a v (may be null) ; LINPREP: D
  int nx, nrows;
  int norm_off;                  // RESID: S[norm_off..+2] = sum F^2, max|F|, max|t|
  int ry;                        // marching kernel: rows per chunk
};

__device__ __forceinline__ const double* sh_row(const double* base, const double* top, const double* bot, int r, int nx,
                                                int nrows) {
  if (r < 0) return top + (size_t)(r + 2) * nx;
  if (r >= nrows) return bot + (size_t)(r - nrows) * nx;
  return base + (size_t)r * nx;
}

// x + a*v exactly as NumPy evaluates `x0 + sc*v` (product rounded, then sum rounded)
__device__ __forceinline__ double combine(double x, double a, double v) { return __dadd_rn(x, __dmul_rn(a, v)); }

struct ShAcc {
  double f2, fmax, xmax;
};

// one output value (returned) plus the optional second output (RESID: x + a v ; LINPREP: D)
template <int OP>
__device__ __forceinline__ double sh_value(const SHParams& P, double scale, double uc, double s1, double sd, double s2,
                                           double dval, double f0val, double& second, ShAcc& acc) {
  if (OP == OP_LAP) return sh_apply5(P, uc, s1);
  double Lu = sh_apply13(P, uc, s1, sd, s2);
  if (OP == OP_L) return Lu;
  if (OP == OP_SETPREV) return sh_prev_const(P, uc, Lu);
  if (OP == OP_RESID) {
    double F = sh_G(P, uc, Lu) - dval;
    second = uc;
    acc.f2 = fma(F, F, acc.f2);
    acc.fmax = fmax(acc.fmax, fabs(F));
    acc.xmax = fmax(acc.xmax, fabs(uc));
    return F;
  }
  if (OP == OP_JVP) {
    double F = sh_G(P, uc, Lu) - dval;
    return (F - f0val) * scale; // scale = 1/div
  }
  if (OP == OP_LINPREP) {
    second = shlin_diag(P, uc, f0val);
    return shlin_rhs(P, uc, Lu);
  }
  return scale * shlin_apply(P, uc, Lu, dval); // OP_LINMV
}

template <int OP>
__device__ __forceinline__ double sh_scale(const ShArgs& A, const double* S) {
  if (OP == OP_JVP) return 1.0 / eval_sref(S, A.div);
  if (OP == OP_LINMV) return eval_sref(S, A.div);
  return 1.0;
}

// ---------------------------------------------------------------------------------------------------
// one thread per point
// ---------------------------------------------------------------------------------------------------
template <int OP, bool HAS_V>
__global__ void __launch_bounds__(256) sh_point_kernel(ShArgs A, SHParams P, double* S, ReduceWs ws) {
  const int nx = A.nx, nrows = A.nrows;
  const size_t n = (size_t)nx * nrows;
  const double a = HAS_V ? eval_sref(S, A.a) : 0.0;
  const double scale = sh_scale<OP>(A, S);
  ShAcc acc = {0.0, 0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    int r = (int)(e / nx), c = (int)(e - (size_t)r * nx);
    int cm1 = c - 1 < 0 ? c - 1 + nx : c - 1, cm2 = c - 2 < 0 ? c - 2 + nx : c - 2;
    int cp1 = c + 1 >= nx ? c + 1 - nx : c + 1, cp2 = c + 2 >= nx ? c + 2 - nx : c + 2;
    auto at = [&](int rr, int cc) -> double {
      double xv = sh_row(A.x, A.xtop, A.xbot, rr, nx, nrows)[cc];
      if (HAS_V) xv = combine(xv, a, sh_row(A.v, A.vtop, A.vbot, rr, nx, nrows)[cc]);
      return xv;
    };
    double uc = at(r, c);
    double a1 = at(r, cm1) + at(r, cp1);
    double a2 = at(r, cm2) + at(r, cp2);
    double a1u = at(r - 1, cm1) + at(r - 1, cp1);
    double a1d = at(r + 1, cm1) + at(r + 1, cp1);
    double s1 = a1 + at(r - 1, c) + at(r + 1, c);
    double sd = a1u + a1d;
    double s2 = a2 + at(r - 2, c) + at(r + 2, c);
    double dval = (OP == OP_RESID || OP == OP_JVP || OP == OP_LINMV) ? A.d[e] : 0.0;
    double f0val = (OP == OP_JVP || OP == OP_LINPREP) ? A.f0[e] : 0.0;
    double second = 0.0;
    A.out[e] = sh_value<OP>(P, scale, uc, s1, sd, s2, dval, f0val, second, acc);
    if ((OP == OP_RESID && A.out2) || OP == OP_LINPREP) A.out2[e] = second;
  }
  if (OP == OP_RESID) {
    double val[3] = {acc.f2, acc.fmax, acc.xmax};
    grid_reduce<3>(val, 0x6u, ws, S + A.norm_off);
  }
}

// ---------------------------------------------------------------------------------------------------
// marching kernel: warp = 64-column strip x ry rows
// ---------------------------------------------------------------------------------------------------
constexpr int kMarchWarps = 4;

template <int OP, bool HAS_V>
__global__ void __launch_bounds__(kMarchWarps * 32) sh_march_kernel(ShArgs A, SHParams P, double* S, ReduceWs ws) {
  __shared__ __align__(16) double smrow[kMarchWarps][2][72];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = A.nx, nrows = A.nrows, ry = A.ry;
  const int strips = (nx + 63) >> 6;
  const int chunks = (nrows + ry - 1) / ry;
  const long long wid = (long long)blockIdx.x * kMarchWarps + warp;
  const bool wact = wid < (long long)strips * chunks;
  const int strip = wact ? (int)(wid % strips) : 0;
  const int chunk = wact ? (int)(wid / strips) : 0;
  const int xs = strip << 6;
  const int W = min(64, nx - xs); // even
  const int y0 = chunk * ry, y1 = min(y0 + ry, nrows);
  const bool act = wact && (2 * lane < W);
  const bool hact = wact && lane == 0;
  const int c = xs + 2 * lane;
  const int cl = xs - 2 < 0 ? xs - 2 + nx : xs - 2;
  const int cr = xs + W >= nx ? xs + W - nx : xs + W;
  const double a = HAS_V ? eval_sref(S, A.a) : 0.0;
  const double scale = sh_scale<OP>(A, S);
  ShAcc acc = {0.0, 0.0, 0.0};

  auto load2 = [&](int r, int col) -> double2 {
    double2 t = ldg2(sh_row(A.x, A.xtop, A.xbot, r, nx, nrows) + col);
    if (HAS_V) {
      double2 vv = ldg2(sh_row(A.v, A.vtop, A.vbot, r, nx, nrows) + col);
      t.x = combine(t.x, a, vv.x);
      t.y = combine(t.y, a, vv.y);
    }
    return t;
  };
  const double2 zero2 = make_double2(0.0, 0.0);
  double2 u0 = zero2, u1 = zero2, u2 = zero2, u3 = zero2, u4 = zero2;
  double2 p0 = zero2, p1 = zero2, p2 = zero2, p3 = zero2; // horizontal +-1 pair sums, rows ra-3..ra
  double2 q0 = zero2, q1 = zero2, q2 = zero2;             // horizontal +-2 pair sums, rows ra-2..ra
  // software pipeline: row ra+1 is in flight while row ra is exchanged and consumed
  double2 nown = zero2, nhl = zero2, nhr = zero2;
  if (act) nown = load2(y0 - 2, c);
  if (hact) { nhl = load2(y0 - 2, cl); nhr = load2(y0 - 2, cr); }
  for (int ra = y0 - 2; ra < y1 + 2; ++ra) {
    double2 own = nown, hl = nhl, hr = nhr;
    if (ra + 1 < y1 + 2) {
      if (act) nown = load2(ra + 1, c);
      if (hact) { nhl = load2(ra + 1, cl); nhr = load2(ra + 1, cr); }
    }
    const int y = ra - 2;
    const bool emit = act && y >= y0;
    const size_t e = (size_t)(emit ? y : 0) * nx + c;
    // operands of the output row: issue their loads before the exchange so they overlap it
    double2 dv = zero2, fv = zero2;
    if (emit) {
      if (OP == OP_RESID || OP == OP_JVP || OP == OP_LINMV) dv = ldg2(A.d + e);
      if (OP == OP_JVP || OP == OP_LINPREP) fv = ldg2(A.f0 + e);
    }
    double* s = smrow[warp][(ra - y0) & 1];
    if (act) *reinterpret_cast<double2*>(s + 2 + 2 * lane) = own;
    if (hact) {
      *reinterpret_cast<double2*>(s) = hl;
      *reinterpret_cast<double2*>(s + 2 + W) = hr;
    }
    __syncwarp();
    double2 L = zero2, R = zero2;
    if (act) {
      L = *reinterpret_cast<const double2*>(s + 2 * lane);
      R = *reinterpret_cast<const double2*>(s + 4 + 2 * lane);
    }
    u0 = u1; u1 = u2; u2 = u3; u3 = u4; u4 = own;
    p0 = p1; p1 = p2; p2 = p3;
    p3.x = L.y + own.y; p3.y = own.x + R.x;
    q0 = q1; q1 = q2;
    q2.x = L.x + R.x; q2.y = L.y + R.y;
    if (emit) {
      // rows: u0..u4 = y-2..y+2 ; p0,p1,p2 = pair sums of rows y-1,y,y+1 ; q0 = +-2 pair sum of row y
      double s1x = p1.x + u1.x + u3.x, s1y = p1.y + u1.y + u3.y;
      double sdx = p0.x + p2.x, sdy = p0.y + p2.y;
      double s2x = q0.x + u0.x + u4.x, s2y = q0.y + u0.y + u4.y;
      double2 o, o2 = zero2;
      o.x = sh_value<OP>(P, scale, u2.x, s1x, sdx, s2x, dv.x, fv.x, o2.x, acc);
      o.y = sh_value<OP>(P, scale, u2.y, s1y, sdy, s2y, dv.y, fv.y, o2.y, acc);
      stg2(A.out + e, o);
      if ((OP == OP_RESID && A.out2) || OP == OP_LINPREP) stg2(A.out2 + e, o2);
    }
  }
  if (OP == OP_RESID) {
    double val[3] = {acc.f2, acc.fmax, acc.xmax};
    grid_reduce<3>(val, 0x6u, ws, S + A.norm_off);
  }
}

} // namespace jfnk
