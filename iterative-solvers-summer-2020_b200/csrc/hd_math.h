// Arithmetic shared verbatim by the CUDA kernels and the CPU test double (tests/hostsim):
// pointwise Swift-Hohenberg formulas and the Hessenberg / Givens / LSQ scalar kernels.
// Everything here is __host__ __device__ so that the formulas the GPU executes are the ones the
// CPU-side logic tests exercise.
#pragma once
#include <math.h>
#include "scalars.h"

#if defined(__CUDACC__)
#define JF_HD __host__ __device__ __forceinline__
#define JF_UNROLL _Pragma("unroll")
#else
#define JF_HD inline
#define JF_UNROLL
#endif

namespace jfnk {

// ---- Swift-Hohenberg ------------------------------------------------------------------------
// L = -Lap*Lap - 2*Lap + (r-1)*I on a periodic grid, e = 1/h^2 (sh_scipy_nk.py:32-39) is the
// 13-point diamond stencil
//   centre  c0 = -20e^2 + 8e + (r-1),  axial +-1  c1 = 8e^2 - 2e,  diagonal c2 = -2e^2,  axial +-2  c3 = -e^2
struct SHParams {
  double c0, c1, c2, c3; // 13-point coefficients of L
  double e;              // 1/h^2 (5-point Laplacian)
  double g, k, r;        // PDE parameters, time step
  double inv_k;          // 1/k
};

JF_HD SHParams make_sh_params(double h, double r, double g, double k) {
  SHParams p;
  double e = 1.0 / (h * h);
  p.e = e;
  p.c0 = -20.0 * e * e + 8.0 * e + (r - 1.0);
  p.c1 = 8.0 * e * e - 2.0 * e;
  p.c2 = -2.0 * e * e;
  p.c3 = -e * e;
  p.g = g; p.k = k; p.r = r; p.inv_k = 1.0 / k;
  return p;
}

// L u at one point from the grouped neighbour sums:
//   s1 = u(y,x-1)+u(y,x+1)+u(y-1,x)+u(y+1,x), sd = the four diagonals, s2 = the four distance-2 axial points.
JF_HD double sh_apply13(const SHParams& p, double uc, double s1, double sd, double s2) {
  return p.c0 * uc + p.c1 * s1 + p.c2 * sd + p.c3 * s2;
}
JF_HD double sh_apply5(const SHParams& p, double uc, double s1) { return p.e * s1 - 4.0 * p.e * uc; }

// N(u) = L u + g u^2 - u^3   (the bracket of sh_scipy_nk.py:49, one time level)
JF_HD double sh_nonlin(const SHParams& p, double u, double Lu) {
  double uu = u * u;
  return Lu + p.g * uu - u * uu;
}
// G(u) = u/k - N(u)/2 ;  F(u) = G(u) - d  with  d = Uo/k + N(Uo)/2   (sh_scipy_nk.py:49 regrouped so the
// previous-time-level terms are one per-step constant field d).
// (1/k is applied as a multiplication by inv_k: one rounding like the division, no fp64 divide in the hot loop.)
JF_HD double sh_G(const SHParams& p, double u, double Lu) { return u * p.inv_k - sh_nonlin(p, u, Lu) * 0.5; }
JF_HD double sh_prev_const(const SHParams& p, double uo, double Luo) { return uo * p.inv_k + sh_nonlin(p, uo, Luo) * 0.5; }

// Linearly-implicit Swift-Hohenberg step (sh_linearised.py:51-57):
//   D = (5U - Uo)^2 k/16 - g k U ;  b = (I + L k/2) U ;  A z = (I + D - L k/2) z
JF_HD double shlin_diag(const SHParams& p, double u, double uo) {
  double t = 5.0 * u - uo;
  return t * t * p.k / 16.0 - p.g * p.k * u;
}
JF_HD double shlin_rhs(const SHParams& p, double u, double Lu) { return u + Lu * p.k / 2.0; }
JF_HD double shlin_apply(const SHParams& p, double z, double Lz, double D) { return z + D * z - Lz * p.k / 2.0; }

// ---- Arnoldi Hessenberg column / Givens update (one thread) -----------------------------------
// Mirrors what qr_insert + |Q[0,-1]| compute in _fgmres (_gcrotmk.py:149-168) for the growing
// Hessenberg matrix, in the standard GMRES form: rotations cs/sn, rotated rhs g (g[0]=1).
// S = scalar arena.  j = Arnoldi index (column), so the column has j+2 entries.
// taken != 0: a second Gram-Schmidt pass was made (RD2, HN2B valid).  rerun != 0: the step for this j already ran
// once after the first pass (the host then decided to re-orthogonalise): restart from the saved rotated rhs.
// h_i = (V_i . w)/||V_i|| with V_i stored unnormalised (first-pass dots plus the second pass's when one was taken)
JF_HD double hess_entry(const double* S, int i, int taken) {
  double h = S[JS_RD + i];
  if (taken) h += S[JS_RD2 + i];
  return h / sqrt(S[JS_VN2 + i]);
}
// hpre: the column entries h_0..h_j already evaluated with hess_entry (a kernel whose threads do that in parallel), or null
JF_HD void hess_givens_step(double* S, int j, int taken, int rerun, const double* hpre = nullptr) {
  const double eps = 2.220446049250313e-16;
  double* rd = S + JS_RD;
  double* vn2 = S + JS_VN2;
  double* cs = S + JS_CS;
  double* sn = S + JS_SN;
  double* g = S + JS_G;
  double* Rcol = S + JS_R + (size_t)j * JF_MAXV;
  int flags = 0;
  double ww = rd[j + 1];
  S[JS_WW] = ww;
  double hn2 = taken ? S[JS_HN2B] : S[JS_HN2A];
  if (taken) flags |= JF_FLAG_REORTH;
  if (j == 0) g[0] = 1.0;
  if (rerun) g[j] = S[JS_GJ_SAVE];
  else S[JS_GJ_SAVE] = g[j];
  double hprev = 0.0;
  for (int i = 0; i <= j; ++i) {
    double h = hpre ? hpre[i] : hess_entry(S, i, taken);
    if (i > 0) {
      // apply rotation i-1 to (hprev, h)
      double t = cs[i - 1] * hprev + sn[i - 1] * h;
      h = -sn[i - 1] * hprev + cs[i - 1] * h;
      Rcol[i - 1] = t;
    }
    hprev = h;
  }
  double hnext = sqrt(hn2);
  vn2[j + 1] = hn2;
  if (!(hnext > eps * sqrt(ww))) flags |= JF_FLAG_BREAKDOWN; // _gcrotmk.py:137-141
  // new rotation annihilating hnext
  double c, s, rr;
  if (hnext == 0.0) { c = 1.0; s = 0.0; rr = hprev; }
  else {
    rr = hypot(hprev, hnext);
    c = hprev / rr; s = hnext / rr;
  }
  cs[j] = c; sn[j] = s;
  Rcol[j] = rr;
  double gj = g[j];
  g[j] = c * gj;
  g[j + 1] = -s * gj;
  const double res = fabs(g[j + 1]);
  S[JS_RES] = res;
  if (!isfinite(rr) || !isfinite(ww)) flags |= JF_FLAG_NONFINITE;
  // the decisions the host would take from these scalars (engine.cpp, Engine::cycle), taken here so that work the host has
  // already enqueued for the next Arnoldi step can be dropped on the device
  const double tau2 = S[JS_TAU2];
  if (!taken && tau2 > 0.0 && hn2 < tau2 * ww) flags |= JF_FLAG_NEED_REORTH;
  const bool stop = (flags & (JF_FLAG_NEED_REORTH | JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) != 0 || res < S[JS_PTOL];
  S[JS_FLAGS] = (double)flags;
  S[JS_STOP] = stop ? 1.0 : 0.0;
  double* rec = S + JS_REC + (size_t)j * JF_REC_STRIDE;
  rec[0] = ww; rec[1] = S[JS_HN2A]; rec[2] = taken ? S[JS_HN2B] : S[JS_HN2A]; rec[3] = res; rec[4] = (double)flags;
}

// y = R^{-1} g (nit x nit), y *= scale ; coef[i] = y[i] / sqrt(S[zn2_idx[i]])  (dx = sum coef_i Z_i with
// unnormalised Z_i).  A zero pivot (Arnoldi breakdown) yields y_i = 0, the minimum-norm choice SciPy's
// lstsq makes (_gcrotmk.py:176-179).
JF_HD void lsq_solve(double* S, int nit, const int* zn2_idx, int scale_n2_idx) {
  double* y = S + JS_Y;
  const double* g = S + JS_G;
  const double* R = S + JS_R;
  for (int i = nit - 1; i >= 0; --i) {
    double acc = g[i];
    for (int c = i + 1; c < nit; ++c) acc -= R[i + (size_t)c * JF_MAXV] * y[c];
    double piv = R[i + (size_t)i * JF_MAXV];
    y[i] = (piv != 0.0) ? acc / piv : 0.0;
  }
  double scale = sqrt(S[scale_n2_idx]);
  for (int i = 0; i < nit; ++i) {
    y[i] *= scale;
    S[JS_COEF + i] = y[i] / sqrt(S[zn2_idx[i]]);
  }
}

} // namespace jfnk
