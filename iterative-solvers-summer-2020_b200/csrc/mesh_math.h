// Matrix-free 4th-order finite differences on the non-periodic computational (ksi, eta) grid of the
// moving-mesh problems, shared verbatim by the CUDA kernels and the CPU test double.
//
// Reference: make_M (PMA2_nk.py:181-233, droplet.py:778-833) builds these operators as sparse
// matrices with kron(); here every row of those matrices is applied on the fly:
//   D1  first derivative  [1,-8,0,8,-1]/(12h),   one-sided closures rows 0,1,n-2,n-1 (droplet.py:801-804)
//   D2  second derivative [-1,16,-30,16,-1]/(12h^2), closures rows 0,1,n-2,n-1      (droplet.py:785-790)
//   Dxy = D1(eta) (x) D1(ksi)                                                        (droplet.py:810)
// Layout: flat index = row*nx + col, row = eta index (0 = bottom edge), col = ksi index (0 = left edge).
#pragma once
#include "device_ops.h"

namespace jfnk {

// 1-D finite-difference weight tables (row type x tap), already divided by 12h / 12h^2.  They live in device
// (or, for the CPU test double, host) memory and are reached through a pointer: indexing a by-value kernel
// parameter with a run-time row type would make every thread copy the whole struct to local memory.
struct MeshTables {
  double d1x[5 * 5], d1y[5 * 5]; // [row type][k]
  double d2x[5 * 6], d2y[5 * 6];
};

struct MeshGeom {
  int nx, ny;
  double dksi, deta, dksi2, deta2;
  double bl, br, bb, bt;
  double c1x[5], c1y[5]; // centred first-derivative weights (row type 2), by value for the interior fast path
  const MeshTables* tab;
};

// row type of index i in a dimension of n points: 0,1 = first two rows, 3,4 = last two rows, 2 = interior
JF_HD int fd_type(int i, int n) { return i == 0 ? 0 : (i == 1 ? 1 : (i == n - 2 ? 3 : (i == n - 1 ? 4 : 2))); }
// first column touched by the 5-wide D1 row / the 6-wide D2 row
JF_HD int d1_start(int i, int n, int type) { return type < 2 ? 0 : (type > 2 ? n - 5 : i - 2); }
JF_HD int d2_start(int i, int n, int type) { return type < 2 ? 0 : (type > 2 ? n - 6 : i - 2); }
JF_HD int d2_count(int type) { return type == 2 ? 5 : 6; }

inline void fill_mesh_tables(const MeshParams& mp, MeshTables& t) {
  const double dksi2 = mp.dksi * mp.dksi, deta2 = mp.deta * mp.deta;
  const double c1[5][5] = {{-25, 48, -36, 16, -3}, {-3, -10, 18, -6, 1}, {1, -8, 0, 8, -1}, {-1, 6, -18, 10, 3}, {3, -16, 36, -48, 25}};
  const double c2[5][6] = {{-415.0 / 6, 96, -36, 32.0 / 3, -1.5, 0}, {10, -15, -4, 14, -6, 1}, {-1, 16, -30, 16, -1, 0},
                           {1, -6, 14, -4, -15, 10}, {0, -1.5, 32.0 / 3, -36, 96, -415.0 / 6}};
  for (int r = 0; r < 5; ++r) {
    for (int k = 0; k < 5; ++k) { t.d1x[r * 5 + k] = c1[r][k] / (12 * mp.dksi); t.d1y[r * 5 + k] = c1[r][k] / (12 * mp.deta); }
    for (int k = 0; k < 6; ++k) { t.d2x[r * 6 + k] = c2[r][k] / (12 * dksi2); t.d2y[r * 6 + k] = c2[r][k] / (12 * deta2); }
  }
}

// `tab` must point to tables filled for the same MeshParams, in memory the executing side can read
inline MeshGeom make_geom(const MeshParams& mp, int nx, int ny, const MeshTables* tab) {
  MeshGeom g;
  g.nx = nx; g.ny = ny;
  g.dksi = mp.dksi; g.deta = mp.deta; g.dksi2 = mp.dksi * mp.dksi; g.deta2 = mp.deta * mp.deta;
  g.bl = mp.bl; g.br = mp.br; g.bb = mp.bb; g.bt = mp.bt;
  const double c1[5] = {1, -8, 0, 8, -1};
  for (int k = 0; k < 5; ++k) { g.c1x[k] = c1[k] / (12 * g.dksi); g.c1y[k] = c1[k] / (12 * g.deta); }
  g.tab = tab;
  return g;
}

// ---- plain derivatives -------------------------------------------------------------------------
JF_HD double d_ksi(const MeshGeom& g, const double* f, int r, int c) {
  int t = fd_type(c, g.nx), s = d1_start(c, g.nx, t);
  const double* p = f + (size_t)r * g.nx + s;
  const double* w = g.tab->d1x + t * 5;
  return w[0] * p[0] + w[1] * p[1] + w[2] * p[2] + w[3] * p[3] + w[4] * p[4];
}
JF_HD double d_eta(const MeshGeom& g, const double* f, int r, int c) {
  int t = fd_type(r, g.ny), s = d1_start(r, g.ny, t);
  const double* p = f + (size_t)s * g.nx + c;
  const double* w = g.tab->d1y + t * 5;
  size_t n = g.nx;
  return w[0] * p[0] + w[1] * p[n] + w[2] * p[2 * n] + w[3] * p[3 * n] + w[4] * p[4 * n];
}
JF_HD double d2_ksi(const MeshGeom& g, const double* f, int r, int c) {
  int t = fd_type(c, g.nx), s = d2_start(c, g.nx, t), cnt = d2_count(t);
  const double* w = g.tab->d2x + t * 6;
  if (t == 4) { s += 1; w += 1; cnt = 5; } // last row touches only the last 5 columns
  const double* p = f + (size_t)r * g.nx + s;
  double acc = 0.0;
  for (int k = 0; k < cnt; ++k) acc += w[k] * p[k];
  return acc;
}
JF_HD double d2_eta(const MeshGeom& g, const double* f, int r, int c) {
  int t = fd_type(r, g.ny), s = d2_start(r, g.ny, t), cnt = d2_count(t);
  const double* w = g.tab->d2y + t * 6;
  if (t == 4) { s += 1; w += 1; cnt = 5; }
  const double* p = f + (size_t)s * g.nx + c;
  double acc = 0.0;
  for (int k = 0; k < cnt; ++k) acc += w[k] * p[(size_t)k * g.nx];
  return acc;
}
// kron(D1 eta, D1 ksi) f
JF_HD double d_ksieta(const MeshGeom& g, const double* f, int r, int c) {
  int ty = fd_type(r, g.ny), sy = d1_start(r, g.ny, ty);
  int tx = fd_type(c, g.nx), sx = d1_start(c, g.nx, tx);
  const double* wy = g.tab->d1y + ty * 5;
  const double* wx = g.tab->d1x + tx * 5;
  double acc = 0.0;
  for (int a = 0; a < 5; ++a) {
    const double* p = f + (size_t)(sy + a) * g.nx + sx;
    for (int b = 0; b < 5; ++b) acc += (wy[a] * wx[b]) * p[b];
  }
  return acc;
}

// ---- mesh metric fields (compute_Q_spatial_ders + J + the A_ij of Laplace_operator) -------------
// PMA2_nk.py:235-248,87,275-277 ; droplet.py:696-711,376,613-615.
// M[0..6] = Q_ksiksi, Q_etaeta, Q_ksieta, J, A11, A22, A12
JF_HD void mesh_metrics_point(const MeshGeom& g, const double* Q, int r, int c, double* const* M) {
  size_t e = (size_t)r * g.nx + c;
  double qxx = d2_ksi(g, Q, r, c);
  if (c == 0) qxx += 25.0 / (6.0 * g.dksi) * fabs(g.bl);
  if (c == g.nx - 1) qxx += 25.0 / (6.0 * g.dksi) * fabs(g.br);
  double qyy = d2_eta(g, Q, r, c);
  if (r == g.ny - 1) qyy += 25.0 / (6.0 * g.deta) * fabs(g.bt);
  if (r == 0) qyy += 25.0 / (6.0 * g.deta) * fabs(g.bb);
  bool bdy = (r == 0 || c == 0 || r == g.ny - 1 || c == g.nx - 1);
  double qxy = bdy ? 0.0 : d_ksieta(g, Q, r, c);
  double J = qxx * qyy - qxy * qxy;
  M[0][e] = qxx; M[1][e] = qyy; M[2][e] = qxy; M[3][e] = J;
  M[4][e] = (qxy * qxy + qyy * qyy) / J;
  M[5][e] = (qxy * qxy + qxx * qxx) / J;
  M[6][e] = -(qxy * (qxx + qyy)) / J;
}

// ---- Laplace_operator (PMA2_nk.py:263-343, droplet.py:601-681) ---------------------------------
// conservative (A v')' along one grid line.  A_(k), V_(k): coefficient and field at index k of the line, i = index of
// the output point along the line, n = line length.  The accessors let the same rows be applied to a field that only
// exists as a combination x + a v (fused FD-JVP) as well as to a plain array.
template <class FA, class FV>
JF_HD double cons_line_g(FA A_, FV V_, int i, int n, double h2) {
  if (i == 0 || i == n - 1) return 0.0; // boundary columns stay 0
  if (i == 1)
    return A_(1) * (10 * V_(0) - 15 * V_(1) - 4 * V_(2) + 14 * V_(3) - 6 * V_(4) + V_(5)) / (12 * h2) +
           (-3 * V_(0) - 10 * V_(1) + 18 * V_(2) - 6 * V_(3) + V_(4)) *
               (-3 * A_(0) - 10 * A_(1) + 18 * A_(2) - 6 * A_(3) + A_(4)) / (144 * h2);
  if (i == 2)
    return A_(2) * (-V_(0) + 16 * V_(1) - 30 * V_(2) + 16 * V_(3) - V_(4)) / (12 * h2) +
           (V_(0) - 8 * V_(1) + 8 * V_(3) - V_(4)) * (A_(0) - 8 * A_(1) + 8 * A_(3) - A_(4)) / (144 * h2);
  int m = n - 1;
  if (i == n - 2)
    return A_(m - 1) * (10 * V_(m) - 15 * V_(m - 1) - 4 * V_(m - 2) + 14 * V_(m - 3) - 6 * V_(m - 4) + V_(m - 5)) / (12 * h2) +
           (3 * V_(m) + 10 * V_(m - 1) - 18 * V_(m - 2) + 6 * V_(m - 3) - V_(m - 4)) *
               (3 * A_(m) + 10 * A_(m - 1) - 18 * A_(m - 2) + 6 * A_(m - 3) - A_(m - 4)) / (144 * h2);
  if (i == n - 3)
    return A_(m - 2) * (-V_(m) + 16 * V_(m - 1) - 30 * V_(m - 2) + 16 * V_(m - 3) - V_(m - 4)) / (12 * h2) +
           (V_(m - 4) - 8 * V_(m - 3) + 8 * V_(m - 1) - V_(m)) * (A_(m - 4) - 8 * A_(m - 3) + 8 * A_(m - 1) - A_(m)) / (144 * h2);
  return (4 * (A_(i - 1) * (V_(i - 3) - 8 * V_(i - 2) + 8 * V_(i) - V_(i + 1))) -
          (-A_(i - 2) + 9 * A_(i - 1) + 9 * A_(i) - A_(i + 1)) * (V_(i - 2) - 27 * V_(i - 1) + 27 * V_(i) - V_(i + 1)) +
          (-A_(i - 1) + 9 * A_(i) + 9 * A_(i + 1) - A_(i + 2)) * (V_(i - 1) - 27 * V_(i) + 27 * V_(i + 1) - V_(i + 2)) -
          4 * (A_(i + 1) * (V_(i - 1) - 8 * V_(i) + 8 * V_(i + 2) - V_(i + 3)))) /
         (288 * h2);
}
// plain arrays: A, v sampled at stride `st`
JF_HD double cons_line(const double* A, const double* v, size_t st, int i, int n, double h2) {
  return cons_line_g([=](int k) { return A[(size_t)k * st]; }, [=](int k) { return v[(size_t)k * st]; }, i, n, h2);
}

// first derivatives of a field given by an accessor f(row, col)
template <class F>
JF_HD double d_ksi_g(const MeshGeom& g, F f, int r, int c) {
  int t = fd_type(c, g.nx), s = d1_start(c, g.nx, t);
  const double* w = g.tab->d1x + t * 5;
  return w[0] * f(r, s) + w[1] * f(r, s + 1) + w[2] * f(r, s + 2) + w[3] * f(r, s + 3) + w[4] * f(r, s + 4);
}
template <class F>
JF_HD double d_eta_g(const MeshGeom& g, F f, int r, int c) {
  int t = fd_type(r, g.ny), s = d1_start(r, g.ny, t);
  const double* w = g.tab->d1y + t * 5;
  return w[0] * f(s, c) + w[1] * f(s + 1, c) + w[2] * f(s + 2, c) + w[3] * f(s + 3, c) + w[4] * f(s + 4, c);
}

// Interior fast path of Laplace_operator: every stencil involved is the centred one (no closure rows/columns within
// reach: 4 <= r < ny-4, 4 <= c < nx-4), so the formula is straight-line code with compile-time weights -- the same
// arithmetic, in the same order, as the general path below.
JF_HD void mesh_laplace_interior(const MeshGeom& g, const double* const* M, const double* v, int r, int c, double& vxx,
                                 double& vyy) {
  const size_t nx = g.nx;
  const size_t e = (size_t)r * nx + c;
  const double* A11 = M[4] + e;
  const double* A22 = M[5] + e;
  const double* A12 = M[6] + e;
  const double* p = v + e;
#define VX(k) p[(k)]
#define VY(k) p[(ptrdiff_t)(k) * (ptrdiff_t)nx]
#define AX(k) A11[(k)]
#define AY(k) A22[(ptrdiff_t)(k) * (ptrdiff_t)nx]
  double xx = (4 * (AX(-1) * (VX(-3) - 8 * VX(-2) + 8 * VX(0) - VX(1))) -
               (-AX(-2) + 9 * AX(-1) + 9 * AX(0) - AX(1)) * (VX(-2) - 27 * VX(-1) + 27 * VX(0) - VX(1)) +
               (-AX(-1) + 9 * AX(0) + 9 * AX(1) - AX(2)) * (VX(-1) - 27 * VX(0) + 27 * VX(1) - VX(2)) -
               4 * (AX(1) * (VX(-1) - 8 * VX(0) + 8 * VX(2) - VX(3)))) /
              (288 * g.dksi2);
  double yy = (4 * (AY(-1) * (VY(-3) - 8 * VY(-2) + 8 * VY(0) - VY(1))) -
               (-AY(-2) + 9 * AY(-1) + 9 * AY(0) - AY(1)) * (VY(-2) - 27 * VY(-1) + 27 * VY(0) - VY(1)) +
               (-AY(-1) + 9 * AY(0) + 9 * AY(1) - AY(2)) * (VY(-1) - 27 * VY(0) + 27 * VY(1) - VY(2)) -
               4 * (AY(1) * (VY(-1) - 8 * VY(0) + 8 * VY(2) - VY(3)))) /
              (288 * g.deta2);
#undef VX
#undef VY
#undef AX
#undef AY
  // centred first derivative weights (row type 2), columns / rows -2,-1,+1,+2 (the weight of 0 is 0)
  const double* wx = g.c1x;
  const double* wy = g.c1y;
  const ptrdiff_t sx = (ptrdiff_t)nx;
  // D_ksi(A12 * D_eta v): sum_k wx[k] * A12(r, c+k-2) * sum_a wy[a] v(r+a-2, c+k-2)
  double accx = 0.0, accy = 0.0;
JF_UNROLL
  for (int k = 0; k < 5; ++k) {
    if (k == 2) continue;
    const double* q = p + (k - 2);
    double ve = wy[0] * q[-2 * sx] + wy[1] * q[-sx] + wy[2] * q[0] + wy[3] * q[sx] + wy[4] * q[2 * sx];
    accx += wx[k] * (A12[k - 2] * ve);
  }
JF_UNROLL
  for (int k = 0; k < 5; ++k) {
    if (k == 2) continue;
    const double* q = p + (ptrdiff_t)(k - 2) * sx;
    double vk = wx[0] * q[-2] + wx[1] * q[-1] + wx[2] * q[0] + wx[3] * q[1] + wx[4] * q[2];
    accy += wy[k] * (A12[(ptrdiff_t)(k - 2) * sx] * vk);
  }
  const double Jv = M[3][e];
  vxx = (xx + accx) / Jv;
  vyy = (yy + accy) / Jv;
}

// General path of Laplace_operator for a field given by an accessor f(row, col): closure rows / columns selected per
// point.  The first derivatives of v fed to the cross terms are what Laplace_operator's callers pass in: deriv_bc = 1
// reproduces compute_u_spatial_ders of droplet.py:719-722 (v_ksi = 0 on the left/right/bottom edges -- incl. its
// `U_dksi[Ibdy.Bottom] = 0` line -- and v_eta = 0 on the top edge).
template <class F>
JF_HD void mesh_laplace_general_g(const MeshGeom& g, const double* const* M, F f, int r, int c, int deriv_bc, double& vxx,
                                  double& vyy) {
  const double* J = M[3];
  const double* A11 = M[4] + (size_t)r * g.nx;
  const double* A22 = M[5] + c;
  const double* A12 = M[6];
  const size_t nx = g.nx;
  size_t e = (size_t)r * g.nx + c;
  // B.1: (A11 v_ksi)_ksi along the row, (A22 v_eta)_eta along the column
  double xx = cons_line_g([=](int k) { return A11[k]; }, [=](int k) { return f(r, k); }, c, g.nx, g.dksi2);
  double yy = cons_line_g([=](int k) { return A22[(size_t)k * nx]; }, [=](int k) { return f(k, c); }, r, g.ny, g.deta2);
  // B.2: D_ksi(A12 v_eta) with left/right columns zeroed ; D_eta(A12 v_ksi) with top/bottom rows zeroed
  if (c != 0 && c != g.nx - 1) {
    int t = fd_type(c, g.nx), s = d1_start(c, g.nx, t);
    const double* w = g.tab->d1x + t * 5;
    double acc = 0.0;
    for (int k = 0; k < 5; ++k) {
      if (w[k] == 0.0) continue;
      int cc = s + k;
      double ve = (deriv_bc && r == g.ny - 1) ? 0.0 : d_eta_g(g, f, r, cc);
      acc += w[k] * (A12[(size_t)r * g.nx + cc] * ve);
    }
    xx += acc;
  }
  if (r != 0 && r != g.ny - 1) {
    int t = fd_type(r, g.ny), s = d1_start(r, g.ny, t);
    const double* w = g.tab->d1y + t * 5;
    double acc = 0.0;
    for (int k = 0; k < 5; ++k) {
      if (w[k] == 0.0) continue;
      int rr = s + k;
      double vk = (deriv_bc && (c == 0 || c == g.nx - 1 || rr == 0)) ? 0.0 : d_ksi_g(g, f, rr, c);
      acc += w[k] * (A12[(size_t)rr * g.nx + c] * vk);
    }
    yy += acc;
  }
  vxx = xx / J[e];
  vyy = yy / J[e];
}

JF_HD void mesh_laplace_point(const MeshGeom& g, const double* const* M, const double* v, int r, int c, int deriv_bc,
                              double& vxx, double& vyy) {
  if (r >= 4 && r < g.ny - 4 && c >= 4 && c < g.nx - 4) { // no closure stencil within reach
    mesh_laplace_interior(g, M, v, r, c, vxx, vyy);
    return;
  }
  const size_t nx = g.nx;
  mesh_laplace_general_g(g, M, [=](int rr, int cc) { return v[(size_t)rr * nx + cc]; }, r, c, deriv_bc, vxx, vyy);
}

// ---- PMA2 pointwise (PMA2_nk.py:131,139,157-159 ; :410-418) -------------------------------------
JF_HD double pma2_rhs_point(const Pma2Params& p, double u, double lap2) {
  double q = 1.0 + u;
  double rhs = -p.lambd / (q * q);
  if (p.epsilon != 0.0) rhs += p.lambd * pow(p.epsilon, (double)(p.m - 2)) / pow(q, (double)p.m);
  rhs -= p.beta * p.beta * lap2;
  return rhs;
}
JF_HD double pma2_combine_point(const Pma2Params& p, double u, double uval, double rhs, double cn) {
  return (u - uval) / p.dt - (rhs + cn) / 2.0;
}

// ---- droplet pointwise (droplet.py:435-473) -------------------------------------------------------
JF_HD double ipow(double x, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }
JF_HD double droplet_PI(const DropletParams& p, double h) {
  double q = p.epsilon / h;
  return (p.n_exp - 1) * (p.m_exp - 1) * (ipow(q, p.m_exp) - ipow(q, p.n_exp)) / (2 * p.epsilon * (p.n_exp - p.m_exp));
}
JF_HD double droplet_pressure_point(const DropletParams& p, double h, double lap) {
  return -lap + droplet_PI(p, h) + p.Bo * cos(p.alpha2) * h;
}
JF_HD void droplet_flux_point(const MeshGeom& g, const DropletParams& dp, const double* const* M, const double* p,
                              const double* h, int r, int c, double& A, double& B) {
  size_t e = (size_t)r * g.nx + c;
  double pk = (c == 0 || c == g.nx - 1) ? 0.0 : d_ksi(g, p, r, c);
  double pe = (r == 0 || r == g.ny - 1) ? 0.0 : d_eta(g, p, r, c);
  double qxx = M[0][e], qyy = M[1][e], qxy = M[2][e], J = M[3][e];
  double pdx = (qyy * pk - qxy * pe) / J;
  double pdy = (-qxy * pk + qxx * pe) / J;
  double h3 = h[e] * h[e] * h[e];
  A = (pdx - dp.Bo * sin(dp.alpha2) / dp.epsilon2) * h3 / 3.0;
  B = pdy * h3 / 3.0;
}
JF_HD double droplet_div_point(const MeshGeom& g, const double* const* M, const double* A, const double* B, int r, int c) {
  size_t e = (size_t)r * g.nx + c;
  double qxx = M[0][e], qyy = M[1][e], qxy = M[2][e], J = M[3][e];
  return (qyy * d_ksi(g, A, r, c) - qxy * d_eta(g, A, r, c) - qxy * d_ksi(g, B, r, c) + qxx * d_eta(g, B, r, c)) / J;
}
JF_HD double droplet_combine_point(const DropletParams& p, double u, double uval, double F2, double Fprev) {
  return (u - uval) - p.dt * (F2 + Fprev) / 2.0;
}


// ---- droplet initial shapes (droplet.py:413-429 compute_U2 / G2 / H2 ; :544-551 compute_U / G / H) ------------
// U = eps + (1 - eps) sum_drops H2(G2(|X - X_drop|, R), R, V) at the PHYSICAL position X = (Q_ksi, Q_eta) of the mesh
// point (compute_Q_spatial_ders, droplet.py:701-705: Dirichlet values on the edges).
JF_HD double droplet_G2(double xx, double R, double a) {
  return R + log((1 + exp(-2 * a * (xx + R))) / (1 + exp(-2 * a * (xx - R)))) / (2 * a);
}
JF_HD double droplet_H2(double psi, double R, double V) { return 4 * V * (1 - psi * psi / (R * R)) / (R * R); }
JF_HD double droplet_shape_point(const MeshGeom& g, const double* Q, int r, int c, const DropList& dl, double a, double eps) {
  double x = (c == 0) ? g.bl : (c == g.nx - 1 ? g.br : d_ksi(g, Q, r, c));
  double y = (r == 0) ? g.bb : (r == g.ny - 1 ? g.bt : d_eta(g, Q, r, c));
  double ret = eps;
  for (int i = 0; i < dl.n; ++i) {
    double dx = x - dl.x[i], dy = y - dl.y[i];
    ret += (1 - eps) * droplet_H2(droplet_G2(sqrt(dx * dx + dy * dy), dl.R[i], a), dl.R[i], dl.V[i]);
  }
  return ret;
}

// ---- moving-mesh relaxation (PMA): monitor function, smoothing filter, spectral solve -------------------
// compute_and_smooth_monitor (PMA2_nk.py:345-391, droplet.py:729-760), solve_PMA (PMA2_nk.py:393-403,
// droplet.py:578-588), loop_pma (droplet.py:590-599).
struct PmaParams {
  double alpha, gamma; // mesh adaption speed / extent of smoothing
  double cnorm;        // Mackenzie regularisation constant (droplet C_ = 0.15; PMA2: 1)
  int smoothing_iters; // 4
  int monitor_mode;    // 0: |u_xx + u_yy|^2 ; 1: 1/(1+u)^6 (PMA2 with epsilon == 0)
};

JF_HD double pma_monitor_point(int mode, double u, double lap) {
  if (mode == 1) { double q = 1.0 + u; double q2 = q * q; return 1.0 / (q2 * q2 * q2); }
  double a = fabs(lap);
  return a * a;
}

// one pass of the fourth-order filter (9-point interior, 6-point edges, 4-point corners)
JF_HD double pma_smooth_point(const MeshGeom& g, const double* t, int r, int c) {
  const int nx = g.nx, ny = g.ny;
#define T_(rr, cc) t[(size_t)(rr) * nx + (cc)]
  const bool top = (r == ny - 1), bot = (r == 0), lef = (c == 0), rig = (c == nx - 1);
  if (!top && !bot && !lef && !rig)
    return T_(r, c) + (T_(r - 1, c) + T_(r + 1, c) + T_(r, c - 1) + T_(r, c + 1)) / 8 +
           (T_(r - 1, c - 1) + T_(r - 1, c + 1) + T_(r + 1, c - 1) + T_(r + 1, c + 1)) / 16;
  if ((top || bot) && (lef || rig)) { // corners
    int r1 = bot ? 1 : ny - 2, c1 = lef ? 1 : nx - 2;
    return (4 * T_(r, c) + 2 * T_(r, c1) + 2 * T_(r1, c) + T_(r1, c1)) / 9;
  }
  if (lef || rig) { // vertical edges
    int c1 = lef ? 1 : nx - 2;
    return (4 * T_(r, c) + 2 * T_(r - 1, c) + 2 * T_(r + 1, c) + 2 * T_(r, c1) + T_(r + 1, c1) + T_(r - 1, c1)) / 12;
  }
  int r1 = bot ? 1 : ny - 2; // horizontal edges
  return (4 * T_(r, c) + 2 * T_(r, c - 1) + 2 * T_(r, c + 1) + 2 * T_(r1, c) + T_(r1, c + 1) + T_(r1, c - 1)) / 12;
#undef T_
}

// eigenvalues of the (DCT-diagonalised) discrete Laplacian used by the reference: M.Leig (droplet.py:831-833)
JF_HD double pma_leig(const MeshGeom& g, int r, int c) {
  const double pi = 3.141592653589793238462643383279502884;
  return ((2 * cos(pi * r / (g.ny - 1)) - 2) + (2 * cos(pi * c / (g.nx - 1)) - 2)) / (g.dksi * g.deta);
}

// orthonormal DCT-II matrix entry C[k][n] of size N: s_k cos(pi (2n+1) k / (2N))  (scipy.fft.dct(norm="ortho"))
JF_HD double dct2_entry(int k, int n, int N) {
  const double pi = 3.141592653589793238462643383279502884;
  long long m = ((long long)(2 * n + 1) * k) % (4LL * N); // exact argument reduction: period 4N
  double cv = cos(pi * (double)m / (2.0 * N));
  return (k == 0 ? sqrt(1.0 / N) : sqrt(2.0 / N)) * cv;
}

} // namespace jfnk
