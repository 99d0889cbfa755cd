// The moving-mesh relaxation loop (loop_pma, droplet.py:590-599; solve_PMA + the explicit Euler update of
// PMA2_nk.py:94,103) as ONE persistent kernel of ONE thread-block cluster with every field resident in (distributed)
// shared memory -- the second generation of pma_relax.cuh.
//
// pma_relax_kernel separates its twelve stages per pass by cluster barriers and hands the fields from stage to stage through
// global memory: every barrier is followed by a cold L2 round trip (ncu: barrier 7.3 + membar 2.7 + long scoreboard 3.7
// stalled warps per issue), 42 us per pass.  Here
//   * every CTA owns a band of rows of the grid; the potential Q, the monitor and the spectral fields live in its shared
//     memory for all `loops` passes, and a stage that needs rows of a neighbour PULLS them over DSMEM (ld.shared::cluster,
//     ~215 cycles) into a local extended band, after which the same point functions as everywhere else (mesh_math.h) run on
//     local shared memory through a pointer biased by the band's first row;
//   * stages without a cross-CTA dependency are merged: the metric fields are computed redundantly on the +-2 rows the
//     Laplacian needs (no exchange of metric fields), the four smoothing sweeps run on a band that shrinks by one row per
//     sweep after ONE pull of 4 halo rows, the row transforms of the 2-D DCT are local to the row band and the column
//     transforms (forward, spectral divide, inverse) local to a column band, so the DCT pair costs two transposes;
//   * five cluster barriers per pass instead of twelve, no global memory traffic inside the loop.
// Arithmetic: the same point functions; the 2-D transforms apply the row factor first (Y = Cy (X Cx^T) instead of
// (Cy X) Cx^T -- equal up to rounding); the weighted sum is reduced per CTA and then in CTA order (deterministic).
#pragma once
#include <cooperative_groups.h>
#include "cuda_common.cuh"
#include "mesh_kernels.cuh"
#include "pma_kernels.cuh"
#include "pma_relax.cuh"

namespace jfnk {

constexpr int kBandThreads = 512;
constexpr int kBandMaxCluster = 16;

// rows [lo, hi) extended by H rows per side, clipped to the grid, and widened to the one-sided closure stencils of the
// 4th-order operators when it touches the first / last three rows (rows 0..2 read rows 0..5, droplet.py:636-668,785-804)
struct RowRange { int lo, hi; };
__host__ __device__ inline RowRange band_ext(int lo, int hi, int H, int ny) {
  RowRange r;
  r.lo = lo - H < 0 ? 0 : lo - H;
  r.hi = hi + H > ny ? ny : hi + H;
  if (r.lo <= 2 && r.hi < 6) r.hi = ny < 6 ? ny : 6;
  if (r.hi >= ny - 2 && r.lo > ny - 6) r.lo = ny - 6 < 0 ? 0 : ny - 6;
  return r;
}
__host__ __device__ inline RowRange band_clip(int lo, int hi, int H, int ny) {
  RowRange r;
  r.lo = lo - H < 0 ? 0 : lo - H;
  r.hi = hi + H > ny ? ny : hi + H;
  return r;
}

// shared-memory layout in doubles, identical in every CTA (sized for the largest band of the cluster)
struct BandLayout {
  int cy, cx, inv, qown, qe, mj, ma11, ma22, ma12, mdummy, ue, aown, ae, be, t1own, colbuf, spec, ucol, urow, mail, tabs;
  int qe_rows, m_rows, u_rows, a_rows, rmax, cmax;
  size_t total;
  __host__ __device__ BandLayout(int nx, int ny, int C) {
    rmax = (ny + C - 1) / C;
    cmax = (nx + C - 1) / C;
    m_rows = rmax + 4 + 3;  // ext(own, 2) (+ closure slack)
    qe_rows = m_rows + 4 + 3;
    u_rows = rmax + 6 + 3;
    a_rows = rmax + 8;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 1) & ~1; return at; };
    cy = take(ny * ny); cx = take(nx * nx); inv = take(ny * cmax);
    qown = take(rmax * nx); qe = take(qe_rows * nx);
    mj = take(m_rows * nx); ma11 = take(m_rows * nx); ma22 = take(m_rows * nx); ma12 = take(m_rows * nx); mdummy = take(m_rows * nx);
    ue = take(u_rows * nx);
    aown = take(rmax * nx); ae = take(a_rows * nx); be = take(a_rows * nx);
    t1own = take(rmax * nx); colbuf = take(ny * cmax); spec = take(ny * cmax); ucol = take(ny * cmax); urow = take(rmax * nx);
    mail = take(kBandMaxCluster);
    // int tables (2 per double): Q pull, monitor pull, column pull, row pull
    tabs = take((qe_rows * nx + a_rows * nx + ny * cmax + rmax * nx + 1) / 2 + 2);
    total = (size_t)o * sizeof(double);
  }
};

__global__ void __launch_bounds__(kBandThreads) pma_relax_band_kernel(const __grid_constant__ RelaxArgs A) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster(); // the whole grid is one cluster
  const int C = (int)cl.num_blocks(), me = (int)cl.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kBandThreads / 32;
  extern __shared__ __align__(16) double band_smem[];
  __shared__ int r0tab[kBandMaxCluster + 1], c0tab[kBandMaxCluster + 1];
  __shared__ double red[NW];
  __shared__ double total_s;
  const MeshGeom& g = A.gm;
  const int nx = g.nx, ny = g.ny;
  const BandLayout L(nx, ny, C);
  double* Cys = band_smem + L.cy;
  double* Cxs = band_smem + L.cx;
  double* inv = band_smem + L.inv;       // the spectral divisor 1 - gamma Leig on the column band [ny][cmax]
  double* Qown = band_smem + L.qown;     // own rows of Q (what the neighbours pull)
  double* QE = band_smem + L.qe;
  double* UE = band_smem + L.ue;
  double* Aown = band_smem + L.aown;     // own rows of the monitor (what the neighbours pull)
  double* AE = band_smem + L.ae;
  double* BE = band_smem + L.be;
  double* T1 = band_smem + L.t1own;      // own rows of b Cx^T
  double* COL = band_smem + L.colbuf;    // its column band [ny][cmax]
  double* SPEC = band_smem + L.spec;
  double* UCOL = band_smem + L.ucol;     // column band of Cy^T spec
  double* UROW = band_smem + L.urow;     // its own rows
  double* mail = band_smem + L.mail;
  if (tid <= C) { r0tab[tid] = (tid * ny) / C; c0tab[tid] = (tid * nx) / C; }
  __syncthreads();
  const int r0 = r0tab[me], r1 = r0tab[me + 1], rows = r1 - r0;
  const int c0 = c0tab[me], ncols = c0tab[me + 1] - c0, cmax = L.cmax;
  const RowRange mr = band_ext(r0, r1, 2, ny);        // metric fields: what the Laplacian of the own rows reads
  const RowRange qr = band_ext(mr.lo, mr.hi, 2, ny);  // potential: what those metric fields read
  const RowRange ur = band_ext(r0, r1, 3, ny);        // old solution: what the Laplacian reads
  const RowRange ar = band_clip(r0, r1, A.pp.smoothing_iters, ny); // monitor: shrinks by one row per smoothing sweep
  auto owner_row = [&](int r) { int o = (r * C) / ny; while (r0tab[o + 1] <= r) ++o; while (r0tab[o] > r) --o; return o; };
  auto owner_col = [&](int c) { int o = (c * C) / nx; while (c0tab[o + 1] <= c) ++o; while (c0tab[o] > c) --o; return o; };
  // pull tables: (owner CTA << 20) | offset inside the owner's array
  int* tabQ = reinterpret_cast<int*>(band_smem + L.tabs);
  int* tabA = tabQ + L.qe_rows * nx;
  int* tabC = tabA + L.a_rows * nx;
  int* tabR = tabC + ny * cmax;
  const int nQ = (qr.hi - qr.lo) * nx, nA = (ar.hi - ar.lo) * nx, nC = ny * ncols, nR = rows * nx;
  for (int e = tid; e < nQ; e += kBandThreads) {
    const int r = qr.lo + e / nx, c = e % nx, o = owner_row(r);
    tabQ[e] = (o << 20) | ((r - r0tab[o]) * nx + c);
  }
  for (int e = tid; e < nA; e += kBandThreads) {
    const int r = ar.lo + e / nx, c = e % nx, o = owner_row(r);
    tabA[e] = (o << 20) | ((r - r0tab[o]) * nx + c);
  }
  for (int e = tid; e < nC; e += kBandThreads) { // COL[k][cc] <- T1 of the owner of row k, column c0 + cc
    const int kk = e / ncols, cc = e % ncols, o = owner_row(kk);
    tabC[e] = (o << 20) | ((kk - r0tab[o]) * nx + c0 + cc);
  }
  for (int e = tid; e < nR; e += kBandThreads) { // UROW[rr][c] <- UCOL of the owner of column c, row r0 + rr
    const int rr = e / nx, c = e % nx, o = owner_col(c);
    tabR[e] = (o << 20) | ((r0 + rr) * cmax + (c - c0tab[o]));
  }
  // constants of the whole run: DCT matrices, spectral divisors of the column band, the old solution with its halo rows
  for (int i = tid; i < ny * ny; i += kBandThreads) Cys[i] = A.dcty[i];
  for (int i = tid; i < nx * nx; i += kBandThreads) Cxs[i] = A.dctx[i];
  for (int e = tid; e < nC; e += kBandThreads) {
    const int kk = e / ncols, cc = e % ncols;
    inv[kk * cmax + cc] = 1.0 - A.pp.gamma * pma_leig(g, kk, c0 + cc);
  }
  for (int e = tid; e < (ur.hi - ur.lo) * nx; e += kBandThreads) UE[e] = A.Uval[(size_t)ur.lo * nx + e];
  for (int e = tid; e < nR; e += kBandThreads) Qown[e] = A.Q[(size_t)r0 * nx + e];
  // pointers biased by the first row of each extended band: the point functions index with global (row, col)
  const double* Qv = QE - (ptrdiff_t)qr.lo * nx;
  const double* Uv = UE - (ptrdiff_t)ur.lo * nx;
  double* Mv[7];
  Mv[0] = Mv[1] = Mv[2] = band_smem + L.mdummy - (ptrdiff_t)mr.lo * nx; // Q_xx, Q_yy, Q_xy: not needed by the relaxation
  Mv[3] = band_smem + L.mj - (ptrdiff_t)mr.lo * nx;
  Mv[4] = band_smem + L.ma11 - (ptrdiff_t)mr.lo * nx;
  Mv[5] = band_smem + L.ma22 - (ptrdiff_t)mr.lo * nx;
  Mv[6] = band_smem + L.ma12 - (ptrdiff_t)mr.lo * nx;
  const double* const* Mc = const_cast<const double* const*>(Mv);
  const double* Jv = Mv[3];
  auto pull = [&](double* dst, const int* tab, int n, const double* src_local) {
    for (int e = tid; e < n; e += kBandThreads) {
      const int o = tab[e] >> 20;
      const double* src = (o == me) ? src_local : cl.map_shared_rank(src_local, o);
      dst[e] = src[tab[e] & 0xfffff];
    }
  };

  for (int it = 0; it < A.loops; ++it) {
    cl.sync(); // [1] every CTA's own rows of Q are current
    pull(QE, tabQ, nQ, Qown);
    __syncthreads();
    // metric fields of the current potential on the rows the Laplacian of the own rows reads
    for (int e = tid; e < (mr.hi - mr.lo) * nx; e += kBandThreads) {
      const int r = mr.lo + e / nx, c = e % nx;
      mesh_metrics_point(g, Qv, r, c, Mv);
    }
    __syncthreads();
    // monitor function of the (old) solution on the own rows: |u_xx + u_yy|^2 or 1/(1+u)^6
    for (int e = tid; e < nR; e += kBandThreads) {
      const int r = r0 + e / nx, c = e % nx;
      double lap = 0.0;
      if (A.pp.monitor_mode == 0) {
        double xx, yy;
        mesh_laplace_point(g, Mc, Uv, r, c, A.deriv_bc, xx, yy);
        lap = xx + yy;
      }
      Aown[e] = pma_monitor_point(A.pp.monitor_mode, Uv[(size_t)r * nx + c], lap);
    }
    cl.sync(); // [2]
    // smoothing sweeps on a band that shrinks by one row per sweep (ping-pong between AE and BE)
    pull(AE, tabA, nA, Aown);
    __syncthreads();
    double* a = AE;
    double* b = BE;
    for (int s = 1; s <= A.pp.smoothing_iters; ++s) {
      const RowRange sr = band_clip(r0, r1, A.pp.smoothing_iters - s, ny);
      const double* av = a - (ptrdiff_t)ar.lo * nx;
      double* bv = b - (ptrdiff_t)ar.lo * nx;
      for (int e = tid; e < (sr.hi - sr.lo) * nx; e += kBandThreads) {
        const int r = sr.lo + e / nx, c = e % nx;
        bv[(size_t)r * nx + c] = pma_smooth_point(g, av, r, c);
      }
      __syncthreads();
      double* tmp = a; a = b; b = tmp;
    }
    const double* mon = a - (ptrdiff_t)ar.lo * nx; // smoothed monitor, valid on the own rows
    // Mackenzie regularisation: sum mon |J| over the grid, then sqrt((mon + C sum |J| dksi deta) |J|) / alpha
    {
      double acc = 0.0;
      for (int e = tid; e < nR; e += kBandThreads) {
        const size_t i = (size_t)r0 * nx + e;
        acc = fma(mon[i], fabs(Jv[i]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) red[warp] = acc;
      __syncthreads();
      if (tid < C) {
        double v = 0.0;
        for (int w = 0; w < NW; ++w) v += red[w];
        double* dst = (tid == me) ? mail : cl.map_shared_rank(mail, tid);
        dst[me] = v;
      }
    }
    cl.sync(); // [3]
    if (tid == 0) {
      double v = 0.0;
      for (int k = 0; k < C; ++k) v += mail[k];
      total_s = v;
    }
    __syncthreads();
    // rhs on the own rows, then the row transform of the forward DCT-II: T1 = rhs . Cx^T
    {
      const double add = A.pp.cnorm * A.cell * total_s;
      for (int e = tid; e < nR; e += kBandThreads) {
        const size_t i = (size_t)r0 * nx + e;
        UROW[e] = sqrt((mon[i] + add) * fabs(Jv[i])) / A.pp.alpha; // (UROW is free until the inverse transform)
      }
    }
    __syncthreads();
    for (int e = tid; e < nR; e += kBandThreads) {
      const int rr = e / nx, c = e % nx;
      const double* xr = UROW + rr * nx;
      const double* cr = Cxs + (size_t)c * nx;
      double acc = 0.0;
      for (int k = 0; k < nx; ++k) acc = fma(xr[k], cr[k], acc);
      T1[e] = acc;
    }
    cl.sync(); // [4] transpose: the column band of T1
    pull(COL, tabC, nC, T1);
    __syncthreads();
    // column transform, spectral divide: spec = (Cy . T1) / (1 - gamma Leig)
    for (int e = tid; e < nC; e += kBandThreads) {
      const int r = e / ncols, cc = e % ncols;
      const double* cyr = Cys + (size_t)r * ny;
      double acc = 0.0;
      for (int k = 0; k < ny; ++k) acc = fma(cyr[k], COL[k * ncols + cc], acc);
      SPEC[r * cmax + cc] = acc / inv[r * cmax + cc];
    }
    __syncthreads();
    // inverse column transform: UCOL = Cy^T . spec
    for (int e = tid; e < nC; e += kBandThreads) {
      const int r = e / ncols, cc = e % ncols;
      double acc = 0.0;
      for (int k = 0; k < ny; ++k) acc = fma(Cys[(size_t)k * ny + r], SPEC[k * cmax + cc], acc);
      UCOL[r * cmax + cc] = acc;
    }
    cl.sync(); // [5] transpose back: the own rows of UCOL
    pull(UROW, tabR, nR, UCOL);
    __syncthreads();
    // inverse row transform and the explicit Euler update: Q += dt (UROW . Cx)
    for (int e = tid; e < nR; e += kBandThreads) {
      const int rr = e / nx, c = e % nx;
      const double* xr = UROW + rr * nx;
      double acc = 0.0;
      for (int k = 0; k < nx; ++k) acc = fma(xr[k], Cxs[(size_t)k * nx + c], acc);
      // Q.val += dt*Q.dt: product rounded, then sum (as the per-stage path and NumPy)
      Qown[e] = combine(1.0 * Qown[e], A.dt, acc);
    }
  }
  __syncthreads();
  for (int e = tid; e < nR; e += kBandThreads) A.Q[(size_t)r0 * nx + e] = Qown[e];
  cl.sync(); // nobody leaves while a neighbour may still pull from its shared memory
}

} // namespace jfnk
