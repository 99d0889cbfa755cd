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
    // int tables (2 per double): Q pull, monitor pull, column pull, row pull, (row, column) of a band offset
    tabs = take((2 * qe_rows * nx + a_rows * nx + ny * cmax + rmax * nx + 1) / 2 + 2);
    total = (size_t)o * sizeof(double);
  }
};

// One 8 x 8 tile of D = A . B on the fp64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA), operands read from shared memory
// with arbitrary strides: A(m, k) = A[m sam + k sak], B(k, n) = B[k sbk + n sbn]; rows / columns / terms beyond M, N, K are
// zero.  Fragment layout (PTX ISA, m8n8k4 .f64): lane = 4 g + t holds A(g, t), B(t, g) and the results D(g, 2t), D(g, 2t+1).
// Two accumulator pairs on alternating k-steps halve the dependent chain.  The dense DCT products of the relaxation are
// shared-memory-bandwidth bound on the CUDA cores (two loads per fma); a DMMA needs two loads per 8 fma.
__device__ __forceinline__ void dmma_tile(const double* A, int sam, int sak, const double* B, int sbk, int sbn, int M, int N,
                                          int K, int m0, int n0, double& d0, double& d1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool mok = (m0 + g) < M, nok = (n0 + g) < N;
  const double* ap = A + (size_t)(m0 + g) * sam + (size_t)t * sak;
  const double* bp = B + (size_t)t * sbk + (size_t)(n0 + g) * sbn;
  double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
  int k = 0;
  for (; k + 8 <= K; k += 8) {
    const double a0 = mok ? ap[(size_t)k * sak] : 0.0, b0 = nok ? bp[(size_t)k * sbk] : 0.0;
    const double a1 = mok ? ap[(size_t)(k + 4) * sak] : 0.0, b1 = nok ? bp[(size_t)(k + 4) * sbk] : 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a0), "d"(b0));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a1), "d"(b1));
  }
  for (; k < K; k += 4) {
    const bool kok = (k + t) < K;
    const double a0 = (mok && kok) ? ap[(size_t)k * sak] : 0.0, b0 = (nok && kok) ? bp[(size_t)k * sbk] : 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a0), "d"(b0));
  }
  d0 = c0 + e0;
  d1 = c1 + e1;
}

// C[M x N] = op(A) . op(B), row-major, op = transpose when the flag is set (the operand conventions of small_gemm_kernel), on the
// fp64 tensor cores: the dense DCT-matrix products of the mesh update on grids that are neither small enough for the cluster
// kernel above nor a power of two (dct_fft.cuh).  One CTA owns a 32 x 32 tile of C (8 warps x two 8 x 8 DMMA tiles), K advances
// in steps of 32 staged in shared memory, zero-filled past the edges; the global reads run along the contiguous index of each
// operand whichever way it is stored.
__global__ void __launch_bounds__(256) dmma_gemm_kernel(int M, int N, int K, const double* __restrict__ A, int ta,
                                                        const double* __restrict__ B, int tb, double* __restrict__ C) {
  __shared__ double sa[32][33], sb[32][33]; // sa[m][k], sb[k][n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int e = tid; e < 1024; e += 256) {
      const int hi = e >> 5, lo = e & 31; // lo runs along the contiguous index of the operand
      {
        const int m = ta ? lo : hi, k = ta ? hi : lo;
        const int row = row0 + m, kk = k0 + k;
        sa[m][k] = (row < M && kk < K) ? (ta ? A[(size_t)kk * M + row] : A[(size_t)row * K + kk]) : 0.0;
      }
      {
        const int k = tb ? lo : hi, n = tb ? hi : lo;
        const int col = col0 + n, kk = k0 + k;
        sb[k][n] = (kk < K && col < N) ? (tb ? B[(size_t)col * K + kk] : B[(size_t)kk * N + col]) : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tt = warp * 2 + j, m0 = (tt >> 2) * 8, n0 = (tt & 3) * 8;
      double d0, d1;
      dmma_tile(&sa[0][0], 33, 1, &sb[0][0], 33, 1, 32, 32, 32, m0, n0, d0, d1);
      acc[j][0] += d0;
      acc[j][1] += d1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int tt = warp * 2 + j;
    const int r = row0 + (tt >> 2) * 8 + (lane >> 2), c = col0 + (tt & 3) * 8 + 2 * (lane & 3);
    if (r < M && c < N) C[(size_t)r * N + c] = acc[j][0];
    if (r < M && c + 1 < N) C[(size_t)r * N + c + 1] = acc[j][1];
  }
}

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
  // the finite-difference weight tables next to the data: a cluster barrier invalidates L1, and a table entry fetched from
  // L2 in front of every derivative would cost more than the derivative
  __shared__ MeshTables tabs_s;
  for (int i = tid; i < (int)(sizeof(MeshTables) / sizeof(double)); i += kBandThreads)
    reinterpret_cast<double*>(&tabs_s)[i] = reinterpret_cast<const double*>(A.gm.tab)[i];
  MeshGeom g = A.gm;
  g.tab = &tabs_s;
  const int nx = g.nx, ny = g.ny;
  const bool timing = (A.prof != nullptr) && tid == 0; // (every CTA keeps its own row of 16 counters)
  long long tlast = timing ? clock64() : 0;
  auto tick = [&](int phase) {
    if (timing) { const long long t = clock64(); A.prof[blockIdx.x * 16 + phase] += t - tlast; tlast = t; }
  };
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
  int* tabRC = tabR + L.rmax * nx;       // e -> (e / nx) << 16 | (e % nx): no integer division inside the passes
  for (int e = tid; e < L.qe_rows * nx; e += kBandThreads) tabRC[e] = ((e / nx) << 16) | (e % nx);
  const int mt_rows = (rows + 7) / 8, nt_x = (nx + 7) / 8, mt_y = (ny + 7) / 8, nt_c = (ncols + 7) / 8; // DMMA tiles
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
    tick(0);
    // metric fields of the current potential on the rows the Laplacian of the own rows reads
    for (int e = tid; e < (mr.hi - mr.lo) * nx; e += kBandThreads) {
      const int r = mr.lo + (tabRC[e] >> 16), c = tabRC[e] & 0xffff;
      mesh_metrics_point(g, Qv, r, c, Mv);
    }
    __syncthreads();
    tick(1);
    // monitor function of the (old) solution on the own rows: |u_xx + u_yy|^2 or 1/(1+u)^6
    for (int e = tid; e < nR; e += kBandThreads) {
      const int r = r0 + (tabRC[e] >> 16), c = tabRC[e] & 0xffff;
      double lap = 0.0;
      if (A.pp.monitor_mode == 0) {
        double xx, yy;
        mesh_laplace_point(g, Mc, Uv, r, c, A.deriv_bc, xx, yy);
        lap = xx + yy;
      }
      Aown[e] = pma_monitor_point(A.pp.monitor_mode, Uv[(size_t)r * nx + c], lap);
    }
    tick(2);
    cl.sync(); // [2]
    tick(7);
    // smoothing sweeps on a band that shrinks by one row per sweep (ping-pong between AE and BE)
    pull(AE, tabA, nA, Aown);
    __syncthreads();
    tick(8);
    double* a = AE;
    double* b = BE;
    for (int s = 1; s <= A.pp.smoothing_iters; ++s) {
      const RowRange sr = band_clip(r0, r1, A.pp.smoothing_iters - s, ny);
      const double* av = a - (ptrdiff_t)ar.lo * nx;
      double* bv = b - (ptrdiff_t)ar.lo * nx;
      for (int e = tid; e < (sr.hi - sr.lo) * nx; e += kBandThreads) {
        const int r = sr.lo + (tabRC[e] >> 16), c = tabRC[e] & 0xffff;
        bv[(size_t)r * nx + c] = pma_smooth_point(g, av, r, c);
      }
      __syncthreads();
      double* tmp = a; a = b; b = tmp;
    }
    const double* mon = a - (ptrdiff_t)ar.lo * nx; // smoothed monitor, valid on the own rows
    tick(3);
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
    for (int tile = warp; tile < mt_rows * nt_x; tile += NW) { // T1 = rhs . Cx^T, one 8 x 8 tile per warp
      const int m0 = (tile / nt_x) * 8, n0 = (tile % nt_x) * 8;
      double d0, d1;
      dmma_tile(UROW, nx, 1, Cxs, 1, nx, rows, nx, nx, m0, n0, d0, d1);
      const int r = m0 + (lane >> 2), c = n0 + 2 * (lane & 3);
      if (r < rows && c < nx) T1[r * nx + c] = d0;
      if (r < rows && c + 1 < nx) T1[r * nx + c + 1] = d1;
    }
    tick(4);
    cl.sync(); // [4] transpose: the column band of T1
    pull(COL, tabC, nC, T1);
    __syncthreads();
    // column transform, spectral divide: spec = (Cy . T1) / (1 - gamma Leig)
    for (int tile = warp; tile < mt_y * nt_c; tile += NW) {
      const int m0 = (tile / nt_c) * 8, n0 = (tile % nt_c) * 8;
      double d0, d1;
      dmma_tile(Cys, ny, 1, COL, ncols, 1, ny, ncols, ny, m0, n0, d0, d1);
      const int r = m0 + (lane >> 2), c = n0 + 2 * (lane & 3);
      if (r < ny && c < ncols) SPEC[r * cmax + c] = d0 / inv[r * cmax + c];
      if (r < ny && c + 1 < ncols) SPEC[r * cmax + c + 1] = d1 / inv[r * cmax + c + 1];
    }
    __syncthreads();
    // inverse column transform: UCOL = Cy^T . spec
    for (int tile = warp; tile < mt_y * nt_c; tile += NW) {
      const int m0 = (tile / nt_c) * 8, n0 = (tile % nt_c) * 8;
      double d0, d1;
      dmma_tile(Cys, 1, ny, SPEC, cmax, 1, ny, ncols, ny, m0, n0, d0, d1);
      const int r = m0 + (lane >> 2), c = n0 + 2 * (lane & 3);
      if (r < ny && c < ncols) UCOL[r * cmax + c] = d0;
      if (r < ny && c + 1 < ncols) UCOL[r * cmax + c + 1] = d1;
    }
    tick(5);
    cl.sync(); // [5] transpose back: the own rows of UCOL
    pull(UROW, tabR, nR, UCOL);
    __syncthreads();
    // inverse row transform and the explicit Euler update: Q += dt (UROW . Cx)
    for (int tile = warp; tile < mt_rows * nt_x; tile += NW) {
      const int m0 = (tile / nt_x) * 8, n0 = (tile % nt_x) * 8;
      double d0, d1;
      dmma_tile(UROW, nx, 1, Cxs, nx, 1, rows, nx, nx, m0, n0, d0, d1);
      const int r = m0 + (lane >> 2), c = n0 + 2 * (lane & 3);
      // Q.val += dt*Q.dt: product rounded, then sum (as the per-stage path and NumPy)
      if (r < rows && c < nx) Qown[r * nx + c] = combine(1.0 * Qown[r * nx + c], A.dt, d0);
      if (r < rows && c + 1 < nx) Qown[r * nx + c + 1] = combine(1.0 * Qown[r * nx + c + 1], A.dt, d1);
    }
    tick(6);
  }
  __syncthreads();
  for (int e = tid; e < nR; e += kBandThreads) A.Q[(size_t)r0 * nx + e] = Qown[e];
  cl.sync(); // nobody leaves while a neighbour may still pull from its shared memory
}

} // namespace jfnk
