// One whole LGMRES inner cycle of the droplet problem (droplet.py:383 -> _fgmres, _gcrotmk.py:16-183, plus the solution
// assembly of lgmres.py:188-208 and the first line-search trial of the Newton step that follows) as ONE launch of one
// 16-CTA thread-block cluster -- the moving-mesh counterpart of sh_cycle.cuh, for the reference's droplet grids
// (61 x 61 ... 91 x 61) and, with a two-stage operator, for PMA2_nk.py's own 51 x 51 grid (KIND 1, below).
//
// The streaming path needs five dependent launches per Arnoldi step (Laplacian + pressure, flux, divergence + quotient,
// multi-dot, update) of ~10 us each on 5551 points: 57 us per residual evaluation, all latency.  Here every CTA owns a band of
// rows and keeps in shared memory for the whole cycle: its band of the Arnoldi basis (<= 41 vectors), of the metric fields
// (A22, A12 with the +-2 rows the Laplacian reads) and of f0.  One Jacobian-vector product
//   w = (F(x0 + sc z) - f0)/omega,  F(t) = (t - uval) - dt (div(flux(p(t, Lap t), t)) + Fprev)/2     (droplet.py:435-450)
// is three stencil stages; a stage that needs rows of a neighbour pulls them over DSMEM into a local extended band, after
// which the same point functions as every other kernel (mesh_math.h) run on local shared memory through a pointer biased
// by the band's first row:
//   T  = x0 + sc z on the own rows +-3   (z pulled from the neighbours' basis bands, x0 from L2)
//   p  = pressure(T, Lap T)  own rows            -> cluster barrier, pull +-2 rows
//   (A, B) = flux(p, T)      own rows            -> cluster barrier, pull +-2 rows
//   w  = ((T - uval) - dt (div(A, B) + Fprev)/2 - f0)/omega
// followed by the classical Gram-Schmidt step of sh_cycle.cuh (all dots in one DSMEM exchange, update + norm in another,
// replicated Givens / stopping decisions from hd_math.h): four cluster barriers per Arnoldi step, no global-memory
// round trip between the stages, one launch and one read-back per Newton iteration.
//
// KIND 1 -- PMA2 (PMA2_nk.py:121-159): F(t) = (t - uval)/dt - (rhs(t, Lap Lap t) + cn)/2 is two stencil stages,
//   lap1 = Lap T own rows -> cluster barrier, pull +-3 rows ; w = (F(T, Lap lap1) - f0)/omega  (boundary rhs zeroed).
#pragma once
#include <cooperative_groups.h>
#include "cuda_common.cuh"
#include "mesh_kernels.cuh"
#include "pma_relax_band.cuh"
#include "sh_cycle.cuh"

namespace jfnk {

constexpr int kMcThreads = 512;
constexpr int kMcCluster = 16;
constexpr int kMcWarps = kMcThreads / 32;

enum MeshCycleKind { MC_DROPLET = 0, MC_PMA2 = 1 };

struct MeshCycleArgs {
  MeshGeom gm;
  DropletParams dp;
  Pma2Params pp;
  int m, k, gs_mode;
  double omega, ptol, tau2, v0n2;
  const double* x0;   // linearisation point
  const double* f0;   // F(x0)
  const double* v0;   // unnormalised start vector
  const double* uval; // previous time level
  const double* fprev;// its Crank-Nicolson term
  MetricCPtrs M;      // Q_xx, Q_yy, Q_xy, J, A11, A22, A12 of the current mesh (global memory)
  const double* ov[JF_MAXOV];
  int ov_zn2[JF_MAXOV];
  double* out;        // dx
  int out_zn2;
  double* S;
  // optional first line-search trial: t = x0 - dx -> trial_x, F(t) -> trial_F, norms -> S[trial_norm_off..+2]
  double *trial_x, *trial_F;
  int trial_norm_off;
};

// shared-memory layout in doubles, identical in every CTA
struct MeshCycleLayout {
  int arena, mailD, mailN, mailX, cf, te, a22e, a12e, own5, pown, pe, aown, bown, be, f0, tab, v;
  int rmax, t_rows, e_rows, pitch;
  size_t total;
  // kind 0 (droplet): p / A / B bands with +-2 rows ; kind 1 (PMA2): `pe` holds lap1 with the +-3 rows of the t band
  __host__ __device__ MeshCycleLayout(int nx, int ny, int m, int kind) {
    rmax = (ny + kMcCluster - 1) / kMcCluster;
    // ext(own, H) spans at most rows + 2H rows; the widening to the closure stencils only applies to ranges that end
    // below row 6 / start above row ny - 6, i.e. to at most 6 rows
    t_rows = rmax + 6;
    e_rows = rmax + 4 < 6 ? 6 : rmax + 4;
    pitch = (rmax * nx + 1) & ~1;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 1) & ~1; return at; };
    arena = take(JS_COUNT);
    mailD = take(kMcCluster * kCycMailW); mailN = take(kMcCluster * kMcWarps); mailX = take(kMcCluster * kMcWarps);
    cf = take(JF_MAXV);
    te = take(t_rows * nx);
    a22e = take(e_rows * nx); a12e = take(e_rows * nx);
    own5 = take(5 * pitch);   // A11, J, Q_xx, Q_yy, Q_xy on the own rows
    pown = take(pitch); pe = take((kind == 1 ? t_rows : e_rows) * nx);   // (pe doubles as the extended band of A)
    aown = take(kind == 1 ? 0 : pitch); bown = take(kind == 1 ? 0 : pitch); be = take(kind == 1 ? 0 : e_rows * nx);
    f0 = take(pitch);
    tab = take((t_rows * nx + 1) / 2 + 1);        // int table of the largest extended band: owner << 20 | offset
    v = take((m + 1) * pitch);
    total = (size_t)o * sizeof(double);
  }
};

template <int KIND>
__global__ void __launch_bounds__(kMcThreads) mesh_cycle_kernel(const __grid_constant__ MeshCycleArgs A) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster(); // the whole grid is one cluster
  const int C = (int)cl.num_blocks(), me = (int)cl.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kMcWarps;
  extern __shared__ __align__(16) double mc_smem[];
  __shared__ int r0tab[kMcCluster + 1];
  __shared__ int znidx[JF_MAXV];
  __shared__ double hpre[JF_MAXV + 1];
  __shared__ MeshTables tabs_s; // (a cluster barrier invalidates L1: keep the finite-difference weights next to the data)
  for (int i = tid; i < (int)(sizeof(MeshTables) / sizeof(double)); i += kMcThreads)
    reinterpret_cast<double*>(&tabs_s)[i] = reinterpret_cast<const double*>(A.gm.tab)[i];
  MeshGeom g = A.gm;
  g.tab = &tabs_s;
  const int nx = g.nx, ny = g.ny, k = A.k, m = A.m;
  const MeshCycleLayout L(nx, ny, m, KIND);
  double* Ssm = mc_smem + L.arena;
  double* mailD = mc_smem + L.mailD;
  double* mailN = mc_smem + L.mailN;
  double* mailX = mc_smem + L.mailX;
  double* cf = mc_smem + L.cf;
  double* TE = mc_smem + L.te;
  double* Pown = mc_smem + L.pown;
  double* PE = mc_smem + L.pe;
  double* Aown = mc_smem + L.aown;
  double* Bown = mc_smem + L.bown;
  double* AE = PE; // p is dead once the flux is formed
  double* BE = mc_smem + L.be;
  double* F0 = mc_smem + L.f0;
  double* V = mc_smem + L.v;
  const int pitch = L.pitch;
  if (tid <= C) r0tab[tid] = (tid * ny) / C;
  __syncthreads();
  const int r0 = r0tab[me], r1 = r0tab[me + 1], rows = r1 - r0;
  const int Pn = rows * nx;
  const size_t goff = (size_t)r0 * nx;
  const RowRange tr = band_ext(r0, r1, 3, ny); // operator input: what the Laplacian of the own rows reads
  const RowRange er = band_ext(r0, r1, 2, ny); // A22, A12 ; p ; (A, B): what the own rows of the next stage read
  const int nT = (tr.hi - tr.lo) * nx, nE = (er.hi - er.lo) * nx;
  // pull table of the t band (the +-2 bands are a contiguous sub-range of it): owner CTA << 20 | offset in the owner's band
  int* tabT = reinterpret_cast<int*>(mc_smem + L.tab);
  for (int e = tid; e < nT; e += kMcThreads) {
    const int r = tr.lo + e / nx, c = e % nx;
    int o = (r * C) / ny;
    while (r0tab[o + 1] <= r) ++o;
    while (r0tab[o] > r) --o;
    tabT[e] = (o << 20) | ((r - r0tab[o]) * nx + c);
  }
  const int* tabE = tabT + (er.lo - tr.lo) * nx;
  // ---- load: augmentation norms, metric fields, f0, start vector ------------------------------------------------------
  for (int i = tid; i < JS_COUNT; i += kMcThreads) Ssm[i] = (i >= JS_ZN2 && i < JS_ZN2 + JF_MAXOV) ? A.S[i] : 0.0;
  double* own5 = mc_smem + L.own5;
  for (int p = tid; p < Pn; p += kMcThreads) {
    own5[0 * pitch + p] = A.M.m[4][goff + p]; // A11
    own5[1 * pitch + p] = A.M.m[3][goff + p]; // J
    own5[2 * pitch + p] = A.M.m[0][goff + p]; // Q_xx
    own5[3 * pitch + p] = A.M.m[1][goff + p]; // Q_yy
    own5[4 * pitch + p] = A.M.m[2][goff + p]; // Q_xy
    F0[p] = A.f0[goff + p];
    V[p] = A.v0[goff + p];
  }
  for (int e = tid; e < nE; e += kMcThreads) {
    mc_smem[L.a22e + e] = A.M.m[5][(size_t)er.lo * nx + e];
    mc_smem[L.a12e + e] = A.M.m[6][(size_t)er.lo * nx + e];
  }
  // pointers biased by the first row of each band: the point functions index with global (row, col)
  const double* Tv = TE - (ptrdiff_t)tr.lo * nx;
  const double* Pv = PE - (ptrdiff_t)er.lo * nx;
  const double* AEv = AE - (ptrdiff_t)er.lo * nx;
  const double* BEv = BE - (ptrdiff_t)er.lo * nx;
  const double* Mv[7];
  Mv[0] = own5 + 2 * pitch - (ptrdiff_t)r0 * nx;
  Mv[1] = own5 + 3 * pitch - (ptrdiff_t)r0 * nx;
  Mv[2] = own5 + 4 * pitch - (ptrdiff_t)r0 * nx;
  Mv[3] = own5 + 1 * pitch - (ptrdiff_t)r0 * nx;
  Mv[4] = own5 + 0 * pitch - (ptrdiff_t)r0 * nx;
  Mv[5] = mc_smem + L.a22e - (ptrdiff_t)er.lo * nx;
  Mv[6] = mc_smem + L.a12e - (ptrdiff_t)er.lo * nx;
  __syncthreads();
  if (tid == 0) {
    Ssm[JS_VN2 + 0] = A.v0n2;
    Ssm[JS_STOP] = 0.0;
    Ssm[JS_PTOL] = A.ptol;
    Ssm[JS_TAU2] = (A.gs_mode == JFNK_GS_CGS_IFNEEDED) ? A.tau2 : 0.0;
  }
  cl.sync();

  auto pull = [&](double* dst, const int* tab, int n, const double* src_local) {
    for (int e = tid; e < n; e += kMcThreads) {
      const int o = tab[e] >> 20;
      const double* src = (o == me) ? src_local : cl.map_shared_rank(src_local, o);
      dst[e] = src[tab[e] & 0xfffff];
    }
  };
  // T = x0 + sc z on the rows [tr.lo, tr.hi): z from global memory (augmentation vectors) or from the basis bands
  auto build_T = [&](const double* zglobal, const double* zband, double sc) {
    for (int e = tid; e < nT; e += kMcThreads) {
      const size_t ge = (size_t)tr.lo * nx + e;
      double zv;
      if (zglobal) zv = zglobal[ge];
      else {
        const int o = tabT[e] >> 20;
        const double* src = (o == me) ? zband : cl.map_shared_rank(zband, o);
        zv = src[tabT[e] & 0xfffff];
      }
      TE[e] = combine(A.x0[ge], sc, zv);
    }
  };
  // the three stencil stages on the own rows.  quot: W = (F - f0) scale ; else F -> global trial_F, t -> trial_x, norms
  auto chain = [&](bool quot, double* W, double scale, double (&nrm)[3]) {
    __syncthreads(); // T complete
    for (int p = tid; p < Pn; p += kMcThreads) {
      const int rr = p / nx, c = p - rr * nx, r = r0 + rr;
      double xx, yy;
      mesh_laplace_point(g, Mv, Tv, r, c, 0, xx, yy);
      Pown[p] = (KIND == MC_PMA2) ? xx + yy : droplet_pressure_point(A.dp, Tv[(size_t)r * nx + c], xx + yy);
    }
    cl.sync();
    if (KIND == MC_PMA2) {
      // second stage: Laplacian of lap1 (pulled with the +-3 rows of the t band), pointwise rhs, Crank-Nicolson combination
      pull(PE, tabT, nT, Pown);
      __syncthreads();
      const double* L1v = PE - (ptrdiff_t)tr.lo * nx;
      for (int p = tid; p < Pn; p += kMcThreads) {
        const int rr = p / nx, c = p - rr * nx, r = r0 + rr;
        double xx, yy;
        mesh_laplace_point(g, Mv, L1v, r, c, 0, xx, yy);
        const bool bdy = (r == 0 || c == 0 || r == ny - 1 || c == nx - 1);
        const double u = Tv[(size_t)r * nx + c];
        const double rhs = bdy ? 0.0 : pma2_rhs_point(A.pp, u, xx + yy);
        const double f = pma2_combine_point(A.pp, u, A.uval[goff + p], rhs, A.fprev[goff + p]);
        if (quot) W[p] = (f - F0[p]) / scale;
        else {
          A.trial_F[goff + p] = f;
          A.trial_x[goff + p] = u;
          nrm[0] = fma(f, f, nrm[0]); nrm[1] = fmax(nrm[1], fabs(f)); nrm[2] = fmax(nrm[2], fabs(u));
        }
      }
      __syncthreads();
      return;
    }
    pull(PE, tabE, nE, Pown);
    __syncthreads();
    for (int p = tid; p < Pn; p += kMcThreads) {
      const int rr = p / nx, c = p - rr * nx, r = r0 + rr;
      double a, b;
      droplet_flux_point(g, A.dp, Mv, Pv, Tv, r, c, a, b);
      Aown[p] = a; Bown[p] = b;
    }
    cl.sync();
    pull(AE, tabE, nE, Aown);
    pull(BE, tabE, nE, Bown);
    __syncthreads();
    for (int p = tid; p < Pn; p += kMcThreads) {
      const int rr = p / nx, c = p - rr * nx, r = r0 + rr;
      const double F2 = droplet_div_point(g, Mv, AEv, BEv, r, c);
      const double u = Tv[(size_t)r * nx + c];
      const double f = droplet_combine_point(A.dp, u, A.uval[goff + p], F2, A.fprev[goff + p]);
      if (quot) W[p] = (f - F0[p]) / scale;
      else {
        A.trial_F[goff + p] = f;
        A.trial_x[goff + p] = u;
        nrm[0] = fma(f, f, nrm[0]); nrm[1] = fmax(nrm[1], fabs(f)); nrm[2] = fmax(nrm[2], fabs(u));
      }
    }
    __syncthreads();
  };
  // ---- Gram-Schmidt machinery (as sh_cycle.cuh) ---------------------------------------------------------------------------
  auto dots = [&](int nd, const double* W, int off) {
    constexpr int Q = (JF_MAXV + NW - 1) / NW;
    double s[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) s[q] = 0.0;
    const int nq = (nd - warp + NW - 1) / NW;
    const double* v0p = V + (size_t)warp * pitch;
    for (int p = lane; p < Pn; p += 32) {
      const double wv = W[p];
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q < nq) s[q] = fma(v0p[(size_t)q * NW * pitch + p], wv, s[q]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int q = 0; q < Q; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    if (lane < C) {
      double* dst = ((lane == me) ? mailD : cl.map_shared_rank(mailD, lane)) + me * kCycMailW + warp;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q < nq) dst[q * NW] = s[q];
    }
    cl.sync();
    if (tid < nd) {
      double t = 0.0;
      for (int o = 0; o < C; ++o) t += mailD[o * kCycMailW + tid];
      Ssm[off + tid] = t;
      if (tid < nd - 1) {
        cf[tid] = -(t / Ssm[JS_VN2 + tid]);
        hpre[tid] = hess_entry(Ssm, tid, off == JS_RD2);
      }
    }
    __syncthreads();
  };
  auto allsum = [&](double part, double* mail) -> double {
    part = warp_sum(part);
    if (lane < C) {
      double* dst = (lane == me) ? mail : cl.map_shared_rank(mail, lane);
      dst[me * NW + warp] = part;
    }
    cl.sync();
    double s = 0.0;
    if (warp == 0) {
      for (int q = lane; q < C * NW; q += 32) s += mail[q];
      s = warp_sum(s);
    }
    return s;
  };
  auto update = [&](int nv, double* W) -> double {
    double acc = 0.0;
    for (int p = tid; p < Pn; p += kMcThreads) {
      double t = W[p];
      for (int i = 0; i < nv; ++i) t = fma(cf[i], V[(size_t)i * pitch + p], t);
      W[p] = t;
      acc = fma(t, t, acc);
    }
    return allsum(acc, mailN);
  };

  int nit = 0, reorth = 0, flags = 0;
  double res = 0.0;
  double unused[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < m; ++j) {
    const double* zg = nullptr;
    const double* zb = nullptr;
    int zi;
    if (j < k) { zg = A.ov[j]; zi = A.ov_zn2[j]; }
    else if (j == k) { zb = V; zi = JS_VN2 + 0; }
    else { zb = V + (size_t)j * pitch; zi = JS_VN2 + j; }
    if (tid == 0) znidx[j] = zi;
    const double zn = sqrt(Ssm[zi]);
    double* W = V + (size_t)(j + 1) * pitch;
    build_T(zg, zb, A.omega / zn);
    chain(true, W, A.omega, unused);
    dots(j + 2, W, JS_RD);
    double hn2 = update(j + 1, W);
    int taken = 0;
    if (A.gs_mode == JFNK_GS_CGS2) {
      if (tid == 0) Ssm[JS_HN2A] = hn2;
      dots(j + 2, W, JS_RD2);
      hn2 = update(j + 1, W);
      if (tid == 0) Ssm[JS_HN2B] = hn2;
      taken = 1;
    } else if (tid == 0) Ssm[JS_HN2A] = hn2;
    if (tid == 0) hess_givens_step(Ssm, j, taken, 0, hpre);
    __syncthreads();
    flags = (int)Ssm[JS_FLAGS];
    if (flags & JF_FLAG_NEED_REORTH) {
      dots(j + 2, W, JS_RD2);
      hn2 = update(j + 1, W);
      if (tid == 0) { Ssm[JS_HN2B] = hn2; hess_givens_step(Ssm, j, 1, 1, hpre); }
      __syncthreads();
      flags = (int)Ssm[JS_FLAGS];
      taken = 1;
    }
    if (taken) ++reorth;
    res = Ssm[JS_RES];
    nit = j + 1;
    if ((flags & (JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) || res < A.ptol) break;
  }

  // least squares by warp 0, one column per step (as sh_cycle.cuh), then dx = sum y_i z_i
  if (warp == 0) {
    const double* R = Ssm + JS_R;
    double a0 = (lane < nit) ? Ssm[JS_G + lane] : 0.0;
    double a1 = (lane + 32 < nit) ? Ssm[JS_G + lane + 32] : 0.0;
    for (int i = nit - 1; i >= 0; --i) {
      const double piv = R[i + (size_t)i * JF_MAXV];
      const double ai = __shfl_sync(0xffffffffu, (i < 32) ? a0 : a1, i & 31);
      const double yi = (piv != 0.0) ? ai / piv : 0.0;
      if (lane == 0) Ssm[JS_Y + i] = yi;
      if (lane < i) a0 = fma(-R[lane + (size_t)i * JF_MAXV], yi, a0);
      if (lane + 32 < i) a1 = fma(-R[lane + 32 + (size_t)i * JF_MAXV], yi, a1);
    }
  }
  __syncthreads();
  if (tid < nit) {
    const double y = Ssm[JS_Y + tid] * sqrt(Ssm[JS_VN2 + 0]);
    Ssm[JS_Y + tid] = y;
    Ssm[JS_COEF + tid] = y / sqrt(Ssm[znidx[tid]]);
  }
  __syncthreads();
  double* DX = V + (size_t)m * pitch; // the slot of the last possible Arnoldi vector: never an input of the assembly
  double acc = 0.0;
  for (int p = tid; p < Pn; p += kMcThreads) {
    double t = 0.0;
    for (int i = 0; i < nit; ++i) {
      const double zv = (i < k) ? A.ov[i][goff + p] : (i == k ? V[p] : V[(size_t)i * pitch + p]);
      t = fma(Ssm[JS_COEF + i], zv, t);
    }
    A.out[goff + p] = t;
    DX[p] = t;
    acc = fma(t, t, acc);
  }
  const double dxn2 = allsum(acc, mailX); // (its cluster barrier also publishes the DX bands)
  double tn[3] = {0.0, 0.0, 0.0};
  if (A.trial_F != nullptr) {
    // F(x0 - dx) with its norms, all-reduced like the dots
    build_T(nullptr, DX, -1.0);
    double nrm[3] = {0.0, 0.0, 0.0};
    chain(false, nullptr, 1.0, nrm);
    const double f2w = warp_sum(nrm[0]), fmw = warp_max(nrm[1]), xmw = warp_max(nrm[2]);
    if (lane < C) {
      double* dst = ((lane == me) ? mailD : cl.map_shared_rank(mailD, lane)) + me * kCycMailW + warp * 3;
      dst[0] = f2w; dst[1] = fmw; dst[2] = xmw;
    }
    cl.sync();
    if (warp == 0) {
      double f2 = 0.0, fm = 0.0, xm = 0.0;
      for (int q = lane; q < C * NW; q += 32) {
        const double* src = mailD + (q / NW) * kCycMailW + (q % NW) * 3;
        f2 += src[0]; fm = fmax(fm, src[1]); xm = fmax(xm, src[2]);
      }
      tn[0] = warp_sum(f2); tn[1] = warp_max(fm); tn[2] = warp_max(xm);
    }
  }
  if (me == 0) {
    if (tid == 0) {
      Ssm[A.out_zn2] = dxn2;
      Ssm[JS_STOP] = 0.0;
      double* cyc = Ssm + JS_CYC;
      cyc[0] = (double)nit; cyc[1] = (double)reorth; cyc[2] = res; cyc[3] = (double)flags; cyc[4] = dxn2;
      cyc[5] = tn[0]; cyc[6] = tn[1]; cyc[7] = tn[2];
      if (A.trial_F != nullptr) { A.S[A.trial_norm_off] = tn[0]; A.S[A.trial_norm_off + 1] = tn[1]; A.S[A.trial_norm_off + 2] = tn[2]; }
    }
    __syncthreads();
    for (int i = JS_WW + tid; i < JS_R; i += kMcThreads) A.S[i] = Ssm[i];
    for (int i = JS_STOP + tid; i < JS_COUNT; i += kMcThreads) A.S[i] = Ssm[i];
  }
}

} // namespace jfnk
