// One whole LGMRES inner cycle (_fgmres, _gcrotmk.py:16-183, plus the solution assembly of lgmres.py:188-208) as ONE
// launch, for the grids the reference itself runs (sh_scipy_nk.py: N = 64; BASELINE config 1: 61 x 61).
//
// At these sizes a vector is 30 KB and an Arnoldi step of the streaming path (operator + multi-dot + update kernels) is three
// dependent ~10 us launches: the cycle is latency bound, 40 us per Arnoldi step.  Here ONE thread-block cluster (16 CTAs,
// 8 where 16 cannot be scheduled) runs the whole cycle:
//   * every CTA owns a band of rows of EVERY vector and keeps its band of the Arnoldi basis (<= 41 vectors), of the
//     linearisation point (with 2 halo rows per side) and of G(x0) in shared memory for the whole cycle -- the basis never
//     touches HBM or L2;
//   * the operator w = (G(x0 + sc z) - G(x0))/omega (KrylovJacobian.matvec, _nonlin.py:1557-1565; or the linearly-implicit
//     operator of sh_linearised.py:54) reads the two halo rows per side of z from the neighbouring CTAs' shared memory
//     (distributed shared memory, ~215 cycles);
//   * the classical Gram-Schmidt dots and the norm are reduced per warp, pushed to every CTA's mailbox with DSMEM stores and
//     summed by every CTA in the same fixed order after a hardware cluster barrier (two barriers per Arnoldi step), so the
//     replicated Hessenberg / Givens / stopping state (hd_math.h, the same functions the streaming path runs) is bit-identical
//     in all CTAs and nobody waits for a broadcast;
//   * the least-squares solve and dx = sum y_i z_i follow in the same launch; the host reads ONE record per cycle.
// The arithmetic per point is sh_value<> (sh_kernels.cuh) with the neighbour sums formed in the order of sh_point_kernel.
#pragma once
#include <cooperative_groups.h>
#include "cuda_common.cuh"
#include "sh_kernels.cuh"

namespace jfnk {

constexpr int kCycThreads = 256;
constexpr int kCycMaxCluster = 16;
constexpr int kCycMailW = JF_MAXV + 4;

struct CycleArgs {
  int nx, ny;
  int m, k;          // Arnoldi steps allowed (inner_m + k), augmentation vectors in use
  int gs_mode;       // JFNK_GS_*
  int pitch;         // doubles per vector band in shared memory (>= the largest band)
  int ext;           // doubles per extended band ((rows + 4) * nx, largest band)
  double omega, ptol, tau2, v0n2;
  const double* x0;  // linearisation point (null for the linear operator)
  const double* g0;  // G(x0) ; linear operator: the diagonal D
  const double* v0;  // unnormalised start vector
  const double* ov[JF_MAXOV]; // augmentation vectors in use, oldest first (global memory)
  int ov_zn2[JF_MAXOV];       // arena index of their squared norms
  double* out;       // dx
  int out_zn2;       // arena index of ||dx||^2
  double* S;         // the context's scalar arena (global memory)
};

// shared-memory layout (doubles); identical in every CTA so that map_shared_rank() addresses match
struct CycleLayout {
  int arena, mailD, mailN, mailX, cf, x0e, g0, t, v;
  size_t total;
  __host__ __device__ CycleLayout(int pitch, int ext, int m) {
    int o = 0;
    arena = o; o += (JS_COUNT + 3) & ~3;
    mailD = o; o += kCycMaxCluster * kCycMailW;
    mailN = o; o += kCycMaxCluster * 8;
    mailX = o; o += kCycMaxCluster * 8;
    cf = o; o += JF_MAXV;
    x0e = o; o += ext;
    g0 = o; o += pitch;
    t = o; o += ext;
    v = o; o += (m + 1) * pitch;
    total = (size_t)o * sizeof(double);
  }
};

__device__ __forceinline__ int cyc_mod(int r, int n) { r %= n; return r < 0 ? r + n : r; }

template <int OP>
__global__ void __launch_bounds__(kCycThreads) sh_cycle_kernel(const __grid_constant__ CycleArgs A, const SHParams P) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster(); // the whole grid is one cluster
  const int C = (int)cl.num_blocks(), me = (int)cl.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kCycThreads / 32;
  extern __shared__ __align__(16) double cyc_smem[];
  __shared__ int r0tab[kCycMaxCluster + 1];
  __shared__ int znidx[JF_MAXV];
  const CycleLayout L(A.pitch, A.ext, A.m);
  double* Ssm = cyc_smem + L.arena;
  double* mailD = cyc_smem + L.mailD; // [src CTA][kCycMailW]
  double* mailN = cyc_smem + L.mailN; // [src CTA][warp]
  double* mailX = cyc_smem + L.mailX;
  double* cf = cyc_smem + L.cf;
  double* X0e = cyc_smem + L.x0e;
  double* G0 = cyc_smem + L.g0;
  double* T = cyc_smem + L.t;
  double* V = cyc_smem + L.v;
  const int nx = A.nx, ny = A.ny, k = A.k, m = A.m;
  if (tid <= C) r0tab[tid] = (int)(((long long)tid * ny) / C);
  __syncthreads();
  const int r0 = r0tab[me], rows = r0tab[me + 1] - r0;
  const int Pn = rows * nx, En = (rows + 4) * nx;
  const size_t g0off = (size_t)r0 * nx;

  // ---- load: scalar arena, linearisation point with halos, G(x0) / D, start vector -----------------------------------
  for (int i = tid; i < JS_COUNT; i += kCycThreads) Ssm[i] = A.S[i];
  for (int e = tid; e < En; e += kCycThreads) {
    if (OP == OP_JVPG) {
      const int rg = cyc_mod(r0 - 2 + e / nx, ny);
      X0e[e] = A.x0[(size_t)rg * nx + (e % nx)];
    }
  }
  for (int p = tid; p < Pn; p += kCycThreads) {
    G0[p] = A.g0[g0off + p];
    V[p] = A.v0[g0off + p];
  }
  __syncthreads();
  if (tid == 0) {
    Ssm[JS_VN2 + 0] = A.v0n2;
    Ssm[JS_STOP] = 0.0;
    Ssm[JS_PTOL] = A.ptol;
    Ssm[JS_TAU2] = (A.gs_mode == JFNK_GS_CGS_IFNEEDED) ? A.tau2 : 0.0;
  }
  cl.sync(); // every CTA's band of v0 is in place before a neighbour reads its halo rows

  // the band of z rows [r0-2, r0+rows+2) as operator input: T = x0 + sc z (FD Jacobian) or z (linear operator)
  auto build_T = [&](const double* zglobal, const double* zband, double sc) {
    for (int e = tid; e < En; e += kCycThreads) {
      const int rg = cyc_mod(r0 - 2 + e / nx, ny), c = e % nx;
      double zv;
      if (zglobal) zv = zglobal[(size_t)rg * nx + c];
      else {
        int o = (int)(((long long)rg * C) / ny);
        while (r0tab[o + 1] <= rg) ++o;
        while (r0tab[o] > rg) --o;
        const double* src = (o == me) ? zband : cl.map_shared_rank(zband, o);
        zv = src[(rg - r0tab[o]) * nx + c];
      }
      T[e] = (OP == OP_JVPG) ? combine(X0e[e], sc, zv) : zv;
    }
  };
  // W = operator(T) on the own band (neighbour sums in the order of sh_point_kernel)
  auto apply = [&](double* W, double scale) {
    ShAcc acc = {0.0, 0.0, 0.0};
    for (int p = tid; p < Pn; p += kCycThreads) {
      const int r = p / nx, c = p - r * nx;
      const int cm1 = c - 1 < 0 ? c - 1 + nx : c - 1, cm2 = c - 2 < 0 ? c - 2 + nx : c - 2;
      const int cp1 = c + 1 >= nx ? c + 1 - nx : c + 1, cp2 = c + 2 >= nx ? c + 2 - nx : c + 2;
      const double* t0 = T + (r + 2) * nx; // row r of the band
      const double uc = t0[c];
      const double a1 = t0[cm1] + t0[cp1];
      const double a2 = t0[cm2] + t0[cp2];
      const double a1u = t0[cm1 - nx] + t0[cp1 - nx];
      const double a1d = t0[cm1 + nx] + t0[cp1 + nx];
      const double s1 = a1 + t0[c - nx] + t0[c + nx];
      const double sd = a1u + a1d;
      const double s2 = a2 + t0[c - 2 * nx] + t0[c + 2 * nx];
      double second = 0.0, third = 0.0;
      W[p] = sh_value<OP>(P, scale, uc, s1, sd, s2, G0[p], 0.0, second, third, acc);
    }
  };
  // all-reduced dots of V_0..V_{nd-1} with W into Ssm[off + i] (every CTA sums the same partials in the same order)
  auto dots = [&](int nd, const double* W, int off) {
    for (int i = warp; i < nd; i += NW) {
      const double* vi = V + (size_t)i * A.pitch;
      double s = 0.0;
      for (int p = lane; p < Pn; p += 32) s = fma(vi[p], W[p], s);
      s = warp_sum(s);
      if (lane < C) {
        double* dst = (lane == me) ? mailD : cl.map_shared_rank(mailD, lane);
        dst[me * kCycMailW + i] = s;
      }
    }
    cl.sync();
    if (tid < nd) {
      double s = 0.0;
      for (int o = 0; o < C; ++o) s += mailD[o * kCycMailW + tid];
      Ssm[off + tid] = s;
      if (tid < nd - 1) cf[tid] = -(s / Ssm[JS_VN2 + tid]); // Gram-Schmidt coefficient of the unnormalised V_i
    }
    __syncthreads();
  };
  // all-reduce of one per-thread partial sum through mailbox `mail`; result returned in every thread of warp 0
  auto allsum = [&](double part, double* mail) -> double {
    part = warp_sum(part);
    if (lane < C) {
      double* dst = (lane == me) ? mail : cl.map_shared_rank(mail, lane);
      dst[me * 8 + warp] = part;
    }
    cl.sync();
    double s = 0.0;
    if (warp == 0) {
      for (int q = lane; q < C * 8; q += 32) s += mail[q];
      s = warp_sum(s);
    }
    return s;
  };
  // W -= sum_{i<nv} (dot_i / ||V_i||^2) V_i ; returns ||W||^2 (warp 0)
  auto update = [&](int nv, double* W) -> double {
    double acc = 0.0;
    for (int p = tid; p < Pn; p += kCycThreads) {
      double t = W[p];
      for (int i = 0; i < nv; ++i) t = fma(cf[i], V[(size_t)i * A.pitch + p], t);
      W[p] = t;
      acc = fma(t, t, acc);
    }
    return allsum(acc, mailN);
  };

  int nit = 0, reorth = 0;
  double res = 0.0;
  int flags = 0;
  for (int j = 0; j < m; ++j) {
    // choice of z (_gcrotmk.py:96-105 with prepend_outer_v=True): augmentation vectors, then v0, then the last Arnoldi vector
    const double* zg = nullptr;
    const double* zb = nullptr;
    int zi;
    if (j < k) { zg = A.ov[j]; zi = A.ov_zn2[j]; }
    else if (j == k) { zb = V; zi = JS_VN2 + 0; }
    else { zb = V + (size_t)j * A.pitch; zi = JS_VN2 + j; }
    if (tid == 0) znidx[j] = zi;
    const double zn = sqrt(Ssm[zi]);
    double* W = V + (size_t)(j + 1) * A.pitch;
    build_T(zg, zb, A.omega / zn);
    __syncthreads();
    apply(W, OP == OP_JVPG ? 1.0 / A.omega : 1.0 / zn);
    __syncthreads();
    // classical Gram-Schmidt against V_0..V_j: all dots (and w.w) in one exchange, the update and its norm in another
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
    if (tid == 0) hess_givens_step(Ssm, j, taken, 0);
    __syncthreads();
    flags = (int)Ssm[JS_FLAGS];
    if (flags & JF_FLAG_NEED_REORTH) {
      // the first pass cancelled more than 1/tau: second pass, then column j again from the saved rotated rhs
      dots(j + 2, W, JS_RD2);
      hn2 = update(j + 1, W);
      if (tid == 0) { Ssm[JS_HN2B] = hn2; hess_givens_step(Ssm, j, 1, 1); }
      __syncthreads();
      flags = (int)Ssm[JS_FLAGS];
      taken = 1;
    }
    if (taken) ++reorth;
    res = Ssm[JS_RES];
    nit = j + 1;
    if ((flags & (JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) || res < A.ptol) break;
  }

  // y = lstsq(R, Q[0,:]) * ||v0|| ; dx = sum y_i z_i (lgmres.py:188,206-208)
  if (tid == 0) lsq_solve(Ssm, nit, znidx, JS_VN2 + 0);
  __syncthreads();
  double acc = 0.0;
  for (int p = tid; p < Pn; p += kCycThreads) {
    double t = 0.0;
    for (int i = 0; i < nit; ++i) {
      const double zv = (i < k) ? A.ov[i][g0off + p] : (i == k ? V[p] : V[(size_t)i * A.pitch + p]);
      t = fma(Ssm[JS_COEF + i], zv, t);
    }
    A.out[g0off + p] = t;
    acc = fma(t, t, acc);
  }
  const double dxn2 = allsum(acc, mailX);
  if (me == 0) {
    if (tid == 0) {
      Ssm[A.out_zn2] = dxn2;
      Ssm[JS_STOP] = 0.0;
      double* cyc = Ssm + JS_CYC;
      cyc[0] = (double)nit; cyc[1] = (double)reorth; cyc[2] = res; cyc[3] = (double)flags; cyc[4] = dxn2;
    }
    __syncthreads();
    for (int i = JS_WW + tid; i < JS_COUNT; i += kCycThreads) A.S[i] = Ssm[i];
  }
}

} // namespace jfnk
