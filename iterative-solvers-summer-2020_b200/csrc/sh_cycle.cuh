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
  // optional: the first line-search trial of the Newton step that follows (nonlin_solve's phi(1), _nonlin.py:300-305)
  // evaluated in the same launch -- t = x0 - dx, F = G(t) - d, G(t), and the three norms at S[trial_norm_off..+2]
  const double* d;   // per-step constant field (null: no trial)
  double *trial_x, *trial_F, *trial_G;
  int trial_norm_off;
  long long* prof;   // optional (JFNK_CYCLE_PROF=1): clock cycles per phase, accumulated by CTA 0
};
enum CyclePhase { CP_LOAD = 0, CP_BUILD, CP_APPLY, CP_DOTS, CP_UPDATE, CP_GIVENS, CP_ASSEMBLE, CP_TRIAL, CP_TOTAL, CP_STEPS, CP_CYCLES, CP_COUNT };

// shared-memory layout (doubles); identical in every CTA so that map_shared_rank() addresses match
struct CycleLayout {
  int arena, mailD, mailN, mailX, cf, idx, x0e, g0, t, v;
  size_t total;
  __host__ __device__ CycleLayout(int pitch, int ext, int m) {
    int o = 0;
    arena = o; o += (JS_COUNT + 3) & ~3;
    mailD = o; o += kCycMaxCluster * kCycMailW;
    mailN = o; o += kCycMaxCluster * 8;
    mailX = o; o += kCycMaxCluster * 8;
    cf = o; o += JF_MAXV;
    idx = o; o += ((ext + 1) / 2) * 2 + (pitch + 1) / 2; // int tables: 2 per extended-band point, 1 per band point
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
  __shared__ double hpre[JF_MAXV + 1];
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
  const bool timing = (A.prof != nullptr) && me == 0 && tid == 0;
  long long tlast = timing ? clock64() : 0;
  const long long tstart = tlast;
  auto tick = [&](int phase) {
    if (timing) { const long long t = clock64(); A.prof[phase] += t - tlast; tlast = t; }
  };
  if (tid <= C) r0tab[tid] = (tid * ny) / C;
  __syncthreads();
  const int r0 = r0tab[me], rows = r0tab[me + 1] - r0;
  const int Pn = rows * nx, En = (rows + 4) * nx;
  const size_t g0off = (size_t)r0 * nx;
  // index tables, fixed for the whole cycle: for every point of the extended band its global offset and where its value
  // lives in the cluster (owner CTA << 20 | offset inside the owner's band); for every band point its column
  int* egoff = reinterpret_cast<int*>(cyc_smem + L.idx);
  int* esrc = egoff + ((A.ext + 1) / 2) * 2;
  int* pcol = esrc + ((A.ext + 1) / 2) * 2;
  for (int e = tid; e < En; e += kCycThreads) {
    const int rg = cyc_mod(r0 - 2 + e / nx, ny), c = e % nx;
    int o = (rg * C) / ny;
    while (r0tab[o + 1] <= rg) ++o;
    while (r0tab[o] > rg) --o;
    egoff[e] = rg * nx + c;
    esrc[e] = (o << 20) | ((rg - r0tab[o]) * nx + c);
  }
  for (int p = tid; p < Pn; p += kCycThreads) pcol[p] = p % nx;

  // ---- load: scalar arena, linearisation point with halos, G(x0) / D, start vector -----------------------------------
  // (everything the cycle reads from the arena it has written itself, except the norms of the augmentation vectors)
  for (int i = tid; i < JS_COUNT; i += kCycThreads) Ssm[i] = (i >= JS_ZN2 && i < JS_ZN2 + JF_MAXOV) ? A.S[i] : 0.0;
  if (OP == OP_JVPG)
    for (int e = tid; e < En; e += kCycThreads) X0e[e] = A.x0[egoff[e]];
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
  tick(CP_LOAD);

  // the band of z rows [r0-2, r0+rows+2) as operator input: T = x0 + sc z (FD Jacobian) or z (linear operator)
  // (t0, ts: first index and stride of the calling thread -- all 256 threads, or the 224 of warps 1..7 while warp 0 runs the
  //  Givens step of the previous column)
  auto build_T = [&](const double* zglobal, const double* zband, double sc, int t0, int ts) {
    for (int e = t0; e < En; e += ts) {
      double zv;
      if (zglobal) zv = zglobal[egoff[e]];
      else {
        const int o = esrc[e] >> 20;
        const double* src = (o == me) ? zband : cl.map_shared_rank(zband, o);
        zv = src[esrc[e] & 0xfffff];
      }
      T[e] = (OP == OP_JVPG) ? combine(X0e[e], sc, zv) : zv;
    }
  };
  // W = operator(T) on the own band (neighbour sums in the order of sh_point_kernel)
  auto apply = [&](double* W, double scale, int t0, int ts) {
    ShAcc acc = {0.0, 0.0, 0.0};
    for (int p = t0; p < Pn; p += ts) {
      const int c = pcol[p];
      const int cm1 = c - 1 < 0 ? c - 1 + nx : c - 1, cm2 = c - 2 < 0 ? c - 2 + nx : c - 2;
      const int cp1 = c + 1 >= nx ? c + 1 - nx : c + 1, cp2 = c + 2 >= nx ? c + 2 - nx : c + 2;
      const double* t0 = T + (p - c) + 2 * nx; // the point's row inside the extended band
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
  // all-reduced dots of V_0..V_{nd-1} with W into Ssm[off + i] (every CTA sums the same partials in the same order).
  // Warp w takes the vectors w, w + 8, ...: up to 6 independent accumulators and interleaved shuffle reductions per warp.
  auto dots = [&](int nd, const double* W, int off) {
    constexpr int Q = (JF_MAXV + NW - 1) / NW;
    double s[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) s[q] = 0.0;
    const int nq = (nd - warp + NW - 1) / NW; // vectors of this warp
    const double* v0p = V + (size_t)warp * A.pitch;
    for (int p = lane; p < Pn; p += 32) {
      const double wv = W[p];
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q < nq) s[q] = fma(v0p[(size_t)q * NW * A.pitch + p], wv, s[q]);
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
        cf[tid] = -(t / Ssm[JS_VN2 + tid]);                 // Gram-Schmidt coefficient of the unnormalised V_i
        hpre[tid] = hess_entry(Ssm, tid, off == JS_RD2);    // Hessenberg column entry, all rows in parallel
      }
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
  bool have_spec = false;
  for (int j = 0; j < m; ++j) {
    // choice of z (_gcrotmk.py:96-105 with prepend_outer_v=True): augmentation vectors, then v0, then the last Arnoldi vector
    const double* zg = nullptr;
    const double* zb = nullptr;
    int zi;
    if (j < k) { zg = A.ov[j]; zi = A.ov_zn2[j]; }
    else if (j == k) { zb = V; zi = JS_VN2 + 0; }
    else { zb = V + (size_t)j * A.pitch; zi = JS_VN2 + j; }
    if (tid == 0) znidx[j] = zi;
    double* W = V + (size_t)(j + 1) * A.pitch;
    if (!have_spec) { // (else W was formed while the Givens step of column j-1 ran, see below)
      const double zn = sqrt(Ssm[zi]);
      build_T(zg, zb, A.omega / zn, tid, kCycThreads);
      __syncthreads();
      tick(CP_BUILD);
      apply(W, OP == OP_JVPG ? 1.0 / A.omega : 1.0 / zn, tid, kCycThreads);
      __syncthreads();
      tick(CP_APPLY);
    }
    have_spec = false;
    // classical Gram-Schmidt against V_0..V_j: all dots (and w.w) in one exchange, the update and its norm in another
    dots(j + 2, W, JS_RD);
    tick(CP_DOTS);
    double hn2 = update(j + 1, W);
    tick(CP_UPDATE);
    int taken = 0;
    if (A.gs_mode == JFNK_GS_CGS2) {
      if (tid == 0) Ssm[JS_HN2A] = hn2;
      dots(j + 2, W, JS_RD2);
      hn2 = update(j + 1, W);
      if (tid == 0) Ssm[JS_HN2B] = hn2;
      taken = 1;
    } else if (tid == 0) Ssm[JS_HN2A] = hn2;
    // The Givens step of column j is one thread's serial work (~2.6 k clocks).  Meanwhile warps 1..7 already apply the operator
    // to the vector just orthogonalised (step j+1's input when j + 1 > k; its norm is all they need from this step) -- dropped
    // if column j ends the process or asks for a second Gram-Schmidt pass, exactly like the streaming loop's speculation.
    const bool spec = (A.gs_mode != JFNK_GS_CGS2) && (j + 1 < m);
    if (spec) {
      if (tid == 0) Ssm[JS_VN2 + j + 1] = hn2; // (what the Givens step stores there as well)
      __syncthreads();
      if (warp == 0) {
        if (lane == 0) hess_givens_step(Ssm, j, taken, 0, hpre);
      } else {
        const int jn = j + 1;
        const double* zg1 = nullptr;
        const double* zb1 = nullptr;
        int zi1;
        if (jn < k) { zg1 = A.ov[jn]; zi1 = A.ov_zn2[jn]; }
        else if (jn == k) { zb1 = V; zi1 = JS_VN2 + 0; }
        else { zb1 = V + (size_t)jn * A.pitch; zi1 = JS_VN2 + jn; }
        const double zn1 = sqrt(Ssm[zi1]);
        build_T(zg1, zb1, A.omega / zn1, tid - 32, kCycThreads - 32);
        asm volatile("bar.sync 1, %0;" ::"n"(kCycThreads - 32) : "memory");
        apply(V + (size_t)(jn + 1) * A.pitch, OP == OP_JVPG ? 1.0 / A.omega : 1.0 / zn1, tid - 32, kCycThreads - 32);
      }
      have_spec = true;
    } else if (tid == 0) hess_givens_step(Ssm, j, taken, 0, hpre);
    __syncthreads();
    flags = (int)Ssm[JS_FLAGS];
    if (flags & JF_FLAG_NEED_REORTH) {
      have_spec = false; // (the second pass changes the vector the operator was applied to)
      // the first pass cancelled more than 1/tau: second pass, then column j again from the saved rotated rhs
      dots(j + 2, W, JS_RD2);
      hn2 = update(j + 1, W);
      if (tid == 0) { Ssm[JS_HN2B] = hn2; hess_givens_step(Ssm, j, 1, 1, hpre); }
      __syncthreads();
      flags = (int)Ssm[JS_FLAGS];
      taken = 1;
    }
    tick(CP_GIVENS);
    if (taken) ++reorth;
    res = Ssm[JS_RES];
    nit = j + 1;
    if ((flags & (JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) || res < A.ptol) break;
  }

  // y = lstsq(R, Q[0,:]) * ||v0|| ; dx = sum y_i z_i (lgmres.py:188,206-208).  Back-substitution by warp 0, one column
  // per step (lane r owns rows r and r + 32); a zero pivot (Arnoldi breakdown) yields y_i = 0 like lsq_solve (hd_math.h).
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
  double* DX = V + (size_t)m * A.pitch; // the slot of the last possible Arnoldi vector: never an input of the assembly
  double acc = 0.0;
  for (int p = tid; p < Pn; p += kCycThreads) {
    double t = 0.0;
    for (int i = 0; i < nit; ++i) {
      const double zv = (i < k) ? A.ov[i][g0off + p] : (i == k ? V[p] : V[(size_t)i * A.pitch + p]);
      t = fma(Ssm[JS_COEF + i], zv, t);
    }
    A.out[g0off + p] = t;
    DX[p] = t;
    acc = fma(t, t, acc);
  }
  const double dxn2 = allsum(acc, mailX); // (its cluster barrier also publishes the DX bands)
  tick(CP_ASSEMBLE);
  double tn[3] = {0.0, 0.0, 0.0};
  if (OP == OP_JVPG && A.d != nullptr) {
    // F(x0 - dx): the operand band with halos, the residual with its norms, all-reduced like the dots
    build_T(nullptr, DX, -1.0, tid, kCycThreads);
    __syncthreads();
    ShAcc ra = {0.0, 0.0, 0.0};
    for (int p = tid; p < Pn; p += kCycThreads) {
      const int c = pcol[p];
      const int cm1 = c - 1 < 0 ? c - 1 + nx : c - 1, cm2 = c - 2 < 0 ? c - 2 + nx : c - 2;
      const int cp1 = c + 1 >= nx ? c + 1 - nx : c + 1, cp2 = c + 2 >= nx ? c + 2 - nx : c + 2;
      const double* t0 = T + (p - c) + 2 * nx;
      const double uc = t0[c];
      const double a1 = t0[cm1] + t0[cp1];
      const double a2 = t0[cm2] + t0[cp2];
      const double a1u = t0[cm1 - nx] + t0[cp1 - nx];
      const double a1d = t0[cm1 + nx] + t0[cp1 + nx];
      const double s1 = a1 + t0[c - nx] + t0[c + nx];
      const double sd = a1u + a1d;
      const double s2 = a2 + t0[c - 2 * nx] + t0[c + 2 * nx];
      double second = 0.0, third = 0.0;
      const double F = sh_value<OP_RESID>(P, 1.0, uc, s1, sd, s2, A.d[g0off + p], 0.0, second, third, ra);
      A.trial_F[g0off + p] = F;
      A.trial_x[g0off + p] = second;
      A.trial_G[g0off + p] = third;
    }
    // sum F^2 (sum), max|F|, max|t| (max): per-warp partials to every CTA's mailbox, every CTA reduces all of them
    const double f2w = warp_sum(ra.f2), fmw = warp_max(ra.fmax), xmw = warp_max(ra.xmax);
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
    tick(CP_TRIAL);
  }
  if (me == 0) {
    if (tid == 0) {
      Ssm[A.out_zn2] = dxn2;
      Ssm[JS_STOP] = 0.0;
      double* cyc = Ssm + JS_CYC;
      cyc[0] = (double)nit; cyc[1] = (double)reorth; cyc[2] = res; cyc[3] = (double)flags; cyc[4] = dxn2;
      cyc[5] = tn[0]; cyc[6] = tn[1]; cyc[7] = tn[2];
      if (OP == OP_JVPG && A.d != nullptr) { A.S[A.trial_norm_off] = tn[0]; A.S[A.trial_norm_off + 1] = tn[1]; A.S[A.trial_norm_off + 2] = tn[2]; }
    }
    __syncthreads();
    // back to the arena: everything but the triangular factor (no later kernel or host code reads it)
    for (int i = JS_WW + tid; i < JS_R; i += kCycThreads) A.S[i] = Ssm[i];
    for (int i = JS_STOP + tid; i < JS_COUNT; i += kCycThreads) A.S[i] = Ssm[i];
  }
  tick(CP_ASSEMBLE);
  if (timing) { A.prof[CP_TOTAL] += clock64() - tstart; A.prof[CP_STEPS] += nit; A.prof[CP_CYCLES] += 1; }
}

} // namespace jfnk
