// Marching kernel for the moving-mesh Laplace_operator (PMA2_nk.py:263-343, droplet.py:601-681) on large grids
// (config 3: PMA2 on 2048^2), with the PMA2 residual / FD-JVP epilogue fused in.
//
// The one-thread-per-point kernels of mesh_kernels.cuh fetch every stencil leg from L1/L2 (~60 loads and ~420
// instructions per point) and are instruction-issue bound at 2048^2.  Here a CTA owns a strip of 122 interior
// columns (+3 halo columns per side = 128 threads, one column per thread) and marches down a chunk of rows:
//   * vertical neighbours live in register windows (rows of t, A22 and g = A12 * D_ksi t), held as circular buffers whose
//     indices are compile-time constants because the row loop is unrolled by the period (8); the shared-memory rings
//     (8 rows of t, 4 operand stages, 2 rows of A11 and f) are indexed by the row count RELATIVE to the chunk, so their
//     addresses are immediates too (no per-row index arithmetic);
//   * horizontal neighbours come from an 8-row shared-memory ring of t and 2-row rings of A11 and f = A12 * D_eta t;
//   * every operand is read from global memory once per point: one elected thread issues a bulk-TMA copy of the strip's
//     row of every field (1 KiB each, completion counted on an mbarrier) into a 4-stage shared-memory ring three rows ahead
//     of its use (requested after the row barrier, into the stage of the row just finished); unaligned / odd grids use one
//     8-byte cp.async per thread and field, one row ahead (the fetch path is a template parameter);
//   * the stencil input may be the combination t = x + a v (FD-JVP / line search), formed when a row is consumed;
//   * divisions by 288 h^2, J, dt and the FD step are multiplications by a correctly rounded reciprocal.
// One block barrier per row.  Only interior points (4 <= r < ny-4, 4 <= c < nx-4: no closure stencil in reach) are
// marched; the frame of 4 rows / columns along each edge is done by a few extra CTAs of the same launch with the
// general per-point formulas (mesh_laplace_general_g), reading t through the same accessor; they run concurrently
// with the marching CTAs (33 us alone against ~65 us for the interior at 2048^2).
//
// Modes: MARCH_LAP        out = Lap t (out2 = t when non-null)
//        MARCH_PMA2_RESID u = px + pa pv ; F = (u - uval)/dt - (rhs(u, Lap t) + cn)/2 ; out = F ; sum F^2, max|F|, max|u|
//        MARCH_PMA2_JVP   out = (F - f0)/div
// Algorithmic bytes per point: LAP 6 fields (t, A11, A22, A12, J in; out) + 1 with v; PMA2 pass 9-11 fields.
// Measured on B200 at 2048^2 (ncu, profiles/ncu_prof_marchp8b_2048_r2.md): 4.6 TB/s of algorithmic traffic
// (0.70 of the measured HBM peak; round 1: 0.5, the one-thread-per-point kernels 0.24): DRAM traffic equals the algorithmic bytes;
// the limit is the instruction stream of the row loop at 16 resident warps per SM (fixed-latency dependencies, the row barrier,
// instruction fetch -- the loop body is kept small on purpose: 309 SASS instructions per row, see the notes at the template).
#pragma once
#include "cuda_common.cuh"
#include "mesh_kernels.cuh"
#include "sh_kernels.cuh" // mbarrier / bulk-TMA helpers

namespace jfnk {

// 128 threads x 4 resident CTAs per SM (<= 128 registers): the same 16 warps per SM as 256 x 2, but four row phases
// in flight instead of two, which hides more of the per-row barrier + fp64 dependency latency (measured at 2048^2:
// 87 vs 93 us per pass; 256 x 3 and 128 x 6 with the register cap at 80 spill and are slower).
#ifndef JFNK_MARCH_THREADS
#define JFNK_MARCH_THREADS 128
#endif
#ifndef JFNK_MARCH_MINCTAS
#define JFNK_MARCH_MINCTAS 4
#endif
constexpr int kMarchThreads = JFNK_MARCH_THREADS;
// halo columns per side / output columns per strip: 3 / 122 with per-thread cp.async operand fetch; 4 / 120 with bulk-TMA
// row copies, whose source must be 16-byte aligned (strip origins at even columns)
__host__ __device__ constexpr int march_halo(bool tma) { return tma ? 4 : 3; }
__host__ __device__ constexpr int march_out(bool tma) { return kMarchThreads - 2 * march_halo(tma); }
constexpr int kMarchRing = 8;                             // rows of t kept in shared memory (5 live; power of 2)
constexpr int kMarchPad = 4;                              // slack columns so that halo threads index in range
constexpr int kMarchMinRows = 8;                          // rows per chunk at least (6 warm-up rows per chunk)

// Rows of operands in flight ahead of the row being computed.  Measured at 2048^2 (profiles/march_components_r2.txt, us per
// launch averaged over the two passes of a PMA2 FD-JVP): per-thread cp.async 1 / 3 / 7 rows ahead 81.6 / 94.6 / 95.5 (a deeper
// cp.async.ca pipeline thrashes what the shared-memory carve-out leaves of L1); bulk TMA 1 / 2 / 3 rows ahead 80.3 / 77.9 / 85.6.
// Round 2, late issue (JFNK_MARCH_LATE, ring of ahead + 1 stages): 3 rows ahead in the shared memory 2 rows used to take, with the
// marching-direction fluxes shared between rows: 79.4 -> 76.3 (profiles/march_components_r2.txt, last block).
#ifndef JFNK_MARCH_AHEAD
#define JFNK_MARCH_AHEAD 3
#endif
// JFNK_MARCH_LATE: the elected thread issues the copies of row r + ahead AFTER the row barrier of iteration r instead of
// at its top.  Every thread has then finished iteration r-1 (post-barrier operands included), so the stage of row r-1 can
// be refilled and the ring needs ahead + 1 stages instead of ahead + 2: one more row in flight in the same shared memory.
#ifndef JFNK_MARCH_LATE
#define JFNK_MARCH_LATE 1
#endif
// JFNK_MARCH_YSHARE: the half-point fluxes of the conservative formula in the marching direction are computed once per row
// and carried in registers (P(r+1), H(r+1/2) are new per iteration; P(r-1), P(r), H(r-1/2) come from the two iterations
// before) instead of being recomputed by the two rows that share them: -13 of ~95 fp64 operations per point.
#ifndef JFNK_MARCH_YSHARE
#define JFNK_MARCH_YSHARE 1
#endif
constexpr int kMarchAhead = JFNK_MARCH_AHEAD;             // bulk-TMA path
constexpr int kMarchAheadCp = 1;                          // per-thread cp.async path
constexpr int kMarchStages = kMarchAhead + (JFNK_MARCH_LATE ? 1 : 2); // operand ring depth (the cp.async path uses kMarchAheadCp + 1 of them)
static_assert(kMarchStages >= kMarchAheadCp + 1, "the cp.async path needs two stages");

enum MarchMode { MARCH_LAP = 0, MARCH_PMA2_RESID = 1, MARCH_PMA2_JVP = 2 };
// operand fields staged per row by cp.async: stencil input (x, v), A12 and A22 of row r+2, A11 and J of row r, and the
// pointwise operands of the PMA2 modes (px, pv, uval, cn, f0).  The PMA2 passes never combine their stencil input
// (HAS_V is false), so pv shares the slot of v: 6 fields per stage for MARCH_LAP, 10 for the PMA2 modes.
enum MarchField { MF_X = 0, MF_V = 1, MF_PV = 1, MF_A12N, MF_A22, MF_A11, MF_J, MF_PX, MF_UVAL, MF_CN, MF_F0, MF_COUNT };
__host__ __device__ constexpr int march_fields(int mode) { return mode == MARCH_LAP ? MF_PX : MF_COUNT; }
constexpr int kMarchW = kMarchThreads + 2 * kMarchPad;
__host__ __device__ constexpr size_t march_smem_bytes(int mode) {
  return sizeof(double) * ((size_t)(kMarchRing + 4) * kMarchW + (size_t)kMarchStages * march_fields(mode) * kMarchThreads + kMarchStages);
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct MarchArgs {
  MeshGeom gm;
  MetricCPtrs M;
  const double *x, *v;   // stencil input t = x + a v (v null: t = x)
  ScalarRef a;
  const double *px, *pv; // pointwise u = px + pa pv (PMA2 modes; pv may be null)
  ScalarRef pa;
  const double *uval, *cn, *f0;
  ScalarRef div;
  Pma2Params pp;
  double *out, *out2;
  int norm_off, deriv_bc;
  int rows_per_chunk, nstrips, nframe_ctas;
  int tma;        // operand rows fetched by one elected thread with bulk-TMA copies (1 KiB per field and row, completion on an
                  // mbarrier) instead of one 8-byte cp.async per thread and field: needs nx even and 16-byte aligned fields
  int skippable;  // part of a speculatively enqueued Arnoldi step: return at once while JS_STOP is set
  int debug_skip; // profiling aid (JFNK_MARCH_DEBUG): 1 = frame CTAs idle, 2 = interior CTAs idle, 3 = no row barrier,
                  // 4 = no output phase (staging only), 6 = no operand fetch (compute only); results are then wrong
};

// pma2_rhs_point (mesh_math.h) with the regularisation term (two pow calls, ~200 instructions, epsilon = 0 in PMA2_nk.py's
// runs) kept OUT of the unrolled row loop: same operations in the same order.
__device__ __noinline__ double march_pma2_eps_term(double lambd, double epsilon, int m, double q) {
  return lambd * pow(epsilon, (double)(m - 2)) / pow(q, (double)m);
}
__device__ __forceinline__ double march_pma2_rhs(const Pma2Params& p, double u, double lap2) {
  const double q = 1.0 + u;
  double rhs = -p.lambd / (q * q);
  if (p.epsilon != 0.0) rhs += march_pma2_eps_term(p.lambd, p.epsilon, p.m, q);
  rhs -= p.beta * p.beta * lap2;
  return rhs;
}

// TMA (compile time): operand rows by bulk-TMA copies (one elected thread) / by one 8-byte cp.async per thread and field.  A
// template parameter rather than a run-time flag: the row loop is unrolled 8x and carried both fetch paths in every copy of
// its body (ncu: 16 % of the warp stalls were instruction-cache misses, `stall_no_inst`).
template <int MODE, bool HAS_V, bool TMA>
__global__ void __launch_bounds__(kMarchThreads, JFNK_MARCH_MINCTAS) mesh_march_kernel(const __grid_constant__ MarchArgs A, double* S,
                                                                       ReduceWs ws) {
  if (A.skippable && S[JS_STOP] != 0.0) return;
  constexpr int W = kMarchW;
  constexpr int NF = march_fields(MODE);
  extern __shared__ __align__(16) unsigned char march_smem[];
  double (*Ts)[W] = reinterpret_cast<double (*)[W]>(march_smem);                  // [kMarchRing][W] rows of t
  double (*A11s)[W] = Ts + kMarchRing;                                             // [2][W]
  double (*Fs)[W] = A11s + 2;                                                      // [2][W]
  double* Pf = reinterpret_cast<double*>(Fs + 2);                                  // [stage][field][thread]
  uint64_t* full = reinterpret_cast<uint64_t*>(Pf + (size_t)kMarchStages * NF * kMarchThreads); // [stage], bulk-TMA path
  const MeshGeom& g = A.gm;
  const int nx = g.nx, ny = g.ny;
  const int tid = threadIdx.x;
  const double a = HAS_V ? eval_sref(S, A.a) : 0.0;
  const bool has_pv = (MODE != MARCH_LAP) && A.pv != nullptr;
  const double pa = has_pv ? eval_sref(S, A.pa) : 0.0;
  const double inv_dv = (MODE == MARCH_PMA2_JVP) ? __drcp_rn(eval_sref(S, A.div)) : 1.0;
  const double inv_dt = (MODE != MARCH_LAP) ? __drcp_rn(A.pp.dt) : 1.0;
  double acc[3] = {0.0, 0.0, 0.0};

  // result of one point: tc = stencil input at the point, lap = Lap t, (u, uval, cn, f0) = pointwise operands
  auto emit = [&](size_t e, bool bdy, double tc, double lap, double u, double uval, double cn, double f0) {
    if (MODE == MARCH_LAP) {
      A.out[e] = lap;
      if (A.out2) A.out2[e] = tc;
    } else {
      double rhs = bdy ? 0.0 : march_pma2_rhs(A.pp, u, lap);
      double F = (u - uval) * inv_dt - (rhs + cn) * 0.5; // pma2_combine_point with 1/dt rounded once
      if (MODE == MARCH_PMA2_RESID) {
        A.out[e] = F;
        acc[0] = fma(F, F, acc[0]); acc[1] = fmax(acc[1], fabs(F)); acc[2] = fmax(acc[2], fabs(u));
      } else {
        A.out[e] = (F - f0) * inv_dv;
      }
    }
  };

  if ((int)blockIdx.x < A.nframe_ctas) {
    if (A.debug_skip == 1) return;
    // ---- frame: rows 0..3, ny-4..ny-1 (8 nx points), then columns 0..3, nx-4..nx-1 of the remaining rows ----
    const long long nrowpts = 8LL * nx, nf = nrowpts + 8LL * (ny - 8);
    auto tf = [&](int r, int c) -> double {
      size_t e = (size_t)r * nx + c;
      double xv = A.x[e];
      if (HAS_V) xv = combine(xv, a, A.v[e]);
      return xv;
    };
    for (long long i = (long long)blockIdx.x * kMarchThreads + tid; i < nf; i += (long long)A.nframe_ctas * kMarchThreads) {
      int r, c;
      if (i < nrowpts) {
        int fr = (int)(i / nx);
        r = fr < 4 ? fr : ny - 8 + fr;
        c = (int)(i - (long long)fr * nx);
      } else {
        long long k = i - nrowpts;
        int cc = (int)(k & 7);
        r = 4 + (int)(k >> 3);
        c = cc < 4 ? cc : nx - 8 + cc;
      }
      double xx, yy;
      mesh_laplace_general_g(g, A.M.m, tf, r, c, A.deriv_bc, xx, yy);
      const size_t e = (size_t)r * nx + c;
      const bool bdy = (r == 0 || c == 0 || r == ny - 1 || c == nx - 1);
      double u = 0.0, uval = 0.0, cn = 0.0, f0 = 0.0;
      if (MODE != MARCH_LAP) {
        u = A.px[e];
        if (has_pv) u = combine(u, pa, A.pv[e]);
        uval = A.uval[e]; cn = A.cn[e];
        if (MODE == MARCH_PMA2_JVP) f0 = A.f0[e];
      }
      emit(e, bdy, MODE == MARCH_LAP ? tf(r, c) : 0.0, xx + yy, u, uval, cn, f0);
    }
  } else {
    // ---- interior: march down rows [r0, r1) of one strip ----
    if (A.debug_skip == 2) return;
    const int id = (int)blockIdx.x - A.nframe_ctas;
    const int strip = id % A.nstrips, chunk = id / A.nstrips;
    const int r0 = 4 + chunk * A.rows_per_chunk;
    const int r1 = min(ny - 4, r0 + A.rows_per_chunk);
    constexpr bool tma = TMA;
    const int halo = march_halo(tma), outw = march_out(tma);
    const int c_strip = 4 + strip * outw - halo; // first column of the strip (even with bulk-TMA fetch)
    const int c_raw = c_strip + tid;
    const int c = min(c_raw, nx - 1); // threads past the right edge load a valid column and compute nothing
    const bool is_out = tid >= halo && tid < halo + outw && c_raw < nx - 4;
    if (tma) {
      if (tid == 0) {
        for (int i = 0; i < kMarchStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
    }
    const int st = tid + kMarchPad;
    const double* __restrict__ Jp = A.M.m[3];
    const double* __restrict__ A11p = A.M.m[4];
    const double* __restrict__ A22p = A.M.m[5];
    const double* __restrict__ A12p = A.M.m[6];
    const double wx0 = g.c1x[0], wx1 = g.c1x[1], wx3 = g.c1x[3], wx4 = g.c1x[4];
    const double wy0 = g.c1y[0], wy1 = g.c1y[1], wy3 = g.c1y[3], wy4 = g.c1y[4];
    // (the centred first-derivative weight of the point itself is 0 and is left out of the sums)

    // Register windows as circular buffers of period 8 (the row loop is unrolled 8x, so every index below is a
    // compile-time constant and no register is ever moved; only the live rows hold registers): with m counted from the
    // first warm-up row,
    //   tw[m % 8]           = t(r0 - 3 + m, c)   rows r-3..r+3 of output row r = r0 + 8 it + u are slots (u + k + 3) % 8
    //   a22w[m % 8], gw[..] = A22, g (r0 - 2 + m, c)   rows r-2..r+2 are slots (u + k + 2) % 8
    // The shared-memory rings use the same relative row count: row r + k of t is Ts[(u + k + 3) % 8], the operand stage of
    // row r is u % stages, the A11 / f slot u % 2 -- every shared-memory address of the loop body is base + immediate.
    constexpr int PER = 8;
    static_assert(kMarchRing == PER, "the t ring and the register windows share the period");
    double tw[PER], a22w[PER], gw[PER];
    double a12_0 = 0.0, a12_1 = 0.0, a12_2 = 0.0;
    const double rx = __drcp_rn(288 * g.dksi2), ry = __drcp_rn(288 * g.deta2); // one rounding each, instead of a
                                                                              // division per point and direction

    // Operand pipeline: the global operands of row r (own column of every field) are copied into a per-thread slot
    // of a shared-memory ring by cp.async kMarchAhead rows before they are used -- ~4 us of loads in flight per
    // thread without holding a register, enough to cover HBM latency under load.  A thread only ever reads the slots
    // it filled itself, so completion is tracked with cp.async groups and needs no barrier.
    // bulk-TMA path: ring of kMarchStages stages indexed by the row count.  Issued at the top of iteration r (ahead + 2 stages),
    // the refilled stage was last read before the barrier of iteration r-1 (its post-barrier operands J, u, uval, cn, f0
    // included); issued after the barrier of iteration r (JFNK_MARCH_LATE, ahead + 1 stages) it is the stage of row r-1.
    const uint32_t row_bytes = (uint32_t)(min(kMarchThreads, nx - c_strip) * 8);
    // qrel = (r - r0) mod 8, a compile-time constant at every call site
    auto issue = [&](int r, int qrel) {
      if (tma) {
        // (the issuing lane rotates over the warps: ~30 instructions per row that would otherwise make warp 0 the last to
        //  reach every row barrier)
        if (tid == ((qrel & (kMarchThreads / 32 - 1)) << 5) && r < r1 && A.debug_skip != 6) {
          const int stg = (PER % kMarchStages == 0) ? qrel % kMarchStages : (r - r0) % kMarchStages;
          double* pf = Pf + (size_t)stg * NF * kMarchThreads;
          uint64_t* bar = &full[stg];
          const size_t o3 = (size_t)(r + 3) * nx + c_strip, o2 = (size_t)(r + 2) * nx + c_strip, o0 = (size_t)r * nx + c_strip;
          int nf = 5 + (HAS_V ? 1 : 0);
          if (MODE != MARCH_LAP) nf += 3 + (has_pv ? 1 : 0) + (MODE == MARCH_PMA2_JVP ? 1 : 0);
          mbar_expect_tx(bar, row_bytes * (uint32_t)nf);
          tma_load(pf + MF_X * kMarchThreads, A.x + o3, row_bytes, bar);
          if (HAS_V) tma_load(pf + MF_V * kMarchThreads, A.v + o3, row_bytes, bar);
          tma_load(pf + MF_A12N * kMarchThreads, A12p + o2, row_bytes, bar);
          tma_load(pf + MF_A22 * kMarchThreads, A22p + o2, row_bytes, bar);
          tma_load(pf + MF_A11 * kMarchThreads, A11p + o0, row_bytes, bar);
          tma_load(pf + MF_J * kMarchThreads, Jp + o0, row_bytes, bar);
          if (MODE != MARCH_LAP) {
            tma_load(pf + MF_PX * kMarchThreads, A.px + o0, row_bytes, bar);
            if (has_pv) tma_load(pf + MF_PV * kMarchThreads, A.pv + o0, row_bytes, bar);
            tma_load(pf + MF_UVAL * kMarchThreads, A.uval + o0, row_bytes, bar);
            tma_load(pf + MF_CN * kMarchThreads, A.cn + o0, row_bytes, bar);
            if (MODE == MARCH_PMA2_JVP) tma_load(pf + MF_F0 * kMarchThreads, A.f0 + o0, row_bytes, bar);
          }
        }
        return;
      }
      if (r < r1 && A.debug_skip != 6) {
        double* pf = Pf + (size_t)(qrel % (kMarchAheadCp + 1)) * NF * kMarchThreads + tid;
        const size_t o3 = (size_t)(r + 3) * nx + c, o2 = (size_t)(r + 2) * nx + c, o0 = (size_t)r * nx + c;
        cp_async8(pf + MF_X * kMarchThreads, A.x + o3);
        if (HAS_V) cp_async8(pf + MF_V * kMarchThreads, A.v + o3);
        cp_async8(pf + MF_A12N * kMarchThreads, A12p + o2);
        cp_async8(pf + MF_A22 * kMarchThreads, A22p + o2);
        cp_async8(pf + MF_A11 * kMarchThreads, A11p + o0);
        cp_async8(pf + MF_J * kMarchThreads, Jp + o0);
        if (MODE != MARCH_LAP) {
          cp_async8(pf + MF_PX * kMarchThreads, A.px + o0);
          if (has_pv) cp_async8(pf + MF_PV * kMarchThreads, A.pv + o0);
          cp_async8(pf + MF_UVAL * kMarchThreads, A.uval + o0);
          cp_async8(pf + MF_CN * kMarchThreads, A.cn + o0);
          if (MODE == MARCH_PMA2_JVP) cp_async8(pf + MF_F0 * kMarchThreads, A.f0 + o0);
        }
      }
      cp_async_commit(); // (an empty group past the last row keeps the group count uniform)
    };
    constexpr int ahead = tma ? kMarchAhead : kMarchAheadCp;
#pragma unroll
    for (int k = 0; k < ahead; ++k) issue(r0 + k, k); // (before the warm-up, so that the first rows land behind it)

    // warm-up: rows r0-3 .. r0+2 of t ; g and A22 of rows r0-2 .. r0+1.  All global loads are issued before the first use and
    // the six rows go into the ring behind ONE barrier (a load -> store -> barrier chain per row cost six serial memory
    // latencies per chunk of ~60 rows).
    {
      double wa12[4];
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const size_t off = (size_t)(r0 - 3 + m) * nx + c;
        double tv = __ldg(A.x + off);
        if (HAS_V) tv = combine(tv, a, __ldg(A.v + off));
        tw[m] = tv;
        if (m >= 1 && m <= 4) {
          wa12[m - 1] = __ldg(A12p + off);
          a22w[m - 1] = __ldg(A22p + off);
        }
      }
#pragma unroll
      for (int m = 0; m < 6; ++m) Ts[m][st] = tw[m];
      __syncthreads();
#pragma unroll
      for (int m = 1; m <= 4; ++m) {
        const double* T = Ts[m] + st;
        const double vk = wx0 * T[-2] + wx1 * T[-1] + wx3 * T[1] + wx4 * T[2];
        gw[m - 1] = wa12[m - 1] * vk;
      }
      a12_1 = wa12[2]; a12_2 = wa12[3]; // A12 of rows r0, r0+1
    }

#if JFNK_MARCH_YSHARE
    // fluxes of the marching direction carried between rows: P(r0-1), P(r0), H(r0-1/2) from the warm-up rows
    double py_m1 = a22w[1] * (tw[0] - 8 * tw[1] + 8 * tw[3] - tw[4]);
    double py_0 = a22w[2] * (tw[1] - 8 * tw[2] + 8 * tw[4] - tw[5]);
    double hy_m = (-a22w[0] + 9 * a22w[1] + 9 * a22w[2] - a22w[3]) * (tw[1] - 27 * tw[2] + 27 * tw[3] - tw[4]);
#endif
    constexpr bool kStagesDivide = (PER % kMarchStages == 0);
    const double* const pf0 = Pf + tid;
    for (int rb = r0; rb < r1; rb += PER) {
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int r = rb + u;
        if (r >= r1) break;
        if (!JFNK_MARCH_LATE || !tma) issue(r + ahead, (u + ahead) % PER);
        const int q = r - r0; // = 8 it + u
        const int stg = tma ? (kStagesDivide ? u % kMarchStages : q % kMarchStages) : u % (kMarchAheadCp + 1);
        if (tma) {
          // (with 2 or 4 stages the phase of stage u % stages in period it is a constant as well)
          const uint32_t par = (kStagesDivide && (PER / kMarchStages) % 2 == 0) ? (uint32_t)((u / kMarchStages) & 1)
                                                                                : (uint32_t)((q / kMarchStages) & 1);
          if (A.debug_skip != 6) mbar_wait(&full[stg], par); // row r has landed
        } else cp_async_wait<kMarchAheadCp>(); // the group of row r has landed
        const double* pf = pf0 + (size_t)stg * NF * kMarchThreads;
        const double tnew = HAS_V ? combine(pf[MF_X * kMarchThreads], a, pf[MF_V * kMarchThreads]) : pf[MF_X * kMarchThreads];
        tw[(u + 6) % PER] = tnew;
        a22w[(u + 4) % PER] = pf[MF_A22 * kMarchThreads];
        a12_0 = a12_1; a12_1 = a12_2; a12_2 = pf[MF_A12N * kMarchThreads]; // A12 of rows r, r+1, r+2
        const double a12r2 = a12_2, a12r = a12_0, a11 = pf[MF_A11 * kMarchThreads];

        const int slot = u & 1;
        Ts[(u + 6) % PER][st] = tnew; // row r + 3
        A11s[slot][st] = a11;
#define TW(k) tw[(u + (k) + 3) % PER]
#define AY(k) a22w[(u + (k) + 2) % PER]
#define GW(k) gw[(u + (k) + 2) % PER]
        // D_eta t at (r, c) from the register window ; f = A12 D_eta t
        const double ve = wy0 * TW(-2) + wy1 * TW(-1) + wy3 * TW(1) + wy4 * TW(2);
        Fs[slot][st] = a12r * ve;
        {
          // D_ksi t at (r+2, c) from the row staged one iteration ago ; g = A12 D_ksi t
          const double* T2 = Ts[(u + 5) % PER] + st;
          const double vk = wx0 * T2[-2] + wx1 * T2[-1] + wx3 * T2[1] + wx4 * T2[2];
          GW(2) = a12r2 * vk;
        }
        if (A.debug_skip != 3 || tma) __syncthreads(); // (the bulk-TMA ring relies on this barrier for its slot reuse)
        if (JFNK_MARCH_LATE && tma) issue(r + ahead, (u + ahead) % PER);  // refills the stage of row r-1: every thread is past iteration r-1
        if (is_out && A.debug_skip != 4) {
          const double* T = Ts[(u + 3) % PER] + st;
          const double* AX = A11s[slot] + st;
          const double* FX = Fs[slot] + st;
          const double xx = (4 * (AX[-1] * (T[-3] - 8 * T[-2] + 8 * T[0] - T[1])) -
                             (-AX[-2] + 9 * AX[-1] + 9 * AX[0] - AX[1]) * (T[-2] - 27 * T[-1] + 27 * T[0] - T[1]) +
                             (-AX[-1] + 9 * AX[0] + 9 * AX[1] - AX[2]) * (T[-1] - 27 * T[0] + 27 * T[1] - T[2]) -
                             4 * (AX[1] * (T[-1] - 8 * T[0] + 8 * T[2] - T[3]))) * rx;
#if JFNK_MARCH_YSHARE
          const double pnew = AY(1) * (TW(-1) - 8 * TW(0) + 8 * TW(2) - TW(3));                                          // P(r+1)
          const double hnew = (-AY(-1) + 9 * AY(0) + 9 * AY(1) - AY(2)) * (TW(-1) - 27 * TW(0) + 27 * TW(1) - TW(2));   // H(r+1/2)
          const double yy = (4 * py_m1 - hy_m + hnew - 4 * pnew) * ry;
          py_m1 = py_0; py_0 = pnew; hy_m = hnew;
#else
          const double yy = (4 * (AY(-1) * (TW(-3) - 8 * TW(-2) + 8 * TW(0) - TW(1))) -
                             (-AY(-2) + 9 * AY(-1) + 9 * AY(0) - AY(1)) * (TW(-2) - 27 * TW(-1) + 27 * TW(0) - TW(1)) +
                             (-AY(-1) + 9 * AY(0) + 9 * AY(1) - AY(2)) * (TW(-1) - 27 * TW(0) + 27 * TW(1) - TW(2)) -
                             4 * (AY(1) * (TW(-1) - 8 * TW(0) + 8 * TW(2) - TW(3)))) * ry;
#endif
          double accx = 0.0, accy = 0.0;
          accx += wx0 * FX[-2]; accx += wx1 * FX[-1]; accx += wx3 * FX[1]; accx += wx4 * FX[2];
          accy += wy0 * GW(-2); accy += wy1 * GW(-1); accy += wy3 * GW(1); accy += wy4 * GW(2);
          const double rJ = __drcp_rn(pf[MF_J * kMarchThreads]);
          const double lap = (xx + accx) * rJ + (yy + accy) * rJ;
          double uu = 0.0, uval = 0.0, cn = 0.0, f0 = 0.0;
          if (MODE != MARCH_LAP) {
            uu = pf[MF_PX * kMarchThreads];
            if (has_pv) uu = combine(uu, pa, pf[MF_PV * kMarchThreads]);
            uval = pf[MF_UVAL * kMarchThreads];
            cn = pf[MF_CN * kMarchThreads];
            if (MODE == MARCH_PMA2_JVP) f0 = pf[MF_F0 * kMarchThreads];
          }
          emit((size_t)r * nx + c, false, TW(0), lap, uu, uval, cn, f0);
        }
#undef TW
#undef AY
#undef GW
      }
    }
  }
  if (MODE == MARCH_PMA2_RESID) grid_reduce<3>(acc, 0x6u, ws, S + A.norm_off);
}

} // namespace jfnk
