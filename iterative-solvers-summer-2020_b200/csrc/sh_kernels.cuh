// Matrix-free Swift-Hohenberg stencil kernels (periodic 5-point Laplacian and the 13-point
// L = -Lap^2 - 2 Lap + (r-1) I), with the nonlinear terms, the Crank-Nicolson combination, the
// finite-difference Jacobian-vector product and the norms fused into the same pass.
//
// Two kernels share one arithmetic core (sh_value):
//  * sh_tma_kernel -- the product path for large even grids: a persistent, warp-specialised marching
//    kernel.  Each CTA owns a contiguous run of rows of a 256-column strip (the global list of
//    strip-rows is cut into equal contiguous ranges, one per resident CTA, so the load balance is exact
//    and only 4 warm-up rows per run are read twice).  One producer thread streams whole rows (strip +
//    2 halo columns per side; operand rows x, v and the pointwise rows d, f0) from HBM into a
//    multi-stage shared-memory ring with TMA bulk copies (cp.async.bulk, completion on mbarriers);
//    four consumer warps (one double2 per thread) read their columns and horizontal neighbours straight
//    from the staged row, keep the vertical neighbours in a register window (5 rows of u, the horizontal
//    pair sums of 4 rows) and write the result with coalesced 128-bit stores.  Every input row is read
//    from HBM once; the input combination t = x + a v of the JVP / line search is formed when the row is
//    consumed, so F(x0 + sc z) never exists in memory.
//  * sh_point_kernel -- one thread per grid point straight from global memory; used for small / odd
//    grids (61 x 61, 91 x 61 ...) where launch latency, not bandwidth, is the bound, and as an independent
//    cross-check of the marching kernel in the parity tests (kernel_variant = 1).
//
// Algorithmic bytes per grid point (fp64): spmv 16, set_prev 16, residual 24 (+8 when x+s dx is stored),
// JVP 40 (x0, z, d, f0 in; w out), linearised matvec 24.
#pragma once
#include "cuda_common.cuh"
#include "p2p_kernels.cuh"

namespace jfnk {

// OP_JVPG: the FD quotient against the stored G(x0) = F(x0) + d -- (G(x0 + sc z) - G(x0))/div, one pointwise operand
// (32 B/point) instead of d and f0 (40 B/point); OP_JVP keeps the two-operand form for callers without G(x0).
enum ShOp { OP_LAP = 0, OP_L = 1, OP_SETPREV = 2, OP_RESID = 3, OP_JVP = 4, OP_LINPREP = 5, OP_LINMV = 6, OP_JVPG = 7 };

struct ShArgs {
  const double *x, *xtop, *xbot; // primary field and its 2-row halos (top = rows -2,-1; bot = rows nrows, nrows+1)
  const double *v, *vtop, *vbot; // optional second field: t = x + a v
  ScalarRef a;                   // combination coefficient
  ScalarRef div;                 // JVP divisor ; LINMV output scale
  const double* d;               // RESID/JVP: per-step constant d ; LINMV: diagonal D ; JVPG: G(x0)
  const double* f0;              // JVP: f0 ; LINPREP: Uo
  double* out;                   // result field
  double* out2;                  // RESID: x + a v (may be null) ; LINPREP: D
  double* out3;                  // RESID: G(x + a v) = F + d (may be null): what OP_JVPG subtracts once this point is accepted
  int nx, nrows;
  int norm_off;                  // RESID: S[norm_off..+2] = sum F^2, max|F|, max|t|
  int mode;                      // sh_box_kernel tuning bits: 1 = streaming (evict-first) stores, 2 = band-major CTA mapping
  // Fused halo exchange over peer memory (marching kernel, slab ranks): the kernel itself stores the first / last two rows of
  // `push_field` (1: x, 2: v) into the neighbours' halo buffers, raises their arrival flags, and its producer lane waits for
  // this rank's own flags only at the moment it needs a halo row -- the exchange overlaps the interior rows.  0 = off.
  int push_field;
  P2PHaloArgs push;
};

__device__ __forceinline__ const double* sh_row(const double* base, const double* top, const double* bot, int r, int nx,
                                                int nrows) {
  if (r < 0) return top + (size_t)(r + 2) * nx;
  if (r >= nrows) return bot + (size_t)(r - nrows) * nx;
  return base + (size_t)r * nx;
}

struct ShAcc {
  double f2, fmax, xmax;
};

// one output value (returned) plus the optional second output (RESID: x + a v ; LINPREP: D)
template <int OP>
__device__ __forceinline__ double sh_value(const SHParams& P, double scale, double uc, double s1, double sd, double s2,
                                           double dval, double f0val, double& second, double& third, ShAcc& acc) {
  if (OP == OP_LAP) return sh_apply5(P, uc, s1);
  double Lu = sh_apply13(P, uc, s1, sd, s2);
  if (OP == OP_L) return Lu;
  if (OP == OP_SETPREV) return sh_prev_const(P, uc, Lu);
  if (OP == OP_JVPG) return (sh_G(P, uc, Lu) - dval) * scale; // dval = G(x0), scale = 1/div
  if (OP == OP_RESID) {
    double G = sh_G(P, uc, Lu);
    double F = G - dval;
    second = uc;
    third = G;
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
  if (OP == OP_JVP || OP == OP_JVPG) return 1.0 / eval_sref(S, A.div);
  if (OP == OP_LINMV) return eval_sref(S, A.div);
  return 1.0;
}

// the passes that are the Jacobian / linear operator inside the Arnoldi loop: they return at once while JS_STOP is set
template <int OP>
__host__ __device__ constexpr bool sh_is_operator() { return OP == OP_JVP || OP == OP_JVPG || OP == OP_LINMV; }

// ---------------------------------------------------------------------------------------------------
// one thread per point
// ---------------------------------------------------------------------------------------------------
template <int OP, bool HAS_V>
__global__ void __launch_bounds__(256) sh_point_kernel(ShArgs A, SHParams P, double* S, ReduceWs ws) {
  if (sh_is_operator<OP>() && S[JS_STOP] != 0.0) return; // speculatively enqueued Arnoldi step after the process stopped
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
    double dval = (OP == OP_RESID || OP == OP_JVP || OP == OP_JVPG || OP == OP_LINMV) ? A.d[e] : 0.0;
    double f0val = (OP == OP_JVP || OP == OP_LINPREP) ? A.f0[e] : 0.0;
    double second = 0.0, third = 0.0;
    A.out[e] = sh_value<OP>(P, scale, uc, s1, sd, s2, dval, f0val, second, third, acc);
    if ((OP == OP_RESID && A.out2) || OP == OP_LINPREP) A.out2[e] = second;
    if (OP == OP_RESID && A.out3) A.out3[e] = third;
  }
  if (OP == OP_RESID) {
    double val[3] = {acc.f2, acc.fmax, acc.xmax};
    grid_reduce<3>(val, 0x6u, ws, S + A.norm_off);
  }
}


// ---------------------------------------------------------------------------------------------------
// TMA-pipelined marching kernel
// ---------------------------------------------------------------------------------------------------
constexpr int kTmaTX = 256;        // columns per strip = 2 per consumer thread
constexpr int kTmaConsumers = 128; // 4 consumer warps
constexpr int kTmaThreads = kTmaConsumers + 32; // + 1 producer warp (one elected lane issues the copies)
constexpr int kShPushCtas = 16;    // CTAs that carry the fused halo push (32 KiB each at nx = 16384)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int OP, bool HAS_V>
struct TmaLayout {
  static constexpr bool kHasD = (OP == OP_RESID || OP == OP_JVP || OP == OP_JVPG || OP == OP_LINMV);
  static constexpr bool kHasF = (OP == OP_JVP || OP == OP_LINPREP);
  static constexpr int kX = 0;                                  // offsets in doubles inside one stage
  static constexpr int kV = kTmaTX + 4;
  static constexpr int kD = kV + (HAS_V ? kTmaTX + 4 : 0);
  static constexpr int kF = kD + (kHasD ? kTmaTX : 0);
  static constexpr int kStageDoubles = kF + (kHasF ? kTmaTX : 0);
  static constexpr int kStages = (kStageDoubles * 8 > 6000) ? 8 : (kStageDoubles * 8 > 3000 ? 12 : 16);
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageDoubles * 8 + 2 * kStages * sizeof(uint64_t);
};

struct TmaWork {
  long long begin, end; // this CTA's range of global strip-row indices (strip-major: idx = strip*nrows + row)
};

template <int OP, bool HAS_V>
__global__ void __launch_bounds__(kTmaThreads) sh_tma_kernel(ShArgs A, SHParams P, double* S, ReduceWs ws) {
  if (sh_is_operator<OP>() && S[JS_STOP] != 0.0) return; // (before the fused halo push: every rank skips alike)
  using LY = TmaLayout<OP, HAS_V>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)LY::kStages * LY::kStageDoubles * 8);
  uint64_t* empty = full + LY::kStages;

  const int nx = A.nx, nrows = A.nrows;
  const int strips = (nx + kTmaTX - 1) / kTmaTX;
  const long long total = (long long)strips * nrows;
  const long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LY::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTmaConsumers / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (A.push_field) {
    // the first few CTAs push this rank's boundary rows to the ring neighbours (16-byte stores over NVLink); the last of
    // them to finish raises the neighbours' flags.  Nothing here waits, so every rank always gets to raise its flags.
    const unsigned npush = min(gridDim.x, (unsigned)kShPushCtas);
    if (blockIdx.x < npush) {
      const size_t n2 = A.push.count >> 1; // double2 elements per message
      const size_t lo = n2 * blockIdx.x / npush, hi = n2 * (blockIdx.x + 1) / npush;
      const double2* s0 = reinterpret_cast<const double2*>(A.push.src_first);
      const double2* s1 = reinterpret_cast<const double2*>(A.push.src_last);
      double2* d0 = reinterpret_cast<double2*>(A.push.dst_prev_bot);
      double2* d1 = reinterpret_cast<double2*>(A.push.dst_next_top);
      for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) { d0[i] = s0[i]; d1[i] = s1[i]; }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(A.push.ticket, 1u) == npush - 1) {
          *A.push.ticket = 0u;
          __threadfence_system();
          st_release_sys(A.push.flag_prev_bot, A.push.epoch);
          st_release_sys(A.push.flag_next_top, A.push.epoch);
        }
      }
    }
  }
  __syncthreads();

  ShAcc acc = {0.0, 0.0, 0.0};
  if (warp == kTmaConsumers / 32) {
    // ------------------------------- producer: one lane streams rows into the ring -----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long idx = begin;
      bool top_ok = A.push_field == 0, bot_ok = A.push_field == 0;
      while (idx < end) {
        const int strip = (int)(idx / nrows), r0 = (int)(idx - (long long)strip * nrows);
        const long long run_end = min(end, (long long)(strip + 1) * nrows);
        const int r1 = r0 + (int)(run_end - idx);
        const int xs = strip * kTmaTX, W = min(kTmaTX, nx - xs);
        const int cl = xs - 2 < 0 ? xs - 2 + nx : xs - 2, cr = xs + W >= nx ? xs + W - nx : xs + W;
        // fused exchange: the neighbours' rows must have arrived before this run fetches a halo row (checked once per
        // run, not per row: the producer lane is the pipeline's critical resource)
        if (!top_ok && r0 < 2) {
          wait_flag(A.push.my_flag_top, A.push.epoch, A.push.err);
          asm volatile("fence.proxy.async;" ::: "memory");
          top_ok = true;
        }
        if (!bot_ok && r1 + 2 > nrows) {
          wait_flag(A.push.my_flag_bot, A.push.epoch, A.push.err);
          asm volatile("fence.proxy.async;" ::: "memory");
          bot_ok = true;
        }
        for (int ra = r0 - 2; ra < r1 + 2; ++ra) {
          const int y = ra - 2;
          const bool pt = y >= r0;
          mbar_wait(&empty[s], ph ^ 1u);
          double* st = stage0 + (size_t)s * LY::kStageDoubles;
          uint32_t bytes = (uint32_t)(W + 4) * 8u * (HAS_V ? 2u : 1u);
          if (pt) bytes += (uint32_t)W * 8u * ((LY::kHasD ? 1u : 0u) + (LY::kHasF ? 1u : 0u));
          mbar_expect_tx(&full[s], bytes);
          {
            const double* row = sh_row(A.x, A.xtop, A.xbot, ra, nx, nrows);
            tma_load(st + LY::kX, row + cl, 16u, &full[s]);
            tma_load(st + LY::kX + 2, row + xs, (uint32_t)W * 8u, &full[s]);
            tma_load(st + LY::kX + 2 + W, row + cr, 16u, &full[s]);
          }
          if (HAS_V) {
            const double* row = sh_row(A.v, A.vtop, A.vbot, ra, nx, nrows);
            tma_load(st + LY::kV, row + cl, 16u, &full[s]);
            tma_load(st + LY::kV + 2, row + xs, (uint32_t)W * 8u, &full[s]);
            tma_load(st + LY::kV + 2 + W, row + cr, 16u, &full[s]);
          }
          if (pt) {
            const size_t e = (size_t)y * nx + xs;
            if (LY::kHasD) tma_load(st + LY::kD, A.d + e, (uint32_t)W * 8u, &full[s]);
            if (LY::kHasF) tma_load(st + LY::kF, A.f0 + e, (uint32_t)W * 8u, &full[s]);
          }
          if (++s == LY::kStages) { s = 0; ph ^= 1u; }
        }
        idx = run_end;
      }
    }
  } else {
    // ------------------------------- consumers: 2 columns per thread ----------------------------------
    const double a = HAS_V ? eval_sref(S, A.a) : 0.0;
    const double scale = sh_scale<OP>(A, S);
    const int t2 = 2 * threadIdx.x; // column offset inside the strip
    const double2 zero2 = make_double2(0.0, 0.0);
    int s = 0;
    uint32_t ph = 0;
    long long idx = begin;
    while (idx < end) {
      const int strip = (int)(idx / nrows), r0 = (int)(idx - (long long)strip * nrows);
      const long long run_end = min(end, (long long)(strip + 1) * nrows);
      const int r1 = r0 + (int)(run_end - idx);
      const int xs = strip * kTmaTX, W = min(kTmaTX, nx - xs);
      const bool act = t2 < W;
      double2 u0 = zero2, u1 = zero2, u2 = zero2, u3 = zero2, u4 = zero2;
      double2 p0 = zero2, p1 = zero2, p2 = zero2, p3 = zero2; // horizontal +-1 pair sums, rows ra-3..ra
      double2 q0 = zero2, q1 = zero2, q2 = zero2;             // horizontal +-2 pair sums, rows ra-2..ra
      for (int ra = r0 - 2; ra < r1 + 2; ++ra) {
        const int y = ra - 2;
        mbar_wait(&full[s], ph);
        const double* st = stage0 + (size_t)s * LY::kStageDoubles;
        double2 L = zero2, own = zero2, R = zero2, dv = zero2, fv = zero2;
        if (act) {
          const double* sx = st + LY::kX + t2;
          L = *reinterpret_cast<const double2*>(sx);
          own = *reinterpret_cast<const double2*>(sx + 2);
          R = *reinterpret_cast<const double2*>(sx + 4);
          if (HAS_V) {
            const double* sv = st + LY::kV + t2;
            double2 vl = *reinterpret_cast<const double2*>(sv);
            double2 vo = *reinterpret_cast<const double2*>(sv + 2);
            double2 vr = *reinterpret_cast<const double2*>(sv + 4);
            L.x = combine(L.x, a, vl.x); L.y = combine(L.y, a, vl.y);
            own.x = combine(own.x, a, vo.x); own.y = combine(own.y, a, vo.y);
            R.x = combine(R.x, a, vr.x); R.y = combine(R.y, a, vr.y);
          }
          if (y >= r0) {
            if (LY::kHasD) dv = *reinterpret_cast<const double2*>(st + LY::kD + t2);
            if (LY::kHasF) fv = *reinterpret_cast<const double2*>(st + LY::kF + t2);
          }
        }
        // this warp is done with the stage: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == LY::kStages) { s = 0; ph ^= 1u; }

        u0 = u1; u1 = u2; u2 = u3; u3 = u4; u4 = own;
        p0 = p1; p1 = p2; p2 = p3;
        p3.x = L.y + own.y; p3.y = own.x + R.x;
        q0 = q1; q1 = q2;
        q2.x = L.x + R.x; q2.y = L.y + R.y;
        if (act && y >= r0) {
          // rows: u0..u4 = y-2..y+2 ; p0,p1,p2 = pair sums of rows y-1,y,y+1 ; q0 = +-2 pair sum of row y
          double s1x = p1.x + u1.x + u3.x, s1y = p1.y + u1.y + u3.y;
          double sdx = p0.x + p2.x, sdy = p0.y + p2.y;
          double s2x = q0.x + u0.x + u4.x, s2y = q0.y + u0.y + u4.y;
          const size_t e = (size_t)y * nx + xs + t2;
          double2 o, o2 = zero2, o3 = zero2;
          o.x = sh_value<OP>(P, scale, u2.x, s1x, sdx, s2x, dv.x, fv.x, o2.x, o3.x, acc);
          o.y = sh_value<OP>(P, scale, u2.y, s1y, sdy, s2y, dv.y, fv.y, o2.y, o3.y, acc);
          stg2(A.out + e, o);
          if ((OP == OP_RESID && A.out2) || OP == OP_LINPREP) stg2(A.out2 + e, o2);
          if (OP == OP_RESID && A.out3) stg2(A.out3 + e, o3);
        }
      }
      idx = run_end;
    }
  }
  if (OP == OP_RESID) {
    double val[3] = {acc.f2, acc.fmax, acc.xmax};
    grid_reduce<3>(val, 0x6u, ws, S + A.norm_off);
  }
}

} // namespace jfnk
