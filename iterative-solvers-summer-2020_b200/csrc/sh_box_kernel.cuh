// sh_box_kernel: the Swift-Hohenberg marching stencil with TENSOR-MAP TMA staging (cp.async.bulk.tensor.2d, SASS
// UTMALDG) -- the product path for large even grids (nx >= 512, even number of rows per slab).
//
// Same arithmetic core (sh_value, sh_kernels.cuh) and the same marching scheme as sh_tma_kernel, re-laid-out so that the
// single producer lane is no longer the critical resource of the pipeline:
//   * strips are 512 columns wide (two 256-column boxes side by side: 256 elements is the tensor-map box limit per
//     dimension) and a pipeline stage holds TWO rows, so one TMA instruction moves a 256 x 2 box = 4 KiB; the two
//     periodic halo columns on each side arrive as 2 x 2 boxes (at wrapped column coordinates: no out-of-bounds column
//     is ever needed, a partial last strip is zero-filled by the TMA unit).  A Jacobian-vector-product stage
//     (2 rows x 512 columns of x, z, d, f0 = 32 KiB) takes 12 instructions where sh_tma_kernel issues 32;
//   * runs start at even rows, so a box is either entirely inside the slab or entirely one of the two 2-row halos (the
//     slab's own opposite edge on one GPU, the neighbours' rows in the peer / NCCL halo buffer with row slabs): the halo
//     rows have their own tensor maps and no box ever straddles an edge;
//   * 256 consumer threads (one double2 = two columns each) + one producer warp per CTA, two CTAs per SM for the
//     four-operand passes (3 stages x 33 KiB), three for the SpMV (8 stages x 8.3 KiB).
// Every input row is read from HBM once per strip (+ 2 warm-up row pairs per run); 64 B of halo columns per 4 KiB row.
#pragma once
#include <cuda.h>
#include "sh_kernels.cuh"

namespace jfnk {

constexpr int kBoxW = 256;          // columns per TMA box
constexpr int kBoxR = 2;            // rows per box = rows per pipeline stage
constexpr int kBoxStrip = 2 * kBoxW; // columns per strip
constexpr int kBoxConsumers = 256;
constexpr int kBoxThreads = kBoxConsumers + 32;
constexpr int kBoxMainDoubles = 2 * kBoxW * kBoxR; // both boxes of one field in one stage
constexpr int kBoxHaloDoubles = 16;                // a 2 x 2 box, padded to the 128-byte TMA destination alignment

struct alignas(64) ShBoxMaps {
  // per stencil field: (256 x 2) boxes over the slab rows / over the 4 halo rows, and (2 x 2) halo-column boxes over the same
  CUtensorMap x_main, x_halo, xc_main, xc_halo;
  CUtensorMap v_main, v_halo, vc_main, vc_halo;
  CUtensorMap d, f0; // pointwise operands, (256 x 2) boxes
};
struct ShBoxRows {
  int x_top, x_bot, v_top, v_bot; // row coordinate inside the halo map of the pair above row 0 / below row nrows-1
};

template <int OP, bool HAS_V>
struct BoxLayout {
  static constexpr bool kHasD = (OP == OP_RESID || OP == OP_JVP || OP == OP_JVPG || OP == OP_LINMV);
  static constexpr bool kHasF = (OP == OP_JVP || OP == OP_LINPREP);
  static constexpr int kFieldDoubles = kBoxMainDoubles + 2 * kBoxHaloDoubles; // main boxes, left halo, right halo
  static constexpr int kX = 0;
  static constexpr int kV = kFieldDoubles;
  static constexpr int kD = kV + (HAS_V ? kFieldDoubles : 0);
  static constexpr int kF = kD + (kHasD ? kBoxMainDoubles : 0);
  static constexpr int kStageDoubles = kF + (kHasF ? kBoxMainDoubles : 0);
  static constexpr int kFit = 102400 / (kStageDoubles * 8);
  static constexpr int kStages = kFit < 3 ? 3 : (kFit > 8 ? 8 : kFit);
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageDoubles * 8 + 2 * kStages * sizeof(uint64_t);
};

__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int OP, bool HAS_V>
__global__ void __launch_bounds__(kBoxThreads, 2)
    sh_box_kernel(const __grid_constant__ ShBoxMaps maps, ShArgs A, ShBoxRows hr, SHParams P, double* S, ReduceWs ws) {
  if (sh_is_operator<OP>() && S[JS_STOP] != 0.0) return; // (before the fused halo push: every rank skips alike)
  using LY = BoxLayout<OP, HAS_V>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)LY::kStages * LY::kStageDoubles * 8);
  uint64_t* empty = full + LY::kStages;

  const int nx = A.nx, nrows = A.nrows;
  const int npairs = nrows / kBoxR;
  const int strips = (nx + kBoxStrip - 1) / kBoxStrip;
  const long long total = (long long)strips * npairs; // work units: (strip, row pair), strip-major
  long long begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  if (A.mode & 2) {
    // band-major mapping: gridDim.x = strips * bands; the CTAs of a band sweep the same rows of all strips side by side,
    // so the rows being read and written at any moment are few and whole (DRAM page locality of a linear sweep)
    const int bands = gridDim.x / strips;
    const int strip = blockIdx.x % strips, band = blockIdx.x / strips;
    begin = (long long)strip * npairs + (long long)npairs * band / bands;
    end = (long long)strip * npairs + (long long)npairs * (band + 1) / bands;
  }
  const bool cs = (A.mode & 1) != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LY::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kBoxConsumers / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (A.push_field) {
    // fused halo exchange over peer memory, as in sh_tma_kernel: the first CTAs push this rank's boundary rows, nothing waits
    const unsigned npush = min(gridDim.x, (unsigned)kShPushCtas);
    if (blockIdx.x < npush) {
      const size_t n2 = A.push.count >> 1;
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
  if (warp == kBoxConsumers / 32) {
    // ------------------------------- producer: one lane streams row pairs into the ring -----------------
    if (lane == 0) {
      tma_prefetch_map(&maps.x_main);
      tma_prefetch_map(&maps.xc_main);
      if (HAS_V) { tma_prefetch_map(&maps.v_main); tma_prefetch_map(&maps.vc_main); }
      if (LY::kHasD) tma_prefetch_map(&maps.d);
      if (LY::kHasF) tma_prefetch_map(&maps.f0);
      int s = 0;
      uint32_t ph = 0;
      long long idx = begin;
      bool top_ok = A.push_field == 0, bot_ok = A.push_field == 0;
      while (idx < end) {
        const int strip = (int)(idx / npairs), p0 = (int)(idx - (long long)strip * npairs);
        const long long run_end = min(end, (long long)(strip + 1) * npairs);
        const int p1 = p0 + (int)(run_end - idx);
        const int xs = strip * kBoxStrip, W = min(kBoxStrip, nx - xs);
        const bool two = W > kBoxW;
        const int cl = xs - 2 < 0 ? xs - 2 + nx : xs - 2, cr = xs + W >= nx ? xs + W - nx : xs + W;
        if (!top_ok && p0 == 0) {
          wait_flag(A.push.my_flag_top, A.push.epoch, A.push.err);
          asm volatile("fence.proxy.async;" ::: "memory");
          top_ok = true;
        }
        if (!bot_ok && p1 == npairs) {
          wait_flag(A.push.my_flag_bot, A.push.epoch, A.push.err);
          asm volatile("fence.proxy.async;" ::: "memory");
          bot_ok = true;
        }
        const uint32_t field_bytes = (two ? 2u : 1u) * (uint32_t)(kBoxW * kBoxR * 8) + 2u * (uint32_t)(2 * kBoxR * 8);
        const uint32_t point_bytes = (two ? 2u : 1u) * (uint32_t)(kBoxW * kBoxR * 8);
        for (int q = p0 - 1; q <= p1; ++q) {
          const bool pt = q - 1 >= p0; // this stage completes the window of output pair q-1
          mbar_wait(&empty[s], ph ^ 1u);
          double* st = stage0 + (size_t)s * LY::kStageDoubles;
          uint32_t bytes = field_bytes * (HAS_V ? 2u : 1u);
          if (pt) bytes += point_bytes * ((LY::kHasD ? 1u : 0u) + (LY::kHasF ? 1u : 0u));
          mbar_expect_tx(&full[s], bytes);
          {
            const bool edge = q < 0 || q >= npairs;
            const CUtensorMap* mm = edge ? &maps.x_halo : &maps.x_main;
            const CUtensorMap* mc = edge ? &maps.xc_halo : &maps.xc_main;
            const int row = q < 0 ? hr.x_top : (q >= npairs ? hr.x_bot : kBoxR * q);
            tma_box(st + LY::kX, mm, xs, row, &full[s]);
            if (two) tma_box(st + LY::kX + kBoxW * kBoxR, mm, xs + kBoxW, row, &full[s]);
            tma_box(st + LY::kX + kBoxMainDoubles, mc, cl, row, &full[s]);
            tma_box(st + LY::kX + kBoxMainDoubles + kBoxHaloDoubles, mc, cr, row, &full[s]);
          }
          if (HAS_V) {
            const bool edge = q < 0 || q >= npairs;
            const CUtensorMap* mm = edge ? &maps.v_halo : &maps.v_main;
            const CUtensorMap* mc = edge ? &maps.vc_halo : &maps.vc_main;
            const int row = q < 0 ? hr.v_top : (q >= npairs ? hr.v_bot : kBoxR * q);
            tma_box(st + LY::kV, mm, xs, row, &full[s]);
            if (two) tma_box(st + LY::kV + kBoxW * kBoxR, mm, xs + kBoxW, row, &full[s]);
            tma_box(st + LY::kV + kBoxMainDoubles, mc, cl, row, &full[s]);
            tma_box(st + LY::kV + kBoxMainDoubles + kBoxHaloDoubles, mc, cr, row, &full[s]);
          }
          if (pt) {
            const int row = kBoxR * (q - 1);
            if (LY::kHasD) {
              tma_box(st + LY::kD, &maps.d, xs, row, &full[s]);
              if (two) tma_box(st + LY::kD + kBoxW * kBoxR, &maps.d, xs + kBoxW, row, &full[s]);
            }
            if (LY::kHasF) {
              tma_box(st + LY::kF, &maps.f0, xs, row, &full[s]);
              if (two) tma_box(st + LY::kF + kBoxW * kBoxR, &maps.f0, xs + kBoxW, row, &full[s]);
            }
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
    const int t = threadIdx.x;
    const int c0 = 2 * t;                                       // column offset inside the strip
    const int own0 = (t >> 7) * (kBoxW * kBoxR) + 2 * (t & 127); // offset of (row 0, my columns) inside a field's main boxes
    const double2 zero2 = make_double2(0.0, 0.0);
    int s = 0;
    uint32_t ph = 0;
    long long idx = begin;
    while (idx < end) {
      const int strip = (int)(idx / npairs), p0 = (int)(idx - (long long)strip * npairs);
      const long long run_end = min(end, (long long)(strip + 1) * npairs);
      const int p1 = p0 + (int)(run_end - idx);
      const int xs = strip * kBoxStrip, W = min(kBoxStrip, nx - xs);
      const bool act = c0 < W;
      // where my left / right neighbour pairs live (offset at row 0, row stride): inside my box, in the other box of the strip,
      // or in the halo-column boxes
      int offL, strL, offR, strR;
      if ((t & 127) != 0) { offL = own0 - 2; strL = kBoxW; }
      else if (t == 128) { offL = kBoxW - 2; strL = kBoxW; }
      else { offL = kBoxMainDoubles; strL = 2; }
      if (c0 + 2 >= W) { offR = kBoxMainDoubles + kBoxHaloDoubles; strR = 2; }
      else if (t == 127) { offR = kBoxW * kBoxR; strR = kBoxW; }
      else { offR = own0 + 2; strR = kBoxW; }
      double2 u0 = zero2, u1 = zero2, u2 = zero2, u3 = zero2, u4 = zero2;
      double2 p0s = zero2, p1s = zero2, p2s = zero2, p3s = zero2; // horizontal +-1 pair sums, rows ra-3..ra
      double2 q0s = zero2, q1s = zero2, q2s = zero2;             // horizontal +-2 pair sums, rows ra-2..ra
      for (int q = p0 - 1; q <= p1; ++q) {
        const bool pt = q - 1 >= p0;
        mbar_wait(&full[s], ph);
        const double* st = stage0 + (size_t)s * LY::kStageDoubles;
#pragma unroll
        for (int r = 0; r < kBoxR; ++r) {
          double2 L = zero2, own = zero2, R = zero2, dv = zero2, fv = zero2;
          if (act) {
            const double* sx = st + LY::kX;
            L = *reinterpret_cast<const double2*>(sx + offL + r * strL);
            own = *reinterpret_cast<const double2*>(sx + own0 + r * kBoxW);
            R = *reinterpret_cast<const double2*>(sx + offR + r * strR);
            if (HAS_V) {
              const double* sv = st + LY::kV;
              double2 vl = *reinterpret_cast<const double2*>(sv + offL + r * strL);
              double2 vo = *reinterpret_cast<const double2*>(sv + own0 + r * kBoxW);
              double2 vr = *reinterpret_cast<const double2*>(sv + offR + r * strR);
              L.x = combine(L.x, a, vl.x); L.y = combine(L.y, a, vl.y);
              own.x = combine(own.x, a, vo.x); own.y = combine(own.y, a, vo.y);
              R.x = combine(R.x, a, vr.x); R.y = combine(R.y, a, vr.y);
            }
            if (pt) {
              if (LY::kHasD) dv = *reinterpret_cast<const double2*>(st + LY::kD + own0 + r * kBoxW);
              if (LY::kHasF) fv = *reinterpret_cast<const double2*>(st + LY::kF + own0 + r * kBoxW);
            }
          }
          u0 = u1; u1 = u2; u2 = u3; u3 = u4; u4 = own;
          p0s = p1s; p1s = p2s; p2s = p3s;
          p3s.x = L.y + own.y; p3s.y = own.x + R.x;
          q0s = q1s; q1s = q2s;
          q2s.x = L.x + R.x; q2s.y = L.y + R.y;
          if (act && pt) {
            // rows: u0..u4 = y-2..y+2 ; p0s,p1s,p2s = pair sums of rows y-1,y,y+1 ; q0s = +-2 pair sum of row y
            const int y = kBoxR * (q - 1) + r;
            double s1x = p1s.x + u1.x + u3.x, s1y = p1s.y + u1.y + u3.y;
            double sdx = p0s.x + p2s.x, sdy = p0s.y + p2s.y;
            double s2x = q0s.x + u0.x + u4.x, s2y = q0s.y + u0.y + u4.y;
            const size_t e = (size_t)y * nx + xs + c0;
            double2 o, o2 = zero2, o3 = zero2;
            o.x = sh_value<OP>(P, scale, u2.x, s1x, sdx, s2x, dv.x, fv.x, o2.x, o3.x, acc);
            o.y = sh_value<OP>(P, scale, u2.y, s1y, sdy, s2y, dv.y, fv.y, o2.y, o3.y, acc);
            if (cs) stg2_cs(A.out + e, o); else stg2(A.out + e, o);
            if ((OP == OP_RESID && A.out2) || OP == OP_LINPREP) stg2(A.out2 + e, o2);
            if (OP == OP_RESID && A.out3) stg2(A.out3 + e, o3);
          }
        }
        // this warp is done with the stage: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == LY::kStages) { s = 0; ph ^= 1u; }
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
