// Layout of the device-resident scalar arena (fp64) shared by all kernels of one context.
//
// Every reduction result (dots, squared norms, max-norms), the Hessenberg/Givens state of the
// Arnoldi process and the LSQ solution live here, so that one Arnoldi step needs no host
// round-trip except the single convergence scalar (JS_RES).  Norm slots hold SQUARED 2-norms
// (raw sums) so that they can be all-reduced across ranks in place; consumers take the sqrt.
#pragma once

namespace jfnk {

constexpr int JF_MAXV = 48;  // max basis vectors (inner_m + outer_k + 1 <= 48)
constexpr int JF_MAXOV = 16; // max augmentation-vector ring slots (outer_k + 1 <= 16)

enum ScalarSlot : int {
  // residual norms, slot set A (current iterate) and B (line-search trial)
  JS_F2_A = 0, JS_FMAX_A, JS_XMAX_A, JS_pad0,
  JS_F2_B, JS_FMAX_B, JS_XMAX_B, JS_pad1,
  JS_TMP0, JS_TMP1, JS_TMP2, JS_TMP3, // misc reductions (||v||^2 for the public jvp, ||dx||^2, ...)
  JS_WW,   // w.w before orthogonalisation (alias of RD[nv], copied by givens)
  JS_HN2A, // ||w||^2 after the first GS pass
  JS_HN2B, // ||w||^2 after the second GS pass (or copy of HN2A when skipped)
  JS_RES,  // |g_{j+1}|: GMRES residual estimate relative to ||v0|| = 1   -> host
  JS_FLAGS, // bit0 breakdown, bit1 non-finite, bit2 second GS pass taken -> host
  JS_DXN2, // ||dx||^2 of the assembled LGMRES correction
  JS_GJ_SAVE, // rotated rhs entry g[j] before the Givens step of column j (for a re-run after a second GS pass)
  JS_pad3,
  JS_RD = 20,                 // [JF_MAXV+1] raw dots V_i.w ; RD[nv] = w.w
  JS_RD2 = JS_RD + JF_MAXV + 1, // [JF_MAXV+1] second-pass raw dots
  JS_VN2 = JS_RD2 + JF_MAXV + 1, // [JF_MAXV+1] squared norms of the (unnormalised) Arnoldi vectors
  JS_ZN2 = JS_VN2 + JF_MAXV + 1, // [JF_MAXOV]  squared norms of the augmentation vectors (ring slots)
  JS_COEF = JS_ZN2 + JF_MAXOV,   // [JF_MAXV]   coefficients for multi-axpy (GS update or dx assembly)
  JS_CS = JS_COEF + JF_MAXV,     // [JF_MAXV]   Givens cosines
  JS_SN = JS_CS + JF_MAXV,       // [JF_MAXV]   Givens sines
  JS_G = JS_SN + JF_MAXV,        // [JF_MAXV+1] rotated right-hand side
  JS_Y = JS_G + JF_MAXV + 1,     // [JF_MAXV]   LSQ solution
  JS_R = JS_Y + JF_MAXV,         // [JF_MAXV*JF_MAXV] upper-triangular factor, column-major R[i + j*JF_MAXV]
  // Device-side control of the Arnoldi loop.  The host enqueues Arnoldi step j+1 BEFORE it has seen the outcome of step j
  // (it reads step j's record one step late, so the stream never drains); the Givens step of column j decides on the
  // device whether the process goes on, and every kernel of the Arnoldi loop returns at once while JS_STOP is set.
  JS_STOP = JS_R + JF_MAXV * JF_MAXV, // != 0: converged / breakdown / non-finite / a second Gram-Schmidt pass is wanted
  JS_PTOL,                       // inner tolerance of the running cycle (host-written; < 0: never stop on the residual)
  JS_TAU2,                       // tau^2 of the "cgs-ifneeded" test ||w_after||^2 < tau^2 w.w (0: never ask for a 2nd pass)
  JS_pad4,
  JS_REC,                        // [JF_MAXV][JF_REC_STRIDE] per-step records {w.w, HN2A, HN2B, RES, FLAGS}
  // outcome of a whole Arnoldi cycle run by ONE launch (sh_cycle.cuh): {inner iterations, second Gram-Schmidt passes,
  // residual estimate, flags of the last column, ||dx||^2, 0, 0, 0} -> the host's single read-back per cycle
  JS_CYC = JS_REC + JF_MAXV * 8,
  JS_COUNT = JS_CYC + 8
};
constexpr int JF_REC_STRIDE = 8;

constexpr int JF_FLAG_BREAKDOWN = 1;
constexpr int JF_FLAG_NONFINITE = 2;
constexpr int JF_FLAG_REORTH = 4;      // a second Gram-Schmidt pass was taken
constexpr int JF_FLAG_NEED_REORTH = 8; // the first pass cancelled more than 1/tau: the host must enqueue the second pass

} // namespace jfnk
