// The device-operation interface the host-side solver logic (engine.cpp) is written against.
//
// The shipped library has exactly ONE implementation: CudaOps (cuda_ops.cu, sm_100a kernels).
// The interface exists so that the scalar control flow of the solver (Newton, Armijo, Eisenstat-
// Walker, LGMRES bookkeeping) can be unit-tested in the GPU-less build container against SciPy
// with a CPU test double that lives under tests/hostsim/ and is never part of the product.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "hd_math.h"

namespace jfnk {

// A scalar that is only known on the device at enqueue time:
//   value = mult * S[mul_idx] * sqrt(S[mul_sqrt_idx]) / sqrt(S[div_sqrt_idx])    (index < 0: factor 1)
struct ScalarRef {
  double mult;
  int mul_idx;
  int mul_sqrt_idx;
  int div_sqrt_idx;
};
inline ScalarRef sref(double mult, int mul_idx = -1, int mul_sqrt_idx = -1, int div_sqrt_idx = -1) {
  ScalarRef r; r.mult = mult; r.mul_idx = mul_idx; r.mul_sqrt_idx = mul_sqrt_idx; r.div_sqrt_idx = div_sqrt_idx; return r;
}
JF_HD double eval_sref(const double* S, const ScalarRef& r) {
  double v = r.mult;
  if (r.mul_idx >= 0) v *= S[r.mul_idx];
  if (r.mul_sqrt_idx >= 0) v *= sqrt(S[r.mul_sqrt_idx]);
  if (r.div_sqrt_idx >= 0) v /= sqrt(S[r.div_sqrt_idx]);
  return v;
}

struct Grid {
  int nx, ny;      // global grid: ny rows x nx columns
  int row0, nrows; // local slab of rows
  int rank, nranks;
  size_t n() const { return (size_t)nx * (size_t)nrows; }
  size_t n_global() const { return (size_t)nx * (size_t)ny; }
};

// Mesh geometry / parameters of the moving-mesh problems (PMA2_nk.py:23-40, droplet.py:23-53).
struct MeshParams {
  double dksi, deta;
  double bl, br, bb, bt; // Dirichlet values of Q_ksi (left,right) and Q_eta (bottom,top)
};
// droplets of an initial state: centre (x, y), radius R, volume V (droplet.py:126 `info`)
constexpr int kMaxDrops = 8;
struct DropList {
  int n;
  double x[kMaxDrops], y[kMaxDrops], R[kMaxDrops], V[kMaxDrops];
};
struct Pma2Params { double lambd, beta, epsilon; int m; double dt; };
struct DropletParams { double epsilon; int n_exp, m_exp; double Bo, alpha2, epsilon2; double dt; };

// per-kernel-class timing (CUDA events on the launch stream) for bench.py's roofline line
struct KernelStat {
  char name[32];
  int64_t launches;
  double ms;    // summed device time of the launches
  double bytes; // summed ALGORITHMIC bytes of the launches (per-kernel table)
};

struct PmaParams; // mesh_math.h

// One whole LGMRES inner cycle in one launch (Engine::cycle -> DeviceOps::cycle_fused; CUDA: sh_cycle.cuh)
struct FusedCycleIn {
  int kind;              // 0: Swift-Hohenberg (fields below) ; 2: droplet, 3: PMA2 (mesh fields at the end)
  int linear;            // 0: FD Jacobian of the Swift-Hohenberg residual about x0 ; 1: the linearly-implicit operator
  const double* x0;      // linearisation point (null when linear)
  const double* g0;      // G(x0) ; linear: the diagonal D
  const double* v0;      // unnormalised start vector
  double v0n2, omega, ptol, tau2;
  int gs_mode, m, k;     // m = inner_m + k Arnoldi steps at most, k augmentation vectors in use
  int m_max;             // inner_m + outer_k
  const double* ov[JF_MAXOV]; // the augmentation vectors in use, oldest first ...
  int ov_zn2[JF_MAXOV];       // ... and the arena index of their squared norms
  double* out;           // dx
  int out_zn2;           // arena index that receives ||dx||^2
  // optional (d != null): also evaluate the full-step trial point of the Newton iteration that follows,
  // t = x0 - dx -> trial_x, F(t) -> trial_F, G(t) -> trial_G, {sum F^2, max|F|, max|t|} -> S[trial_norm_off..+2]
  const double* d;
  double *trial_x, *trial_F, *trial_G;
  int trial_norm_off;
  // droplet (kind 2): geometry, parameters, the 7 metric fields, previous time level, its Crank-Nicolson term, F(x0)
  const MeshParams* mp;
  const DropletParams* dp;
  const Pma2Params* pp;
  const double* const* M;
  const double *uval, *fprev, *f0;
};
struct FusedCycleOut {
  int nit, reorth, flags;
  double res, dxn2;
  int has_trial;         // the trial point was evaluated; its norms:
  double trial_nrm[3];
};

class DeviceOps {
 protected:
  double posted_[JF_MAXV + 2][8]; // default post_read / wait_read storage (synchronous backends)
 public:
  virtual ~DeviceOps() {}
  virtual int64_t launches() const = 0;
  virtual void profile_enable(bool on) = 0;
  // synchronises, aggregates and clears the records; returns the number of classes written (<= cap)
  virtual int profile_read(KernelStat* out, int cap) = 0;
  // 0 when no asynchronous device error is pending, else JFNK_CUDA_ERROR / JFNK_NCCL_ERROR (message via last_error()).
  virtual int status() = 0;
  virtual const char* last_error() const = 0;

  // ---- launch-sequence capture (CUDA graphs) for launch-bound inner loops --------------------------------------
  // graph_begin(): start recording the operations enqueued from now on instead of running them; returns false when
  // the backend cannot (then the caller simply keeps enqueuing eagerly).  graph_end_launch(replays): stop recording
  // and run the recorded sequence `replays` times.  Nothing that synchronises or allocates may be called in between.
  virtual bool graph_begin() { return false; }
  virtual void graph_end_launch(int /*replays*/) {}

  // ---- scalar arena ---------------------------------------------------------------------------
  virtual void read_scalars(int off, int cnt, double* host) = 0; // synchronises the stream
  virtual void write_scalars(int off, int cnt, const double* host) = 0;
  virtual void allreduce_sum(int off, int cnt) = 0; // across slab ranks; no-op on one rank
  virtual void allreduce_max(int off, int cnt) = 0;
  // all-reduce (sum) followed by the Hessenberg/Givens step of Arnoldi column j on the reduced values; the CUDA
  // backend does both in one launch when the ranks share peer memory
  virtual void allreduce_sum_givens(int off, int cnt, int j, int taken, int rerun) {
    allreduce_sum(off, cnt);
    givens(j, taken, rerun);
  }

  // ---- the Arnoldi loop without host round trips ------------------------------------------------------
  // true: the operator / multi-dot / update kernels of this backend return at once while S[JS_STOP] is set, so the
  // engine may enqueue Arnoldi step j+1 before it has seen the outcome of step j (Engine::cycle)
  virtual bool can_speculate() const { return false; }
  // ... and so do the kernels behind pma2_eval / droplet_eval in operator mode (f0 != null)
  virtual bool can_speculate_mesh() const { return false; }
  // enqueue a copy of S[off..off+cnt) into host slot `slot` (0 <= slot < JF_MAXV + 1, cnt <= 8) ...
  virtual void post_read(int slot, int off, int cnt) { read_scalars(off, cnt, posted_[slot]); }
  // ... and wait for exactly that copy (later work may still be running on the device)
  virtual void wait_read(int slot, int cnt, double* host) { for (int i = 0; i < cnt; ++i) host[i] = posted_[slot][i]; }
  // out[i] = V_i . w, out[nv] = w . w, summed over the slab ranks
  virtual void mdot_reduced(int nv, const double* const* V, const double* w, int out_off) {
    mdot(nv, V, w, out_off);
    allreduce_sum(out_off, nv + 1);
  }
  // Gram-Schmidt update (as gs_update), ||w||^2 summed over the slab ranks into S[n2_off], then the Hessenberg / Givens
  // step of column j (hd_math.h); `skippable`: part of a speculatively enqueued step
  virtual void gs_update_givens(int nv, const double* const* V, double* w, int rd_off, int n2_off, int j, int taken,
                                int rerun) {
    gs_update(nv, V, w, rd_off, n2_off, -1);
    allreduce_sum_givens(n2_off, 1, j, taken, rerun);
  }

  // The whole cycle -- Arnoldi process with classical Gram-Schmidt, Givens QR, least squares, dx = sum y_i z_i -- as one
  // launch, when the backend has such a kernel for this grid (small single-rank Swift-Hohenberg grids); false = not done.
  virtual bool cycle_fused(const FusedCycleIn& /*in*/, FusedCycleOut& /*out*/) { return false; }

  // S[norm_off..+2] = {sum F^2, max|F|, max|t|} of a residual pass, combined over the slab ranks (sum, max, max).  A backend
  // whose sh_residual already did that inside its kernels returns true from residual_norms_global().
  virtual bool residual_norms_global() { return false; }
  virtual void reduce_residual_norms(int norm_off) {
    if (residual_norms_global()) return;
    allreduce_sum(norm_off, 1);
    allreduce_max(norm_off + 1, 2);
  }
  // out = sum_i S[JS_COEF+i] Z_i ; S[n2_off] = ||out||^2 summed over the slab ranks
  virtual void maxpy_reduced(int nz, const double* const* Z, double* out, int n2_off) {
    maxpy(nz, Z, out, n2_off);
    allreduce_sum(n2_off, 1);
  }

  // ---- BLAS-1 on slab-local vectors of length grid.n() -------------------------------------------
  // out[i] = V_i . w (i < nv), out[nv] = w . w
  virtual void mdot(int nv, const double* const* V, const double* w, int out_off) = 0;
  // w -= sum_i (S[rd_off+i] / S[JS_VN2+i]) V_i ; S[n2_off] = ||w||^2.
  // fuse_givens_j >= 0 (single rank only): the kernel's last CTA also performs the Hessenberg/Givens step of
  // column j (first pass, not a re-run) right after finalising the norm -- one launch less per Arnoldi step.
  virtual void gs_update(int nv, const double* const* V, double* w, int rd_off, int n2_off, int fuse_givens_j) = 0;
  // w -= sum_i S[JS_COEF+i] V_i ; S[n2_off] = ||w||^2   (public microbenchmark form)
  virtual void maxpy_sub(int nv, const double* const* V, double* w, int n2_off) = 0;
  // out = sum_i S[JS_COEF+i] Z_i ; S[n2_off] = ||out||^2
  virtual void maxpy(int nz, const double* const* Z, double* out, int n2_off) = 0;
  // out = a x + b y (y may be null; out may alias x or y) ; S[n2_off] = ||out||^2 when n2_off >= 0
  virtual void lincomb(double* out, ScalarRef a, const double* x, ScalarRef b, const double* y, int n2_off) = 0;
  // out = (x - y)/div
  virtual void diff_scale(double* out, const double* x, const double* y, ScalarRef div) = 0;
  virtual void copy(double* dst, const double* src) = 0;
  virtual void maxabs(const double* v, int out_off) = 0;
  // Hessenberg column + Givens update of Arnoldi step j; LSQ back-substitution (hd_math.h).
  virtual void givens(int j, int taken, int rerun) = 0;
  virtual void lsq(int nit, const int* zn2_idx, int scale_n2_idx) = 0;

  // ---- Swift-Hohenberg stencil operators --------------------------------------------------------
  virtual void sh_setup(const SHParams& p) = 0;
  virtual void sh_spmv(int which, const double* x, double* y) = 0; // 0: Lap (5-pt), 1: L (13-pt)
  virtual void sh_set_prev(const double* uo, double* d) = 0;       // d = Uo/k + N(Uo)/2
  // t = x + a v (v may be null) ; F = G(t) - d ; xt_out (may be null) = t ;
  // S[norm_off+0] = sum F^2, S[norm_off+1] = max|F|, S[norm_off+2] = max|t|  (local partials; caller all-reduces)
  // g_out (may be null) = G(t) = F + d, the part of the residual that depends on t: kept for the accepted iterate so that
  // the Jacobian-vector products read ONE pointwise operand, G(x0), instead of d and f0
  virtual void sh_residual(const double* x, const double* v, ScalarRef a, const double* d, double* xt_out, double* F,
                           double* g_out, int norm_off) = 0;
  // cache the row halos of the linearisation point (multi-rank); no-op on one rank
  virtual void sh_bind_x0(const double* x0) = 0;
  // w = (G(x0 + sc z) - d - f0)/div ; with g0 = G(x0) given: w = (G(x0 + sc z) - g0)/div (d, f0 unused)
  virtual void sh_jvp(const double* x0, const double* z, ScalarRef sc, ScalarRef div, const double* d, const double* f0,
                      const double* g0, double* w) = 0;
  // linearly-implicit SH: D = (5U-Uo)^2 k/16 - g k U ; b = U + k/2 L U
  virtual void shlin_prepare(const double* U, const double* Uo, double* D, double* b) = 0;
  // w = a (z + D z - k/2 L z)
  virtual void shlin_matvec(const double* z, ScalarRef a, const double* D, double* w) = 0;

  // ---- moving-mesh operators (PMA2 / droplet) ----------------------------------------------------
  // Q -> metric fields M[0..6] = Q_ksiksi, Q_etaeta, Q_ksieta, J, A11, A22, A12 (each grid.n() long)
  virtual void mesh_metrics(const MeshParams& mp, const double* Q, double* const* M) = 0;
  // (v_xx, v_yy) = Laplace_operator(v); `sum_only`: vxx receives v_xx + v_yy and vyy is ignored
  // deriv_bc = 1: the first derivatives of v fed to the cross terms are boundary-zeroed the way
  // droplet.py:719-722 does it (v_ksi = 0 on left/right/bottom edges, v_eta = 0 on the top edge).
  virtual void mesh_laplace(const MeshParams& mp, const double* const* M, const double* v, double* vxx, double* vyy,
                            int sum_only, int deriv_bc) = 0;
  // PMA2 rhs(u) given lap2 = Laplace(Laplace(u)):  out = -lambd/(1+u)^2 + lambd eps^(m-2)/(1+u)^m - beta^2 lap2, boundary -> 0
  virtual void pma2_rhs(const Pma2Params& pp, const double* u, const double* lap2, double* out) = 0;
  // F = (u - uval)/dt - (rhs + cn)/2 ; norms as sh_residual (max|t| is max|u|)
  virtual void pma2_combine(const Pma2Params& pp, const double* u, const double* uval, const double* rhs,
                            const double* cn, double* F, int norm_off) = 0;
  // Whole PMA2 residual / FD-JVP (PMA2_nk.py:121-159) for t = x + a v (v may be null):
  //   F = (t - uval)/dt - (rhs(t, Laplace(Laplace(t))) + cn)/2
  //   f0 == null: out = F, xt_out (may be null) = t, norms at S[norm_off..+2] as sh_residual
  //   f0 != null: out = (F - f0)/div                                   (KrylovJacobian.matvec)
  // scratch: 4 work vectors.  The CUDA backend runs this as two fused marching passes on large grids
  // (mesh_march.cuh); the default composes the primitives above.
  virtual void pma2_eval(const MeshParams& mp, const Pma2Params& pp, const double* const* M, const double* x,
                         const double* v, ScalarRef a, const double* uval, const double* cn, const double* f0,
                         ScalarRef div, double* const* scratch, double* xt_out, double* out, int norm_off) {
    pma2_eval_unfused(mp, pp, M, x, v, a, uval, cn, f0, div, scratch, xt_out, out, norm_off);
  }
  void pma2_eval_unfused(const MeshParams& mp, const Pma2Params& pp, const double* const* M, const double* x,
                         const double* v, ScalarRef a, const double* uval, const double* cn, const double* f0,
                         ScalarRef div, double* const* scratch, double* xt_out, double* out, int norm_off) {
    const double* t = x;
    if (v) {
      double* tt = (xt_out && !f0) ? xt_out : scratch[3];
      lincomb(tt, sref(1.0), x, a, v, -1);
      t = tt;
    }
    mesh_laplace(mp, M, t, scratch[0], nullptr, 1, 0);
    mesh_laplace(mp, M, scratch[0], scratch[1], nullptr, 1, 0);
    pma2_rhs(pp, t, scratch[1], scratch[2]);
    if (f0) {
      pma2_combine(pp, t, uval, scratch[2], cn, scratch[0], norm_off);
      diff_scale(out, scratch[0], f0, div);
    } else {
      pma2_combine(pp, t, uval, scratch[2], cn, out, norm_off);
    }
  }
  // Whole droplet residual / FD-JVP (droplet.py:435-450) for t = x + a v (v may be null):
  //   F = (t - uval) - dt (div(flux(p(t, Laplace(t)), t)) + fprev)/2
  //   f0 == null: out = F, xt_out (may be null) = t when v is given, norms at S[norm_off..+2] as sh_residual
  //   f0 != null: out = (F - f0)/div                                   (KrylovJacobian.matvec)
  // scratch: 7 work vectors.  The CUDA backend fuses the chain into three launches; the default composes the primitives.
  virtual void droplet_eval(const MeshParams& mp, const DropletParams& dp, const double* const* M, const double* x,
                            const double* v, ScalarRef a, const double* uval, const double* fprev, const double* f0,
                            ScalarRef div, double* const* scratch, double* xt_out, double* out, int norm_off) {
    const double* t = x;
    if (v) {
      double* tt = (xt_out && !f0) ? xt_out : scratch[5];
      lincomb(tt, sref(1.0), x, a, v, -1);
      t = tt;
    }
    mesh_laplace(mp, M, t, scratch[0], nullptr, 1, 0);
    droplet_pressure(dp, t, scratch[0], scratch[1]);
    droplet_flux(mp, dp, M, scratch[1], t, scratch[2], scratch[3]);
    droplet_div(mp, M, scratch[2], scratch[3], scratch[4]);
    if (f0) {
      droplet_combine(dp, t, uval, scratch[4], fprev, scratch[6], norm_off);
      diff_scale(out, scratch[6], f0, div);
    } else {
      droplet_combine(dp, t, uval, scratch[4], fprev, out, norm_off);
    }
  }
  // droplet: p = -(lap) + PI(h) + Bo cos(alpha2) h
  virtual void droplet_pressure(const DropletParams& dp, const double* h, const double* lap, double* p) = 0;
  // (A, B) = flux components from pressure p and height h  (droplet.py:439-447)
  virtual void droplet_flux(const MeshParams& mp, const DropletParams& dp, const double* const* M, const double* p,
                            const double* h, double* A, double* B) = 0;
  // out = (Q_etaeta D_ksi A - Q_ksieta D_eta A - Q_ksieta D_ksi B + Q_ksiksi D_eta B)/J   (droplet.py:448-449)
  virtual void droplet_div(const MeshParams& mp, const double* const* M, const double* A, const double* B, double* out) = 0;
  // F = (u - uval) - dt (F2 + Fprev)/2 ; norms as sh_residual
  virtual void droplet_combine(const DropletParams& dp, const double* u, const double* uval, const double* F2,
                               const double* Fprev, double* F, int norm_off) = 0;

  // droplet initial shape on the current mesh (compute_U2 / compute_U, droplet.py:413-429,544-551):
  // out = eps + (1 - eps) sum_i H2(G2(|(Q_ksi, Q_eta) - (x_i, y_i)|, R_i), R_i, V_i)
  virtual void droplet_shape(const MeshParams& mp, const double* Q, const DropList& drops, double a, double eps,
                             double* out) = 0;

  // ---- moving-mesh relaxation (loop_pma / solve_PMA; PMA2_nk.py:345-403, droplet.py:578-599,729-760) ------
  // The whole relaxation loop (`loops` passes of: metrics of Q, monitor of Uval, smoothing, regularisation, DCT solve,
  // Q += dt Q_t) in one launch, when the backend has such a kernel for this grid; false = not done, use the stages below.
  // M: the 7 metric fields (left holding the values of the last pass's Q); a, b, t, spec: scratch fields.
  virtual bool mesh_relax_fused(const MeshParams& /*mp*/, const PmaParams& /*pp*/, double* /*Q*/, const double* /*Uval*/,
                                double /*dt*/, int /*loops*/, int /*deriv_bc*/, double* const* /*M*/, double* /*a*/,
                                double* /*b*/, double* /*t*/, double* /*spec*/) {
    return false;
  }
  // out = monitor(u, lap): mode 0 |lap|^2, mode 1 1/(1+u)^6
  virtual void pma_monitor(int mode, const double* u, const double* lap, double* out) = 0;
  // one pass of the 9-point (edges 6-point, corners 4-point) smoothing filter
  virtual void pma_smooth(const MeshParams& mp, const double* in, double* out) = 0;
  // S[out_off] = sum mon |J|
  virtual void pma_wsum(const double* mon, const double* J, int out_off) = 0;
  // out = sqrt((mon + add) |J|) / alpha
  virtual void pma_rhs(const double* mon, const double* J, ScalarRef add, double alpha, double* out) = 0;
  // 2-D orthonormal DCT-II of the ny x nx field (inverse != 0: the inverse transform); tmp is scratch
  virtual void pma_dct2(const double* in, double* tmp, double* out, int inverse) = 0;
  // Y /= (1 - gamma Leig)
  virtual void pma_spectral_divide(const MeshParams& mp, double gamma, double* Y) = 0;
};

} // namespace jfnk
