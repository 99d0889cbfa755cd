// Host-side solver logic of the JFNK engine.  See engine.h for the SciPy functions each part follows.
#include "engine.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <limits>

namespace jfnk {

static const double kEps = 2.220446049250313e-16;
static const int kNumScratch = 8;

static bool is_mesh_problem(int p) { return p == JFNK_PROBLEM_PMA2 || p == JFNK_PROBLEM_DROPLET; }

const char* Engine::validate(const jfnk_config& c) {
  if (c.abi_version != JFNK_ABI_VERSION) return "jfnk_config.abi_version mismatch";
  if (c.problem < JFNK_PROBLEM_SH || c.problem > JFNK_PROBLEM_DROPLET) return "unknown problem";
  if (c.nx < 1 || c.ny < 1) return "nx, ny must be positive";
  if (c.nranks < 1 || c.rank < 0 || c.rank >= c.nranks) return "bad rank / nranks";
  if (c.row0 < 0 || c.nrows < 1 || c.row0 + c.nrows > c.ny) return "slab [row0, row0+nrows) outside the grid";
  if (c.nranks == 1 && (c.row0 != 0 || c.nrows != c.ny)) return "single rank must own the whole grid";
  if (c.inner_m < 1 || c.outer_k < 0) return "inner_m must be >= 1 and outer_k >= 0";
  if (c.inner_m + c.outer_k + 1 > JF_MAXV) return "inner_m + outer_k + 1 exceeds JF_MAXV (48)";
  if (c.outer_k + 1 > JF_MAXOV) return "outer_k + 1 exceeds JF_MAXOV (16)";
  if (c.gs_mode < JFNK_GS_CGS || c.gs_mode > JFNK_GS_CGS2) return "unknown gs_mode";
  if ((c.problem == JFNK_PROBLEM_SH || c.problem == JFNK_PROBLEM_SH_LINEAR)) {
    if (c.nx < 5 || c.ny < 5) return "Swift-Hohenberg grid must be at least 5x5 (13-point periodic stencil)";
    if (c.nranks > 1 && c.nrows < 2) return "Swift-Hohenberg slab needs at least 2 rows per rank";
  }
  if (is_mesh_problem(c.problem)) {
    if (c.nx < 7 || c.ny < 7) return "moving-mesh grid must be at least 7x7 (4th-order closures)";
    if (c.nranks > 1) return "PMA2 / droplet problems are single-GPU in this build";
  }
  return nullptr;
}

size_t Engine::vector_stride(const jfnk_config& c) {
  size_t n = (size_t)c.nx * (size_t)c.nrows;
  return (n + 31) / 32 * 32; // 256-byte aligned vectors
}

int Engine::vector_count(const jfnk_config& c) {
  int cnt = 4;                        // XT, FX, FT, D
  if (c.problem == JFNK_PROBLEM_SH) cnt += 2; // GX, GT: G = F + d of the current iterate / of the line-search trial
  cnt += c.inner_m + c.outer_k;       // Arnoldi slots 1..m
  cnt += c.outer_k + 1;               // augmentation ring
  if (is_mesh_problem(c.problem)) cnt += kNumScratch + 7 + 2; // scratch, metric fields, UVAL, CN
  return cnt;
}

size_t Engine::workspace_doubles(const jfnk_config& c) { return vector_stride(c) * (size_t)vector_count(c); }

Engine::Engine(const jfnk_config& cfg, DeviceOps* ops, double* ws, size_t /*ws_doubles*/)
    : cfg_(cfg), ops_(ops), ws_(ws) {
  grid_.nx = cfg.nx; grid_.ny = cfg.ny; grid_.row0 = cfg.row0; grid_.nrows = cfg.nrows;
  grid_.rank = cfg.rank; grid_.nranks = cfg.nranks;
  vstride_ = vector_stride(cfg);
  size_t i = 0;
  auto next = [&]() { return ws_ + vstride_ * (i++); };
  XT_ = next(); FX_ = next(); FT_ = next(); D_ = next();
  GX_ = GT_ = nullptr;
  if (cfg.problem == JFNK_PROBLEM_SH) { GX_ = next(); GT_ = next(); }
  for (int j = 0; j < cfg.inner_m + cfg.outer_k; ++j) VS_.push_back(next());
  for (int j = 0; j < cfg.outer_k + 1; ++j) OV_.push_back(next());
  if (is_mesh_problem(cfg.problem)) {
    scratch_.resize(kNumScratch);
    for (int j = 0; j < kNumScratch; ++j) scratch_[j] = next();
    for (int j = 0; j < 7; ++j) MF_[j] = next();
    UVAL_ = next(); CN_ = next();
  } else {
    for (int j = 0; j < 7; ++j) MF_[j] = nullptr;
    UVAL_ = CN_ = nullptr;
  }
  linear_op_ = (cfg.problem == JFNK_PROBLEM_SH_LINEAR);
}

Engine::~Engine() {}

// ------------------------------------------------------------------------------------------------
// problem setup
// ------------------------------------------------------------------------------------------------
int Engine::sh_setup(double h, double r, double g, double k) {
  if (cfg_.problem != JFNK_PROBLEM_SH && cfg_.problem != JFNK_PROBLEM_SH_LINEAR)
    return fail(JFNK_INVALID, "jfnk_sh_setup: context was not created for a Swift-Hohenberg problem");
  if (!(h > 0) || !(k > 0)) return fail(JFNK_INVALID, "jfnk_sh_setup: h and k must be positive");
  shp_ = make_sh_params(h, r, g, k);
  ops_->sh_setup(shp_);
  sh_ready_ = true;
  prev_ready_ = false;
  return ops_->status();
}

int Engine::spmv(int which, const double* x, double* y) {
  if (!sh_ready_) return fail(JFNK_INVALID, "jfnk_spmv_*: call jfnk_sh_setup first");
  ops_->sh_spmv(which, x, y);
  return ops_->status();
}

int Engine::set_prev(const double* uo) {
  if (cfg_.problem != JFNK_PROBLEM_SH) return fail(JFNK_INVALID, "jfnk_set_prev: Swift-Hohenberg JFNK problem only");
  if (!sh_ready_) return fail(JFNK_INVALID, "jfnk_set_prev: call jfnk_sh_setup first");
  ops_->sh_set_prev(uo, D_);
  prev_ready_ = true;
  return ops_->status();
}

int Engine::shlin_prepare(const double* U, const double* Uo, double* b) {
  if (cfg_.problem != JFNK_PROBLEM_SH_LINEAR) return fail(JFNK_INVALID, "jfnk_shlin_prepare: SH_LINEAR problem only");
  if (!sh_ready_) return fail(JFNK_INVALID, "jfnk_shlin_prepare: call jfnk_sh_setup first");
  ops_->shlin_prepare(U, Uo, D_, b);
  prev_ready_ = true;
  return ops_->status();
}

int Engine::shlin_step(double* U, double* Uo, int nsteps, double rtol, int maxiter, int* info, int64_t* matvecs) {
  if (cfg_.problem != JFNK_PROBLEM_SH_LINEAR) return fail(JFNK_INVALID, "jfnk_shlin_step: SH_LINEAR problem only");
  if (!sh_ready_) return fail(JFNK_INVALID, "jfnk_shlin_step: call jfnk_sh_setup first");
  int worst = 0;
  int64_t mv = 0;
  for (int s = 0; s < nsteps; ++s) {
    // sh_linearised.py:51-57: D from (U[s], U[s-1]); Uo <- U[s]; U[s+1] = solve(I + D - L k/2, (I + L k/2) Uo)
    ops_->shlin_prepare(U, Uo, D_, FX_);
    prev_ready_ = true;
    ops_->copy(Uo, U);
    ov_slots_.clear();
    int inf = 0, inner = 0;
    double res = 0;
    int rc = lgmres_general(FX_, U, rtol, maxiter, &inf, &res, &inner);
    if (rc) return rc;
    worst = std::max(worst, inf);
    mv += inner;
    if (inf != 0) {
      // sh_linearised.py:57 solves with a direct solver; an unconverged LGMRES must not become the next step's state silently
      if (info) *info = worst;
      if (matvecs) *matvecs = mv;
      return fail(JFNK_NO_CONVERGENCE, "jfnk_shlin_step: LGMRES did not reach rtol within maxiter outer cycles (step " +
                                           std::to_string(s) + ")");
    }
  }
  if (info) *info = worst;
  if (matvecs) *matvecs = mv;
  return ops_->status();
}

int Engine::mesh_setup(const MeshParams& mp) {
  if (!is_mesh_problem(cfg_.problem)) return fail(JFNK_INVALID, "jfnk_mesh_setup: PMA2 / droplet problems only");
  if (!(mp.dksi > 0) || !(mp.deta > 0)) return fail(JFNK_INVALID, "jfnk_mesh_setup: spacings must be positive");
  mp_ = mp;
  mesh_ready_ = true;
  metrics_ready_ = false;
  prev_ready_ = false;
  return JFNK_OK;
}

int Engine::mesh_set_potential(const double* Q) {
  if (!mesh_ready_) return fail(JFNK_INVALID, "jfnk_mesh_set_potential: call jfnk_mesh_setup first");
  ops_->mesh_metrics(mp_, Q, MF_);
  metrics_ready_ = true;
  prev_ready_ = false;
  return ops_->status();
}

int Engine::mesh_laplace(const double* v, double* vxx, double* vyy) {
  if (!metrics_ready_) return fail(JFNK_INVALID, "jfnk_mesh_laplace: call jfnk_mesh_set_potential first");
  ops_->mesh_laplace(mp_, MF_, v, vxx, vyy, 0, 0);
  return ops_->status();
}

int Engine::pma2_setup(const Pma2Params& pp) {
  if (cfg_.problem != JFNK_PROBLEM_PMA2) return fail(JFNK_INVALID, "jfnk_pma2_setup: PMA2 problem only");
  if (!(pp.dt > 0)) return fail(JFNK_INVALID, "jfnk_pma2_setup: dt must be positive");
  pp_ = pp;
  pp_ready_ = true;
  prev_ready_ = false;
  return JFNK_OK;
}

int Engine::pma2_set_prev(const double* uval) {
  if (!pp_ready_ || !metrics_ready_)
    return fail(JFNK_INVALID, "jfnk_pma2_set_prev: call jfnk_pma2_setup and jfnk_mesh_set_potential first");
  // PMA2_nk.py:83 U.val ; :97 CN_term = compute_rhs_pde() (:405-419)
  ops_->copy(UVAL_, uval);
  ops_->mesh_laplace(mp_, MF_, UVAL_, scratch_[0], nullptr, 1, 0);
  ops_->mesh_laplace(mp_, MF_, scratch_[0], scratch_[1], nullptr, 1, 0);
  ops_->pma2_rhs(pp_, UVAL_, scratch_[1], CN_);
  prev_ready_ = true;
  return ops_->status();
}

int Engine::droplet_setup(const DropletParams& dp) {
  if (cfg_.problem != JFNK_PROBLEM_DROPLET) return fail(JFNK_INVALID, "jfnk_droplet_setup: droplet problem only");
  dp_ = dp;
  dp_ready_ = true;
  prev_ready_ = false;
  return JFNK_OK;
}

int Engine::droplet_set_prev(const double* uval, double dt) {
  if (!dp_ready_ || !metrics_ready_)
    return fail(JFNK_INVALID, "jfnk_droplet_set_prev: call jfnk_droplet_setup and jfnk_mesh_set_potential first");
  if (!(dt > 0)) return fail(JFNK_INVALID, "jfnk_droplet_set_prev: dt must be positive");
  dp_.dt = dt;
  // droplet.py:373-381: U.val ; compute_u_spatial_ders (boundary-zeroed first derivatives, :716-727) ;
  // P.val = pressure(U.val, U.xx, U.yy) ; compute_P_spatial_ders ; F = pde_rhs(...)
  ops_->copy(UVAL_, uval);
  ops_->mesh_laplace(mp_, MF_, UVAL_, scratch_[0], nullptr, 1, 1);
  ops_->droplet_pressure(dp_, UVAL_, scratch_[0], scratch_[1]);
  ops_->droplet_flux(mp_, dp_, MF_, scratch_[1], UVAL_, scratch_[2], scratch_[3]);
  ops_->droplet_div(mp_, MF_, scratch_[2], scratch_[3], CN_);
  prev_ready_ = true;
  return ops_->status();
}

// loop_pma (droplet.py:590-599) / the single solve_PMA + explicit Euler update of PMA2_nk.py:94,103, on the device.
// Every pass: metrics of the current Q, Laplacian of the (old) solution, monitor + smoothing + Mackenzie
// regularisation, spectral solve, Q += dt Q_t.  (The script skips the derivative refresh in its first pass because
// they are still current from the start of the step; recomputing them gives the same values.)
int Engine::mesh_relax(double* Q, const double* Uval, double dt, int loops, const PmaParams& pp) {
  if (!mesh_ready_) return fail(JFNK_INVALID, "jfnk_mesh_relax: call jfnk_mesh_setup first");
  if (loops < 1) return fail(JFNK_INVALID, "jfnk_mesh_relax: loops must be >= 1");
  if (!(pp.alpha > 0) || pp.smoothing_iters < 0 || pp.monitor_mode < 0 || pp.monitor_mode > 1)
    return fail(JFNK_INVALID, "jfnk_mesh_relax: bad parameters");
  const int deriv_bc = (cfg_.problem == JFNK_PROBLEM_DROPLET) ? 1 : 0;
  const double cell = mp_.dksi * mp_.deta;
  double *lap = scratch_[0], *a = scratch_[1], *b = scratch_[2], *t = scratch_[3], *spec = scratch_[4];
  // small grids: the whole loop as one persistent kernel (stages separated by grid barriers, no launches at all)
  if (ops_->mesh_relax_fused(mp_, pp, Q, Uval, dt, loops, deriv_bc, MF_, a, b, t, spec)) {
    metrics_ready_ = false;
    prev_ready_ = false;
    return ops_->status();
  }
  auto one_pass = [&]() {
    ops_->mesh_metrics(mp_, Q, MF_);
    if (pp.monitor_mode == 0) ops_->mesh_laplace(mp_, MF_, Uval, lap, nullptr, 1, deriv_bc);
    ops_->pma_monitor(pp.monitor_mode, Uval, lap, a);
    for (int s = 0; s < pp.smoothing_iters; ++s) { ops_->pma_smooth(mp_, a, b); std::swap(a, b); }
    ops_->pma_wsum(a, MF_[3], JS_TMP0);
    ops_->pma_rhs(a, MF_[3], sref(pp.cnorm * cell, JS_TMP0), pp.alpha, b); // sqrt((mon + C sum|J| dksi deta)|J|)/alpha
    ops_->pma_dct2(b, t, spec, 0);
    ops_->pma_spectral_divide(mp_, pp.gamma, spec);
    ops_->pma_dct2(spec, t, b, 1);
    ops_->lincomb(Q, sref(1.0), Q, sref(dt), b, -1);
  };
  // The pass is a fixed sequence of ~14 small launches with no host decision in it, repeated `loops` (400 in
  // droplet.py:384) times: launch-bound on the reference's grids.  The first pass runs eagerly (it may allocate the DCT
  // matrices / upload the weight tables); the rest is recorded once and replayed as a CUDA graph.  With an odd number
  // of smoothing sweeps the ping-pong buffers swap roles every pass, so the recorded unit is two passes.
  int it = 0;
  one_pass();
  it = 1;
  const int unit = (pp.smoothing_iters % 2 == 0) ? 1 : 2;
  if (loops - it >= 4 * unit && ops_->graph_begin()) {
    for (int u = 0; u < unit; ++u) one_pass();
    const int reps = (loops - it) / unit;
    ops_->graph_end_launch(reps);
    it += reps * unit;
  }
  for (; it < loops; ++it) one_pass();
  metrics_ready_ = false; // Q moved on: the next step must call jfnk_mesh_set_potential (compute_Q_spatial_ders)
  prev_ready_ = false;
  return ops_->status();
}

// compute_U2 (droplet.py:413-423): the analytic droplet shapes on the current mesh; info = ndrops x (x, y, R, V).
int Engine::droplet_shape(const double* Q, int ndrops, const double* info, double a, double* U) {
  if (cfg_.problem != JFNK_PROBLEM_DROPLET) return fail(JFNK_INVALID, "jfnk_droplet_shape: droplet problem only");
  if (!mesh_ready_ || !dp_ready_) return fail(JFNK_INVALID, "jfnk_droplet_shape: call jfnk_mesh_setup and jfnk_droplet_setup first");
  if (ndrops < 0 || ndrops > kMaxDrops || (ndrops > 0 && !info)) return fail(JFNK_INVALID, "jfnk_droplet_shape: 0..8 droplets");
  if (!(a > 0)) return fail(JFNK_INVALID, "jfnk_droplet_shape: the interface steepness a must be positive");
  DropList dl;
  memset(&dl, 0, sizeof(dl));
  dl.n = ndrops;
  for (int i = 0; i < ndrops; ++i) {
    dl.x[i] = info[4 * i]; dl.y[i] = info[4 * i + 1]; dl.R[i] = info[4 * i + 2]; dl.V[i] = info[4 * i + 3];
    if (!(dl.R[i] > 0)) return fail(JFNK_INVALID, "jfnk_droplet_shape: radii must be positive");
  }
  ops_->droplet_shape(mp_, Q, dl, a, dp_.epsilon, U);
  return ops_->status();
}

bool Engine::problem_ready(std::string& why) const {
  switch (cfg_.problem) {
    case JFNK_PROBLEM_SH:
      if (!sh_ready_) { why = "call jfnk_sh_setup first"; return false; }
      if (!prev_ready_) { why = "call jfnk_set_prev first"; return false; }
      return true;
    case JFNK_PROBLEM_SH_LINEAR:
      if (!sh_ready_) { why = "call jfnk_sh_setup first"; return false; }
      if (!prev_ready_) { why = "call jfnk_shlin_prepare first"; return false; }
      return true;
    case JFNK_PROBLEM_PMA2:
      if (!prev_ready_) { why = "call jfnk_mesh_setup, jfnk_mesh_set_potential, jfnk_pma2_setup, jfnk_pma2_set_prev first"; return false; }
      return true;
    case JFNK_PROBLEM_DROPLET:
      if (!prev_ready_) { why = "call jfnk_mesh_setup, jfnk_mesh_set_potential, jfnk_droplet_setup, jfnk_droplet_set_prev first"; return false; }
      return true;
  }
  why = "unknown problem";
  return false;
}

// ------------------------------------------------------------------------------------------------
// residual / operator
// ------------------------------------------------------------------------------------------------
void Engine::generic_residual(const double* u, double* F, int norm_off) {
  if (cfg_.problem == JFNK_PROBLEM_PMA2) {
    // PMA2_nk.py:121-159
    ops_->pma2_eval(mp_, pp_, MF_, u, nullptr, sref(1.0), UVAL_, CN_, nullptr, sref(1.0), scratch_.data(), nullptr, F,
                    norm_off);
  } else {
    // droplet.py:435-450
    ops_->droplet_eval(mp_, dp_, MF_, u, nullptr, sref(1.0), UVAL_, CN_, nullptr, sref(1.0), scratch_.data(), nullptr, F,
                       norm_off);
  }
}

int Engine::eval_residual(const double* x, const double* v, ScalarRef a, double* xt_out, double* F, double* g_out,
                          int norm_off, double nrm[3]) {
  if (cfg_.problem == JFNK_PROBLEM_SH) {
    ops_->sh_residual(x, v, a, D_, xt_out, F, g_out, norm_off);
  } else if (cfg_.problem == JFNK_PROBLEM_PMA2) {
    // t = x + a v is formed inside the evaluation (never stored unless the caller wants the trial iterate)
    if (!v && xt_out && xt_out != x) ops_->copy(xt_out, x);
    ops_->pma2_eval(mp_, pp_, MF_, x, v, a, UVAL_, CN_, nullptr, sref(1.0), scratch_.data(), v ? xt_out : nullptr, F,
                    norm_off);
  } else {
    if (!v && xt_out && xt_out != x) ops_->copy(xt_out, x);
    ops_->droplet_eval(mp_, dp_, MF_, x, v, a, UVAL_, CN_, nullptr, sref(1.0), scratch_.data(), v ? xt_out : nullptr, F,
                       norm_off);
  }
  ops_->reduce_residual_norms(norm_off);
  ops_->read_scalars(norm_off, 3, nrm);
  nfev_++;
  return ops_->status();
}

int Engine::residual(const double* u, double* F) {
  std::string why;
  if (cfg_.problem == JFNK_PROBLEM_SH_LINEAR) return fail(JFNK_INVALID, "jfnk_residual: not defined for SH_LINEAR");
  if (!problem_ready(why)) return fail(JFNK_INVALID, "jfnk_residual: " + why);
  double nrm[3];
  return eval_residual(u, nullptr, sref(1.0), nullptr, F, nullptr, JS_F2_A, nrm);
}

int Engine::linearize(const double* x0, double rdiff) {
  std::string why;
  if (linear_op_) return fail(JFNK_INVALID, "jfnk_linearize: SH_LINEAR has a fixed linear operator");
  if (!problem_ready(why)) return fail(JFNK_INVALID, "jfnk_linearize: " + why);
  if (!(rdiff > 0)) rdiff = sqrt(kEps);
  ops_->copy(XT_, x0);
  double nrm[3];
  int rc = eval_residual(XT_, nullptr, sref(1.0), nullptr, FX_, GX_, JS_F2_A, nrm);
  if (rc) return rc;
  x0_ = XT_; f0_ = FX_; g0_ = GX_;
  omega_ = rdiff * std::max(1.0, nrm[2]) / std::max(1.0, nrm[1]); // _nonlin.py:1552-1555
  if (cfg_.problem == JFNK_PROBLEM_SH) ops_->sh_bind_x0(x0_);
  return ops_->status();
}

void Engine::apply_operator(const double* z, int zn2_idx, double* w, bool unit_input) {
  if (linear_op_) {
    ops_->shlin_matvec(z, unit_input ? sref(1.0, -1, -1, zn2_idx) : sref(1.0), D_, w);
    return;
  }
  // KrylovJacobian.matvec (_nonlin.py:1557-1565): sc = omega/||z|| ; w = (F(x0 + sc z) - f0)/sc.
  // For an Arnoldi vector stored unnormalised (z = ||z|| zhat) the SciPy call is matvec(zhat): the same
  // perturbation x0 + (omega/||z||) z, divided by omega.
  ScalarRef sc = sref(omega_, -1, -1, zn2_idx);
  ScalarRef div = unit_input ? sref(omega_) : sc;
  if (cfg_.problem == JFNK_PROBLEM_SH) {
    ops_->sh_jvp(x0_, z, sc, div, D_, f0_, g0_, w);
  } else if (cfg_.problem == JFNK_PROBLEM_PMA2) {
    ops_->pma2_eval(mp_, pp_, MF_, x0_, z, sc, UVAL_, CN_, f0_, div, scratch_.data(), nullptr, w, JS_TMP0 /*unused*/);
  } else {
    ops_->droplet_eval(mp_, dp_, MF_, x0_, z, sc, UVAL_, CN_, f0_, div, scratch_.data(), nullptr, w, JS_TMP0 /*unused*/);
  }
  nfev_++;
}

int Engine::jvp(const double* v, double* Jv) {
  if (linear_op_) {
    if (!prev_ready_) return fail(JFNK_INVALID, "jfnk_jvp: call jfnk_shlin_prepare first");
    ops_->shlin_matvec(v, sref(1.0), D_, Jv);
    return ops_->status();
  }
  if (!x0_) return fail(JFNK_INVALID, "jfnk_jvp: call jfnk_linearize first");
  // ||v||^2 -> TMP3 ; the operator divides by ||v||, scale back by ||v|| afterwards (0*v when ||v|| == 0).
  // (Not TMP0..TMP2: the unfused mesh residuals park their unused norms there before the quotient is formed.)
  ops_->mdot(0, nullptr, v, JS_TMP3);
  ops_->allreduce_sum(JS_TMP3, 1);
  double vn2;
  ops_->read_scalars(JS_TMP3, 1, &vn2);
  if (vn2 == 0.0) {
    ops_->lincomb(Jv, sref(0.0), v, sref(0.0), nullptr, -1);
    return ops_->status();
  }
  apply_operator(v, JS_TMP3, Jv, false);
  return ops_->status();
}

// ------------------------------------------------------------------------------------------------
// LGMRES
// ------------------------------------------------------------------------------------------------
int Engine::lgmres_reset() { ov_slots_.clear(); return JFNK_OK; }

// One _fgmres Arnoldi process + solution assembly (_gcrotmk.py:16-183, lgmres.py:179-227).
// v0vec is the UNNORMALISED start vector (v0 = v0vec/sqrt(v0n2)); Arnoldi vectors are stored unnormalised
// with their squared norms in the scalar arena ("lazy normalisation": no scaling pass over the vectors).
int Engine::cycle(const double* v0vec, double v0n2, double ptol, CycleOut& out) {
  const int k = (int)ov_slots_.size();
  const int m = cfg_.inner_m + k;
  const double* vs[JF_MAXV + 1];
  const double* zs[JF_MAXV];
  int znidx[JF_MAXV];
  vs[0] = v0vec;
  ops_->write_scalars(JS_VN2 + 0, 1, &v0n2);
  const double tau2 = cfg_.gs_tau * cfg_.gs_tau;
  const bool single = (grid_.nranks == 1);
  int nit = 0;
  double res = NAN;
  // choice of z (_gcrotmk.py:96-105 with prepend_outer_v=True); the new Arnoldi vector goes to VS_[j]
  auto choose = [&](int j) {
    if (j < k) { zs[j] = OV_[ov_slots_[j]]; znidx[j] = JS_ZN2 + ov_slots_[j]; }
    else if (j == k) { zs[j] = vs[0]; znidx[j] = JS_VN2 + 0; }
    else { zs[j] = vs[j]; znidx[j] = JS_VN2 + j; }
    vs[j + 1] = VS_[j];
  };
  auto clear_stop = [&]() { const double zero = 0.0; ops_->write_scalars(JS_STOP, 1, &zero); };
  // the ring slot the assembled correction goes to
  int slot = -1;
  for (int s = 0; s < (int)OV_.size(); ++s) {
    if (std::find(ov_slots_.begin(), ov_slots_.end(), s) == ov_slots_.end()) { slot = s; break; }
  }
  double dxn2 = 0.0;
  // Reference-sized Swift-Hohenberg grids: the whole cycle (Arnoldi process, least squares, dx assembly) is ONE launch of a
  // single thread-block cluster that keeps the basis in shared memory (sh_cycle.cuh); the host reads one record.
  bool fused = false;
  const bool droplet = (cfg_.problem == JFNK_PROBLEM_DROPLET), pma2 = (cfg_.problem == JFNK_PROBLEM_PMA2);
  if (!psolve_ && (((cfg_.problem == JFNK_PROBLEM_SH || linear_op_) && (linear_op_ || g0_)) || ((droplet || pma2) && f0_))) {
    FusedCycleIn in;
    memset(&in, 0, sizeof(in));
    in.kind = droplet ? 2 : (pma2 ? 3 : 0);
    in.linear = linear_op_ ? 1 : 0;
    in.x0 = linear_op_ ? nullptr : x0_;
    in.g0 = linear_op_ ? D_ : g0_;
    if (droplet || pma2) { in.mp = &mp_; in.dp = &dp_; in.pp = &pp_; in.M = MF_; in.uval = UVAL_; in.fprev = CN_; in.f0 = f0_; }
    in.v0 = v0vec; in.v0n2 = v0n2; in.omega = omega_; in.ptol = ptol; in.tau2 = tau2;
    in.gs_mode = cfg_.gs_mode; in.m = m; in.k = k; in.m_max = cfg_.inner_m + cfg_.outer_k;
    for (int j = 0; j < k; ++j) { in.ov[j] = OV_[ov_slots_[j]]; in.ov_zn2[j] = JS_ZN2 + ov_slots_[j]; }
    in.out = OV_[slot]; in.out_zn2 = JS_ZN2 + slot;
    trial_.valid = false;
    if (trial_.want && !linear_op_) {
      in.d = D_; in.trial_x = trial_.xt; in.trial_F = trial_.Ft; in.trial_G = trial_.Gt; in.trial_norm_off = JS_F2_B;
    }
    FusedCycleOut fo;
    memset(&fo, 0, sizeof(fo));
    if (ops_->cycle_fused(in, fo)) {
      int st = ops_->status();
      if (st) return st;
      fused = true;
      nit = fo.nit; res = fo.res; dxn2 = fo.dxn2;
      nfev_ += nit; inner_total_ += nit; reorth_total_ += fo.reorth;
      if (fo.flags & JF_FLAG_NONFINITE) return fail(JFNK_NONFINITE, "Function returned non-finite results");
      if (fo.has_trial) { trial_.valid = true; for (int i = 0; i < 3; ++i) trial_.nrm[i] = fo.trial_nrm[i]; }
    }
  }
  trial_.want = false;
  if (!fused) {
  // Speculative mode: the Givens step of column j decides ON THE DEVICE whether the process goes on (converged to ptol,
  // breakdown, non-finite, second Gram-Schmidt pass wanted -> JS_STOP), the host enqueues step j+1 before it waits for
  // step j's record, and the kernels of a step enqueued after the stop return at once.  The stream never drains inside
  // the loop and nothing that was not needed runs.  Needs kernels that honour JS_STOP (DeviceOps::can_speculate; the
  // multi-kernel mesh residuals and preconditioner callbacks do not).
  const bool spec = ops_->can_speculate() && !psolve_ && (!is_mesh_problem(cfg_.problem) || ops_->can_speculate_mesh());
  {
    const double ctl[3] = {0.0, spec ? ptol : -1.0, (spec && cfg_.gs_mode == JFNK_GS_CGS_IFNEEDED) ? tau2 : 0.0};
    ops_->write_scalars(JS_STOP, 3, ctl); // JS_STOP, JS_PTOL, JS_TAU2
  }
  if (spec) {
    auto enqueue_step = [&](int j) {
      choose(j);
      double* w = VS_[j];
      apply_operator(zs[j], znidx[j], w, true);
      // classical Gram-Schmidt against vs[0..j]: all dots in one fused reduction (all-reduced by the kernel's last CTA
      // with slab ranks), then one fused update whose last CTA (all-reduces the norm and) runs the Givens step
      ops_->mdot_reduced(j + 1, vs, w, JS_RD);
      if (cfg_.gs_mode == JFNK_GS_CGS2) {
        ops_->gs_update(j + 1, vs, w, JS_RD, JS_HN2A, -1);
        ops_->allreduce_sum(JS_HN2A, 1);
        ops_->mdot_reduced(j + 1, vs, w, JS_RD2);
        ops_->gs_update_givens(j + 1, vs, w, JS_RD2, JS_HN2B, j, 1, 0);
      } else {
        ops_->gs_update_givens(j + 1, vs, w, JS_RD, JS_HN2A, j, 0, 0);
      }
      ops_->post_read(j, JS_REC + j * JF_REC_STRIDE, 5);
    };
    enqueue_step(0);
    for (int j = 0; j < m; ++j) {
      bool ahead = (j + 1 < m);
      if (ahead) enqueue_step(j + 1); // dropped on the device if step j stops the process
      double sc[5]; // w.w, ||w||^2 after pass 1, after pass 2, residual estimate, flags
      ops_->wait_read(j, 5, sc);
      int st = ops_->status();
      if (st) { clear_stop(); return st; }
      int flags = (int)sc[4];
      bool second = (cfg_.gs_mode == JFNK_GS_CGS2);
      if (flags & JF_FLAG_NEED_REORTH) {
        // the first pass cancelled more than 1/tau: step j+1 was dropped; take the second pass of step j, redo its column
        if (ahead) { nfev_--; ahead = false; }
        clear_stop();
        ops_->mdot_reduced(j + 1, vs, VS_[j], JS_RD2);
        ops_->gs_update_givens(j + 1, vs, VS_[j], JS_RD2, JS_HN2B, j, 1, 1);
        ops_->post_read(j, JS_REC + j * JF_REC_STRIDE, 5);
        ops_->wait_read(j, 5, sc);
        st = ops_->status();
        if (st) { clear_stop(); return st; }
        flags = (int)sc[4];
        second = true;
        const bool stops = (flags & (JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) || sc[3] < ptol;
        if (!stops && j + 1 < m) { enqueue_step(j + 1); ahead = true; }
      }
      res = sc[3];
      nit = j + 1;
      inner_total_++;
      if (second) reorth_total_++;
      const bool stops = (flags & (JF_FLAG_BREAKDOWN | JF_FLAG_NONFINITE)) || res < ptol;
      if (stops && ahead) nfev_--; // the step that was enqueued ahead never ran
      if (flags & JF_FLAG_NONFINITE) { clear_stop(); return fail(JFNK_NONFINITE, "Function returned non-finite results"); }
      if (stops) break;
    }
    clear_stop();
  } else {
  for (int j = 0; j < m; ++j) {
    choose(j);
    const double* z = zs[j];
    const int zi = znidx[j];
    double* w = VS_[j];
    if (psolve_ && op_tmp_) {
      // left preconditioning (_gcrotmk.py:113-121): w = M (A z)
      apply_operator(z, zi, op_tmp_, true);
      int st0 = ops_->status();
      if (st0) return st0;
      if (psolve_(psolve_user_, op_tmp_, w) != 0) return fail(JFNK_INVALID, "the preconditioner callback reported an error");
    } else {
      apply_operator(z, zi, w, true);
    }
    // classical Gram-Schmidt against vs[0..j]: all dots in one fused reduction, then one fused update.
    // One host round-trip per Arnoldi step: {w.w, ||w||^2 after pass 1, residual estimate, flags}.  The host
    // decides from it whether a second (re-orthogonalisation) pass is needed; only then more work is enqueued.
    ops_->mdot_reduced(j + 1, vs, w, JS_RD);
    bool second = (cfg_.gs_mode == JFNK_GS_CGS2);
    double sc[5]; // JS_WW, JS_HN2A, JS_HN2B, JS_RES, JS_FLAGS
    if (!second) {
      ops_->gs_update_givens(j + 1, vs, w, JS_RD, JS_HN2A, j, 0, 0);
      ops_->read_scalars(JS_WW, 5, sc);
      second = (cfg_.gs_mode == JFNK_GS_CGS_IFNEEDED) && (sc[1] < tau2 * sc[0]);
      if (second) {
        ops_->mdot_reduced(j + 1, vs, w, JS_RD2);
        ops_->gs_update_givens(j + 1, vs, w, JS_RD2, JS_HN2B, j, 1, 1); // redo column j with h = (RD + RD2)/||V||, ||w|| from pass 2
        ops_->read_scalars(JS_WW, 5, sc);
      }
    } else {
      ops_->gs_update(j + 1, vs, w, JS_RD, JS_HN2A, -1);
      ops_->allreduce_sum(JS_HN2A, 1);
      ops_->mdot_reduced(j + 1, vs, w, JS_RD2);
      ops_->gs_update_givens(j + 1, vs, w, JS_RD2, JS_HN2B, j, 1, 0);
      ops_->read_scalars(JS_WW, 5, sc);
    }
    int st = ops_->status();
    if (st) { clear_stop(); return st; }
    res = sc[3];
    int flags = (int)sc[4];
    nit = j + 1;
    inner_total_++;
    if (second) reorth_total_++;
    if (flags & JF_FLAG_NONFINITE) { clear_stop(); return fail(JFNK_NONFINITE, "Function returned non-finite results"); }
    if (res < ptol || (flags & JF_FLAG_BREAKDOWN)) break;
  }
  clear_stop(); // (a breakdown sets it even in this mode)
  }
  (void)single;
  // y = lstsq(R, Q[0,:]) * inner_res_0 ; dx = sum zs_i y_i (lgmres.py:188,206-208)
  ops_->lsq(nit, znidx, JS_VN2 + 0);
  ops_->maxpy_reduced(nit, zs, OV_[slot], JS_ZN2 + slot);
  ops_->read_scalars(JS_ZN2 + slot, 1, &dxn2);
  } // !fused
  out.sol = OV_[slot];
  out.sol_slot = slot;
  out.sol_n2 = dxn2;
  out.res = res;
  out.inner = nit;
  if (!isfinite(dxn2)) return fail(JFNK_NONFINITE, "LGMRES produced a non-finite correction");
  // store the augmentation vector dx/||dx|| (lazily normalised) and trim to outer_k (lgmres.py:211-224)
  if (dxn2 > 0) ov_slots_.push_back(slot);
  while ((int)ov_slots_.size() > cfg_.outer_k) ov_slots_.pop_front();
  return ops_->status();
}

// scipy.sparse.linalg.lgmres with x0 = 0 (lgmres.py:124-232), general number of outer cycles.
int Engine::lgmres_general(const double* b, double* x, double rtol, int maxiter, int* info, double* res_out,
                           int* inner_out) {
  { const double zero = 0.0; ops_->write_scalars(JS_STOP, 1, &zero); } // (stale after an aborted solve)
  ops_->mdot(0, nullptr, b, JS_TMP0);
  ops_->allreduce_sum(JS_TMP0, 1);
  double bn2;
  ops_->read_scalars(JS_TMP0, 1, &bn2);
  if (!isfinite(bn2)) return fail(JFNK_INVALID, "RHS must contain only finite numbers");
  double bnorm = sqrt(bn2);
  double atol = std::max(0.0, rtol * bnorm);
  int inner_sum = 0;
  double res = 0.0;
  if (bnorm == 0.0) {
    ops_->copy(x, b);
    if (info) *info = 0;
    if (res_out) *res_out = 0.0;
    if (inner_out) *inner_out = 0;
    return ops_->status();
  }
  double ptol_max_factor = 1.0;
  bool x_is_zero = true;
  int result = maxiter;
  for (int k_outer = 0; k_outer < maxiter; ++k_outer) {
    const double* v0vec;
    double v0n2, r_norm;
    if (x_is_zero) { v0vec = b; v0n2 = bn2; r_norm = bnorm; }
    else {
      // r_outer = A x - b ; v0 = -r_outer.  ||x||^2 lives in TMP3: the unfused mesh residuals behind apply_operator park
      // their (unused) norms in TMP0..TMP2, exactly as Engine::jvp has to allow for.
      ops_->mdot(0, nullptr, x, JS_TMP3);
      ops_->allreduce_sum(JS_TMP3, 1);
      apply_operator(x, JS_TMP3, FT_, true);
      ops_->lincomb(FT_, sref(1.0), b, sref(-1.0, -1, JS_TMP3, -1), FT_, JS_TMP2);
      ops_->allreduce_sum(JS_TMP2, 1);
      ops_->read_scalars(JS_TMP2, 1, &v0n2);
      r_norm = sqrt(v0n2);
      v0vec = FT_;
    }
    if (r_norm <= std::max(atol, rtol * bnorm)) { result = 0; break; }
    double ptol = std::min(ptol_max_factor, std::max(atol, rtol * bnorm) / r_norm);
    CycleOut co;
    int rc = cycle(v0vec, v0n2, ptol, co);
    if (rc) return rc;
    inner_sum += co.inner;
    res = co.res;
    if (co.res > ptol) ptol_max_factor = std::min(1.0, 1.5 * ptol_max_factor);
    else ptol_max_factor = std::max(1e-16, 0.25 * ptol_max_factor);
    if (x_is_zero) ops_->copy(x, co.sol);
    else ops_->lincomb(x, sref(1.0), x, sref(1.0), co.sol, -1);
    x_is_zero = false;
  }
  if (info) *info = result;
  if (res_out) *res_out = res;
  if (inner_out) *inner_out = inner_sum;
  return ops_->status();
}

int Engine::lgmres(const double* b, double* x, double rtol, int maxiter, int* info, double* res, int* inner) {
  if (linear_op_) {
    if (!prev_ready_) return fail(JFNK_INVALID, "jfnk_lgmres: call jfnk_shlin_prepare first");
  } else if (!x0_) {
    return fail(JFNK_INVALID, "jfnk_lgmres: call jfnk_linearize first");
  }
  if (maxiter < 1) return fail(JFNK_INVALID, "jfnk_lgmres: maxiter must be >= 1");
  if (psolve_) return fail(JFNK_INVALID, "jfnk_lgmres: the preconditioner hook applies to jfnk_newton only; remove it first");
  return lgmres_general(b, x, rtol, maxiter, info, res, inner);
}

// ------------------------------------------------------------------------------------------------
// Newton-Krylov (nonlin_solve, _nonlin.py:134-277)
// ------------------------------------------------------------------------------------------------
int Engine::newton(double* u, const jfnk_newton_opts* o, jfnk_history* hist) {
  std::string why;
  if (linear_op_) return fail(JFNK_INVALID, "jfnk_newton: not defined for SH_LINEAR (use jfnk_shlin_step)");
  if (!problem_ready(why)) return fail(JFNK_INVALID, "jfnk_newton: " + why);
  jfnk_newton_opts dflt;
  memset(&dflt, 0, sizeof(dflt));
  dflt.line_search = 1;
  if (!o) o = &dflt;
  const double inf = std::numeric_limits<double>::infinity();
  const double f_tol = o->f_tol > 0 ? o->f_tol : pow(kEps, 1.0 / 3.0);
  const double f_rtol = o->f_rtol > 0 ? o->f_rtol : inf;
  const double x_tol = o->x_tol > 0 ? o->x_tol : inf;
  const double x_rtol = o->x_rtol > 0 ? o->x_rtol : inf;
  const double rdiff = o->rdiff > 0 ? o->rdiff : sqrt(kEps);
  const bool need_dx_norm = isfinite(x_tol) || isfinite(x_rtol);
  int64_t maxiter = o->maxiter;
  if (maxiter <= 0) maxiter = (o->iter > 0) ? (int64_t)o->iter + 1 : 100 * ((int64_t)grid_.n_global() + 1);

  nfev_ = 0; inner_total_ = 0; reorth_total_ = 0;
  { const double zero = 0.0; ops_->write_scalars(JS_STOP, 1, &zero); } // (stale after an aborted solve)
  if (hist) { hist->count = 0; hist->nfev = 0; hist->inner_iters = 0; hist->reorth = 0; }

  double* x = u;
  double* xt = XT_;
  double* Fx = FX_;
  double* Ft = FT_;
  double* Gx = GX_; // (null for the mesh problems)
  double* Gt = GT_;
  double nrm[3];
  int rc = eval_residual(x, nullptr, sref(1.0), nullptr, Fx, Gx, JS_F2_A, nrm);
  if (rc) return rc;
  double f2 = nrm[0], fmax = nrm[1], xmax = nrm[2];
  double Fx_norm = sqrt(f2);
  if (hist) { hist->f0_max = fmax; hist->f0_l2 = Fx_norm; }

  ov_slots_.clear(); // KrylovJacobian is created per newton_krylov call: outer_v = [] (_nonlin.py:1508)
  x0_ = x; f0_ = Fx; g0_ = Gx;
  omega_ = rdiff * std::max(1.0, xmax) / std::max(1.0, fmax);
  if (cfg_.problem == JFNK_PROBLEM_SH) ops_->sh_bind_x0(x0_);

  const double gamma = 0.9, eta_max = 0.9999, eta_treshold = 0.1;
  double eta = 1e-3;
  double dxmax = inf; // dx = full(inf) before the first step (_nonlin.py:188)
  double f0_norm = -1.0;
  int64_t iteration = 0;
  int status = 0;
  int result = JFNK_NO_CONVERGENCE;
  int64_t n = 0;
  for (n = 0; n < maxiter; ++n) {
    // TerminationCondition.check (_nonlin.py:361-385)
    iteration++;
    if (f0_norm < 0) f0_norm = fmax;
    if (fmax == 0.0) status = 1;
    else if (o->iter > 0) status = 2 * (iteration > o->iter);
    else status = ((fmax <= f_tol && fmax / f_rtol <= f0_norm) && (dxmax <= x_tol && dxmax / x_rtol <= xmax)) ? 1 : 0;
    if (status) { result = JFNK_OK; break; }

    double tol = std::min(eta, eta * Fx_norm);
    // jacobian.solve(Fx, tol) -> lgmres(op, Fx, rtol=tol, maxiter=1, atol=0, ...): one cycle from x=0
    if (!isfinite(f2)) return fail(JFNK_INVALID, "RHS must contain only finite numbers");
    if (Fx_norm == 0.0) return fail(JFNK_ZERO_STEP, "Jacobian inversion yielded zero vector.");
    double atol_l = std::max(0.0, tol * Fx_norm);
    double ptol = std::min(1.0, std::max(atol_l, tol * Fx_norm) / Fx_norm);
    CycleOut co;
    if (Fx_norm <= std::max(atol_l, tol * Fx_norm)) {
      // lgmres breaks before the first cycle (lgmres.py:166-167): solution stays 0
      return fail(JFNK_ZERO_STEP, "Jacobian inversion yielded zero vector.");
    }
    if (psolve_) {
      // lgmres.py:169-178: v0 = M b (sign carried by dx = -sol), inner_res_0 = ||v0||; the trial buffers are free until
      // the line search: Ft holds v0 for the whole cycle, xt receives A z before M is applied
      if (psolve_(psolve_user_, Fx, Ft) != 0) return fail(JFNK_INVALID, "the preconditioner callback reported an error");
      ops_->mdot(0, nullptr, Ft, JS_TMP3);
      ops_->allreduce_sum(JS_TMP3, 1);
      double v0n2 = 0.0;
      ops_->read_scalars(JS_TMP3, 1, &v0n2);
      if (!isfinite(v0n2)) return fail(JFNK_NONFINITE, "the preconditioner returned non-finite values");
      if (v0n2 == 0.0) return fail(JFNK_INVALID, "Preconditioner returned a zero vector");
      op_tmp_ = xt;
      rc = cycle(Ft, v0n2, ptol, co);
      op_tmp_ = nullptr;
    } else {
      // (a backend that runs the whole cycle in one launch may evaluate the full-step trial point in the same launch)
      trial_.want = true; trial_.xt = xt; trial_.Ft = Ft; trial_.Gt = Gt;
      rc = cycle(Fx, f2, ptol, co);
    }
    trial_.want = false;
    if (rc) { trial_.valid = false; return rc; }
    if (co.sol_n2 == 0.0)
      return fail(JFNK_ZERO_STEP, "Jacobian inversion yielded zero vector. This indicates a bug in the Jacobian approximation.");
    const double* sol = co.sol; // Newton direction dx = -sol
    if (need_dx_norm) {
      ops_->maxabs(sol, JS_TMP3);
      ops_->allreduce_max(JS_TMP3, 1);
      ops_->read_scalars(JS_TMP3, 1, &dxmax);
    } else dxmax = 0.0;

    // line search (_nonlin_line_search :283-325, scalar_search_armijo _linesearch.py:698-753)
    double s = 1.0;
    double tn[3] = {0, 0, 0};
    double last_s = NAN;
    auto phi = [&](double sv, int& err) -> double {
      if (trial_.valid && sv == 1.0) { // F(x - dx) came with the cycle's launch
        trial_.valid = false;
        for (int i = 0; i < 3; ++i) tn[i] = trial_.nrm[i];
        nfev_++;
        err = ops_->status();
      } else {
        trial_.valid = false;
        err = eval_residual(x, sol, sref(-sv), xt, Ft, Gt, JS_F2_B, tn);
      }
      last_s = sv;
      if (!isfinite(tn[0])) return inf; // _safe_norm
      double nv = sqrt(tn[0]);
      return nv * nv;
    };
    int err = 0;
    if (o->line_search) {
      const double phi0 = Fx_norm * Fx_norm, derphi0 = -phi0, c1 = 1e-4, amin = 1e-2;
      double alpha0 = 1.0;
      double phi_a0 = phi(alpha0, err);
      if (err) return err;
      bool found = false;
      if (phi_a0 <= phi0 + c1 * alpha0 * derphi0) { s = alpha0; found = true; }
      if (!found) {
        double alpha1 = -(derphi0)*alpha0 * alpha0 / 2.0 / (phi_a0 - phi0 - derphi0 * alpha0);
        double phi_a1 = phi(alpha1, err);
        if (err) return err;
        if (phi_a1 <= phi0 + c1 * alpha1 * derphi0) { s = alpha1; found = true; }
        while (!found && alpha1 > amin) {
          double factor = alpha0 * alpha0 * alpha1 * alpha1 * (alpha1 - alpha0);
          double a = alpha0 * alpha0 * (phi_a1 - phi0 - derphi0 * alpha1) -
                     alpha1 * alpha1 * (phi_a0 - phi0 - derphi0 * alpha0);
          a = a / factor;
          double b = -alpha0 * alpha0 * alpha0 * (phi_a1 - phi0 - derphi0 * alpha1) +
                     alpha1 * alpha1 * alpha1 * (phi_a0 - phi0 - derphi0 * alpha0);
          b = b / factor;
          double alpha2 = (-b + sqrt(fabs(b * b - 3 * a * derphi0))) / (3.0 * a);
          double phi_a2 = phi(alpha2, err);
          if (err) return err;
          if (phi_a2 <= phi0 + c1 * alpha2 * derphi0) { s = alpha2; found = true; break; }
          if ((alpha1 - alpha2) > alpha1 / 2.0 || (1 - alpha2 / alpha1) < 0.96) alpha2 = alpha1 / 2.0;
          alpha0 = alpha1; alpha1 = alpha2; phi_a0 = phi_a1; phi_a1 = phi_a2;
        }
        if (!found) s = 1.0; // "No suitable step length found. Take the full Newton step" (:311-314)
      }
      if (!(s == last_s)) { phi(s, err); if (err) return err; }
    } else {
      phi(1.0, err);
      if (err) return err;
      s = 1.0;
    }
    // accept: x <- x + s dx, Fx <- F(x)
    std::swap(x, xt);
    std::swap(Fx, Ft);
    std::swap(Gx, Gt);
    double Fx_norm_new = sqrt(tn[0]);
    f2 = tn[0]; fmax = tn[1]; xmax = tn[2];
    // jacobian.update(x, Fx)
    x0_ = x; f0_ = Fx; g0_ = Gx;
    omega_ = rdiff * std::max(1.0, xmax) / std::max(1.0, fmax);
    if (cfg_.problem == JFNK_PROBLEM_SH) ops_->sh_bind_x0(x0_);
    if (cb_) {
      stop_requested_ = false;
      cb_(cb_user_, (int32_t)n, x, Fx, fmax, Fx_norm_new);
      if (stop_requested_) {
        stop_requested_ = false;
        if (x != u) ops_->copy(u, x);
        x0_ = nullptr; f0_ = nullptr; g0_ = nullptr;
        return fail(JFNK_INVALID, "newton_krylov: stopped by the callback");
      }
    }

    // Eisenstat-Walker forcing (:246-250)
    double eta_A = gamma * (Fx_norm_new * Fx_norm_new) / (Fx_norm * Fx_norm);
    if (gamma * eta * eta < eta_treshold) eta = std::min(eta_max, eta_A);
    else eta = std::min(eta_max, std::max(eta_A, gamma * eta * eta));
    Fx_norm = Fx_norm_new;

    if (hist) {
      if (hist->count < hist->capacity) {
        int i = hist->count;
        if (hist->f_max) hist->f_max[i] = fmax;
        if (hist->f_l2) hist->f_l2[i] = Fx_norm;
        if (hist->step) hist->step[i] = s;
        if (hist->inner) hist->inner[i] = co.inner;
      }
      hist->count++;
    }
  }
  if (x != u) ops_->copy(u, x);
  x0_ = nullptr; f0_ = nullptr; g0_ = nullptr; // the iterate buffers are about to be reused
  if (hist) { hist->nfev = nfev_; hist->inner_iters = inner_total_; hist->reorth = reorth_total_; }
  int st = ops_->status();
  if (st) return st;
  if (result == JFNK_NO_CONVERGENCE) return fail(JFNK_NO_CONVERGENCE, "newton_krylov: maximum number of iterations reached");
  return JFNK_OK;
}

int Engine::sh_step(double* u, int nsteps, const jfnk_newton_opts* opts, jfnk_history* hist) {
  if (cfg_.problem != JFNK_PROBLEM_SH) return fail(JFNK_INVALID, "jfnk_sh_step: Swift-Hohenberg JFNK problem only");
  if (!sh_ready_) return fail(JFNK_INVALID, "jfnk_sh_step: call jfnk_sh_setup first");
  for (int s = 0; s < nsteps; ++s) {
    // sh_scipy_nk.py:56-61: Uo = U ; U = newton_krylov(residual, Uo)
    ops_->sh_set_prev(u, D_);
    prev_ready_ = true;
    int rc = newton(u, opts, hist ? hist + s : nullptr);
    if (rc) return rc;
  }
  return ops_->status();
}

} // namespace jfnk
