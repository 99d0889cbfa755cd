// Host-side solver logic: LGMRES outer/inner bookkeeping, Newton iteration, Armijo line search,
// Eisenstat-Walker forcing, termination -- the scalar control flow of
//   scipy/optimize/_nonlin.py:134-277   nonlin_solve
//   scipy/optimize/_nonlin.py:283-325   _nonlin_line_search  (+ _linesearch.py:698-753 scalar_search_armijo)
//   scipy/optimize/_nonlin.py:328-385   TerminationCondition
//   scipy/optimize/_nonlin.py:1477-1598 KrylovJacobian
//   scipy/sparse/linalg/_isolve/lgmres.py:124-232, _gcrotmk.py:16-183 (_fgmres)
// All vector work is enqueued through DeviceOps; only convergence scalars come back to the host.
#pragma once
#include <deque>
#include <string>
#include <vector>
#include "../../include/jfnk.h"
#include "device_ops.h"
#include "mesh_math.h"

namespace jfnk {

struct CycleOut {
  const double* sol = nullptr; // assembled correction dx (lives in an augmentation ring slot)
  int sol_slot = -1;
  double sol_n2 = 0.0;
  double res = 0.0;
  int inner = 0;
};

class Engine {
 public:
  Engine(const jfnk_config& cfg, DeviceOps* ops, double* workspace, size_t workspace_doubles);
  ~Engine();

  static size_t vector_stride(const jfnk_config& cfg);
  static int vector_count(const jfnk_config& cfg);
  static size_t workspace_doubles(const jfnk_config& cfg);
  static const char* validate(const jfnk_config& cfg);

  DeviceOps* ops() { return ops_; }
  const Grid& grid() const { return grid_; }
  const std::string& error() const { return err_; }
  void set_callback(jfnk_callback cb, void* user) { cb_ = cb; cb_user_ = user; }
  void set_preconditioner(jfnk_psolve_fn fn, void* user) { psolve_ = fn; psolve_user_ = user; }
  void request_stop() { stop_requested_ = true; }

  // problem setup
  int sh_setup(double h, double r, double g, double k);
  int set_prev(const double* uo);
  int spmv(int which, const double* x, double* y);
  int shlin_prepare(const double* U, const double* Uo, double* b);
  int shlin_step(double* U, double* Uo, int nsteps, double rtol, int maxiter, int* info, int64_t* matvecs);
  int mesh_setup(const MeshParams& mp);
  int mesh_set_potential(const double* Q);
  int mesh_laplace(const double* v, double* vxx, double* vyy);
  int pma2_setup(const Pma2Params& pp);
  int pma2_set_prev(const double* uval);
  int droplet_setup(const DropletParams& dp);
  int droplet_set_prev(const double* uval, double dt);
  int mesh_relax(double* Q, const double* Uval, double dt, int loops, const PmaParams& pp);
  int droplet_shape(const double* Q, int ndrops, const double* info, double a, double* U);

  // operator level
  int residual(const double* u, double* F);
  int linearize(const double* x0, double rdiff);
  int jvp(const double* v, double* Jv);
  int lgmres_reset();
  int lgmres(const double* b, double* x, double rtol, int maxiter, int* info, double* res, int* inner);

  // solver level
  int newton(double* u, const jfnk_newton_opts* opts, jfnk_history* hist);
  int sh_step(double* u, int nsteps, const jfnk_newton_opts* opts, jfnk_history* hist);

 private:
  int fail(int code, const std::string& msg) { err_ = msg; return code; }
  // F(x + a v) with norms -> slot set norm_off; all-reduces the norms; host copy in nrm[3] = {sum F^2, max|F|, max|x|}
  // g_out (Swift-Hohenberg only, may be null): G = F + d of the evaluated point, see DeviceOps::sh_residual
  int eval_residual(const double* x, const double* v, ScalarRef a, double* xt_out, double* F, double* g_out, int norm_off,
                    double nrm[3]);
  // w = A (z * inv_norm) for the current linear operator
  // unit_input: z is treated as z/||z|| (Arnoldi vectors are stored unnormalised); else the SciPy matvec(v)
  void apply_operator(const double* z, int zn2_idx, double* w, bool unit_input);
  int cycle(const double* v0vec, double v0n2, double ptol, CycleOut& out);
  int lgmres_general(const double* b, double* x, double rtol, int maxiter, int* info, double* res, int* inner);
  void generic_residual(const double* u, double* F, int norm_off);
  bool problem_ready(std::string& why) const;

  jfnk_config cfg_;
  Grid grid_;
  DeviceOps* ops_;
  std::string err_;
  jfnk_callback cb_ = nullptr;
  void* cb_user_ = nullptr;
  jfnk_psolve_fn psolve_ = nullptr; // left preconditioner of the inner solve (inner_M); used by newton() only
  void* psolve_user_ = nullptr;
  bool stop_requested_ = false;     // set from inside the callback (jfnk_request_stop)
  double* op_tmp_ = nullptr;        // where A z is formed before M is applied (a buffer newton() has free)

  // workspace vectors
  double* ws_;
  size_t vstride_;
  double *XT_, *FX_, *FT_, *D_;
  double *GX_, *GT_; // Swift-Hohenberg: G(x) = F(x) + d of the iterate / the trial point (the FD quotient's subtrahend)
  std::vector<double*> scratch_; // scratch fields for the multi-kernel (mesh) residuals
  double* MF_[7];  // metric fields
  double *UVAL_, *CN_; // previous state and Crank-Nicolson term of the mesh problems
  std::vector<double*> VS_; // Arnoldi slots 1..m
  std::vector<double*> OV_; // augmentation ring slots
  std::deque<int> ov_slots_; // ring slots in use, oldest first

  // problem state
  bool sh_ready_ = false, prev_ready_ = false, mesh_ready_ = false, metrics_ready_ = false, pp_ready_ = false,
       dp_ready_ = false;
  SHParams shp_;
  MeshParams mp_;
  Pma2Params pp_;
  DropletParams dp_;

  // linearisation point of the Jacobian operator
  const double* x0_ = nullptr;
  const double* f0_ = nullptr;
  const double* g0_ = nullptr; // G(x0) when the problem keeps it (Swift-Hohenberg), else null
  double omega_ = 0.0;
  bool linear_op_ = false; // true: SH_LINEAR operator instead of the FD Jacobian

  // The first line-search trial F(x - dx) of a Newton iteration, evaluated by the launch that ran the inner cycle
  // (DeviceOps::cycle_fused): newton() says where it wants it before calling cycle(), cycle() reports whether it was done.
  struct Trial {
    bool want = false, valid = false;
    double *xt = nullptr, *Ft = nullptr, *Gt = nullptr;
    double nrm[3] = {0, 0, 0};
  } trial_;

  // statistics of the running solve
  int64_t nfev_ = 0, inner_total_ = 0, reorth_total_ = 0;
};

} // namespace jfnk
