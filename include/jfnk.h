/*
 * jfnk.h -- C ABI of the B200-native Jacobian-free Newton-Krylov engine (libjfnk.so).
 *
 * This is the drop-in boundary for the reference's hot path
 *   newton_krylov(residual, u0, ...)            python_work/sh_scipy_nk.py:61, sh_vscode_nk.py:59,
 *                                               PMA2_nk.py:100, droplet.py:383
 *   nonlin_solve(residual, Uo, 6e-6, inf,inf,inf)   cpp_work/.../Project1/main.cpp:104
 * and for the SciPy internals that call stack runs through
 *   scipy/optimize/_nonlin.py:134  nonlin_solve          -> jfnk_newton
 *   scipy/optimize/_nonlin.py:1557 KrylovJacobian.matvec -> jfnk_linearize + jfnk_jvp
 *   scipy/sparse/linalg/_isolve/lgmres.py:17 lgmres      -> jfnk_lgmres
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every entry point returns an int status (JFNK_OK == 0); no exception crosses the ABI;
 *     jfnk_last_error() returns a thread-local description of the last failure.
 *   - pointers named d* are DEVICE pointers to fp64 data owned by the caller (PyTorch tensors in
 *     the Python host layer).  The engine never frees caller memory.  The Krylov workspace is also
 *     caller-owned: ask jfnk_workspace_bytes(), allocate, hand it to jfnk_create().
 *   - fields are flat fp64 arrays in C order, index = row*nx + col (row = eta/y, col = ksi/x), the
 *     reference's layout (np.meshgrid default, droplet.py:60; kron(eyeY, Dx), droplet.py:805).
 *   - multi-GPU: one context per (process, GPU); rank r owns rows [row0, row0+nrows) of every field.
 *   - a context is not thread-safe; all work is enqueued on cfg.stream.
 *   - there is NO CPU fallback: every compute entry point fails with JFNK_CUDA_ERROR when no
 *     CUDA device / kernel image is available.
 */
#ifndef JFNK_H_
#define JFNK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JFNK_ABI_VERSION 1

/* ---- status codes (mapped 1:1 to Python exceptions by the host shim) ------------------------- */
enum {
  JFNK_OK = 0,
  JFNK_NO_CONVERGENCE = 1, /* scipy NoConvergence(x)            _nonlin.py:258-260 */
  JFNK_ZERO_STEP = 2,      /* ValueError "Jacobian inversion yielded zero vector"  _nonlin.py:225-228 */
  JFNK_NONFINITE = 3,      /* ValueError "Function returned non-finite results"    _nonlin.py:1563-1564 */
  JFNK_INVALID = 4,        /* bad argument / wrong call order (ValueError) */
  JFNK_CUDA_ERROR = 5,     /* RuntimeError */
  JFNK_NCCL_ERROR = 6      /* RuntimeError */
};

/* ---- problems (the residual operators F(u) on the hot path) ---------------------------------- */
enum {
  JFNK_PROBLEM_SH = 0,        /* Swift-Hohenberg CN residual, periodic      sh_scipy_nk.py:47-49 */
  JFNK_PROBLEM_SH_LINEAR = 1, /* linearly-implicit SH operator              sh_linearised.py:51-57 */
  JFNK_PROBLEM_PMA2 = 2,      /* MEMS / PMA moving-mesh residual (p=2)      PMA2_nk.py:121-159 */
  JFNK_PROBLEM_DROPLET = 3    /* thin-film droplet residual                 droplet.py:435-450 */
};

/* Gram-Schmidt variants of the Arnoldi process (SciPy uses modified GS, _gcrotmk.py:123-127;
 * the device path batches all dots of one Arnoldi step into one reduction = classical GS). */
enum {
  JFNK_GS_CGS = 0,          /* classical GS, one pass */
  JFNK_GS_CGS_IFNEEDED = 1, /* second pass when ||w_after|| < gs_tau * ||w_before||; decided by the host from the
                               scalars it reads back once per Arnoldi step anyway */
  JFNK_GS_CGS2 = 2          /* always two passes */
};

typedef struct jfnk_ctx jfnk_ctx;

typedef struct {
  int32_t abi_version; /* JFNK_ABI_VERSION */
  int32_t problem;     /* JFNK_PROBLEM_* */
  int32_t nx, ny;      /* global grid: ny rows (slow axis) x nx columns (fast axis) */
  int32_t row0, nrows; /* this rank's slab of rows; (0, ny) on one GPU */
  int32_t rank, nranks;
  int32_t inner_m;        /* LGMRES inner_m  (scipy default 30; lgmres.py:17) */
  int32_t outer_k;        /* LGMRES outer_k  (KrylovJacobian default 10; _nonlin.py:1477) */
  int32_t gs_mode;        /* JFNK_GS_* */
  int32_t kernel_variant; /* Swift-Hohenberg stencil kernel selection: 0 = product rule (TMA-pipelined marching kernel
                             for even grids >= 256 x 32, one-thread-per-point kernel otherwise); 1 = always the
                             per-point kernel; 2 = the marching kernel wherever it is legal (even nx, 16-byte aligned
                             fields).  1 and 2 exist so that the parity tests can cross-check the two kernels. */
  double gs_tau;          /* threshold for JFNK_GS_CGS_IFNEEDED (Python default 0.25) */
  void* stream;           /* cudaStream_t; NULL = default stream */
} jfnk_config;

/* Options of one nonlinear solve; the fields and defaults are scipy.optimize.newton_krylov's
 * (_nonlin.py:134-137, :328-348).  A negative / zero value selects SciPy's default. */
typedef struct {
  double f_tol;        /* <=0: eps**(1/3) */
  double f_rtol;       /* <=0: inf */
  double x_tol;        /* <=0: inf */
  double x_rtol;       /* <=0: inf */
  double rdiff;        /* <=0: sqrt(eps)            (_nonlin.py:1589-1590) */
  int64_t maxiter;     /* <=0: 100*(n+1)            (_nonlin.py:200-204) */
  int32_t iter;        /* <=0: unset; else exactly that many iterations (_nonlin.py:371-373) */
  int32_t line_search; /* 1 = armijo (default), 0 = none */
} jfnk_newton_opts;

/* Per-Newton-iteration trace, caller-allocated with `capacity` entries per array (may be 0/NULL).
 * Entry i describes the state after Newton iteration i (what SciPy's verbose line prints,
 * _nonlin.py:255-257).  f0_* describe F(u0). */
typedef struct {
  int32_t capacity;
  int32_t count;       /* out: Newton iterations taken (may exceed capacity; only capacity are stored) */
  int64_t nfev;        /* out: residual-equivalent evaluations (F, JVP, line-search trials) */
  int64_t inner_iters; /* out: total Arnoldi iterations */
  int64_t reorth;      /* out: Arnoldi iterations that took the second Gram-Schmidt pass */
  double f0_max, f0_l2;
  double* f_max;       /* ||F||_inf after the step */
  double* f_l2;        /* ||F||_2 after the step */
  double* step;        /* accepted line-search step s */
  int32_t* inner;      /* Arnoldi iterations of that Newton iteration */
} jfnk_history;

/* Newton-iteration callback (SciPy's callback(x, Fx), _nonlin.py:242-243); device pointers. */
typedef void (*jfnk_callback)(void* user, int32_t iter, const double* dx, const double* dF, double f_max, double f_l2);

/* ---- lifetime -------------------------------------------------------------------------------- */
int jfnk_abi_version(void);
const char* jfnk_last_error(void);
/* 1 when a CUDA device is visible and the sm_100a kernel image loads, else 0 (never falls back). */
int jfnk_device_ok(void);
size_t jfnk_workspace_bytes(const jfnk_config* cfg);
int jfnk_create(const jfnk_config* cfg, void* dworkspace, size_t workspace_bytes, jfnk_ctx** out);
int jfnk_destroy(jfnk_ctx* ctx);
int jfnk_set_callback(jfnk_ctx* ctx, jfnk_callback cb, void* user);
/* Callable from inside a jfnk_callback: makes the running jfnk_newton / jfnk_sh_step return JFNK_INVALID right after the
 * callback returns (the host layer uses it to surface an exception raised by the user's callback, which SciPy would
 * propagate, _nonlin.py:240-243 -- no exception may cross this ABI). */
int jfnk_request_stop(jfnk_ctx* ctx);
/* Preconditioner of the inner LGMRES in jfnk_newton: scipy.optimize.newton_krylov's inner_M (_nonlin.py:1398-1410, :1516),
 * which SciPy hands to lgmres as M, i.e. a LEFT preconditioner (lgmres.py:169 v0 = -psolve(r_outer); _gcrotmk.py:113-121
 * w = lpsolve(matvec(z))).  dout = M din on device vectors of the context's local length; the callee enqueues its work on the
 * context's stream (or synchronises) and returns 0 on success.  fn = NULL removes it.  The reference never sets inner_M;
 * this is the SURVEY.md section 8f hook that makes the fixed-domain large-N Swift-Hohenberg / PMA2 cases tractable. */
typedef int (*jfnk_psolve_fn)(void* user, const double* din, double* dout);
int jfnk_set_preconditioner(jfnk_ctx* ctx, jfnk_psolve_fn fn, void* user);

/* ---- multi-GPU plumbing (slab decomposition; NCCL send/recv halos + allreduce) ---------------- */
/* rank 0 fills a 128-byte NCCL unique id; the host layer broadcasts it (torch.distributed). */
int jfnk_comm_unique_id(void* id128);
int jfnk_comm_init(jfnk_ctx* ctx, const void* id128);
/* 1 when the ranks of this communicator exchange halos and reduce the Krylov scalars through peer memory (CUDA IPC over
 * NVLink: direct stores into the neighbours' halo buffers, one-shot all-reduce), 0 when they use the NCCL send/recv +
 * ncclAllReduce path (single rank, IPC unavailable, or JFNK_P2P=0 in the environment). */
int jfnk_comm_peer_memory(jfnk_ctx* ctx);

/* ---- Swift-Hohenberg (sh_scipy_nk.py:15-39) --------------------------------------------------- */
/* h = mesh spacing d/N, r,g = PDE parameters, k = time step. */
int jfnk_sh_setup(jfnk_ctx* ctx, double h, double r, double g, double k);
/* y = Lap x  (periodic 5-point; sh_scipy_nk.py:34-35) and y = L x (13-point; :39). */
int jfnk_spmv_lap(jfnk_ctx* ctx, const double* dx, double* dy);
int jfnk_spmv_sh(jfnk_ctx* ctx, const double* dx, double* dy);
/* Per-step constant from the previous state Uo (sh_scipy_nk.py:56-58). */
int jfnk_set_prev(jfnk_ctx* ctx, const double* dUo);
/* Linearly-implicit SH: set D = diag((5U-Uo)^2 k/16 - g k U) and b = (I + k/2 L) U (sh_linearised.py:51-57);
 * then jfnk_lgmres solves (I + D - k/2 L) x = b.  db receives b. */
int jfnk_shlin_prepare(jfnk_ctx* ctx, const double* dU, const double* dUo, double* db);
/* One linearly-implicit step: dU <- solve, dUo <- old U. info/res as jfnk_lgmres. */
int jfnk_shlin_step(jfnk_ctx* ctx, double* dU, double* dUo, int nsteps, double rtol, int maxiter,
                    int* info, int64_t* matvecs);

/* ---- residual / JVP (any problem) ------------------------------------------------------------- */
int jfnk_residual(jfnk_ctx* ctx, const double* du, double* dF);
/* Fix the linearisation point: x0 = dx0, f0 = F(x0), omega = rdiff*max(1,|x0|inf)/max(1,|f0|inf)
 * (KrylovJacobian.setup/update, _nonlin.py:1552-1555,:1574-1593).  rdiff<=0: sqrt(eps). */
int jfnk_linearize(jfnk_ctx* ctx, const double* dx0, double rdiff);
/* dJv = (F(x0 + sc v) - f0)/sc, sc = omega/||v||_2 ; 0 when ||v|| = 0 (_nonlin.py:1557-1565). */
int jfnk_jvp(jfnk_ctx* ctx, const double* dv, double* dJv);

/* ---- LGMRES on the current linear operator (J at the linearisation point, or the SH_LINEAR
 *      operator): scipy.sparse.linalg.lgmres(op, b, x0=0, rtol, atol=0, maxiter, inner_m, outer_k,
 *      outer_v=<ctx list>, prepend_outer_v=True, store_outer_Av=False) (lgmres.py:17-232).
 *      info: 0 converged, >0 = maxiter reached.  res: last inner residual estimate.  */
int jfnk_lgmres_reset(jfnk_ctx* ctx); /* outer_v = [] */
int jfnk_lgmres(jfnk_ctx* ctx, const double* db, double* dx, double rtol, int maxiter, int* info, double* res,
                int* inner_iters);

/* ---- Newton-Krylov and time stepping ---------------------------------------------------------- */
/* du_inout: initial guess in, solution out (also on JFNK_NO_CONVERGENCE: last iterate, as
 * NoConvergence.args[0]). */
int jfnk_newton(jfnk_ctx* ctx, double* du_inout, const jfnk_newton_opts* opts, jfnk_history* hist);
/* nsteps implicit SH time steps: { set_prev(U); U = newton(U) } (sh_scipy_nk.py:53-61).
 * hist (optional) is an array of nsteps histories. */
int jfnk_sh_step(jfnk_ctx* ctx, double* du_inout, int nsteps, const jfnk_newton_opts* opts, jfnk_history* hist);

/* ---- PMA2 / droplet (non-periodic 4th-order FD on a moving mesh) ------------------------------- */
/* Geometry: dksi, deta = computational mesh spacings; bl,br,bb,bt = the Dirichlet values of
 * Q_ksi / Q_eta on the left/right/bottom/top edges (PMA2_nk.py:240-241: -1,1,-1,1; droplet.py:704-705:
 * endl,endr,endb,endt). */
int jfnk_mesh_setup(jfnk_ctx* ctx, double dksi, double deta, double bl, double br, double bb, double bt);
/* Mesh potential Q -> Q_ksi.., J and the metric fields A11,A22,A12 (compute_Q_spatial_ders,
 * PMA2_nk.py:235-248 / droplet.py:696-711; J: PMA2_nk.py:87 / droplet.py:376). */
int jfnk_mesh_set_potential(jfnk_ctx* ctx, const double* dQ);
/* (v_xx, v_yy) = Laplace_operator(v) (PMA2_nk.py:263-343, droplet.py:601-681). */
int jfnk_mesh_laplace(jfnk_ctx* ctx, const double* dv, double* dvxx, double* dvyy);
/* PMA2: parameters (PMA2_nk.py:23-40) and per-step precompute U.val, CN_term (PMA2_nk.py:83-97). */
int jfnk_pma2_setup(jfnk_ctx* ctx, double lambd, double beta, double epsilon, int m, double dt);
int jfnk_pma2_set_prev(jfnk_ctx* ctx, const double* dUval);
/* Droplet: parameters (droplet.py:23-53) and per-step precompute U.val, F=pde_rhs, dt (droplet.py:373-381). */
int jfnk_droplet_setup(jfnk_ctx* ctx, double epsilon, int n_exp, int m_exp, double Bo, double alpha2, double epsilon2);
int jfnk_droplet_set_prev(jfnk_ctx* ctx, const double* dUval, double dt);
/* compute_U2 / compute_U (droplet.py:413-429, :544-551): the analytic initial shape
 * dU = eps + (1 - eps) sum_i H2(G2(|(Q_ksi, Q_eta) - (x_i, y_i)|, R_i), R_i, V_i) on the mesh given by the potential dQ;
 * info_host = ndrops x (x, y, R, V) (<= 8 droplets), a = steepness of the smoothed contact line (a_ = 100, :24). */
int jfnk_droplet_shape(jfnk_ctx* ctx, const double* dQ, int ndrops, const double* info_host, double a, double* dU);

/* Moving-mesh relaxation on the device: `loops` passes of { compute_Q_spatial_ders; J; Laplace_operator(U.val);
 * compute_and_smooth_monitor; solve_PMA (2-D DCT-II solve); Q += dt_mesh * Q_t } -- loop_pma of droplet.py:590-599
 * (loops = 400, C = cnorm = 0.15, monitor_mode 0 = |u_xx+u_yy|^2) and the single solve_PMA + explicit update of
 * PMA2_nk.py:94,103 (loops = 1, cnorm = 1, monitor_mode 1 = 1/(1+u)^6 when epsilon == 0).  dQ_inout is advanced in
 * place; dUval is the OLD solution (U.val), as in the scripts.  Afterwards call jfnk_mesh_set_potential again. */
int jfnk_mesh_relax(jfnk_ctx* ctx, double* dQ_inout, const double* dUval, double dt_mesh, int loops, double alpha,
                    double gamma, double cnorm, int smoothing_iters, int monitor_mode);

/* ---- introspection for benches / tests --------------------------------------------------------- */
/* number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t jfnk_launch_count(jfnk_ctx* ctx);
/* per-kernel-class device timing (CUDA events on the launch stream) with the ALGORITHMIC bytes each launch
 * moves; bench.py derives the roofline line from it.  jfnk_profile_read synchronises, fills up to `cap`
 * classes, and clears the records. */
typedef struct {
  char name[32];
  int64_t launches;
  double ms;
  double bytes;
} jfnk_kernel_stat;
int jfnk_profile_enable(jfnk_ctx* ctx, int on);
int jfnk_profile_read(jfnk_ctx* ctx, jfnk_kernel_stat* out, int cap, int* count);
/* BLAS-1 building blocks exposed for the roofline microbenchmarks and unit parity tests:
 * out[i] = V_i . w (i<nv), out[nv] = w.w ; V = nv vectors `stride` doubles apart starting at dV. */
int jfnk_multi_dot(jfnk_ctx* ctx, int nv, const double* dV, size_t stride, const double* dw, double* out_host);
/* w -= sum_i coef_host[i] V_i ; returns ||w||^2 in *nrm2_host. */
/* Micro-benchmark of the slab collectives (multi-rank contexts; every rank must call it): enqueues `reps` back-to-back
 * all-reduces of `count` (<= 48) fp64 scalars (what = 0) or halo exchanges of the 2-row halos of dfield (what = 1) and
 * returns the host-clock time per operation in microseconds (stream drained before and after). */
int jfnk_comm_bench(jfnk_ctx* ctx, int what, int count, const double* dfield, int reps, double* usec_host);
int jfnk_multi_axpy(jfnk_ctx* ctx, int nv, const double* dV, size_t stride, const double* coef_host, double* dw,
                    double* nrm2_host);

#ifdef __cplusplus
}
#endif
#endif /* JFNK_H_ */
