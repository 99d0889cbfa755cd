// extern "C" boundary of libjfnk.so (include/jfnk.h).  No exception crosses it.
#include <chrono>
#include <math.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/jfnk.h"
#include "backend.h"
#include "engine.h"

using namespace jfnk;

struct jfnk_ctx {
  Engine* eng;
  DeviceOps* ops;
};

static thread_local std::string g_err;

static int set_err(int code, const std::string& msg) { g_err = msg; return code; }

// translate an engine status into the thread-local message
static int done(jfnk_ctx* c, int rc) {
  if (rc == JFNK_OK) return rc;
  if (rc == JFNK_CUDA_ERROR || rc == JFNK_NCCL_ERROR) {
    const char* m = c->ops->last_error();
    g_err = (m && *m) ? m : c->eng->error();
  } else {
    g_err = c->eng->error();
  }
  return rc;
}

#define JF_TRY try {
#define JF_CATCH                                                                       \
  }                                                                                    \
  catch (const std::bad_alloc&) { return set_err(JFNK_INVALID, "out of host memory"); } \
  catch (const std::exception& e) { return set_err(JFNK_INVALID, e.what()); }           \
  catch (...) { return set_err(JFNK_INVALID, "unknown C++ exception"); }

#define JF_CHECK_CTX(c) \
  if (!(c) || !(c)->eng) return set_err(JFNK_INVALID, "null jfnk_ctx")

extern "C" {

int jfnk_abi_version(void) { return JFNK_ABI_VERSION; }
const char* jfnk_last_error(void) { return g_err.c_str(); }

int jfnk_device_ok(void) {
  std::string why;
  int ok = backend_device_ok(why);
  if (!ok) g_err = why;
  return ok;
}

size_t jfnk_workspace_bytes(const jfnk_config* cfg) {
  if (!cfg || Engine::validate(*cfg)) return 0;
  return Engine::workspace_doubles(*cfg) * sizeof(double);
}

int jfnk_create(const jfnk_config* cfg, void* dworkspace, size_t workspace_bytes, jfnk_ctx** out) {
  JF_TRY
  if (!cfg || !out) return set_err(JFNK_INVALID, "jfnk_create: null argument");
  *out = nullptr;
  if (const char* why = Engine::validate(*cfg)) return set_err(JFNK_INVALID, std::string("jfnk_create: ") + why);
  size_t need = Engine::workspace_doubles(*cfg) * sizeof(double);
  if (!dworkspace || workspace_bytes < need)
    return set_err(JFNK_INVALID, "jfnk_create: workspace missing or smaller than jfnk_workspace_bytes()");
  if (((uintptr_t)dworkspace) % 256 != 0) return set_err(JFNK_INVALID, "jfnk_create: workspace must be 256-byte aligned");
  std::string why;
  int code = JFNK_CUDA_ERROR;
  DeviceOps* ops = backend_make_ops(*cfg, why, code);
  if (!ops) return set_err(code, "jfnk_create: " + why);
  jfnk_ctx* c = new jfnk_ctx;
  c->ops = ops;
  c->eng = new Engine(*cfg, ops, (double*)dworkspace, need / sizeof(double));
  *out = c;
  return JFNK_OK;
  JF_CATCH
}

int jfnk_destroy(jfnk_ctx* ctx) {
  if (!ctx) return JFNK_OK;
  delete ctx->eng;
  delete ctx->ops;
  delete ctx;
  return JFNK_OK;
}

int jfnk_set_callback(jfnk_ctx* ctx, jfnk_callback cb, void* user) {
  JF_CHECK_CTX(ctx);
  ctx->eng->set_callback(cb, user);
  return JFNK_OK;
}

int jfnk_request_stop(jfnk_ctx* ctx) {
  JF_CHECK_CTX(ctx);
  ctx->eng->request_stop();
  return JFNK_OK;
}

int jfnk_set_preconditioner(jfnk_ctx* ctx, jfnk_psolve_fn fn, void* user) {
  JF_TRY
  JF_CHECK_CTX(ctx);
  ctx->eng->set_preconditioner(fn, user);
  return JFNK_OK;
  JF_CATCH
}

int jfnk_comm_unique_id(void* id128) {
  std::string why;
  int rc = backend_unique_id(id128, why);
  if (rc) g_err = why;
  return rc;
}

int jfnk_comm_init(jfnk_ctx* ctx, const void* id128) {
  JF_TRY
  JF_CHECK_CTX(ctx);
  std::string why;
  int rc = backend_comm_init(ctx->ops, id128, why);
  if (rc) g_err = why;
  return rc;
  JF_CATCH
}

int jfnk_comm_peer_memory(jfnk_ctx* ctx) {
  if (!ctx || !ctx->ops) return 0;
  return backend_peer_memory(ctx->ops);
}

int jfnk_sh_setup(jfnk_ctx* ctx, double h, double r, double g, double k) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->sh_setup(h, r, g, k)); JF_CATCH
}
int jfnk_spmv_lap(jfnk_ctx* ctx, const double* dx, double* dy) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->spmv(0, dx, dy)); JF_CATCH
}
int jfnk_spmv_sh(jfnk_ctx* ctx, const double* dx, double* dy) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->spmv(1, dx, dy)); JF_CATCH
}
int jfnk_set_prev(jfnk_ctx* ctx, const double* dUo) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->set_prev(dUo)); JF_CATCH
}
int jfnk_shlin_prepare(jfnk_ctx* ctx, const double* dU, const double* dUo, double* db) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->shlin_prepare(dU, dUo, db)); JF_CATCH
}
int jfnk_shlin_step(jfnk_ctx* ctx, double* dU, double* dUo, int nsteps, double rtol, int maxiter, int* info,
                    int64_t* matvecs) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->shlin_step(dU, dUo, nsteps, rtol, maxiter, info, matvecs)); JF_CATCH
}
int jfnk_residual(jfnk_ctx* ctx, const double* du, double* dF) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->residual(du, dF)); JF_CATCH
}
int jfnk_linearize(jfnk_ctx* ctx, const double* dx0, double rdiff) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->linearize(dx0, rdiff)); JF_CATCH
}
int jfnk_jvp(jfnk_ctx* ctx, const double* dv, double* dJv) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->jvp(dv, dJv)); JF_CATCH
}
int jfnk_lgmres_reset(jfnk_ctx* ctx) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->lgmres_reset()); JF_CATCH
}
int jfnk_lgmres(jfnk_ctx* ctx, const double* db, double* dx, double rtol, int maxiter, int* info, double* res,
                int* inner_iters) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->lgmres(db, dx, rtol, maxiter, info, res, inner_iters)); JF_CATCH
}
int jfnk_newton(jfnk_ctx* ctx, double* du_inout, const jfnk_newton_opts* opts, jfnk_history* hist) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->newton(du_inout, opts, hist)); JF_CATCH
}
int jfnk_sh_step(jfnk_ctx* ctx, double* du_inout, int nsteps, const jfnk_newton_opts* opts, jfnk_history* hist) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->sh_step(du_inout, nsteps, opts, hist)); JF_CATCH
}

int jfnk_mesh_setup(jfnk_ctx* ctx, double dksi, double deta, double bl, double br, double bb, double bt) {
  JF_TRY JF_CHECK_CTX(ctx);
  MeshParams mp; mp.dksi = dksi; mp.deta = deta; mp.bl = bl; mp.br = br; mp.bb = bb; mp.bt = bt;
  return done(ctx, ctx->eng->mesh_setup(mp));
  JF_CATCH
}
int jfnk_mesh_set_potential(jfnk_ctx* ctx, const double* dQ) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->mesh_set_potential(dQ)); JF_CATCH
}
int jfnk_mesh_laplace(jfnk_ctx* ctx, const double* dv, double* dvxx, double* dvyy) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->mesh_laplace(dv, dvxx, dvyy)); JF_CATCH
}
int jfnk_pma2_setup(jfnk_ctx* ctx, double lambd, double beta, double epsilon, int m, double dt) {
  JF_TRY JF_CHECK_CTX(ctx);
  Pma2Params pp; pp.lambd = lambd; pp.beta = beta; pp.epsilon = epsilon; pp.m = m; pp.dt = dt;
  return done(ctx, ctx->eng->pma2_setup(pp));
  JF_CATCH
}
int jfnk_pma2_set_prev(jfnk_ctx* ctx, const double* dUval) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->pma2_set_prev(dUval)); JF_CATCH
}
int jfnk_droplet_setup(jfnk_ctx* ctx, double epsilon, int n_exp, int m_exp, double Bo, double alpha2, double epsilon2) {
  JF_TRY JF_CHECK_CTX(ctx);
  DropletParams dp; dp.epsilon = epsilon; dp.n_exp = n_exp; dp.m_exp = m_exp; dp.Bo = Bo; dp.alpha2 = alpha2;
  dp.epsilon2 = epsilon2; dp.dt = 0.0;
  return done(ctx, ctx->eng->droplet_setup(dp));
  JF_CATCH
}
int jfnk_droplet_shape(jfnk_ctx* ctx, const double* dQ, int ndrops, const double* info_host, double a, double* dU) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->droplet_shape(dQ, ndrops, info_host, a, dU)); JF_CATCH
}
int jfnk_droplet_set_prev(jfnk_ctx* ctx, const double* dUval, double dt) {
  JF_TRY JF_CHECK_CTX(ctx); return done(ctx, ctx->eng->droplet_set_prev(dUval, dt)); JF_CATCH
}

int jfnk_mesh_relax(jfnk_ctx* ctx, double* dQ_inout, const double* dUval, double dt_mesh, int loops, double alpha,
                    double gamma, double cnorm, int smoothing_iters, int monitor_mode) {
  JF_TRY JF_CHECK_CTX(ctx);
  PmaParams pp; pp.alpha = alpha; pp.gamma = gamma; pp.cnorm = cnorm; pp.smoothing_iters = smoothing_iters;
  pp.monitor_mode = monitor_mode;
  return done(ctx, ctx->eng->mesh_relax(dQ_inout, dUval, dt_mesh, loops, pp));
  JF_CATCH
}

int jfnk_profile_enable(jfnk_ctx* ctx, int on) {
  JF_CHECK_CTX(ctx);
  ctx->ops->profile_enable(on != 0);
  return JFNK_OK;
}
int jfnk_profile_read(jfnk_ctx* ctx, jfnk_kernel_stat* out, int cap, int* count) {
  JF_TRY JF_CHECK_CTX(ctx);
  static_assert(sizeof(jfnk_kernel_stat) == sizeof(KernelStat), "jfnk_kernel_stat layout");
  int k = ctx->ops->profile_read(reinterpret_cast<KernelStat*>(out), cap);
  if (count) *count = k;
  return done(ctx, ctx->ops->status());
  JF_CATCH
}

int64_t jfnk_launch_count(jfnk_ctx* ctx) { return (ctx && ctx->ops) ? ctx->ops->launches() : 0; }

int jfnk_multi_dot(jfnk_ctx* ctx, int nv, const double* dV, size_t stride, const double* dw, double* out_host) {
  JF_TRY JF_CHECK_CTX(ctx);
  if (nv < 0 || nv > JF_MAXV) return set_err(JFNK_INVALID, "jfnk_multi_dot: nv out of range");
  const double* V[JF_MAXV + 1];
  for (int i = 0; i < nv; ++i) V[i] = dV + stride * (size_t)i;
  ctx->ops->mdot(nv, V, dw, JS_RD);
  ctx->ops->allreduce_sum(JS_RD, nv + 1);
  if (out_host) ctx->ops->read_scalars(JS_RD, nv + 1, out_host);
  return done(ctx, ctx->ops->status());
  JF_CATCH
}

int jfnk_comm_bench(jfnk_ctx* ctx, int what, int count, const double* dfield, int reps, double* usec_host) {
  JF_TRY JF_CHECK_CTX(ctx);
  if (reps < 1 || !usec_host || (what == 0 && (count < 1 || count > JF_MAXV)) || (what == 1 && !dfield) || what < 0 || what > 1)
    return set_err(JFNK_INVALID, "jfnk_comm_bench: bad arguments");
  double dummy;
  ctx->ops->read_scalars(JS_TMP0, 1, &dummy); // drain the stream
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) {
    if (what == 0) ctx->ops->allreduce_sum(JS_RD2, count);
    else ctx->ops->sh_bind_x0(dfield);
  }
  ctx->ops->read_scalars(JS_TMP0, 1, &dummy);
  auto t1 = std::chrono::steady_clock::now();
  *usec_host = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
  return done(ctx, ctx->ops->status());
  JF_CATCH
}

int jfnk_multi_axpy(jfnk_ctx* ctx, int nv, const double* dV, size_t stride, const double* coef_host, double* dw,
                    double* nrm2_host) {
  JF_TRY JF_CHECK_CTX(ctx);
  if (nv < 0 || nv > JF_MAXV || !coef_host) return set_err(JFNK_INVALID, "jfnk_multi_axpy: bad arguments");
  const double* V[JF_MAXV + 1];
  for (int i = 0; i < nv; ++i) V[i] = dV + stride * (size_t)i;
  ctx->ops->write_scalars(JS_COEF, nv, coef_host);
  ctx->ops->maxpy_sub(nv, V, dw, JS_TMP0);
  ctx->ops->allreduce_sum(JS_TMP0, 1);
  if (nrm2_host) ctx->ops->read_scalars(JS_TMP0, 1, nrm2_host);
  return done(ctx, ctx->ops->status());
  JF_CATCH
}

} // extern "C"
