// CPU TEST DOUBLE of the device backend -- test infrastructure, never shipped, never loaded by the
// product package.  It implements jfnk::DeviceOps with plain loops over host memory using the SAME
// __host__ __device__ arithmetic (csrc/hd_math.h, csrc/mesh_math.h) the CUDA kernels use, so that the
// host-side solver logic in csrc/engine.cpp (Newton, Armijo, Eisenstat-Walker, LGMRES bookkeeping)
// and the stencil formulas can be checked against SciPy / the reference in the GPU-less container.
// Built by tests/hostsim/build.py into tests/hostsim/_build/libjfnk_hostsim.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "backend.h"
#include "mesh_math.h"

namespace jfnk {

typedef void (*hostsim_allreduce_fn)(void* user, double* buf, int cnt, int op /*0 sum, 1 max*/);
// exchange `rows` rows of nx doubles with the ring neighbours: send_first -> previous rank's bottom halo,
// send_last -> next rank's top halo
typedef void (*hostsim_halo_fn)(void* user, const double* send_first, const double* send_last, double* recv_top,
                                double* recv_bot, int count);

class HostOps : public DeviceOps {
 public:
  explicit HostOps(const jfnk_config& c) {
    g_.nx = c.nx; g_.ny = c.ny; g_.row0 = c.row0; g_.nrows = c.nrows; g_.rank = c.rank; g_.nranks = c.nranks;
    S_.assign(JS_COUNT, 0.0);
  }
  int64_t launches() const override { return launches_; }
  void profile_enable(bool) override {}
  int profile_read(KernelStat*, int) override { return 0; }
  int status() override { return 0; }
  const char* last_error() const override { return ""; }

  void read_scalars(int off, int cnt, double* host) override { memcpy(host, &S_[off], sizeof(double) * cnt); }
  void write_scalars(int off, int cnt, const double* host) override { memcpy(&S_[off], host, sizeof(double) * cnt); }
  // (a collective of a speculatively enqueued Arnoldi step is dropped with the step: every rank sees the same JS_STOP)
  void allreduce_sum(int off, int cnt) override { if (g_.nranks > 1 && !stopped()) ar_(user_, &S_[off], cnt, 0); }
  void allreduce_max(int off, int cnt) override { if (g_.nranks > 1) ar_(user_, &S_[off], cnt, 1); }

  bool can_speculate() const override { return !(getenv("JFNK_SPECULATE") && atoi(getenv("JFNK_SPECULATE")) == 0); }
  bool stopped() const { return S_[JS_STOP] != 0.0; } // what the CUDA kernels of the Arnoldi loop check first
  void mdot(int nv, const double* const* V, const double* w, int out_off) override {
    if (stopped()) return;
    launches_++;
    size_t n = g_.n();
    for (int i = 0; i < nv; ++i) {
      double acc = 0;
      for (size_t e = 0; e < n; ++e) acc += V[i][e] * w[e];
      S_[out_off + i] = acc;
    }
    double acc = 0;
    for (size_t e = 0; e < n; ++e) acc += w[e] * w[e];
    S_[out_off + nv] = acc;
  }
  void gs_update(int nv, const double* const* V, double* w, int rd_off, int n2_off, int fuse_givens_j) override {
    if (stopped()) return;
    launches_++;
    size_t n = g_.n();
    std::vector<double> c(nv);
    for (int i = 0; i < nv; ++i) c[i] = S_[rd_off + i] / S_[JS_VN2 + i];
    double acc = 0;
    for (size_t e = 0; e < n; ++e) {
      double t = w[e];
      for (int i = 0; i < nv; ++i) t -= c[i] * V[i][e];
      w[e] = t;
      acc += t * t;
    }
    S_[n2_off] = acc;
    if (fuse_givens_j >= 0 && g_.nranks == 1) hess_givens_step(S_.data(), fuse_givens_j, 0, 0);
  }
  void maxpy_sub(int nv, const double* const* V, double* w, int n2_off) override {
    launches_++;
    size_t n = g_.n();
    double acc = 0;
    for (size_t e = 0; e < n; ++e) {
      double t = w[e];
      for (int i = 0; i < nv; ++i) t -= S_[JS_COEF + i] * V[i][e];
      w[e] = t;
      acc += t * t;
    }
    S_[n2_off] = acc;
  }
  void maxpy(int nz, const double* const* Z, double* out, int n2_off) override {
    launches_++;
    size_t n = g_.n();
    double acc = 0;
    for (size_t e = 0; e < n; ++e) {
      double t = S_[JS_COEF + 0] * Z[0][e];
      for (int i = 1; i < nz; ++i) t += S_[JS_COEF + i] * Z[i][e];
      out[e] = t;
      acc += t * t;
    }
    S_[n2_off] = acc;
  }
  void lincomb(double* out, ScalarRef a, const double* x, ScalarRef b, const double* y, int n2_off) override {
    launches_++;
    size_t n = g_.n();
    double av = eval_sref(S_.data(), a), bv = eval_sref(S_.data(), b);
    double acc = 0;
    for (size_t e = 0; e < n; ++e) {
      double t = av * x[e];
      if (y) t += bv * y[e];
      out[e] = t;
      acc += t * t;
    }
    if (n2_off >= 0) S_[n2_off] = acc;
  }
  void diff_scale(double* out, const double* x, const double* y, ScalarRef div) override {
    launches_++;
    size_t n = g_.n();
    double dv = eval_sref(S_.data(), div);
    for (size_t e = 0; e < n; ++e) out[e] = (x[e] - y[e]) / dv;
  }
  void copy(double* dst, const double* src) override {
    launches_++;
    if (dst != src) memmove(dst, src, sizeof(double) * g_.n());
  }
  void maxabs(const double* v, int out_off) override {
    launches_++;
    double m = 0;
    for (size_t e = 0; e < g_.n(); ++e) m = fmax(m, fabs(v[e]));
    S_[out_off] = m;
  }
  void givens(int j, int taken, int rerun) override {
    if (stopped()) return;
    launches_++;
    hess_givens_step(S_.data(), j, taken, rerun);
  }
  void lsq(int nit, const int* zn2_idx, int scale_n2_idx) override { launches_++; lsq_solve(S_.data(), nit, zn2_idx, scale_n2_idx); }

  // ---- Swift-Hohenberg ---------------------------------------------------------------------------
  void sh_setup(const SHParams& p) override { shp_ = p; }

  struct Field { // slab-local field with 2-row halos and periodic columns
    const double* v; const double* top; const double* bot; int nx, nrows;
    double at(int r, int c) const {
      c = (c % nx + nx) % nx;
      if (r < 0) return top[(r + 2) * nx + c];
      if (r >= nrows) return bot[(r - nrows) * nx + c];
      return v[(size_t)r * nx + c];
    }
  };
  void halos(const double* v, std::vector<double>& top, std::vector<double>& bot) {
    int nx = g_.nx, nr = g_.nrows;
    top.resize(2 * nx); bot.resize(2 * nx);
    if (g_.nranks == 1) {
      memcpy(top.data(), v + (size_t)(nr - 2) * nx, sizeof(double) * 2 * nx);
      memcpy(bot.data(), v, sizeof(double) * 2 * nx);
    } else {
      halo_(user_, v, v + (size_t)(nr - 2) * nx, top.data(), bot.data(), 2 * nx);
    }
  }
  // combined field t = x + a v with halos
  void combined(const double* x, const double* v, double a, std::vector<double>& t, std::vector<double>& top,
                std::vector<double>& bot) {
    size_t n = g_.n();
    t.resize(n);
    for (size_t e = 0; e < n; ++e) t[e] = v ? x[e] + a * v[e] : x[e];
    halos(t.data(), top, bot);
  }
  static void sums(const Field& f, int r, int c, double& uc, double& s1, double& sd, double& s2) {
    uc = f.at(r, c);
    // same grouping as the marching CUDA kernel: horizontal pairs first
    double a1 = f.at(r, c - 1) + f.at(r, c + 1);
    double a2 = f.at(r, c - 2) + f.at(r, c + 2);
    double a1u = f.at(r - 1, c - 1) + f.at(r - 1, c + 1);
    double a1d = f.at(r + 1, c - 1) + f.at(r + 1, c + 1);
    s1 = a1 + f.at(r - 1, c) + f.at(r + 1, c);
    sd = a1u + a1d;
    s2 = a2 + f.at(r - 2, c) + f.at(r + 2, c);
  }
  void sh_spmv(int which, const double* x, double* y) override {
    launches_++;
    std::vector<double> top, bot;
    halos(x, top, bot);
    Field f{x, top.data(), bot.data(), g_.nx, g_.nrows};
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        y[(size_t)r * g_.nx + c] = which ? sh_apply13(shp_, uc, s1, sd, s2) : sh_apply5(shp_, uc, s1);
      }
  }
  void sh_set_prev(const double* uo, double* d) override {
    launches_++;
    std::vector<double> top, bot;
    halos(uo, top, bot);
    Field f{uo, top.data(), bot.data(), g_.nx, g_.nrows};
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        d[(size_t)r * g_.nx + c] = sh_prev_const(shp_, uc, sh_apply13(shp_, uc, s1, sd, s2));
      }
  }
  void sh_residual(const double* x, const double* v, ScalarRef a, const double* d, double* xt_out, double* F,
                   double* g_out, int norm_off) override {
    launches_++;
    std::vector<double> t, top, bot;
    combined(x, v, eval_sref(S_.data(), a), t, top, bot);
    Field f{t.data(), top.data(), bot.data(), g_.nx, g_.nrows};
    double f2 = 0, fm = 0, xm = 0;
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        size_t e = (size_t)r * g_.nx + c;
        double Gv = sh_G(shp_, uc, sh_apply13(shp_, uc, s1, sd, s2));
        double Fv = Gv - d[e];
        F[e] = Fv;
        if (g_out) g_out[e] = Gv;
        f2 += Fv * Fv; fm = fmax(fm, fabs(Fv)); xm = fmax(xm, fabs(uc));
        if (!(fabs(Fv) <= 1.79e308)) fm = INFINITY; // NaN/inf poison like np.abs(F).max()
      }
    if (xt_out) memcpy(xt_out, t.data(), sizeof(double) * g_.n());
    S_[norm_off] = f2; S_[norm_off + 1] = fm; S_[norm_off + 2] = xm;
  }
  void sh_bind_x0(const double*) override {}
  void sh_jvp(const double* x0, const double* z, ScalarRef sc, ScalarRef div, const double* d, const double* f0,
              const double* g0, double* w) override {
    if (stopped()) return;
    launches_++;
    std::vector<double> t, top, bot;
    combined(x0, z, eval_sref(S_.data(), sc), t, top, bot);
    double dv = eval_sref(S_.data(), div);
    Field f{t.data(), top.data(), bot.data(), g_.nx, g_.nrows};
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        size_t e = (size_t)r * g_.nx + c;
        double Gv = sh_G(shp_, uc, sh_apply13(shp_, uc, s1, sd, s2));
        w[e] = g0 ? (Gv - g0[e]) * (1.0 / dv) : ((Gv - d[e]) - f0[e]) / dv;
      }
  }
  void shlin_prepare(const double* U, const double* Uo, double* D, double* b) override {
    launches_++;
    std::vector<double> top, bot;
    halos(U, top, bot);
    Field f{U, top.data(), bot.data(), g_.nx, g_.nrows};
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        size_t e = (size_t)r * g_.nx + c;
        D[e] = shlin_diag(shp_, uc, Uo[e]);
        b[e] = shlin_rhs(shp_, uc, sh_apply13(shp_, uc, s1, sd, s2));
      }
  }
  void shlin_matvec(const double* z, ScalarRef a, const double* D, double* w) override {
    if (stopped()) return;
    launches_++;
    std::vector<double> top, bot;
    halos(z, top, bot);
    double av = eval_sref(S_.data(), a);
    Field f{z, top.data(), bot.data(), g_.nx, g_.nrows};
    for (int r = 0; r < g_.nrows; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double uc, s1, sd, s2;
        sums(f, r, c, uc, s1, sd, s2);
        size_t e = (size_t)r * g_.nx + c;
        w[e] = av * shlin_apply(shp_, uc, sh_apply13(shp_, uc, s1, sd, s2), D[e]);
      }
  }

  // ---- moving mesh ---------------------------------------------------------------------------------
  void mesh_metrics(const MeshParams& mp, const double* Q, double* const* M) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) mesh_metrics_point(gm, Q, r, c, M);
  }
  void mesh_laplace(const MeshParams& mp, const double* const* M, const double* v, double* vxx, double* vyy,
                    int sum_only, int deriv_bc) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        double xx, yy;
        mesh_laplace_point(gm, M, v, r, c, deriv_bc, xx, yy);
        size_t e = (size_t)r * g_.nx + c;
        if (sum_only) vxx[e] = xx + yy;
        else { vxx[e] = xx; vyy[e] = yy; }
      }
  }
  void pma2_rhs(const Pma2Params& pp, const double* u, const double* lap2, double* out) override {
    launches_++;
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        size_t e = (size_t)r * g_.nx + c;
        bool bdy = (r == 0 || c == 0 || r == g_.ny - 1 || c == g_.nx - 1);
        out[e] = bdy ? 0.0 : pma2_rhs_point(pp, u[e], lap2[e]);
      }
  }
  void norms(const double* F, const double* u, int norm_off) {
    double f2 = 0, fm = 0, xm = 0;
    for (size_t e = 0; e < g_.n(); ++e) {
      f2 += F[e] * F[e]; fm = fmax(fm, fabs(F[e])); xm = fmax(xm, fabs(u[e]));
      if (!(fabs(F[e]) <= 1.79e308)) fm = INFINITY;
    }
    S_[norm_off] = f2; S_[norm_off + 1] = fm; S_[norm_off + 2] = xm;
  }
  void pma2_combine(const Pma2Params& pp, const double* u, const double* uval, const double* rhs, const double* cn,
                    double* F, int norm_off) override {
    launches_++;
    for (size_t e = 0; e < g_.n(); ++e) F[e] = pma2_combine_point(pp, u[e], uval[e], rhs[e], cn[e]);
    norms(F, u, norm_off);
  }
  void droplet_shape(const MeshParams& mp, const double* Q, const DropList& drops, double a, double eps, double* out) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) out[(size_t)r * g_.nx + c] = droplet_shape_point(gm, Q, r, c, drops, a, eps);
  }
  void droplet_pressure(const DropletParams& dp, const double* h, const double* lap, double* p) override {
    launches_++;
    for (size_t e = 0; e < g_.n(); ++e) p[e] = droplet_pressure_point(dp, h[e], lap[e]);
  }
  void droplet_flux(const MeshParams& mp, const DropletParams& dp, const double* const* M, const double* p,
                    const double* h, double* A, double* B) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) {
        size_t e = (size_t)r * g_.nx + c;
        droplet_flux_point(gm, dp, M, p, h, r, c, A[e], B[e]);
      }
  }
  void droplet_div(const MeshParams& mp, const double* const* M, const double* A, const double* B, double* out) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) out[(size_t)r * g_.nx + c] = droplet_div_point(gm, M, A, B, r, c);
  }
  void droplet_combine(const DropletParams& dp, const double* u, const double* uval, const double* F2,
                       const double* Fprev, double* F, int norm_off) override {
    launches_++;
    for (size_t e = 0; e < g_.n(); ++e) F[e] = droplet_combine_point(dp, u[e], uval[e], F2[e], Fprev[e]);
    norms(F, u, norm_off);
  }

  // ---- moving-mesh relaxation ------------------------------------------------------------------------
  void pma_monitor(int mode, const double* u, const double* lap, double* out) override {
    launches_++;
    for (size_t e = 0; e < g_.n(); ++e) out[e] = pma_monitor_point(mode, u[e], mode == 1 ? 0.0 : lap[e]);
  }
  void pma_smooth(const MeshParams& mp, const double* in, double* out) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) out[(size_t)r * g_.nx + c] = pma_smooth_point(gm, in, r, c);
  }
  void pma_wsum(const double* mon, const double* J, int out_off) override {
    launches_++;
    double acc = 0;
    for (size_t e = 0; e < g_.n(); ++e) acc += mon[e] * fabs(J[e]);
    S_[out_off] = acc;
  }
  void pma_rhs(const double* mon, const double* J, ScalarRef add, double alpha, double* out) override {
    launches_++;
    double av = eval_sref(S_.data(), add);
    for (size_t e = 0; e < g_.n(); ++e) out[e] = sqrt((mon[e] + av) * fabs(J[e])) / alpha;
  }
  static void gemm(int M, int N, int K, const double* A, int ta, const double* B, int tb, double* C) {
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (ta ? A[(size_t)k * M + i] : A[(size_t)i * K + k]) * (tb ? B[(size_t)j * K + k] : B[(size_t)k * N + j]);
        C[(size_t)i * N + j] = acc;
      }
  }
  void pma_dct2(const double* in, double* tmp, double* out, int inverse) override {
    launches_++;
    const int nx = g_.nx, ny = g_.ny;
    if (dctx_.empty()) {
      dctx_.resize((size_t)nx * nx); dcty_.resize((size_t)ny * ny);
      for (int k = 0; k < nx; ++k) for (int j = 0; j < nx; ++j) dctx_[(size_t)k * nx + j] = dct2_entry(k, j, nx);
      for (int k = 0; k < ny; ++k) for (int j = 0; j < ny; ++j) dcty_[(size_t)k * ny + j] = dct2_entry(k, j, ny);
    }
    if (!inverse) { gemm(ny, nx, ny, dcty_.data(), 0, in, 0, tmp); gemm(ny, nx, nx, tmp, 0, dctx_.data(), 1, out); }
    else { gemm(ny, nx, ny, dcty_.data(), 1, in, 0, tmp); gemm(ny, nx, nx, tmp, 0, dctx_.data(), 0, out); }
  }
  void pma_spectral_divide(const MeshParams& mp, double gamma, double* Y) override {
    launches_++;
    MeshGeom gm = geom(mp);
    for (int r = 0; r < g_.ny; ++r)
      for (int c = 0; c < g_.nx; ++c) { size_t e = (size_t)r * g_.nx + c; Y[e] = Y[e] / (1.0 - gamma * pma_leig(gm, r, c)); }
  }

  void set_comm(hostsim_allreduce_fn ar, hostsim_halo_fn halo, void* user) { ar_ = ar; halo_ = halo; user_ = user; }
  std::vector<double> dctx_, dcty_;
  MeshTables htab_;
  MeshGeom geom(const MeshParams& mp) { fill_mesh_tables(mp, htab_); return make_geom(mp, g_.nx, g_.ny, &htab_); }

 private:
  Grid g_;
  std::vector<double> S_;
  SHParams shp_;
  int64_t launches_ = 0;
  hostsim_allreduce_fn ar_ = nullptr;
  hostsim_halo_fn halo_ = nullptr;
  void* user_ = nullptr;
};

static HostOps* g_last_ops = nullptr;

int backend_device_ok(std::string& why) { why = "hostsim test double (no device)"; return 0; }
DeviceOps* backend_make_ops(const jfnk_config& cfg, std::string&, int&) {
  g_last_ops = new HostOps(cfg);
  return g_last_ops;
}
int backend_unique_id(void* id128, std::string&) { memset(id128, 0, 128); return JFNK_OK; }
int backend_comm_init(DeviceOps*, const void*, std::string&) { return JFNK_OK; }
int backend_peer_memory(DeviceOps*) { return 0; }

} // namespace jfnk

// test-only hook: attach gloo-backed collectives (implemented in Python) to the most recently created context
extern "C" void hostsim_set_comm(jfnk::hostsim_allreduce_fn ar, jfnk::hostsim_halo_fn halo, void* user) {
  if (jfnk::g_last_ops) jfnk::g_last_ops->set_comm(ar, halo, user);
}
