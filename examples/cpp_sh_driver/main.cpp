// The reference's C++ Swift-Hohenberg driver (cpp_work/NewtonKrylov_Implementation/Project1/main.cpp) against the
// C ABI of libjfnk.so -- no Python, no torch: plain CUDA runtime for the buffers, jfnk.h for everything else.
//
//   main.cpp:3-11    d = 2, N = 5, h = d/N, k = 0.2, r = 0.01, g = 1
//   main.cpp:39-81   Lap / L assembled with Eigen triplets          -> matrix-free inside the engine (jfnk_sh_setup)
//   main.cpp:93-104  for s: Uo = U; UoUo; UoUoUo; U = nonlin_solve(residual, Uo, 6e-6, inf, inf, inf)
//                                                                   -> jfnk_set_prev + jfnk_newton(f_tol = 6e-6)
//
// Build:  g++ -O2 -std=c++17 main.cpp -I../../include -I/usr/local/cuda/include -L<dir of libjfnk.so> -ljfnk \
//             -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,<dir of libjfnk.so> -o sh_driver
// Usage:  sh_driver [N] [d] [steps]     prints one line per step: step, Newton iterations, |F|inf, checksum of U
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "jfnk.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != JFNK_OK) {                                                        \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, jfnk_last_error());    \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 5;
  const double d = argc > 2 ? atof(argv[2]) : 2.0;
  const int steps = argc > 3 ? atoi(argv[3]) : 10;
  const double h = d / N, k = 0.2, r = 0.01, g = 1.0;
  const size_t nn = (size_t)N * N;

  // deterministic initial condition (the reference uses Vec::Random; any state does)
  std::vector<double> U(nn);
  unsigned long long s = 88172645463325252ULL;
  for (size_t i = 0; i < nn; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    U[i] = 2.0 * ((double)(s >> 11) / 9007199254740992.0) - 1.0; // uniform in [-1, 1)
  }

  jfnk_config cfg = {};
  cfg.abi_version = JFNK_ABI_VERSION;
  cfg.problem = JFNK_PROBLEM_SH;
  cfg.nx = cfg.ny = N;
  cfg.row0 = 0; cfg.nrows = N;
  cfg.rank = 0; cfg.nranks = 1;
  cfg.inner_m = 30; cfg.outer_k = 10;
  cfg.gs_mode = JFNK_GS_CGS_IFNEEDED; cfg.gs_tau = 0.25;
  cfg.kernel_variant = 0;
  cfg.stream = nullptr;

  size_t wsb = jfnk_workspace_bytes(&cfg);
  void* ws = nullptr;
  double* dU = nullptr;
  if (cudaMalloc(&ws, wsb) != cudaSuccess || cudaMalloc((void**)&dU, nn * sizeof(double)) != cudaSuccess) {
    fprintf(stderr, "cudaMalloc failed\n");
    return 1;
  }
  cudaMemcpy(dU, U.data(), nn * sizeof(double), cudaMemcpyHostToDevice);

  jfnk_ctx* ctx = nullptr;
  CHECK(jfnk_create(&cfg, ws, wsb, &ctx));
  CHECK(jfnk_sh_setup(ctx, h, r, g, k));
  jfnk_newton_opts o = {};
  o.f_tol = 6e-6; // nonlin_solve(residual, Uo, 6e-6, inf, inf, inf)
  o.line_search = 1;
  std::vector<double> fmax(64), fl2(64), step(64);
  std::vector<int32_t> inner(64);
  for (int st = 0; st < steps; ++st) {
    jfnk_history hist = {};
    hist.capacity = 64;
    hist.f_max = fmax.data(); hist.f_l2 = fl2.data(); hist.step = step.data(); hist.inner = inner.data();
    CHECK(jfnk_set_prev(ctx, dU));
    CHECK(jfnk_newton(ctx, dU, &o, &hist));
    cudaMemcpy(U.data(), dU, nn * sizeof(double), cudaMemcpyDeviceToHost);
    double sum = 0.0, sq = 0.0;
    for (double v : U) { sum += v; sq += v * v; }
    printf("%d %d %.6e %.17g %.17g\n", st + 1, hist.count, hist.count ? fmax[hist.count - 1] : hist.f0_max, sum, sqrt(sq));
  }
  jfnk_destroy(ctx);
  cudaFree(dU);
  cudaFree(ws);
  return 0;
}
