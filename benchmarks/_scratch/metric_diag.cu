#include <cstdio>
#include <vector>
#include <cmath>
#include "mesh_kernels.cuh"
using namespace jfnk;
int main() {
  const int N = 300;
  MeshParams mp; mp.dksi = mp.deta = 2.0 / (N - 1); mp.bl = -1; mp.br = 1; mp.bb = -1; mp.bt = 1;
  MeshTables ht; fill_mesh_tables(mp, ht);
  MeshTables* dt; cudaMalloc(&dt, sizeof(ht)); cudaMemcpy(dt, &ht, sizeof(ht), cudaMemcpyHostToDevice);
  std::vector<double> Q((size_t)N * N);
  for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) {
    double x = -1 + 2.0 * c / (N - 1), y = -1 + 2.0 * r / (N - 1);
    Q[(size_t)r * N + c] = 0.5 * x * x + 0.5 * y * y + 0.02 * cos(2.1 * x + 0.3) * sin(1.7 * y - 0.2);
  }
  size_t n = (size_t)N * N;
  double* dQ; cudaMalloc(&dQ, n * 8); cudaMemcpy(dQ, Q.data(), n * 8, cudaMemcpyHostToDevice);
  MetricPtrs P; std::vector<std::vector<double>> hm(7, std::vector<double>(n)), gm_(7, std::vector<double>(n));
  double* hp[7];
  for (int i = 0; i < 7; ++i) { cudaMalloc(&P.m[i], n * 8); hp[i] = hm[i].data(); }
  MeshGeom gd = make_geom(mp, N, N, dt), gh = make_geom(mp, N, N, &ht);
  mesh_metrics_kernel<<<148 * 4, 256>>>(gd, dQ, P);
  cudaDeviceSynchronize();
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < 7; ++i) cudaMemcpy(gm_[i].data(), P.m[i], n * 8, cudaMemcpyDeviceToHost);
  for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) mesh_metrics_point(gh, Q.data(), r, c, hp);
  const char* names[7] = {"qxx", "qyy", "qxy", "J", "A11", "A22", "A12"};
  for (int i = 0; i < 7; ++i) {
    double mx = 0, mxv = 0; int mr = 0, mc = 0;
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) {
      size_t e = (size_t)r * N + c;
      double d = fabs(gm_[i][e] - hm[i][e]);
      if (d > mx) { mx = d; mr = r; mc = c; mxv = hm[i][e]; }
    }
    printf("%s max abs diff %.3e at (%d,%d) host val %.17g gpu val %.17g\n", names[i], mx, mr, mc, mxv, gm_[i][(size_t)mr * N + mc]);
  }
  return 0;
}
