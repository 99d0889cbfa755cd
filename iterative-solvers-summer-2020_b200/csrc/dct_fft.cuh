// Fast orthonormal DCT-II / DCT-III along the rows of a field (scipy.fft.dct / idct with norm="ortho", the transforms of
// solve_PMA: PMA2_nk.py:401-402, droplet.py:586-587) for power-of-two row lengths -- the large synthetic grids
// (BASELINE config 3: 2048^2), where the dense DCT-matrix products of pma_kernels.cuh would cost four 2048^3 fp64 products.
//
// Makhoul's N-point algorithm: v[n] = x[2n], v[N-1-n] = x[2n+1] ; V = FFT_N(v) ; X[k] = 2 Re(e^{-i pi k/(2N)} V[k]), scaled
// by sqrt(1/(4N)) (k = 0) / sqrt(1/(2N)).  Inverse: V[k] = e^{+i pi k/(2N)} (X[k] - i X[N-k])/2 (X[N] = 0), v = IFFT(V),
// un-permute.  One CTA transforms one row entirely in shared memory: the permuted row is stored in bit-reversed order, the
// log2 N radix-2 butterfly stages run in place with twiddles from a shared-memory table, the pre/post factors come from an
// L2-resident table, loads and stores of the row are coalesced.  HBM traffic: one read and one write of the field per pass.
// The column direction is done as transpose -> row pass -> transpose (transpose_kernel, 32 x 32 tiles through shared memory).
#pragma once
#include "cuda_common.cuh"

namespace jfnk {

constexpr int kFftThreads = 256;

// W[j] = e^{-2 pi i j/N}, j < N/2 ; Wq[k] = e^{-i pi k/(2N)}, k < N   (sincospi: accurate to an ulp for every index)
__global__ void __launch_bounds__(256) dct_twiddle_kernel(int N, double2* W, double2* Wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N / 2) {
    double s, c;
    sincospi(-2.0 * (double)i / (double)N, &s, &c);
    W[i] = make_double2(c, s);
  }
  if (i < N) {
    double s, c;
    sincospi(-(double)i / (2.0 * (double)N), &s, &c);
    Wq[i] = make_double2(c, s);
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// rows x N field, N = 2^logN ; INV = false: out = DCT-II(in) per row ; true: out = DCT-III(in) per row (both orthonormal)
template <bool INV>
__global__ void __launch_bounds__(kFftThreads) dct_fft_rows_kernel(int N, int logN, int rows, const double* __restrict__ in,
                                                                   double* __restrict__ out, const double2* __restrict__ W,
                                                                   const double2* __restrict__ Wq) {
  extern __shared__ __align__(16) double2 fft_smem[];
  double2* buf = fft_smem;     // [N]
  double2* tw = fft_smem + N;  // [N/2]
  const int tid = threadIdx.x;
  for (int i = tid; i < N / 2; i += kFftThreads) tw[i] = W[i];
  const double f0 = sqrt(1.0 / (4.0 * N)), f1 = sqrt(1.0 / (2.0 * N));
  const unsigned shift = 32u - (unsigned)logN;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const double* x = in + (size_t)row * N;
    double* y = out + (size_t)row * N;
    __syncthreads(); // (twiddles in place ; the previous row's epilogue done)
    if (!INV) {
      for (int n = tid; n < N; n += kFftThreads) {
        const int m = (n & 1) ? N - 1 - (n >> 1) : (n >> 1);
        buf[__brev((unsigned)m) >> shift] = make_double2(x[n], 0.0);
      }
    } else {
      for (int k = tid; k < N; k += kFftThreads) {
        const double xu = x[k] / (k == 0 ? f0 : f1);
        const double xr = (k == 0) ? 0.0 : x[N - k] / f1;
        const double2 wq = Wq[k];
        // conj(V[k]) with V[k] = conj(wq) (xu - i xr)/2: the inverse FFT is run as conj(FFT(conj(V)))/N
        const double2 v = cmul(make_double2(wq.x, -wq.y), make_double2(0.5 * xu, -0.5 * xr));
        buf[__brev((unsigned)k) >> shift] = make_double2(v.x, -v.y);
      }
    }
    __syncthreads();
    for (int s = 0; s < logN; ++s) {
      const int half = 1 << s;
      for (int b = tid; b < N / 2; b += kFftThreads) {
        const int pos = b & (half - 1);
        const int i0 = ((b >> s) << (s + 1)) + pos, i1 = i0 + half;
        const double2 t = cmul(tw[pos << (logN - 1 - s)], buf[i1]);
        const double2 u = buf[i0];
        buf[i0] = make_double2(u.x + t.x, u.y + t.y);
        buf[i1] = make_double2(u.x - t.x, u.y - t.y);
      }
      __syncthreads();
    }
    if (!INV) {
      for (int k = tid; k < N; k += kFftThreads) {
        const double2 p = cmul(Wq[k], buf[k]);
        y[k] = (k == 0 ? f0 : f1) * (2.0 * p.x);
      }
    } else {
      const double invN = 1.0 / (double)N;
      for (int n = tid; n < N; n += kFftThreads) {
        const int m = (n & 1) ? N - 1 - (n >> 1) : (n >> 1);
        y[n] = buf[m].x * invN;
      }
    }
  }
}

// out[c][r] = in[r][c], in: rows x cols
__global__ void __launch_bounds__(256) transpose_kernel(int rows, int cols, const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
  const int tiles_x = (cols + 31) / 32, tiles_y = (rows + 31) / 32;
  for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
    const int c0 = (t % tiles_x) * 32, r0 = (t / tiles_x) * 32;
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
      if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(size_t)(r0 + j) * cols + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
      if (c0 + j < cols && r0 + tx < rows) out[(size_t)(c0 + j) * rows + r0 + tx] = tile[tx][j];
  }
}

} // namespace jfnk
