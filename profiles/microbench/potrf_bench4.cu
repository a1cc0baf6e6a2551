#include <cstdio>
#include <cuda_runtime.h>
constexpr int IB = 64, PB = 16, LTD = 66;
template <int VAR>
__global__ void __launch_bounds__(128) k(const double* A, double* out, long long* cyc, int pw) {
  __shared__ __align__(16) double Lt[IB * LTD];
  __shared__ double col[2 * IB];
  __shared__ double dinv[IB];
  const int tid = threadIdx.x;
  const double* S = A;
  __syncthreads();
  long long t0 = clock64();
  const int r = tid & 63, h = tid >> 6;
  for (int c0 = 0; c0 < pw; c0 += PB) {
    const int cbase = c0 + h * 8;
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (r < pw && cbase + j < pw && cbase + j <= r) ? S[r * IB + cbase + j] : 0.0;
    if (VAR != 3 && r >= c0 && r < pw) {
#pragma unroll 4
      for (int c = 0; c < c0; ++c) {
        double lrc = Lt[c * LTD + r];
        const double2* lc = reinterpret_cast<const double2*>(Lt + c * LTD + cbase);
#pragma unroll
        for (int j = 0; j < 4; ++j) { double2 l2 = lc[j]; a[2 * j] -= lrc * l2.x; a[2 * j + 1] -= lrc * l2.y; }
      }
    }
#pragma unroll
    for (int kk = 0; kk < PB; ++kk) {
      const int k = c0 + kk;
      if (k < pw) {
        double* cb = col + (kk & 1) * IB;
        const int hh = kk >> 3;
        if (h == hh && r >= k) cb[r] = a[kk & 7];
        if (VAR != 4) asm volatile("bar.sync 1, 128;" ::: "memory");
        if (h >= hh) {
          double akk = cb[k];
          double me = cb[r];
          double cj[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cj[j] = cb[cbase + j];
          double s = (VAR == 1) ? akk * 0.001 : rsqrt(akk);
          double tt = me * (s * s);
          if (VAR != 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] -= tt * cj[j];
          } else a[(kk + 1) & 7] -= tt * cj[(kk + 1) & 7];
          if (h == hh) {
            if (r >= k && r < pw) Lt[k * LTD + r] = me * s;
            if (r == k) dinv[k] = s;
          }
        }
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  for (int i = tid; i < IB * IB; i += 128) out[i] = Lt[(i % IB) * LTD + i / IB];
}
int main() {
  double *A, *o; long long* c; cudaMalloc(&A, 8 * 4096); cudaMalloc(&o, 8 * 4096); cudaMalloc(&c, 64);
  double h[4096]; for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) h[i * 64 + j] = (i == j) ? 70.0 : 1.0 / (1 + abs(i - j));
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  long long hc;
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 128>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V0 full            %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<1><<<1, 128>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V1 no rsqrt        %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<2><<<1, 128>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V2 no col update   %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<3><<<1, 128>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V3 no left-looking %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<4><<<1, 128>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V4 no column barrier %lld cycles (%.0f/col)\n", hc, hc / 64.0);
  }
  double ho[4096]; k<0><<<1, 128>>>(A, o, c, 64); cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
  double maxerr = 0; for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int q = 0; q <= j; ++q) s += ho[i * 64 + q] * ho[j * 64 + q]; double e = fabs(s - h[i * 64 + j]); if (e > maxerr) maxerr = e; }
  printf("max |LL^T - A| = %.3e  (%s)\n", maxerr, cudaGetErrorString(cudaGetLastError()));
}
