#include <cstdio>
#include <cuda_runtime.h>
constexpr int IB = 64, PB = 16, PLD = 65, LTD = 66;
template <int VAR>
__global__ void k(const double* A, double* out, long long* cyc, int pw) {
  __shared__ __align__(16) double Lt[IB * LTD];
  __shared__ double col[2 * IB];
  __shared__ double dinv[IB];
  const int lane = threadIdx.x;
  const double* S = A; // read straight from global (row stride PLD emulated below)

  long long t0 = clock64();
  const int r0 = lane, r1 = lane + 32;
  for (int c0 = 0; c0 < pw; c0 += PB) {
    double a0[PB], a1[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      a0[j] = (r0 < pw && c0 + j < pw && c0 + j <= r0) ? S[r0 * IB + c0 + j] : 0.0;
      a1[j] = (r1 < pw && c0 + j < pw && c0 + j <= r1) ? S[r1 * IB + c0 + j] : 0.0;
    }
    const bool u0 = r0 >= c0 && r0 < pw, u1 = r1 >= c0 && r1 < pw;
    if (VAR != 3)
#pragma unroll 4
    for (int c = 0; c < c0; ++c) {
      double l0 = u0 ? Lt[c * LTD + r0] : 0.0, l1 = u1 ? Lt[c * LTD + r1] : 0.0;
      const double2* lc = reinterpret_cast<const double2*>(Lt + c * LTD + c0);
#pragma unroll
      for (int j = 0; j < PB / 2; ++j) {
        double2 l2 = lc[j];
        a0[2 * j] -= l0 * l2.x; a0[2 * j + 1] -= l0 * l2.y;
        a1[2 * j] -= l1 * l2.x; a1[2 * j + 1] -= l1 * l2.y;
      }
    }
#pragma unroll
    for (int kk = 0; kk < PB; ++kk) {
      const int k = c0 + kk;
      if (k < pw) {
        double* cb = col + (kk & 1) * IB;
        if (r0 >= k) cb[r0] = a0[kk];
        if (r1 >= k) cb[r1] = a1[kk];
        __syncwarp();
        double akk = cb[k];
        double cj[PB];
        if (VAR != 2) {
#pragma unroll
        for (int jj = kk + 1; jj < PB; ++jj) cj[jj] = cb[c0 + jj];
        }
        double s;
        if (VAR == 1) s = akk * 0.001; else s = rsqrt(akk);
        double inv = s * s;
        double t0_ = a0[kk] * inv, t1_ = a1[kk] * inv;
        if (VAR != 2) {
#pragma unroll
        for (int jj = kk + 1; jj < PB; ++jj) { a0[jj] -= t0_ * cj[jj]; a1[jj] -= t1_ * cj[jj]; }
        } else { a0[(kk + 1) & 15] -= t0_ * akk; a1[(kk + 1) & 15] -= t1_ * akk; }
        if (r0 >= k && r0 < pw) Lt[k * LTD + r0] = a0[kk] * s;
        if (r1 >= k && r1 < pw) Lt[k * LTD + r1] = a1[kk] * s;
        if (lane == 0) dinv[k] = s;
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  for (int i = lane; i < IB * IB; i += 32) out[i] = Lt[(i % IB) * LTD + i / IB];
}
int main() {
  double *A, *o; long long* c; cudaMalloc(&A, 8 * 4096); cudaMalloc(&o, 8 * 4096); cudaMalloc(&c, 64);
  double h[4096]; for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) h[i * 64 + j] = (i == j) ? 70.0 : 1.0 / (1 + abs(i - j));
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  long long hc;
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 32>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V0 full            %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<1><<<1, 32>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V1 no rsqrt        %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<2><<<1, 32>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V2 no col update   %lld cycles (%.0f/col)\n", hc, hc / 64.0);
    k<3><<<1, 32>>>(A, o, c, 64); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("V3 no left-looking %lld cycles (%.0f/col)\n", hc, hc / 64.0);
  }
  double ho[4096]; k<0><<<1, 32>>>(A, o, c, 64); cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
  // check L L^T = A on a few entries
  double maxerr = 0; for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int q = 0; q <= j; ++q) s += ho[i * 64 + q] * ho[j * 64 + q]; double e = fabs(s - h[i * 64 + j]); if (e > maxerr) maxerr = e; }
  printf("max |LL^T - A| = %.3e  (%s)\n", maxerr, cudaGetErrorString(cudaGetLastError()));
}
