#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  __shared__ double sh[64];
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // DMUL chain
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // rsqrt chain
  x = fabs(x) + 1.0;
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // sqrt chain
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // div chain
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  // float rsqrt + 2 newton
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) { double s = (double)rsqrtf((float)x); double h = 0.5 * x; s = s * fma(-h * s, s, 1.5); s = s * fma(-h * s, s, 1.5); x = s + 1.0; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  // smem round trip: sts -> syncwarp -> lds dependent
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) { sh[threadIdx.x & 31] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 31] + 1e-9; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  // shfl double
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31) + 1e-9;
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  // __syncthreads pair with smem
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) { sh[threadIdx.x & 63] = x; __syncthreads(); x = sh[(threadIdx.x + 1) & 63] + 1e-9; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
  out[threadIdx.x] = x;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 256); cudaMalloc(&c, 8 * 16);
  for (int threads : {32, 128}) {
    k<<<1, threads>>>(o, c, 1.0); k<<<1, threads>>>(o, c, 1.0); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, 8 * 16, cudaMemcpyDeviceToHost);
    const char* nm[] = {"dfma", "dmul", "rsqrt+add", "sqrt+add", "div+add", "frsqrt+2newton+add", "sts-syncwarp-lds", "shfl64+add", "sts-syncthreads-lds"};
    for (int i = 0; i < 9; ++i) printf("threads %d %-22s %.1f cycles/op\n", threads, nm[i], h[i] / 256.0);
  }
  return 0;
}
