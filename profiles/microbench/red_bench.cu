#include <cstdio>
#include <cuda_runtime.h>
// every thread issues `per` REDs; a warp covers 8 rows x 4 consecutive doubles per instruction
template <int MODE>
__global__ void k(double* dst, long long ld, int per, long long span) {
  int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long base = ((long long)warp * 8 * ld * 7919) % span;
  for (int i = 0; i < per; ++i) {
    long long a = (base + (long long)(lane >> 2) * ld + (lane & 3) + (long long)i * 4) % span;
    if (MODE == 0) atomicAdd(dst + a, 1.0);
    else if (MODE == 1) dst[a] += 1.0;
    else dst[a] = 1.0;
  }
}
int main() {
  long long span = 1ll << 28;  // 2 GiB of doubles
  double* d; cudaMalloc(&d, span * 8); cudaMemset(d, 0, span * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) for (int blocks : {148 * 2, 148 * 8}) {
    int per = 256;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, 256>>>(d, 4096, per, span); else if (mode == 1) k<1><<<blocks, 256>>>(d, 4096, per, span); else k<2><<<blocks, 256>>>(d, 4096, per, span);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * 256 * per;
    printf("mode %d (%s) blocks %d: %.1f G elem/s (%s)\n", mode, mode == 0 ? "RED.F64" : mode == 1 ? "ld+add+st" : "st", blocks, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
}
