// CUDA error handling of the library: a failed runtime call raises spllt::CudaFailure, which the
// extern "C" layer (capi.cu) turns into the reference's error codes -- info%flag =
// SPLLT_ERROR_ALLOCATION (-1) for out-of-memory, SPLLT_ERROR_UNKNOWN (-99) otherwise
// (src/spllt_data_mod.F90:31-35) -- after printing the CUDA error.  Nothing aborts the host
// process and no exception crosses the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>

namespace spllt {

struct CudaFailure {
  cudaError_t err;
  const char* file;
  int line;
  bool oom() const { return err == cudaErrorMemoryAllocation; }
};
struct NoDevice {};   // no CUDA device: the numerical phase has no CPU fallback

inline void cuda_check(cudaError_t e, const char* file, int line) {
  if (e == cudaSuccess) return;
  fprintf(stderr, "spllt_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line);
  cudaGetLastError();   // clear the sticky-less error state for the next call
  throw CudaFailure{e, file, line};
}

}  // namespace spllt

#define CK(x) ::spllt::cuda_check((x), __FILE__, __LINE__)
