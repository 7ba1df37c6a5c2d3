// common.cuh — shared helpers of libsafeincave_cuda (error reporting, warp reductions).
#ifndef SIC_COMMON_CUH_
#define SIC_COMMON_CUH_
#include <cuda_runtime.h>

// host: record an error message (thread-local) and return -1 / check the last launch
int sic_fail(const char* msg);
int sic_check_launch(const char* what);
int sic_check_cuda(cudaError_t e, const char* what);

namespace sic {
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace sic
#endif
