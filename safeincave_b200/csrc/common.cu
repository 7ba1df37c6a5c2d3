// common.cu — error channel and device queries of the C ABI.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/safeincave_cuda.h"
#include "common.cuh"

static thread_local char g_err[512] = "";

int sic_fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
int sic_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
int sic_check_launch(const char* what) { return sic_check_cuda(cudaGetLastError(), what); }

extern "C" const char* sic_last_error(void) { return g_err; }
extern "C" int sic_abi_version(void) { return SIC_ABI_VERSION; }
extern "C" int sic_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (int rc = sic_check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  cudaDeviceProp prop;
  if (int rc = sic_check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return rc;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
