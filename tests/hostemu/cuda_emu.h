// cuda_emu.h — minimal single-threaded emulation of the CUDA constructs used by safeincave_b200/csrc.
//
// TEST INFRASTRUCTURE ONLY.  It lets the CPU test-suite execute the product's device code (the very
// same .cu/.cuh sources, passed through tests/hostemu/translate.py) on the host, so that kernels can
// be checked against the oracle in a container without a GPU.  It is never linked into, loaded by or
// shipped with the product: safeincave_b200 only ever loads libsafeincave_cuda.so and raises without a
// CUDA device.
//
// Model: a launch runs its blocks one after the other; every thread of a block runs on its own fiber
// (hand-rolled x86-64 stack switch), and a pass-based scheduler advances all live threads from collective
// to collective (__syncthreads, __shfl_*_sync) in lock step; a thread without collectives simply runs to
// completion in its first slice.
// Atomics are plain read-modify-writes; "device memory" is host memory.
#ifndef SIC_CUDA_EMU_H_
#define SIC_CUDA_EMU_H_

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define SIC_HOSTEMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local   /* one block at a time PER emulated rank (= host thread) */
#define __align__(n) alignas(n)
#define __constant__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

namespace sic_emu {
void launch(unsigned grid, unsigned block, const std::function<void()>& thread_body);
void collective();                 // yield to the scheduler until every live thread of the block got here
double* shfl_slot(int parity_out); // per-thread exchange slots of the current block
int collective_count();
}  // namespace sic_emu

#define SIC_EMU_LAUNCH(kern, grid, block, smem, stream, ...) \
  ::sic_emu::launch((unsigned)(grid), (unsigned)(block), [&]() { kern(__VA_ARGS__); })

static inline void __syncthreads() { ::sic_emu::collective(); }
static inline void __threadfence() {}
static inline void __syncwarp(unsigned = 0xffffffffu) { ::sic_emu::collective(); }

// every lane deposits its value, all lanes pass the collective, then each reads its partner's
static inline double sic_emu_shfl(double v, int src_lane_in_warp) {
  const int par = ::sic_emu::collective_count() & 1;
  double* slot = ::sic_emu::shfl_slot(par);
  slot[threadIdx.x] = v;
  ::sic_emu::collective();
  const int base = (int)(threadIdx.x & ~31u);
  int src = base + src_lane_in_warp;
  if (src >= (int)blockDim.x) src = (int)threadIdx.x;
  return slot[src];
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  return sic_emu_shfl(v, (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask));
}
static inline double __shfl_down_sync(unsigned, double v, unsigned delta) {
  const int l = (int)(threadIdx.x & 31u) + (int)delta;
  return sic_emu_shfl(v, l < 32 ? l : (int)(threadIdx.x & 31u));
}
static inline double __shfl_sync(unsigned, double v, int src) { return sic_emu_shfl(v, src & 31); }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }

static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0u : o + 1u; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }

static inline long long __double_as_longlong(double x) { long long u; memcpy(&u, &x, 8); return u; }
static inline double __longlong_as_double(long long u) { double x; memcpy(&x, &u, 8); return x; }

// ---- the slice of the runtime API the host drivers use ----------------------------------------------
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { int multiProcessorCount = 148, major = 10, minor = 0; };
static inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n); return *p ? cudaSuccess : cudaErrorUnknown; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(1, n); return *p ? cudaSuccess : cudaErrorUnknown; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }

// stand-ins for what ebe_tma.cuh (compiled out: SIC_EBE_IMPL == 3) would have declared
#define SIC_TILE 128
template <class K> static inline cudaError_t ebe_allow_smem(K, size_t) { return cudaSuccess; }

#endif  // SIC_CUDA_EMU_H_
