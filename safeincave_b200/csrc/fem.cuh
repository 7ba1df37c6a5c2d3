// fem.cuh — element-by-element (matrix-free) P1 tetrahedron kernels shared by fem.cu / solver.cu.
//
// Weak form of the reference (MomentumEquation.py:1008-1020 with Utils.py:111-227):
//   a(u,v) = int (C_T : eps(u)) : eps(v) dx    =>  K_e = V_e B^T W C_T B,  W = diag(1,1,1,2,2,2),
// B the 6x12 TENSORIAL strain-displacement matrix of the P1 tet (constant per cell).  Written here
// in "stress form": eps = sym(sum_a u_a (x) g_a), sigma = C_T eps (Voigt, C_T used as stored, not
// symmetrised), f_a = V_e sigma . g_a  (the full double contraction sigma : sym(e_j (x) g_a)).
#ifndef SIC_FEM_CUH_
#define SIC_FEM_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/safeincave_cuda.h"
#include "common.cuh"

namespace sic {

#define SIC_EBE_THREADS 128
#define SIC_VEC_THREADS 256

struct CellGeom {
  int node[4];
  double g[12];
  double vol;
};

__device__ __forceinline__ void load_geom(const sic_problem_t& P, int i, CellGeom& c) {
  const int ns = P.cell_stride;
#pragma unroll
  for (int a = 0; a < 4; ++a) c.node[a] = __ldg(P.conn + (size_t)a * ns + i);
#pragma unroll
  for (int k = 0; k < 12; ++k) c.g[k] = __ldg(P.grad + (size_t)k * ns + i);
  c.vol = __ldg(P.vol + i);
}

__device__ __forceinline__ void strain_from_nodal(const CellGeom& c, const double ua[12], double eps[6]) {
  double exx = 0, eyy = 0, ezz = 0, exy = 0, exz = 0, eyz = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = c.g[3 * a], gy = c.g[3 * a + 1], gz = c.g[3 * a + 2];
    const double ux = ua[3 * a], uy = ua[3 * a + 1], uz = ua[3 * a + 2];
    exx += ux * gx; eyy += uy * gy; ezz += uz * gz;
    exy += ux * gy + uy * gx; exz += ux * gz + uz * gx; eyz += uy * gz + uz * gy;
  }
  eps[0] = exx; eps[1] = eyy; eps[2] = ezz; eps[3] = 0.5 * exy; eps[4] = 0.5 * exz; eps[5] = 0.5 * eyz;
}

// sigma = C_T eps with C_T streamed from its SoA rows
__device__ __forceinline__ void stress_from_CT(const sic_problem_t& P, int i, const double eps[6], double sig[6]) {
  const int ns = P.cell_stride;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += __ldg(P.CT + (size_t)(r * 6 + k) * ns + i) * eps[k];
    sig[r] = s;
  }
}

// nodal forces f_a = V sigma . g_a
__device__ __forceinline__ void forces(const CellGeom& c, const double s[6], double f[12]) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = c.g[3 * a], gy = c.g[3 * a + 1], gz = c.g[3 * a + 2];
    f[3 * a + 0] = c.vol * (s[0] * gx + s[3] * gy + s[4] * gz);
    f[3 * a + 1] = c.vol * (s[3] * gx + s[1] * gy + s[5] * gz);
    f[3 * a + 2] = c.vol * (s[4] * gx + s[5] * gy + s[2] * gz);
  }
}

// Deterministic grid-wide reduction of NV values: every block deposits its partial sums, the last
// block to arrive (atomic ticket) adds them up in block order and calls fin(sums).
template <int NV, int THREADS, class Fin>
__device__ __forceinline__ void grid_reduce(double v[NV], double* __restrict__ partials, unsigned* counter, Fin fin) {
  __shared__ double sh[NV][THREADS / 32];
  __shared__ bool is_last;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (l == 0) sh[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < THREADS / 32; ++j) s += sh[k][j];
      partials[(size_t)blockIdx.x * NV + k] = s;
    }
    __threadfence();
    unsigned t = atomicInc(counter, gridDim.x - 1);  // wraps back to 0: reusable next launch
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS) {
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += __ldcg(partials + (size_t)b * NV + k);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(acc[k]);
    if (l == 0) sh[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < THREADS / 32; ++j) s += sh[k][j];
      tot[k] = s;
    }
    fin(tot);
  }
}

}  // namespace sic
#endif
