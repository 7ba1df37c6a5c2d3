// ebe_tma.cuh — the matrix-free P1 operator, version 2: persistent CTAs fed by the TMA unit.
//
// Version 1 (one thread per cell, plain LDGs) was latency-bound: ptxas interleaves the 53 independent
// per-cell loads (36 C_T rows, 12 grad rows, volume, 4 node ids) with the FP64 math that consumes them,
// so each warp keeps only a few loads in flight (ncu: 44 % occupancy, long-scoreboard stalls, DRAM at
// 34 % of peak; profiles/r1_k_ebe_dot_v1_ncu_full_summary.txt).
// Here C_T and the cell geometry are stored TILED: the 36 C_T rows of 128 consecutive cells are one
// contiguous 36 KB block, their grad/vol/conn another 15 KB block (a first attempt with 53 separate 1 KB
// row copies per tile was SLOWER than version 1: the TMA unit's per-request cost dominated).  Two
// cp.async.bulk requests (TMA 1-D bulk copies, SASS UBLKCP) bring a tile into shared memory,
// completion is signalled on an mbarrier, and a
// two-stage ring keeps the next tile in flight while the 128 threads compute the current one from
// shared memory (conflict-free: thread i reads element i of every row).  Two CTAs per SM (2 x 104 KB
// of shared memory) => ~200 KB of loads in flight per SM, independent of register allocation.
#ifndef SIC_EBE_TMA_CUH_
#define SIC_EBE_TMA_CUH_

#include "fem.cuh"

namespace sic {

#define SIC_TILE 128
#define SIC_TMA_STAGES 2

template <int MODE>
struct __align__(128) TileStage;

template <>
struct __align__(128) TileStage<0> {   // y += K x
  double CT[36][SIC_TILE];
  sic_geom_tile_t geo;
};
template <>
struct __align__(128) TileStage<1> {   // r += sum V B^T W CT (eps_rhs - B x0)
  double CT[36][SIC_TILE];
  sic_geom_tile_t geo;
  double er[6][SIC_TILE];
};

template <int MODE>
struct __align__(128) TileSmem {
  TileStage<MODE> st[SIC_TMA_STAGES];
  unsigned long long full[SIC_TMA_STAGES];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Issued by warp 0.  C_T and the geometry of a tile are each ONE contiguous block in HBM (tiled layout,
// include/safeincave_cuda.h), i.e. two bulk copies of 36 KB and 15 KB; eps_rhs (residual only, once per
// solve) is still SoA: six 1 KB rows.
template <int MODE>
__device__ __forceinline__ void issue_tile(const sic_problem_t& P, int tile, TileStage<MODE>& st,
                                           unsigned long long* bar, int lane) {
  constexpr unsigned CT_BYTES = 36 * SIC_TILE * 8, GEO_BYTES = sizeof(sic_geom_tile_t);
  constexpr unsigned BYTES = CT_BYTES + GEO_BYTES + (MODE == 1 ? 6 * SIC_TILE * 8 : 0);
  if (lane == 0) mbar_expect_tx(bar, BYTES);
  __syncwarp();
  if (lane == 0) bulk_g2s(&st.CT[0][0], P.CT + (size_t)tile * 36 * SIC_TILE, CT_BYTES, bar);
  if (lane == 1) bulk_g2s(&st.geo, P.geom_tiles + tile, GEO_BYTES, bar);
  if (MODE == 1 && lane >= 2 && lane < 8) {
    const size_t ns = (size_t)P.cell_stride, c0 = (size_t)tile * SIC_TILE;
    bulk_g2s(&((TileStage<1>&)st).er[lane - 2][0], P.eps_rhs + (lane - 2) * ns + c0, SIC_TILE * 8, bar);
  }
}

// One cell from a staged tile.  Returns V eps:sigma (the cell's share of x^T K x).
template <int MODE>
__device__ __forceinline__ double tile_cell(const TileStage<MODE>& st, int t, const double* __restrict__ x,
                                            double* __restrict__ y) {
  int node[4];
  double g[12], ua[12];
#pragma unroll
  for (int a = 0; a < 4; ++a) node[a] = st.geo.conn[a][t];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int j = 0; j < 3; ++j) ua[3 * a + j] = __ldg(x + 3 * (size_t)node[a] + j);
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) g[k] = st.geo.grad[k][t];
  const double vol = st.geo.vol[t];
  double exx = 0, eyy = 0, ezz = 0, exy = 0, exz = 0, eyz = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = g[3 * a], gy = g[3 * a + 1], gz = g[3 * a + 2];
    const double ux = ua[3 * a], uy = ua[3 * a + 1], uz = ua[3 * a + 2];
    exx += ux * gx; eyy += uy * gy; ezz += uz * gz;
    exy += ux * gy + uy * gx; exz += ux * gz + uz * gx; eyz += uy * gz + uz * gy;
  }
  double eps[6] = {exx, eyy, ezz, 0.5 * exy, 0.5 * exz, 0.5 * eyz};
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) eps[k] = ((const TileStage<1>&)st).er[k][t] - eps[k];
  }
  double s[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += st.CT[r * 6 + k][t] * eps[k];
    s[r] = acc;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = g[3 * a], gy = g[3 * a + 1], gz = g[3 * a + 2];
    double* ya = y + 3 * (size_t)node[a];
    atomicAdd(ya + 0, vol * (s[0] * gx + s[3] * gy + s[4] * gz));
    atomicAdd(ya + 1, vol * (s[3] * gx + s[1] * gy + s[5] * gz));
    atomicAdd(ya + 2, vol * (s[4] * gx + s[5] * gy + s[2] * gz));
  }
  return vol * ((eps[0] * s[0] + eps[1] * s[1] + eps[2] * s[2]) + 2.0 * (eps[3] * s[3] + eps[4] * s[4] + eps[5] * s[5]));
}

// The persistent tile loop.  blockDim.x == SIC_TILE.  Returns this thread's sum of V eps:sigma.
template <int MODE>
__device__ __forceinline__ double ebe_tiles(const sic_problem_t& P, const double* __restrict__ x,
                                            double* __restrict__ y, TileSmem<MODE>& sm) {
  const int n_tiles = (P.n_cells + SIC_TILE - 1) / SIC_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SIC_TMA_STAGES; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double energy = 0.0;
  int tile = blockIdx.x;
  if (warp == 0 && tile < n_tiles) issue_tile<MODE>(P, tile, sm.st[0], &sm.full[0], lane);
  for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
    const int s = it & 1;
    const int next = tile + gridDim.x;
    if (warp == 0 && next < n_tiles) issue_tile<MODE>(P, next, sm.st[s ^ 1], &sm.full[s ^ 1], lane);
    mbar_wait(&sm.full[s], (it >> 1) & 1);
    const int cell = tile * SIC_TILE + tid;
    if (cell < P.n_cells) energy += tile_cell<MODE>(sm.st[s], tid, x, y);
    __syncthreads();   // every thread is done with stage s before it is refilled two iterations later
  }
  return energy;
}

// ---- host-side launch helpers ---------------------------------------------------------------------
inline int ebe_grid(int n_cells) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int n_tiles = (n_cells + SIC_TILE - 1) / SIC_TILE;
  const int g = 2 * sms;   // two resident CTAs per SM
  return n_tiles < g ? (n_tiles > 0 ? n_tiles : 1) : g;
}
template <class Kernel>
inline cudaError_t ebe_allow_smem(Kernel k, size_t bytes) {
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace sic
#endif
