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

#include <type_traits>

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
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += __ldg(P.CT + SIC_CT_INDEX(r * 6 + k, i)) * eps[k];
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

// ---- operator, version 3: one thread per cell with FRONT-BATCHED loads ---------------------------
// ptxas, left alone, interleaves the 53 independent per-cell loads with the FP64 math that consumes
// them (version 1: few loads in flight per warp, DRAM at 34 % of peak).  Here every load is an
// `asm volatile` statement, which the compiler may not reorder against the others: the 4 node ids,
// 12 gradients, the volume and all 36 C_T entries are issued back to back (53 loads = 13.5 KB in
// flight per warp), then the 12 gathers of x, and only then the arithmetic.
__device__ __forceinline__ double ldg_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// coherent (L2) load: for vectors that OTHER CTAs of the same launch have written before a grid-wide barrier -- the
// non-coherent path above may only be used for data that is read-only for the whole kernel
__device__ __forceinline__ double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int MODE>   // 0: y += K x ; 1: r += sum V B^T W CT (eps_rhs - B x0).  Returns V eps:sigma.
__device__ __forceinline__ double ebe_cell_batched(const sic_problem_t& P, int i, const double* __restrict__ x,
                                                   double* __restrict__ y, double* f_out = nullptr) {
  const size_t ns = (size_t)P.cell_stride;
  int node[4];
  double g[12], CT[36], er[6], ua[12];
#pragma unroll
  for (int a = 0; a < 4; ++a) node[a] = ldg_s32(P.conn + a * ns + i);
#pragma unroll
  for (int k = 0; k < 12; ++k) g[k] = ldg_f64(P.grad + k * ns + i);
  const double vol = ldg_f64(P.vol + i);
  const double* ct = P.CT + SIC_CT_INDEX(0, i);
#pragma unroll
  for (int k = 0; k < 36; ++k) CT[k] = ldg_f64(ct + k * SIC_TILE_CELLS);
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) er[k] = ldg_f64(P.eps_rhs + k * ns + i);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
#ifdef SIC_DBG_NOGATHER
    for (int j = 0; j < 3; ++j) ua[3 * a + j] = 1e-3 * (double)(node[a] & 7) + j;
#else
    for (int j = 0; j < 3; ++j) ua[3 * a + j] = ldg_f64(x + 3 * (size_t)node[a] + j);
#endif
  }
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
    for (int k = 0; k < 6; ++k) eps[k] = er[k] - eps[k];
  }
  double s[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += CT[r * 6 + k] * eps[k];
    s[r] = acc;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = g[3 * a], gy = g[3 * a + 1], gz = g[3 * a + 2];
    const double fx = vol * (s[0] * gx + s[3] * gy + s[4] * gz);
    const double fy = vol * (s[3] * gx + s[1] * gy + s[5] * gz);
    const double fz = vol * (s[4] * gx + s[5] * gy + s[2] * gz);
    if (f_out) {                       // staged scatter: the caller sums per unique node in shared memory
      f_out[3 * a + 0] = fx; f_out[3 * a + 1] = fy; f_out[3 * a + 2] = fz;
    } else {
      double* ya = y + 3 * (size_t)node[a];
      atomicAdd(ya + 0, fx); atomicAdd(ya + 1, fy); atomicAdd(ya + 2, fz);
    }
  }
  return vol * ((eps[0] * s[0] + eps[1] * s[1] + eps[2] * s[2]) + 2.0 * (eps[3] * s[3] + eps[4] * s[4] + eps[5] * s[5]));
}

// ---- operator, version 4: version 3 + shared-memory staged scatter ------------------------------
// A/B measurement on 918k cells (scripts/apply_microbench.py): the streaming part of version 3 runs at
// 6.3 TB/s (96 % of the measured HBM peak) when the 12 global FP64 atomics per cell are removed, and at
// 3.9 TB/s with them: the scatter, not the loads, bounds the kernel.  Here a CTA handles one tile of 128
// cells, parks the 12 nodal forces of each cell in shared memory, and a second phase sums them per UNIQUE
// node of the tile (precomputed plan, sic_problem_t.tile_*), issuing one plain store per tile-interior
// node component and one atomic per shared node component: ~5-6x fewer global RMWs, and none contended
// inside the tile.  blockDim.x must be SIC_TILE_CELLS; f_s is double[12][128] in shared memory.
struct TileGather {                       // + 12 KB for the compressed operator: x of the tile's unique nodes
  double xs[3 * 4 * SIC_TILE_CELLS];
};
struct TileScratch {                      // shared memory of one operator CTA (17.4 KB)
  double f[12][SIC_TILE_CELLS];           // nodal forces of the tile's cells
  unsigned short ent[4 * SIC_TILE_CELLS]; // the tile's (cell,slot) references, grouped by unique node
  int nodes[4 * SIC_TILE_CELLS];          // unique node ids (tile-interior first)
  int eoff[4 * SIC_TILE_CELLS + 1];       // per unique node: offset into ent
};

// Compressed operator of the multigrid PRECONDITIONER (PC = true): S = sym(W C_T), W = diag(1,1,1,2,2,2) (the form in
// which the tangent of tensorial strains is symmetric, see k_mg_ct_compress), as 21 floats per cell
// (tiled like C_T: [tile][21][128]) and the gradients + volume as 13 floats (tiled the same way: [tile][13][128]): 152 B per cell
// instead of 408.  The arithmetic stays FP64 (the vectors are).  A preconditioner only has to approximate K: the
// non-symmetry of the finite-difference tangent is round-off (1e-6 relative) and float storage perturbs the entries by
// 6e-8, neither of which moves the Krylov iteration count; the OUTER operator of the solve is always the exact one.
#ifndef SIC_EBE_GEOM_TILES
#define SIC_EBE_GEOM_TILES 0      /* 1: the exact tile kernels read the tiled copy of the geometry (sic_problem_t.geom_tiles);
                                     measured no faster than the SoA rows (4.93 vs 4.78 ms at 58.8 M cells, profiles/r2_ab4_*) */
#endif
#ifndef SIC_PC_F32MATH
#define SIC_PC_F32MATH 1
#endif
#define SIC_PC_CT_ROWS 21
#define SIC_PC_GEOM_ROWS 13
#define SIC_PC_CT_INDEX(e, i) (((size_t)((i) / SIC_TILE_CELLS) * SIC_PC_CT_ROWS + (e)) * SIC_TILE_CELLS + ((i) % SIC_TILE_CELLS))
__host__ __device__ constexpr int sic_sym_index(int r, int k) {      // upper triangle, row-major
  return (r <= k) ? (r * 6 - (r * (r - 1)) / 2 + (k - r)) : (k * 6 - (k * (k - 1)) / 2 + (r - k));
}

// XCOH: gather x through the coherent path (x was written earlier in the SAME launch, k_mg_coarse_fused).
// LIDX (compressed operator): x is gathered ONCE PER UNIQUE NODE of the tile into shared memory (xg) and every cell
// picks its four nodes there through pc_lidx ([4][cell_stride] uint16: index of the cell's node in the tile's list of
// unique nodes, sic_problem_t.tile_nodes) -- ncu showed the per-cell gathers (12 requests x 32 sectors per warp, 8 of
// every 32 bytes used) keeping L1TEX at 77 % while DRAM idled at 47 %; the connectivity is not read at all.
template <int MODE, bool XCOH = false, bool PC = false, bool LIDX = false>
__device__ __forceinline__ double ebe_tile_scatter(const sic_problem_t& P, const double* __restrict__ x,
                                                   double* __restrict__ y, TileScratch& sc,
                                                   const int* done_flag = nullptr, const float* __restrict__ pc_ct = nullptr,
                                                   const float* __restrict__ pc_geom = nullptr,
                                                   const uint16_t* __restrict__ pc_lidx = nullptr, TileGather* xg = nullptr,
                                                   int tile_of_block = -1) {
  static_assert(!(PC && MODE != 0), "the compressed operator only applies K");
  static_assert(!LIDX || (PC && !XCOH), "the unique-node gather belongs to the compressed operator");
  const int tile = tile_of_block >= 0 ? tile_of_block : (int)blockIdx.x, tid = threadIdx.x;
  const int i = tile * SIC_TILE_CELLS + tid;
  const size_t ns = (size_t)P.cell_stride;
  // ---- phase 0: every load of the tile is issued up front (asm volatile keeps program order) ----------
  int q0, q1, nint, done = 0;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(q0) : "l"(P.tile_ptr + tile));
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(q1) : "l"(P.tile_ptr + tile + 1));
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(nint) : "l"(P.tile_nint + tile));
  if (done_flag) asm volatile("ld.global.s32 %0, [%1];" : "=r"(done) : "l"(done_flag));
  unsigned ent_lo, ent_hi;   // this thread's 4 of the tile's 512 scatter references (8 bytes)
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(ent_lo), "=r"(ent_hi)
               : "l"(P.ent + (size_t)tile * 4 * SIC_TILE_CELLS + 4 * tid));
  // PC: the float operands stay floats in registers and are widened where they enter the FP64 arithmetic
  using store_t = typename std::conditional<PC, float, double>::type;
  int node[4];
  store_t g[12], CT[PC ? SIC_PC_CT_ROWS : 36], vol;
  double er[6], ua[12];
  if constexpr (LIDX) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      unsigned short v;
      asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(pc_lidx + a * ns + i));
      node[a] = (int)v;
    }
  } else if (SIC_EBE_GEOM_TILES && !PC) {
    // connectivity, gradients and volume of the tile from ONE contiguous 15 KB block (sic_geom_tile_t) instead of 17
    // rows that lie cell_stride apart: constant offsets, one DRAM page
    const int* ct_conn = &P.geom_tiles[tile].conn[0][0];
#pragma unroll
    for (int a = 0; a < 4; ++a) node[a] = ldg_s32(ct_conn + a * SIC_TILE_CELLS + tid);
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) node[a] = ldg_s32(P.conn + a * ns + i);
  }
  if constexpr (PC) {
    const float* gm = pc_geom + (size_t)tile * (SIC_PC_GEOM_ROWS * SIC_TILE_CELLS) + tid;   // tiled like pc_ct: [tile][13][128]
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = ldg_f32(gm + k * SIC_TILE_CELLS);
    vol = ldg_f32(gm + 12 * SIC_TILE_CELLS);
    const float* ct = pc_ct + SIC_PC_CT_INDEX(0, i);
#pragma unroll
    for (int k = 0; k < SIC_PC_CT_ROWS; ++k) CT[k] = ldg_f32(ct + k * SIC_TILE_CELLS);
  } else {
    if (SIC_EBE_GEOM_TILES) {
      const double* gt = &P.geom_tiles[tile].grad[0][0] + tid;
#pragma unroll
      for (int k = 0; k < 12; ++k) g[k] = ldg_f64(gt + k * SIC_TILE_CELLS);
      vol = ldg_f64(gt + 12 * SIC_TILE_CELLS);
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) g[k] = ldg_f64(P.grad + k * ns + i);
      vol = ldg_f64(P.vol + i);
    }
    const double* ct = P.CT + SIC_CT_INDEX(0, i);
#pragma unroll
    for (int k = 0; k < 36; ++k) CT[k] = ldg_f64(ct + k * SIC_TILE_CELLS);
    if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 6; ++k) er[k] = ldg_f64(P.eps_rhs + k * ns + i);
    }
  }
  if constexpr (!LIDX) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        ua[3 * a + j] = XCOH ? ldcg_f64(x + 3 * (size_t)node[a] + j) : ldg_f64(x + 3 * (size_t)node[a] + j);
    }
  }
  // scatter plan of the tile -> shared memory (depends only on q0, requested first)
  const int nq = q1 - q0, ebase = tile * 4 * SIC_TILE_CELLS;
  for (int k = tid; k < nq; k += SIC_TILE_CELLS) {
    int nd, eo;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(nd) : "l"(P.tile_nodes + q0 + k));
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(eo) : "l"(P.ent_ptr + q0 + k));
    sc.nodes[k] = nd;
    sc.eoff[k] = eo - ebase;
    if constexpr (LIDX) {          // x of unique node k, once for the whole tile
      const double* xn = x + 3 * (size_t)nd;
      const double x0 = ldg_f64(xn), x1 = ldg_f64(xn + 1), x2 = ldg_f64(xn + 2);
      xg->xs[3 * k] = x0; xg->xs[3 * k + 1] = x1; xg->xs[3 * k + 2] = x2;
    }
  }
  if (tid == 0) sc.eoff[nq] = 4 * SIC_TILE_CELLS;
  reinterpret_cast<uint2*>(sc.ent)[tid] = make_uint2(ent_lo, ent_hi);
  if constexpr (LIDX) {
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int j = 0; j < 3; ++j) ua[3 * a + j] = xg->xs[3 * node[a] + j];
    }
  }
  // ---- phase 1: the cell's arithmetic (cells of the padding compute zeros: their C_T, grad, vol are 0) --
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
    for (int k = 0; k < 6; ++k) eps[k] = er[k] - eps[k];
  }
  // F32 (compressed operator, SIC_PC_F32MATH): the strain is formed in FP64 (differences of displacements), the stress, the
  // nodal forces, their staging in shared memory and their per-node sums in FP32 -- the operands are floats already and
  // the preconditioner tolerates 1e-7 relative; full-rate FFMA instead of half-rate DFMA, half the shared-memory traffic.
  constexpr bool F32 = PC && (SIC_PC_F32MATH != 0);
  float (*ff)[SIC_TILE_CELLS] = reinterpret_cast<float (*)[SIC_TILE_CELLS]>(&sc.f[0][0]);
  double energy = 0.0;
  if constexpr (F32) {
    float ef[6], sf[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ef[k] = (float)eps[k];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 6; ++k) acc += (float)CT[sic_sym_index(r, k)] * ef[k];
      sf[r] = (r >= 3) ? 0.5f * acc : acc;
    }
    const float vf = (float)vol;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float gx = (float)g[3 * a], gy = (float)g[3 * a + 1], gz = (float)g[3 * a + 2];
      ff[3 * a + 0][tid] = vf * (sf[0] * gx + sf[3] * gy + sf[4] * gz);
      ff[3 * a + 1][tid] = vf * (sf[3] * gx + sf[1] * gy + sf[5] * gz);
      ff[3 * a + 2][tid] = vf * (sf[4] * gx + sf[5] * gy + sf[2] * gz);
    }
  } else {
    double s[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) acc += CT[PC ? sic_sym_index(r, k) : r * 6 + k] * eps[k];
      s[r] = (PC && r >= 3) ? 0.5 * acc : acc;       // PC: CT holds S = sym(W C_T), sigma = W^-1 S eps
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double gx = g[3 * a], gy = g[3 * a + 1], gz = g[3 * a + 2];
      sc.f[3 * a + 0][tid] = vol * (s[0] * gx + s[3] * gy + s[4] * gz);
      sc.f[3 * a + 1][tid] = vol * (s[3] * gx + s[1] * gy + s[5] * gz);
      sc.f[3 * a + 2][tid] = vol * (s[4] * gx + s[5] * gy + s[2] * gz);
    }
    energy = vol * ((eps[0] * s[0] + eps[1] * s[1] + eps[2] * s[2]) + 2.0 * (eps[3] * s[3] + eps[4] * s[4] + eps[5] * s[5]));
  }
  __syncthreads();
  if (done) return 0.0;   // uniform over the grid: nothing is written once the solve has converged
  // ---- phase 2: one global write per unique node of the tile, everything read from shared memory ------
  for (int k = tid; k < nq; k += SIC_TILE_CELLS) {
    const int e0 = sc.eoff[k], e1 = sc.eoff[k + 1];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    if constexpr (F32) {
      float fx = 0.0f, fy = 0.0f, fz = 0.0f;
      for (int e = e0; e < e1; ++e) {
        const int c = (int)sc.ent[e];
        const int cell = c >> 2, slot = c & 3;
        fx += ff[3 * slot + 0][cell];
        fy += ff[3 * slot + 1][cell];
        fz += ff[3 * slot + 2][cell];
      }
      sx = fx; sy = fy; sz = fz;
    } else {
      for (int e = e0; e < e1; ++e) {
        const int c = (int)sc.ent[e];
        const int cell = c >> 2, slot = c & 3;
        sx += sc.f[3 * slot + 0][cell];
        sy += sc.f[3 * slot + 1][cell];
        sz += sc.f[3 * slot + 2][cell];
      }
    }
    double* yn = y + 3 * (size_t)sc.nodes[k];
    if (k < nint) { yn[0] = sx; yn[1] = sy; yn[2] = sz; }                 // only this tile touches the node
    else { atomicAdd(yn + 0, sx); atomicAdd(yn + 1, sy); atomicAdd(yn + 2, sz); }
  }
  return energy;
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

// Block partial sums only (no ticket): the operator kernel runs ~50 waves of small CTAs, and making each
// CTA wait for an atomic round trip (grid_reduce's ticket) held its SM slot ~1.5 us per wave.  The partials
// are summed by k_sum_partials below, in block order (deterministic).
template <int NV, int THREADS>
__device__ __forceinline__ void block_partials(double v[NV], double* __restrict__ partials) {
  __shared__ double shp[NV][THREADS / 32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (l == 0) shp[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < THREADS / 32; ++j) s += shp[k][j];
      partials[(size_t)blockIdx.x * NV + k] = s;
    }
  }
}

}  // namespace sic
#endif
