// fem.cu — hot-path part (2): matrix-free tangent operator, RHS, block-Jacobi blocks, Neumann loads.
#include "ebe_tma.cuh"

namespace sic {

// y += K x  (MODE 0)   or   r += sum_e V B^T W CT (eps_rhs - B x0)  (MODE 1); TMA-staged tiles (ebe_tma.cuh)
#ifndef SIC_EBE_IMPL
#define SIC_EBE_IMPL 3
#endif
#if SIC_EBE_IMPL == 3
template <int MODE>
#ifndef SIC_EBE_MINBLOCKS
#define SIC_EBE_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(128, SIC_EBE_MINBLOCKS) k_ebe(sic_problem_t P, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ TileScratch sc;
  ebe_tile_scatter<MODE>(P, x, y, sc);
}
#define SIC_EBE_LAUNCH(MODE, p, xin, yout, st) \
  k_ebe<MODE><<<blocks_for((p)->n_cells, 128), 128, 0, st>>>(*(p), xin, yout)
#else
template <int MODE>
__global__ void __launch_bounds__(SIC_TILE, 2) k_ebe(sic_problem_t P, const double* __restrict__ x,
                                                    double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<MODE>& sm = *reinterpret_cast<TileSmem<MODE>*>(smem_raw);
  ebe_tiles<MODE>(P, x, y, sm);
}
#define SIC_EBE_LAUNCH(MODE, p, xin, yout, st)                                                     \
  do {                                                                                             \
    ebe_allow_smem(k_ebe<MODE>, sizeof(TileSmem<MODE>));                                           \
    k_ebe<MODE><<<ebe_grid((p)->n_cells), SIC_TILE, sizeof(TileSmem<MODE>), st>>>(*(p), xin, yout); \
  } while (0)
#endif

__global__ void k_mask_copy(int n, double* __restrict__ y, const double* __restrict__ x,
                            const uint8_t* __restrict__ fixed) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n && fixed[d]) y[d] = x ? x[d] : 0.0;
}

// r = fixed ? 0 : r + b
__global__ void k_add_mask(int n, double* __restrict__ r, const double* __restrict__ b,
                           const uint8_t* __restrict__ fixed) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n) r[d] = (fixed && fixed[d]) ? 0.0 : r[d] + b[d];
}

// nodal 3x3 diagonal blocks of K: column l of K_aa is the force on node a caused by a unit
// displacement of node a along l.
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_diag_blocks(sic_problem_t P, double* __restrict__ dblk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  CellGeom c;
  load_geom(P, i, c);
  double CT[36];
#pragma unroll
  for (int j = 0; j < 36; ++j) CT[j] = __ldg(P.CT + SIC_CT_INDEX(j, i));
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = c.g[3 * a], gy = c.g[3 * a + 1], gz = c.g[3 * a + 2];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      double e[6] = {0, 0, 0, 0, 0, 0};
      if (l == 0) { e[0] = gx; e[3] = 0.5 * gy; e[4] = 0.5 * gz; }
      if (l == 1) { e[1] = gy; e[3] = 0.5 * gx; e[5] = 0.5 * gz; }
      if (l == 2) { e[2] = gz; e[4] = 0.5 * gx; e[5] = 0.5 * gy; }
      double s[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) t += CT[r * 6 + k] * e[k];
        s[r] = t;
      }
      const double fx = c.vol * (s[0] * gx + s[3] * gy + s[4] * gz);
      const double fy = c.vol * (s[3] * gx + s[1] * gy + s[5] * gz);
      const double fz = c.vol * (s[4] * gx + s[5] * gy + s[2] * gz);
      double* blk = dblk + 9 * (size_t)c.node[a];
      atomicAdd(blk + 0 * 3 + l, fx);
      atomicAdd(blk + 1 * 3 + l, fy);
      atomicAdd(blk + 2 * 3 + l, fz);
    }
  }
}

// fixed dofs are decoupled (unit diagonal, zero row/column as assemble_matrix(bcs) does,
// MomentumEquation.py:1010), then the 3x3 block is inverted in place.
__global__ void k_invert_blocks(int n_nodes, double* __restrict__ dblk, const uint8_t* __restrict__ fixed) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  double a[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = dblk[9 * (size_t)n + k];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (fixed[3 * (size_t)n + j]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { a[3 * j + k] = 0.0; a[3 * k + j] = 0.0; }
      a[3 * j + j] = 1.0;
    }
  }
  const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  double b[9];
  b[0] = c00 * id; b[1] = (a[2] * a[7] - a[1] * a[8]) * id; b[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  b[3] = c01 * id; b[4] = (a[0] * a[8] - a[2] * a[6]) * id; b[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  b[6] = c02 * id; b[7] = (a[1] * a[6] - a[0] * a[7]) * id; b[8] = (a[0] * a[4] - a[1] * a[3]) * id;
#pragma unroll
  for (int k = 0; k < 9; ++k) dblk[9 * (size_t)n + k] = b[k];
}

// Neumann load (MomentumBC.py:247-277): f(x) = p + rho g (H - x_dir) is linear on the facet, the
// integral of f * phi_a over the triangle is A (f_a/6 + (f_b + f_c)/12); traction = f * n.
__global__ void k_neumann(int n_tri, const int32_t* __restrict__ tri, const double* __restrict__ area_n,
                          const int32_t* __restrict__ bc_of_tri, const double* __restrict__ coords,
                          const double* __restrict__ bc_par, double* __restrict__ b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tri) return;
  const int bc = bc_of_tri[t];
  if (bc < 0) return;
  const double pval = bc_par[4 * bc], rho_g = bc_par[4 * bc + 1], H = bc_par[4 * bc + 2];
  const int dir = (int)bc_par[4 * bc + 3];
  int nd[3];
  double fv[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    nd[a] = tri[(size_t)a * n_tri + t];
    fv[a] = pval + rho_g * (H - coords[3 * (size_t)nd[a] + dir]);
  }
  const double nx = area_n[t], ny = area_n[(size_t)n_tri + t], nz = area_n[2 * (size_t)n_tri + t];
  const double tot = fv[0] + fv[1] + fv[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double w = (tot + fv[a]) / 12.0;  // f_a/6 + (f_b+f_c)/12
    atomicAdd(b + 3 * (size_t)nd[a] + 0, w * nx);
    atomicAdd(b + 3 * (size_t)nd[a] + 1, w * ny);
    atomicAdd(b + 3 * (size_t)nd[a] + 2, w * nz);
  }
}

template __global__ void k_ebe<0>(sic_problem_t, const double*, double*);
template __global__ void k_ebe<1>(sic_problem_t, const double*, double*);

}  // namespace sic

using namespace sic;

static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

extern "C" int sic_apply(const sic_problem_t* p, const double* x, double* y, const uint8_t* fixed, void* stream) {
  if (!p || !x || !y) return sic_fail("sic_apply: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nd = 3 * p->n_nodes;
  if (int rc = sic_check_cuda(cudaMemsetAsync(y, 0, sizeof(double) * nd, st), "memset y")) return rc;
  if (p->n_cells > 0) SIC_EBE_LAUNCH(0, p, x, y, st);
  if (fixed && nd > 0) k_mask_copy<<<blocks_for(nd, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(nd, y, x, fixed);
  return sic_check_launch("sic_apply");
}

extern "C" int sic_residual0(const sic_problem_t* p, const double* b_ext, const double* x0, double* r,
                             const uint8_t* fixed, const sic_halo_t* halo, void* stream) {
  if (!p || !b_ext || !x0 || !r) return sic_fail("sic_residual0: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nd = 3 * p->n_nodes;
  // element part first (partial on interface nodes -> summed over ranks), then the consistent b_ext
  if (int rc = sic_check_cuda(cudaMemsetAsync(r, 0, sizeof(double) * nd, st), "memset r")) return rc;
  if (p->n_cells > 0) SIC_EBE_LAUNCH(1, p, x0, r, st);
  if (int rc = sic_check_launch("k_ebe<1>")) return rc;
  if (int rc = sic_exchange(halo, r, 3, nullptr, 0, stream)) return rc;
  if (nd > 0) k_add_mask<<<blocks_for(nd, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(nd, r, b_ext, fixed);
  return sic_check_launch("sic_residual0");
}

extern "C" int sic_block_jacobi(const sic_problem_t* p, double* dinv, const uint8_t* fixed, const sic_halo_t* halo,
                                void* stream) {
  if (!p || !dinv || !fixed) return sic_fail("sic_block_jacobi: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = sic_check_cuda(cudaMemsetAsync(dinv, 0, sizeof(double) * 9 * p->n_nodes, st), "memset dinv")) return rc;
  if (p->n_cells > 0) k_diag_blocks<<<blocks_for(p->n_cells, SIC_EBE_THREADS), SIC_EBE_THREADS, 0, st>>>(*p, dinv);
  if (int rc = sic_check_launch("k_diag_blocks")) return rc;
  if (int rc = sic_exchange(halo, dinv, 9, nullptr, 0, stream)) return rc;
  if (p->n_nodes > 0)
    k_invert_blocks<<<blocks_for(p->n_nodes, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(p->n_nodes, dinv, fixed);
  return sic_check_launch("sic_block_jacobi");
}

extern "C" int sic_neumann(int n_tri, const int32_t* tri, const double* area_n, const int32_t* bc_of_tri,
                           const double* coords, int n_bc, const double* bc_par, double* b, void* stream) {
  if (n_tri == 0 || n_bc == 0) return 0;
  if (!tri || !area_n || !bc_of_tri || !coords || !bc_par || !b) return sic_fail("sic_neumann: null argument");
  k_neumann<<<blocks_for(n_tri, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, (cudaStream_t)stream>>>(
      n_tri, tri, area_n, bc_of_tri, coords, bc_par, b);
  return sic_check_launch("k_neumann");
}
