// mg.cu — hot-path part (3b): geometric multigrid preconditioned CG on the nested red-refinement hierarchy.
//
// Replaces the PETSc preconditioner of `self.solver.solve(b, X)` (MomentumEquation.py:1023-1025; the
// reference's examples use `asm`, `gamg` once: nobian/run_interlayer.py:2114-2116).  Block-Jacobi CG
// (solver.cu) needs 300 / 660 / 1400 / 2800 iterations on the 14k / 115k / 918k / 7.3M-cell levels of
// cavern_regular; one V(2,2) cycle per CG iteration keeps the count at ~25 on every level for ~6 operator
// applications per iteration (scripts/mg_prototype.py).
//
// B200 design: every level is a sic_problem_t of its own and runs the SAME operator kernel
// (ebe_tile_scatter, fem.cuh).  Coarse operators are Galerkin by construction: on nested P1 tets with
// equal-volume children, P^T K P is the operator evaluated with the mean of the eight children's C_T, so
// "coarsening" is one streaming pass over C_T per tangent (k_ct_coarsen).  The smoother is a Chebyshev
// polynomial in (block-Jacobi)^-1 K: no dot products, hence no grid-wide reduction and no host
// synchronisation anywhere inside the cycle; its coefficients are kernel arguments computed on the host
// from lambda_max (a few power iterations per level per tangent).  Transfers are gather-only
// (deterministic, no atomics): restriction walks a CSR of the fine nodes each coarse node receives from,
// prolongation reads the two parents of each fine node.  Dirichlet dofs are kept at zero on every level.
// The outer CG keeps its scalars on the device (as solver.cu does) and the host looks at them every
// `check_every` iterations; every kernel of the cycle is a no-op once the `done` flag is up.
//
// Several GPUs (levels[l].halo != NULL): the levels from some level `lc` up to the finest are partitioned by cells
// exactly as in solver.cu (interface nodes duplicated, vectors kept consistent, operator results completed by
// sic_exchange over NVLink, dot products with owner weights + scalar exchange), NESTED: a cell lives on the rank of
// its ancestor on level lc, so children, prolongation parents and the Galerkin coarsening of C_T are rank-local
// between two partitioned levels, and the restriction (summed by the owners of the fine nodes) is completed by one
// halo sum on the coarse level.  The levels below lc are replicated on every rank: the transfers of level lc index
// level lc-1 GLOBALLY, its right-hand side is completed by one all-reduce per cycle and its C_T by one all-reduce per
// tangent (small: the host keeps lc-1 at <= a few 100k cells).
#include <math.h>
#include <string.h>

#include "fem.cuh"
#ifndef SIC_HOSTEMU
#include "comm.cuh"
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#endif

namespace sic {

struct MgScal {
  double rz, pq, rr, rr0, rr_ref, alpha, beta, tol2;
  double pw;            // power iteration: ||Dinv K v||^2 with ||v|| = 1
  double sum[2];        // this rank's partial sum of the running reduction (summed over ranks by sic_exchange)
  int done, iters, nanflag, reason;
};
static_assert(sizeof(MgScal) <= 64 * sizeof(double), "MgScal must fit the reserved workspace header");

#define SIC_MG_HEADER 64    /* doubles reserved for MgScal */
#define SIC_MG_COUNTERS 8   /* doubles reserved for the ticket counters of grid_reduce */

enum { MG_OP_REF = 0, MG_OP_INIT_RR, MG_OP_INIT_RZ, MG_OP_RR, MG_OP_RZ, MG_OP_PW, MG_OP_PQ };

struct MgFin {   // what the last block of a reducing kernel does with the grid total
  MgScal* S; int op; double rtol, atol; int guess; int multi;
  // one GPU: run the scalar recurrence at once; several: park the partial sum, k_mg_scal runs it after the exchange
  __device__ __forceinline__ void run(double tot) const {
    if (multi) { S->sum[0] = tot; return; }
    step(tot);
  }
  __device__ __forceinline__ void step(double tot) const {
    switch (op) {
      case MG_OP_REF: S->rr_ref = tot; break;
      case MG_OP_INIT_RR: {
        S->rr = tot; S->rr0 = tot;
        const double ref = guess ? S->rr_ref : tot;
        const double t = rtol * rtol * ref, a2 = atol * atol;
        S->tol2 = (t > a2) ? t : a2;
        S->iters = 0; S->nanflag = 0; S->reason = 0; S->done = 0;
        if (!(tot == tot) || isinf(tot)) { S->nanflag = 1; S->done = 1; S->reason = -9; }
        else if (tot <= S->tol2 || tot == 0.0) { S->done = 1; S->reason = (tot <= a2) ? 3 : 2; }
        break;
      }
      case MG_OP_INIT_RZ: S->rz = tot; S->beta = 0.0; break;
      case MG_OP_RR:
        S->rr = tot; S->iters += 1;
        if (!(tot == tot) || isinf(tot)) { S->nanflag = 1; S->done = 1; S->reason = -9; }
        else if (tot <= S->tol2) { S->done = 1; }
        break;
      case MG_OP_RZ: S->beta = tot / S->rz; S->rz = tot; break;
      case MG_OP_PW: S->pw = tot; break;
      case MG_OP_PQ: S->pq = tot; S->alpha = S->rz / tot; break;
    }
  }
};

// several GPUs: the scalar recurrence after the partial sums have been added over the ranks
__global__ void k_mg_scal(MgFin fin, int skip_if_done) {
  if (skip_if_done && fin.S->done) return;
  fin.step(fin.S->sum[0]);
}

// ---- operator (the kernel of fem.cuh) ---------------------------------------------------------------
__global__ void __launch_bounds__(SIC_TILE_CELLS, 3) k_mg_ebe(sic_problem_t P, const double* __restrict__ x,
                                                            double* __restrict__ y, const int* done) {
  __shared__ TileScratch sc;
  ebe_tile_scatter<0>(P, x, y, sc, done);
}

// the compressed operator of the preconditioner (fem.cuh, PC = true): symmetric float C_T, float geometry
// SIC_PC_LIDX=1 (variant, measured SLOWER: 0.304 vs 0.272 ms per launch at 7.35 M cells, profiles/r2_ab2_*): gather x once
// per unique node of the tile into shared memory through 16-bit tile-local node indices instead of once per cell through
// the connectivity.  It takes L1TEX from 77 % to ~35 % and the DRAM bytes from 152 to 144 per cell, but the extra barrier
// and shared-memory hop lengthen a CTA's dependent chain, and latency x resident warps is what bounds this kernel.
#ifndef SIC_PC_LIDX
#define SIC_PC_LIDX 0
#endif
#ifndef SIC_PC_MINBLOCKS
#define SIC_PC_MINBLOCKS 6      /* resident CTAs per SM it is compiled for (latency-bound at 4: ncu, profiles/) */
#endif
__global__ void __launch_bounds__(SIC_TILE_CELLS, SIC_PC_MINBLOCKS) k_mg_ebe_pc(sic_problem_t P, const float* __restrict__ pc_ct,
                                                               const float* __restrict__ pc_geom,
                                                               const uint16_t* __restrict__ pc_lidx,
                                                               const double* __restrict__ x, double* __restrict__ y,
                                                               const int* done) {
  __shared__ TileScratch sc;
#if SIC_PC_LIDX
  __shared__ TileGather xg;
  ebe_tile_scatter<0, false, true, true>(P, x, y, sc, done, pc_ct, pc_geom, pc_lidx, &xg);
#else
  (void)pc_lidx;
  ebe_tile_scatter<0, false, true, false>(P, x, y, sc, done, pc_ct, pc_geom);
#endif
}

#ifndef SIC_HOSTEMU
// Operator + halo exchange in ONE launch (several GPUs, P2P mailboxes).  Launch order: the tiles that touch an interface
// node (H.tile_order[0 .. n_iface)), then n_peers * bpp COMMUNICATION CTAs, then the interior tiles.  The communication
// CTAs wait until every interface tile of this launch has scattered its forces, then run the halo exchange of comm.cuh on
// y -- send the interface sums into the neighbours' mailboxes over NVLink, wait for theirs, add them -- while the interior
// tiles (which touch no interface node) are still streaming: the exchange latency hides behind them and one launch
// replaces two.  Everything it synchronises on lives in device memory (comm.cuh: epochs[3] counts these launches,
// epochs[4] the interface tiles that have finished), so the launch replays from the CUDA graph like the others.
__global__ void __launch_bounds__(SIC_TILE_CELLS, SIC_PC_MINBLOCKS) k_mg_ebe_pc_x(sic_problem_t P, const float* __restrict__ pc_ct,
                                                                 const float* __restrict__ pc_geom, sic_halo_t H, P2P ctx,
                                                                 const double* __restrict__ x, double* __restrict__ y,
                                                                 const int* done) {
  __shared__ TileScratch sc;
  const int n_iface = H.n_iface_tiles, n_comm = H.n_peers * ctx.bpp, b = (int)blockIdx.x;
  if (b >= n_iface && b < n_iface + n_comm) {
    // ---- communication CTA -----------------------------------------------------------------------------------
    const unsigned long long epoch = ((volatile unsigned long long*)ctx.epochs)[0];
    const unsigned long long fused = ((volatile unsigned long long*)ctx.epochs)[3];
    const long long timeout = (*(volatile int*)ctx.error) ? 0ll : SIC_P2P_TIMEOUT_CYCLES;
    if (threadIdx.x == 0) {
      const unsigned long long target = (fused + 1ull) * (unsigned long long)n_iface;
      const long long t0 = clock64();
      while (((volatile unsigned long long*)ctx.epochs)[4] < target) {
        if (clock64() - t0 > timeout) { atomicExch(ctx.error, 1); break; }
      }
      __threadfence();
    }
    __syncthreads();
    const int c = b - n_iface;
    p2p_halo_block(H, ctx, y, 3, c / ctx.bpp, c % ctx.bpp, epoch, timeout);
    __syncthreads();
    if (threadIdx.x == 0) {          // the last communication CTA advances the device-side epochs for the next launch
      const unsigned long long t = atomicAdd(ctx.epochs + 2, 1ull);
      if (t == (unsigned long long)n_comm - 1ull) {
        ctx.epochs[2] = 0ull;
        ctx.epochs[0] = epoch + 1ull;
        ctx.epochs[3] = fused + 1ull;
        __threadfence();
      }
    }
    return;
  }
  // ---- operator CTA ------------------------------------------------------------------------------------------
  const int slot = (b < n_iface) ? b : b - n_comm;
  ebe_tile_scatter<0, false, true, false>(P, x, y, sc, done, pc_ct, pc_geom, nullptr, nullptr, H.tile_order[slot]);
  if (b < n_iface) {                 // tell the communication CTAs that this interface tile's sums are in y
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(ctx.epochs + 4, 1ull);
  }
}
#endif

__global__ void k_mg_to_float(int n, const double* __restrict__ a, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (float)a[k];
}

// pc_ct = float(sym(W C_T)), W = diag(1,1,1,2,2,2), of one level (once per set-up; both tiled by 128 cells).
// C_T maps TENSORIAL strains to stresses, so the cell energy is eps^T W C_T eps and the operator K = B^T W C_T B is
// symmetric iff W C_T is (C_T itself is not: for the creep tangents C_T[normal][shear] = 2 C_T[shear][normal], the
// reference's doubled shear columns).  The compressed operator applies sigma = W^-1 S eps with S = sym(W C_T):
// exactly K when the tangent has major symmetry, its symmetric part otherwise.
__global__ void __launch_bounds__(128) k_mg_ct_compress(int n_cells, const double* __restrict__ CT, float* __restrict__ pc_ct) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int k = r; k < 6; ++k) {
      const double wr = (r < 3) ? 1.0 : 2.0, wk = (k < 3) ? 1.0 : 2.0;
      const double v = (r == k) ? wr * __ldg(CT + SIC_CT_INDEX(r * 6 + k, i))
                                : 0.5 * (wr * __ldg(CT + SIC_CT_INDEX(r * 6 + k, i)) + wk * __ldg(CT + SIC_CT_INDEX(k * 6 + r, i)));
      pc_ct[SIC_PC_CT_INDEX(sic_sym_index(r, k), i)] = (float)v;
    }
  }
}

// q = K p with the per-cell energies p.Kp summed per block (finished by k_mg_sum_pq)
#ifndef SIC_EBE_DOT_MINBLOCKS
#define SIC_EBE_DOT_MINBLOCKS 3      /* 4 (128 registers, 64 bytes of spills) measured the same: 4.825 vs 4.834 ms at 58.8 M cells */
#endif
__global__ void __launch_bounds__(SIC_TILE_CELLS, SIC_EBE_DOT_MINBLOCKS) k_mg_ebe_dot(sic_problem_t P, const double* __restrict__ x,
                                                                double* __restrict__ y, double* __restrict__ partials,
                                                                const int* done) {
  __shared__ TileScratch sc;
  double v[1] = {ebe_tile_scatter<0>(P, x, y, sc, done)};
  block_partials<1, SIC_TILE_CELLS>(v, partials);
}

__global__ void __launch_bounds__(1024) k_mg_sum_pq(const double* __restrict__ partials, int n, MgFin fin) {
  if (fin.S->done) return;
  __shared__ double sh[32];
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 1024) a += partials[k];
  a = warp_sum(a);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 32; ++k) tot += sh[k];
    fin.run(tot);       // MG_OP_PQ
  }
}

__device__ __forceinline__ void mg_precond3(const double* __restrict__ dinv, size_t n, const double r[3], double z[3]) {
  const double* d = dinv + 9 * n;
#pragma unroll
  for (int j = 0; j < 3; ++j) z[j] = __ldg(d + 3 * j) * r[0] + __ldg(d + 3 * j + 1) * r[1] + __ldg(d + 3 * j + 2) * r[2];
}

// ---- outer CG ---------------------------------------------------------------------------------------
// z = fixed ? x : 0
__global__ void k_mg_zero_free(int nd, double* __restrict__ z, const double* __restrict__ x,
                               const uint8_t* __restrict__ fixed) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < nd) z[d] = fixed[d] ? x[d] : 0.0;
}

// The two reducing vector kernels of the CG recurrence run a GRID-STRIDE loop on about one wave of blocks
// (mg_reduce_blocks): with one block per 256 DOFs every block paid the ticket's atomic round trip and two barriers of
// grid_reduce for a few hundred bytes of work (ncu: k_mg_dot 33 us for 62 MB, "barrier" the top stall).
// sum a.b over all dofs (w: owner weights on several GPUs, every node counted once; NULL on one)
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_dot(int nd, const double* __restrict__ a,
                                                           const double* __restrict__ b, const double* __restrict__ w,
                                                           MgFin fin, int skip_if_done,
                                                           double* __restrict__ partials, unsigned* counter) {
  if (skip_if_done && fin.S->done) return;
  double v[1] = {0.0};
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += gridDim.x * blockDim.x)
    v[0] += (w ? w[k / 3] : 1.0) * a[k] * b[k];
  grid_reduce<1, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot[0]); });
}

// x += alpha p ; r -= alpha q (fixed dofs: r = 0) ; rr = r.r -> convergence test
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_cg_update(int nd, double* __restrict__ x,
                                                                 double* __restrict__ r, const double* __restrict__ p,
                                                                 const double* __restrict__ q,
                                                                 const uint8_t* __restrict__ fixed,
                                                                 const double* __restrict__ w, MgFin fin,
                                                                 double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const double alpha = fin.S->alpha;
  double v[1] = {0.0};
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += gridDim.x * blockDim.x) {
    double rn = 0.0;
    if (!fixed[k]) {
      x[k] += alpha * p[k];
      rn = r[k] - alpha * q[k];
    }
    r[k] = rn;
    v[0] += (w ? w[k / 3] : 1.0) * rn * rn;
  }
  grid_reduce<1, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot[0]); });
}

// p = z + beta p (first: p = z ; fixed dofs: 0) ; q = 0
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_cg_p(int nd, double* __restrict__ p, const double* __restrict__ z,
                                                            double* __restrict__ q, const uint8_t* __restrict__ fixed,
                                                            const MgScal* S, int first) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const double v = first ? z[d] : z[d] + S->beta * p[d];
  p[d] = fixed[d] ? 0.0 : v;
  q[d] = 0.0;
}

// ---- Chebyshev smoother -------------------------------------------------------------------------------
// First step.  zero_guess: r = b, x = d = Dinv r / theta.  Otherwise t holds K x: r = b - t, d = Dinv r / theta,
// x += d.  Leaves t = 0 for the next operator application.
__device__ __forceinline__ void mg_cheb_first_node(int n, const double* __restrict__ b, double* r, double* d, double* x,
                                                   double* t, const double* __restrict__ dinv,
                                                   const uint8_t* __restrict__ fixed, double inv_theta, int zero_guess) {
  double rn[3], zn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t k = 3 * (size_t)n + j;
    double v = zero_guess ? b[k] : b[k] - t[k];
    if (fixed[k]) v = 0.0;
    rn[j] = v;
    r[k] = v;
    t[k] = 0.0;
  }
  mg_precond3(dinv, n, rn, zn);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t k = 3 * (size_t)n + j;
    const double dn = fixed[k] ? 0.0 : zn[j] * inv_theta;
    d[k] = dn;
    x[k] = zero_guess ? dn : x[k] + dn;
  }
}
// The stand-alone smoother kernels run one thread per DOF (coalesced 8-byte accesses to r, d, x, t, b; the row of the
// node's 3x3 block is the three consecutive doubles at dinv[3 k]); the node's other two residual components come
// through shared memory.  SIC_DOF_THREADS is a multiple of 3, so a node never straddles two blocks.  Same expressions,
// in the same order, as the per-node functions the fused coarsest-level kernel uses.
#define SIC_DOF_THREADS 192
// DT: the blocks are read as doubles (exact preconditioner) or from the float copy sic_mg_setup keeps for the compressed
// one (pc_dinv: 12 instead of 24 of a step's 89 bytes per DOF)
template <class DT>
__global__ void __launch_bounds__(SIC_DOF_THREADS) k_mg_cheb_first(int nd, const double* __restrict__ b,
                                                                  double* __restrict__ r, double* __restrict__ d,
                                                                  double* __restrict__ x, double* __restrict__ t,
                                                                  const DT* __restrict__ dinv,
                                                                  const uint8_t* __restrict__ fixed, double inv_theta,
                                                                  int zero_guess, const int* done) {
  if (*done) return;
  __shared__ double rs[SIC_DOF_THREADS];
  const int k = blockIdx.x * SIC_DOF_THREADS + threadIdx.x;
  const bool in = k < nd;
  bool fx = true;
  double v = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0, xk = 0.0;
  if (in) {
    fx = fixed[k] != 0;
    v = zero_guess ? b[k] : b[k] - t[k];
    if (fx) v = 0.0;
    d0 = (double)__ldg(dinv + 3 * (size_t)k); d1 = (double)__ldg(dinv + 3 * (size_t)k + 1); d2 = (double)__ldg(dinv + 3 * (size_t)k + 2);
    if (!zero_guess) xk = x[k];
    r[k] = v;
    t[k] = 0.0;
  }
  rs[threadIdx.x] = v;
  __syncthreads();
  if (!in) return;
  const int base = threadIdx.x - (threadIdx.x % 3);
  const double z = d0 * rs[base] + d1 * rs[base + 1] + d2 * rs[base + 2];
  const double dn = fx ? 0.0 : z * inv_theta;
  d[k] = dn;
  x[k] = zero_guess ? dn : xk + dn;
}

// Step k >= 1.  t holds K d: r -= t ; d = a d + c Dinv r ; x += d ; t = 0.
// TCOH: t was accumulated by OTHER CTAs of the same launch (k_mg_coarse_fused): read it at L2, not through this SM's L1
template <bool TCOH = false>
__device__ __forceinline__ void mg_cheb_step_node(int n, double* r, double* d, double* x, double* t,
                                                  const double* __restrict__ dinv, const uint8_t* __restrict__ fixed,
                                                  double a, double c) {
  double rn[3], zn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t k = 3 * (size_t)n + j;
    const double tk = TCOH ? __ldcg(t + k) : t[k];
    const double v = fixed[k] ? 0.0 : r[k] - tk;
    rn[j] = v;
    r[k] = v;
    t[k] = 0.0;
  }
  mg_precond3(dinv, n, rn, zn);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t k = 3 * (size_t)n + j;
    const double dn = fixed[k] ? 0.0 : a * d[k] + c * zn[j];
    d[k] = dn;
    x[k] += dn;
  }
}
template <class DT>
__global__ void __launch_bounds__(SIC_DOF_THREADS) k_mg_cheb_step(int nd, double* __restrict__ r,
                                                                 double* __restrict__ d, double* __restrict__ x,
                                                                 double* __restrict__ t, const DT* __restrict__ dinv,
                                                                 const uint8_t* __restrict__ fixed, double a, double c,
                                                                 const int* done) {
  if (*done) return;
  __shared__ double rs[SIC_DOF_THREADS];
  const int k = blockIdx.x * SIC_DOF_THREADS + threadIdx.x;
  const bool in = k < nd;
  bool fx = true;
  double v = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0, dk = 0.0, xk = 0.0;
  if (in) {
    fx = fixed[k] != 0;
    v = fx ? 0.0 : r[k] - t[k];
    d0 = (double)__ldg(dinv + 3 * (size_t)k); d1 = (double)__ldg(dinv + 3 * (size_t)k + 1); d2 = (double)__ldg(dinv + 3 * (size_t)k + 2);
    dk = d[k];
    xk = x[k];
    r[k] = v;
    t[k] = 0.0;
  }
  rs[threadIdx.x] = v;
  __syncthreads();
  if (!in) return;
  const int base = threadIdx.x - (threadIdx.x % 3);
  const double z = d0 * rs[base] + d1 * rs[base + 1] + d2 * rs[base + 2];
  const double dn = fx ? 0.0 : a * dk + c * z;
  d[k] = dn;
  x[k] = xk + dn;
}

#ifndef SIC_HOSTEMU
// ---- coarsest level in ONE launch (sic_mg_opts_t.fused_coarse) ---------------------------------------------
// The coarsest level (the gmsh grid: 14 346 cells = 113 CTAs, < one wave of a B200) runs `coarse_its` Chebyshev steps per
// V-cycle; as separate launches every step is an operator kernel (~10 us) plus a vector kernel (~6 us), both pure
// latency: ~0.5 ms per cycle, 10-15 % of a Krylov iteration at 7.35 M cells and the largest part of one on 8 GPUs.
// Here all steps run in one cooperative launch (one CTA per tile, grid-wide barriers of cooperative groups between
// the operator and the vector phase); the phases are the SAME device functions as the kernels above
// (ebe_tile_scatter, mg_cheb_first_node, mg_cheb_step_node).  Vectors written in one phase and gathered in the next
// (d, t) are read through the coherent path (XCOH / plain loads): the read-only cache is not covered by grid.sync().
// Verified on a B200 against the oracle's V-cycle (tests/test_gpu_mg.py); the host emulation has no cooperative launch
// and always runs the launch-per-step sweep.
#define SIC_MG_MAX_COARSE_ITS 64
struct MgCoarseCoef { double a[SIC_MG_MAX_COARSE_ITS], c[SIC_MG_MAX_COARSE_ITS]; double inv_theta; int its; };

__global__ void __launch_bounds__(SIC_TILE_CELLS, 3) k_mg_coarse_fused(sic_problem_t P, const double* __restrict__ b, double* r,
                                                                     double* d, double* x, double* t,
                                                                     const double* __restrict__ dinv,
                                                                     const uint8_t* __restrict__ fixed, MgCoarseCoef cf,
                                                                     const int* done) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  if (*done) return;                                   // uniform over the grid: no barrier is reached by anybody
  __shared__ TileScratch sc;
  const int nn = P.n_nodes, gsz = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int n = gtid; n < nn; n += gsz) mg_cheb_first_node(n, b, r, d, x, t, dinv, fixed, cf.inv_theta, 1);
  grid.sync();
  for (int k = 1; k < cf.its; ++k) {
    ebe_tile_scatter<0, true>(P, d, t, sc, nullptr);   // t += K d (t is zero on entry: both node functions leave it so)
    grid.sync();
    for (int n = gtid; n < nn; n += gsz) mg_cheb_step_node<true>(n, r, d, x, t, dinv, fixed, cf.a[k], cf.c[k]);
    grid.sync();
  }
}
#endif

// r -= t (t = K d of the last smoothing step): the true residual of x ; t = 0
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_resid(int nd, double* __restrict__ r, double* __restrict__ t,
                                                             const uint8_t* __restrict__ fixed, const int* done) {
  if (*done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  r[k] = fixed[k] ? 0.0 : r[k] - t[k];
  t[k] = 0.0;
}

// ---- transfers ----------------------------------------------------------------------------------------
// b_c = P^T r_f: every coarse node sums the fine nodes that interpolate from it, weight 1/2 per entry (a fine
// node that IS the coarse node appears twice)
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_restrict(int n_coarse, const int32_t* __restrict__ ptr,
                                                                const int32_t* __restrict__ idx,
                                                                const double* __restrict__ r_f, double* __restrict__ b_c,
                                                                const uint8_t* __restrict__ fixed_c,
                                                                const double* __restrict__ w_f, const int* done) {
  if (*done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per coarse DOF: (node c, component j)
  if (k >= 3 * n_coarse) return;
  const int c = k / 3, j = k - 3 * c;
  double s = 0.0;
  const int e1 = __ldg(ptr + c + 1);
  for (int e = __ldg(ptr + c); e < e1; ++e) {
    const int fn = __ldg(idx + e);
    const double wn = w_f ? w_f[fn] : 1.0;      // several GPUs: a fine node is summed by its owner only
    s += wn * r_f[3 * (size_t)fn + j];
  }
  b_c[k] = fixed_c[k] ? 0.0 : 0.5 * s;
}

// x_f += P x_c
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_prolong_add(int n_fine, const int32_t* __restrict__ pa,
                                                                   const int32_t* __restrict__ pb,
                                                                   const double* __restrict__ x_c, double* __restrict__ x_f,
                                                                   const uint8_t* __restrict__ fixed_f, const int* done) {
  if (*done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per fine DOF
  if (k >= 3 * n_fine) return;
  const int n = k / 3, j = k - 3 * n;
  if (!fixed_f[k]) x_f[k] += 0.5 * (x_c[3 * (size_t)__ldg(pa + n) + j] + x_c[3 * (size_t)__ldg(pb + n) + j]);
}

// CT_c = mean of the eight children's CT_f (Galerkin coarse operator; both in the tiled SIC_CT_INDEX layout)
__global__ void __launch_bounds__(128) k_mg_ct_coarsen(int n_coarse, const int32_t* __restrict__ children,
                                                      const double* __restrict__ CT_f, double* __restrict__ CT_c) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_coarse) return;
  int ch[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ch[j] = __ldg(children + (size_t)j * n_coarse + c);
  for (int rc = 0; rc < 36; ++rc) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)      // a child held by another rank is -1 here: partial sums, all-reduced by the caller
      if (ch[j] >= 0) s += __ldg(CT_f + SIC_CT_INDEX(rc, ch[j]));
    CT_c[SIC_CT_INDEX(rc, c)] = 0.125 * s;
  }
}

// ---- power iteration for lambda_max(Dinv K) --------------------------------------------------------------
__global__ void k_mg_pw_init(int nd, double* __restrict__ v, double* __restrict__ t, const uint8_t* __restrict__ fixed,
                             double scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  unsigned h = (unsigned)k * 2654435761u;      // deterministic pseudo-random start vector, all frequencies
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  const double u = (double)(h & 0xffffu) / 65536.0 - 0.5;
  v[k] = fixed[k] ? 0.0 : u * scale;
  t[k] = 0.0;
}

// w = Dinv t (t = K v), fixed dofs 0, stored over t's companion vector w ; sums w.w
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_mg_pw_step(int n_nodes, const double* __restrict__ t,
                                                               double* __restrict__ w, const double* __restrict__ dinv,
                                                               const uint8_t* __restrict__ fixed,
                                                               const double* __restrict__ ow, MgFin fin,
                                                               double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (n < n_nodes) {
    const double own = ow ? ow[n] : 1.0;
    double tn[3], zn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) tn[j] = fixed[3 * (size_t)n + j] ? 0.0 : t[3 * (size_t)n + j];
    mg_precond3(dinv, n, tn, zn);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t k = 3 * (size_t)n + j;
      const double wn = fixed[k] ? 0.0 : zn[j];
      w[k] = wn;
      v[0] += own * wn * wn;
    }
  }
  grid_reduce<1, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot[0]); });
}

// v = w / ||w|| ; t = 0
__global__ void k_mg_pw_scale(int nd, double* __restrict__ v, const double* __restrict__ w, double* __restrict__ t,
                              const MgScal* S) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  const double nrm = sqrt(S->pw);
  v[k] = (nrm > 0.0) ? w[k] / nrm : 0.0;
  t[k] = 0.0;
}

}  // namespace sic

using namespace sic;

static inline int mg_blocks(int n, int t) { return (n + t - 1) / t; }

// blocks of a reducing grid-stride kernel: one wave of 8 resident 256-thread blocks per SM, or fewer for small vectors
static int mg_reduce_blocks(int n) {
  static int wave = 0;
  if (!wave) {
    int sms = 8;
#ifndef SIC_HOSTEMU
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 64;
#endif
    wave = 8 * sms;
  }
  const int b = mg_blocks(n, SIC_VEC_THREADS);
  return b < wave ? (b > 0 ? b : 1) : wave;
}

static int mg_check_levels(const sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o) {
  if (!lv || !o) return sic_fail("multigrid: null argument");
  if (n_levels < 1 || n_levels > SIC_MG_MAX_LEVELS) return sic_fail("multigrid: bad number of levels");
  if (o->nu < 1 || o->coarse_its < 1) return sic_fail("multigrid: nu and coarse_its must be >= 1");
  if (!(o->smooth_lo > 0.0 && o->smooth_lo < 1.0 && o->coarse_lo > 0.0 && o->coarse_lo < 1.0))
    return sic_fail("multigrid: smooth_lo / coarse_lo must lie in (0,1)");
  for (int l = 0; l < n_levels; ++l) {
    const sic_mg_level_t& L = lv[l];
    if (L.prob.abi_version != SIC_ABI_VERSION) return sic_fail("multigrid: sic_problem_t.abi_version mismatch");
    if (!L.fixed || !L.dinv || !L.x || !L.b || !L.r || !L.d || !L.t || !L.prob.CT)
      return sic_fail("multigrid: level with a null buffer");
    if ((L.pc_ct != nullptr) != (L.pc_geom != nullptr) || (L.pc_ct != nullptr) != (L.pc_lidx != nullptr))
      return sic_fail("multigrid: pc_ct, pc_geom and pc_lidx go together");
    const bool part = L.halo && L.halo->n_ranks > 1;
    if (part && l == 0) return sic_fail("multigrid: the coarsest level must be replicated (not partitioned)");
    if (part && l + 1 < n_levels && !(lv[l + 1].halo && lv[l + 1].halo->n_ranks > 1))
      return sic_fail("multigrid: every level above a partitioned level must be partitioned");
    if (part && (!L.halo->owner_w || !L.halo->comm)) return sic_fail("multigrid: halo without owner weights / communicator");
    if (part && n_levels < 2) return sic_fail("multigrid: a partitioned run needs at least two levels");
    if (l > 0) {
      if (!L.parent_a || !L.parent_b || !L.rst_ptr || !L.rst_idx || !L.children)
        return sic_fail("multigrid: level without transfer tables");
      const bool cpart = lv[l - 1].halo && lv[l - 1].halo->n_ranks > 1;
      // nested 1:8: on one GPU, and between two partitioned levels (a cell lives on its parent's rank)
      if ((!part || cpart) && L.prob.n_cells != 8 * lv[l - 1].prob.n_cells) return sic_fail("multigrid: levels are not nested 1:8");
    }
  }
  return 0;
}

static inline const sic_halo_t* mg_halo(const sic_mg_level_t& L) { return (L.halo && L.halo->n_ranks > 1) ? L.halo : nullptr; }

// The compressed operator and the halo exchange of its result as ONE launch (k_mg_ebe_pc_x): several GPUs with P2P
// mailboxes, a tile order in the halo plan, and the option switched on.  Returns false when the caller has to launch the
// operator and the exchange separately.
// OFF by default: measured on 2 x B200 (profiles/r2_fx*_n2_*.json) it is 1.5 % faster than two launches at 7.35 M cells
// (13.9 k interface nodes) and 5.6 % SLOWER at 58.8 M (55 k interface nodes: the exchange's gathers, NVLink stores and
// atomics compete with six streaming operator CTAs per SM, and without a common end of the exchange the ranks drift
// apart until the next blocking one).  Kept, verified by tests/test_gpu_ranks.py, as the starting point for a version
// that sends from the interface tiles themselves.
static int g_mg_fuse_exchange = 0;
extern "C" void sic_mg_set_fused_exchange(int on) { g_mg_fuse_exchange = on ? 1 : 0; }
static long long g_fused_exchange_launches = 0;
extern "C" long long sic_mg_fused_exchange_launches(void) { return g_fused_exchange_launches; }

static bool mg_apply_fused(const sic_mg_level_t& L, const double* x, double* t, const int* done, cudaStream_t st) {
#ifdef SIC_HOSTEMU
  (void)L; (void)x; (void)t; (void)done; (void)st;
  return false;
#else
  const sic_halo_t* h = mg_halo(L);
  if (!g_mg_fuse_exchange || !h || !h->p2p || !h->tile_order || !L.pc_ct || h->n_peers <= 0) return false;
  const int cb = mg_blocks(L.prob.n_cells, SIC_TILE_CELLS);
  if (cb <= 0 || h->n_iface_tiles <= 0 || h->n_iface_tiles > cb) return false;
  const P2P* c = (const P2P*)h->p2p;
  int cap_needed = 0;
  for (int p = 0; p < h->n_peers; ++p) {
    const int cnt = h->peer_off[p + 1] - h->peer_off[p];
    if (cnt > cap_needed) cap_needed = cnt;
  }
  if (cap_needed > c->cap) return false;
  k_mg_ebe_pc_x<<<cb + h->n_peers * c->bpp, SIC_TILE_CELLS, 0, st>>>(L.prob, L.pc_ct, L.pc_geom, *h, *c, x, t, done);
  g_fused_exchange_launches += 1;
  return true;
#endif
}

// t = K x on level L (t must be zero on entry); several GPUs: completed on the interface nodes by the halo sum
static int mg_apply(const sic_mg_level_t& L, const double* x, double* t, const int* done, cudaStream_t st) {
  const int cb = mg_blocks(L.prob.n_cells, SIC_TILE_CELLS);
  if (mg_apply_fused(L, x, t, done, st)) return sic_check_launch("k_mg_ebe_pc_x");
  if (cb > 0) {
    if (L.pc_ct) k_mg_ebe_pc<<<cb, SIC_TILE_CELLS, 0, st>>>(L.prob, L.pc_ct, L.pc_geom, L.pc_lidx, x, t, done);
    else k_mg_ebe<<<cb, SIC_TILE_CELLS, 0, st>>>(L.prob, x, t, done);
  }
  if (const sic_halo_t* h = mg_halo(L)) return sic_exchange(h, t, 3, nullptr, 0, (void*)st);
  return 0;
}

// scratch shared by the reducing kernels of one call (header + counters + block partials)
struct MgWork {
  MgScal* S; unsigned* counter; double* partials;
};
static MgWork mg_work(double* work) {
  return MgWork{(MgScal*)work, (unsigned*)(work + SIC_MG_HEADER), work + SIC_MG_HEADER + SIC_MG_COUNTERS};
}
static int64_t mg_partial_slots(int n_cells, int n_nodes) {
  return (int64_t)n_cells / SIC_TILE_CELLS + 3 * (int64_t)n_nodes / SIC_VEC_THREADS + 8;
}

extern "C" int64_t sic_mg_workspace_doubles(int n_cells, int n_nodes) {
  return SIC_MG_HEADER + SIC_MG_COUNTERS + mg_partial_slots(n_cells, n_nodes) + 3 * 3 * (int64_t)n_nodes;
}

static MgScal* g_mg_host = nullptr;   // pinned mirror of the device scalars
static cudaEvent_t g_mg_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [0,1] V-cycle operator, [2,3] Krylov operator,
                                                                                   // [1,4] the halo exchange after the V-cycle operator

static int mg_host_mirror() {
  if (g_mg_host) return 0;
  return sic_check_cuda(cudaMallocHost((void**)&g_mg_host, sizeof(MgScal)), "cudaMallocHost");
}

// One Chebyshev sweep of `its` steps on level L for eigenvalues in [lo, 1] * lambda_max; b: right-hand side.
// zero_guess != 0: x starts from 0.  On return r is the residual BEFORE the last update d: the caller that needs
// the true residual of x applies K to d once more (k_mg_resid).
static int mg_chebyshev(const sic_mg_level_t& L, const double* b, int its, double lo, int zero_guess, const int* done,
                        cudaStream_t st) {
  const int nn = L.prob.n_nodes, nc = L.prob.n_cells;
  const int nb = mg_blocks(3 * nn, SIC_DOF_THREADS);
  (void)nc;
  const double lmax = L.lambda_max, lmin = lo * lmax;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  if (!zero_guess)
    if (int rc = mg_apply(L, L.x, L.t, done, st)) return rc;                         // t = K x (t is 0 on entry)
  if (L.pc_dinv)
    k_mg_cheb_first<float><<<nb, SIC_DOF_THREADS, 0, st>>>(3 * nn, b, L.r, L.d, L.x, L.t, L.pc_dinv, L.fixed, 1.0 / theta,
                                                           zero_guess, done);
  else
    k_mg_cheb_first<double><<<nb, SIC_DOF_THREADS, 0, st>>>(3 * nn, b, L.r, L.d, L.x, L.t, L.dinv, L.fixed, 1.0 / theta,
                                                            zero_guess, done);
  for (int k = 1; k < its; ++k) {
    if (int rc = mg_apply(L, L.d, L.t, done, st)) return rc;
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    if (L.pc_dinv)
      k_mg_cheb_step<float><<<nb, SIC_DOF_THREADS, 0, st>>>(3 * nn, L.r, L.d, L.x, L.t, L.pc_dinv, L.fixed, rho_new * rho,
                                                            2.0 * rho_new / delta, done);
    else
      k_mg_cheb_step<double><<<nb, SIC_DOF_THREADS, 0, st>>>(3 * nn, L.r, L.d, L.x, L.t, L.dinv, L.fixed, rho_new * rho,
                                                             2.0 * rho_new / delta, done);
    rho = rho_new;
  }
  return sic_check_launch("multigrid: Chebyshev sweep");
}

// The coarsest-level sweep as one cooperative launch (k_mg_coarse_fused).  Returns true if it was launched; false means
// "use mg_chebyshev": not asked for, not a single-GPU replicated level, too many steps, or the grid is not co-resident.
static long long g_fused_coarse_launches = 0;
extern "C" long long sic_mg_fused_coarse_launches(void) { return g_fused_coarse_launches; }

static int g_fused_refused = 0;        // the cooperative launch failed once: launch-per-step sweep from then on
static int g_fused_not_in_capture = 0;
static bool mg_coarse_fused(const sic_mg_level_t& L, const double* b, int its, double lo, const int* done, cudaStream_t st,
                            int wanted) {
#ifdef SIC_HOSTEMU
  (void)L; (void)b; (void)its; (void)lo; (void)done; (void)st; (void)wanted;
  return false;
#else
  if (!wanted || g_fused_refused || mg_halo(L) || its < 2 || its > SIC_MG_MAX_COARSE_ITS || L.prob.n_cells <= 0) return false;
  if (g_fused_not_in_capture) {       // a capture with the cooperative launch in it failed once: keep it out of graphs
    cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cst) != cudaSuccess) cudaGetLastError();
    if (cst != cudaStreamCaptureStatusNone) return false;
  }
  const int cb = mg_blocks(L.prob.n_cells, SIC_TILE_CELLS);
  static int max_blocks = -1;
  if (max_blocks < 0) {
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mg_coarse_fused, SIC_TILE_CELLS, 0);
    max_blocks = coop ? sms * per_sm : 0;
  }
  if (cb > max_blocks) return false;
  MgCoarseCoef cf;
  const double lmax = L.lambda_max, lmin = lo * lmax;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  cf.inv_theta = 1.0 / theta;
  cf.its = its;
  cf.a[0] = cf.c[0] = 0.0;
  for (int k = 1; k < its; ++k) {       // the recurrence of mg_chebyshev
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    cf.a[k] = rho_new * rho;
    cf.c[k] = 2.0 * rho_new / delta;
    rho = rho_new;
  }
  sic_problem_t P = L.prob;
  const double* bb = b;
  double *r = L.r, *d = L.d, *x = L.x, *t = L.t;
  const double* dinv = L.dinv;
  const uint8_t* fixed = L.fixed;
  const int* dn = done;
  void* args[] = {&P, &bb, &r, &d, &x, &t, &dinv, &fixed, &cf, &dn};
  if (cudaLaunchCooperativeKernel((const void*)k_mg_coarse_fused, dim3(cb), dim3(SIC_TILE_CELLS), args, 0, st) != cudaSuccess) {
    cudaGetLastError();      // clear; fall back to the launch-per-step sweep for good
    g_fused_refused = 1;
    return false;
  }
  g_fused_coarse_launches += 1;
  return true;
#endif
}

// V(nu,nu) cycle: reads `b_top` as the right-hand side of the finest level, leaves the result in levels[top].x.
// `done` points at a device int: every kernel is a no-op once it is non-zero.
static int mg_vcycle(sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o, const double* b_top, const int* done,
                     cudaStream_t st, cudaEvent_t* time_top_apply) {
  const int top = n_levels - 1;
  for (int l = top; l >= 1; --l) {
    const sic_mg_level_t& L = lv[l];
    const sic_mg_level_t& C = lv[l - 1];
    const double* b = (l == top) ? b_top : L.b;
    if (int rc = mg_chebyshev(L, b, o->nu, o->smooth_lo, 1, done, st)) return rc;
    const int nn = L.prob.n_nodes, nd = 3 * nn, cb = mg_blocks(L.prob.n_cells, SIC_TILE_CELLS);
    const sic_halo_t* h = mg_halo(L);
    // (the timed launch keeps the operator and its exchange apart so that each can be bracketed by events)
    const bool timed = time_top_apply && l == top;
    if (timed || !mg_apply_fused(L, L.d, L.t, done, st)) {
      if (timed) cudaEventRecord(time_top_apply[0], st);
      if (cb > 0) {
        if (L.pc_ct) k_mg_ebe_pc<<<cb, SIC_TILE_CELLS, 0, st>>>(L.prob, L.pc_ct, L.pc_geom, L.pc_lidx, L.d, L.t, done);
        else k_mg_ebe<<<cb, SIC_TILE_CELLS, 0, st>>>(L.prob, L.d, L.t, done);
      }
      if (timed) cudaEventRecord(time_top_apply[1], st);
      if (h) if (int rc = sic_exchange(h, L.t, 3, nullptr, 0, (void*)st)) return rc;
      if (timed && h) cudaEventRecord(time_top_apply[4], st);   // [1]..[4]: one finest-level halo exchange
    }
    k_mg_resid<<<mg_blocks(nd, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(nd, L.r, L.t, L.fixed, done);
    // several GPUs: every fine node is restricted by its owner only; the coarse right-hand side is then completed by a
    // halo sum when the coarse level is partitioned too (nested partition: the parents of an owned fine node are
    // local), by one all-reduce over NVLink when it is replicated
    k_mg_restrict<<<mg_blocks(3 * C.prob.n_nodes, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(
        C.prob.n_nodes, L.rst_ptr, L.rst_idx, L.r, C.b, C.fixed, h ? h->owner_w : nullptr, done);
    if (h) {
      if (const sic_halo_t* hc = mg_halo(C)) { if (int rc = sic_exchange(hc, C.b, 3, nullptr, 0, (void*)st)) return rc; }
      else if (int rc = sic_allreduce_sum(h->comm, C.b, 3 * C.prob.n_nodes, (void*)st)) return rc;
    }
  }
  {
    const sic_mg_level_t& L = lv[0];
    const double* b = (top == 0) ? b_top : L.b;
    if (!mg_coarse_fused(L, b, o->coarse_its, o->coarse_lo, done, st, o->fused_coarse))
      if (int rc = mg_chebyshev(L, b, o->coarse_its, o->coarse_lo, 1, done, st)) return rc;
  }
  for (int l = 1; l <= top; ++l) {
    const sic_mg_level_t& L = lv[l];
    const sic_mg_level_t& C = lv[l - 1];
    const double* b = (l == top) ? b_top : L.b;
    const int nn = L.prob.n_nodes;
    k_mg_prolong_add<<<mg_blocks(3 * nn, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(nn, L.parent_a, L.parent_b, C.x, L.x,
                                                                                 L.fixed, done);
    if (int rc = mg_chebyshev(L, b, o->nu, o->smooth_lo, 0, done, st)) return rc;
  }
  return sic_check_launch("multigrid: V-cycle");
}

extern "C" int sic_mg_setup(sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o, double* work, void* stream) {
  if (int rc = mg_check_levels(lv, n_levels, o)) return rc;
  if (!work) return sic_fail("sic_mg_setup: null workspace");
  if (o->power_its == 1) return sic_fail("sic_mg_setup: power_its must be 0 or >= 2");
  if (o->power_its_warm < 0) return sic_fail("sic_mg_setup: power_its_warm must be >= 0");
  if (int rc = mg_host_mirror()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // 1. Galerkin coarse tangents, fine to coarse (several GPUs: rank-local between partitioned levels; into the first
  //    replicated level every rank sums the children it holds and one all-reduce completes it; the levels below are
  //    computed redundantly)
  for (int l = n_levels - 1; l >= 1; --l) {
    const int ncoarse = lv[l - 1].prob.n_cells;
    if (ncoarse > 0)
      k_mg_ct_coarsen<<<mg_blocks(ncoarse, 128), 128, 0, st>>>(ncoarse, lv[l].children, lv[l].prob.CT, lv[l - 1].prob.CT);
    if (int rc = sic_check_launch("k_mg_ct_coarsen")) return rc;
    const sic_halo_t* h = mg_halo(lv[l]);
    if (h && !mg_halo(lv[l - 1])) {      // replicated coarse level: partial sums of the children each rank holds
      const int64_t cnt = 36 * (int64_t)lv[l - 1].prob.cell_stride;
      if (cnt > 2147483647) return sic_fail("sic_mg_setup: coarse C_T too large for one all-reduce");
      if (int rc = sic_allreduce_sum(h->comm, lv[l - 1].prob.CT, (int)cnt, stream)) return rc;
    }
  }
  for (int l = 0; l < n_levels; ++l) {     // compressed copies for the operator applications inside the V-cycle
    const int nc = lv[l].prob.n_cells;
    if (lv[l].pc_ct && nc > 0) k_mg_ct_compress<<<mg_blocks(nc, 128), 128, 0, st>>>(nc, lv[l].prob.CT, lv[l].pc_ct);
  }
  if (int rc = sic_check_launch("k_mg_ct_compress")) return rc;
  MgWork W = mg_work(work);
  // 2. block-Jacobi blocks and 3. lambda_max(Dinv K) by power iteration, level by level (x: v, d: w, t: K v)
  for (int l = 0; l < n_levels; ++l) {
    sic_mg_level_t& L = lv[l];
    const sic_halo_t* h = mg_halo(L);
    const int multi = h ? 1 : 0;
    const double* ow = h ? h->owner_w : nullptr;
    if (int rc = sic_block_jacobi(&L.prob, L.dinv, L.fixed, h, stream)) return rc;
    if (L.pc_dinv && L.prob.n_nodes > 0) {
      const int n9 = 9 * L.prob.n_nodes;
      k_mg_to_float<<<mg_blocks(n9, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(n9, L.dinv, L.pc_dinv);
    }
    if (o->power_its <= 0) {
      if (!(L.lambda_max > 0.0)) return sic_fail("sic_mg_setup: power_its = 0 needs lambda_max from an earlier call");
      continue;
    }
    const int nn = L.prob.n_nodes, nd = 3 * nn, nc = L.prob.n_cells;
    if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_MG_HEADER + SIC_MG_COUNTERS), st),
                                "memset mg scalars"))
      return rc;
    const int db = mg_blocks(nd, SIC_VEC_THREADS), nb = mg_blocks(nn, SIC_VEC_THREADS), cb = mg_blocks(nc, SIC_TILE_CELLS);
    const int* never = &W.S->done;    // stays 0
    // The iterate lives in L.pv when the caller provides that persistent vector: the next setup then restarts
    // from it (the tangent changes little between Newton iterations) and needs only power_its_warm passes.
    double* v = L.pv ? L.pv : L.x;
    const bool warm = L.pv && L.lambda_max > 0.0 && o->power_its_warm >= 1;
    const int its = warm ? o->power_its_warm : o->power_its;
    if (!warm) k_mg_pw_init<<<db, SIC_VEC_THREADS, 0, st>>>(nd, v, L.t, L.fixed, 1.0);
    else if (int rc = sic_check_cuda(cudaMemsetAsync(L.t, 0, sizeof(double) * nd, st), "memset t")) return rc;
    (void)cb;
    const MgFin pw{W.S, MG_OP_PW, 0.0, 0.0, 0, multi};
    for (int it = 0; it < its; ++it) {
      if (int rc = mg_apply(L, v, L.t, never, st)) return rc;
      // w = Dinv K v ; pw = w.w (= lambda^2 once v has unit length: from the second pass on, or at once when warm)
      k_mg_pw_step<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, L.t, L.d, L.dinv, L.fixed, ow, pw, W.partials, W.counter);
      if (multi) {
        if (int rc = sic_exchange(h, nullptr, 0, W.S->sum, 1, stream)) return rc;
        k_mg_scal<<<1, 1, 0, st>>>(pw, 0);
      }
      if (it + 1 < its || L.pv) k_mg_pw_scale<<<db, SIC_VEC_THREADS, 0, st>>>(nd, v, L.d, L.t, W.S);
    }
    if (int rc = sic_check_launch("multigrid: power iteration")) return rc;
    cudaMemcpyAsync(g_mg_host, W.S, sizeof(MgScal), cudaMemcpyDeviceToHost, st);
    if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "sic_mg_setup sync")) return rc;
    const double lam = sqrt(g_mg_host->pw);
    if (!(lam > 0.0) || isinf(lam)) return sic_fail("sic_mg_setup: power iteration failed (lambda_max not positive/finite)");
    L.lambda_max = o->safety * lam;
    if (int rc = sic_check_cuda(cudaMemsetAsync(L.t, 0, sizeof(double) * nd, st), "memset t")) return rc;
  }
  return 0;
}

extern "C" int sic_mg_vcycle(sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o, double* work, void* stream) {
  if (int rc = mg_check_levels(lv, n_levels, o)) return rc;
  if (!work) return sic_fail("sic_mg_vcycle: null workspace");
  cudaStream_t st = (cudaStream_t)stream;
  for (int l = 0; l < n_levels; ++l) {
    if (!(lv[l].lambda_max > 0.0)) return sic_fail("sic_mg_vcycle: call sic_mg_setup first (lambda_max not set)");
    if (int rc = sic_check_cuda(cudaMemsetAsync(lv[l].t, 0, sizeof(double) * 3 * lv[l].prob.n_nodes, st), "memset t"))
      return rc;
  }
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_MG_HEADER + SIC_MG_COUNTERS), st),
                              "memset mg header"))
    return rc;
  MgWork W = mg_work(work);
  return mg_vcycle(lv, n_levels, o, lv[n_levels - 1].b, &W.S->done, st, nullptr);
}

static long long g_mg_graph_captures = 0;      // stays 0 in the host emulation (no graphs there)
extern "C" long long sic_mg_graph_captures(void) { return g_mg_graph_captures; }

#ifndef SIC_HOSTEMU
// ---- CUDA graph of one MG-CG iteration -----------------------------------------------------------------------
// One iteration is ~110 launches on a 5-level hierarchy (170 with the launch-per-step coarsest sweep), most of them
// 5-10 us kernels of the coarse levels: launched one by one they are bound by the launch rate of the host thread.
// The iteration is captured ONCE per (hierarchy, tangent set-up, solve arguments) and replayed with one
// cudaGraphLaunch per iteration.  Everything that changes between replays lives in device memory (CG scalars, the
// `done` flag, the epochs of the P2P exchange); what is baked into the nodes -- pointers, sizes, the Chebyshev
// coefficients that follow lambda_max -- is the cache key, so a new set-up re-captures (about a millisecond).
struct MgGraphKey {
  sic_mg_level_t lv[SIC_MG_MAX_LEVELS];
  sic_mg_opts_t o;
  const double* b_ext; double* x; double* work;
  double rtol, atol;
  double lam[SIC_MG_MAX_LEVELS];      // the levels' lambda_max (kept out of lv: a new set-up changes only these)
  int n_levels, guess, fused_refused;
  int kind;                           // 0: one CG iteration, 1: the first cycle of a solve (z = M^-1 r0, rz, p = z)
  cudaStream_t st;
};
struct MgGraphEntry { MgGraphKey key; cudaGraphExec_t exec; unsigned long long used; };
#define SIC_MG_GRAPH_SLOTS 8
static MgGraphEntry g_mg_graphs[SIC_MG_GRAPH_SLOTS];
static unsigned long long g_mg_graph_clock = 0;
static int g_mg_graph_broken = 0;      // a capture failed once: launch kernel by kernel from then on

static cudaStream_t mg_capture_stream() {
  static cudaStream_t cs = nullptr;
  if (!cs && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); cs = nullptr; }
  return cs;
}

template <class Body>
static cudaGraphExec_t mg_iteration_graph(const sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o, const double* b_ext,
                                          double* x, double* work, double rtol, double atol, int guess, cudaStream_t st,
                                          Body& body, int kind = 0) {
  if (g_mg_graph_broken) return nullptr;
  MgGraphKey key;
  memset(&key, 0, sizeof(key));
  memcpy(key.lv, lv, sizeof(sic_mg_level_t) * n_levels);
  memcpy(&key.o, o, sizeof(sic_mg_opts_t));
  key.b_ext = b_ext; key.x = x; key.work = work; key.rtol = rtol; key.atol = atol;
  key.n_levels = n_levels; key.guess = guess; key.fused_refused = g_fused_refused | (g_fused_not_in_capture << 1); key.st = st;
  key.kind = kind;
  for (int l = 0; l < n_levels; ++l) { key.lam[l] = lv[l].lambda_max; key.lv[l].lambda_max = 0.0; }
  MgGraphEntry* slot = nullptr;
  MgGraphEntry* lru = &g_mg_graphs[0];
  for (int i = 0; i < SIC_MG_GRAPH_SLOTS; ++i) {
    MgGraphEntry& e = g_mg_graphs[i];
    if (e.exec && memcmp(&e.key, &key, sizeof(key)) == 0) { e.used = ++g_mg_graph_clock; return e.exec; }
    if (e.exec && memcmp(e.key.lv, key.lv, sizeof(key.lv)) == 0 && e.key.x == key.x && e.key.work == key.work &&
        e.key.kind == key.kind)
      slot = &e;                             // same hierarchy and vectors: its graph has the same topology
    if (e.used < lru->used) lru = &e;        // least recently used (empty slots have used == 0)
  }
  if (!slot) slot = lru;
  cudaGraph_t graph = nullptr;
  for (int attempt = 0; attempt < 2 && !graph; ++attempt) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      fprintf(stderr, "safeincave_cuda: cudaStreamBeginCapture failed (%s); launching kernel by kernel\n",
              cudaGetErrorString(cudaGetLastError()));
      g_mg_graph_broken = 1;
      return nullptr;
    }
    const int rc = body(nullptr);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc == 0 && ce == cudaSuccess && graph) break;
    cudaGetLastError();
    if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
    if (attempt == 0 && o->fused_coarse && !g_fused_not_in_capture) {   // retry without the cooperative launch
      g_fused_not_in_capture = 1;
      key.fused_refused = g_fused_refused | (g_fused_not_in_capture << 1);
      continue;
    }
    g_mg_graph_broken = 1;
    fprintf(stderr, "safeincave_cuda: CUDA graph capture of the MG-CG iteration failed (%s); launching kernel by kernel\n",
            ce != cudaSuccess ? cudaGetErrorString(ce) : sic_last_error());
    return nullptr;
  }
  cudaGraphExec_t exec = nullptr;
  // same hierarchy as the slot's previous graph (a new tangent set-up changed only kernel arguments): update in place
  if (slot->exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(slot->exec, graph, &info) == cudaSuccess) exec = slot->exec;
    else { cudaGetLastError(); cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
  }
  if (!exec && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    cudaGraphDestroy(graph);
    g_mg_graph_broken = 1;
    return nullptr;
  }
  cudaGraphDestroy(graph);
  slot->key = key; slot->exec = exec; slot->used = ++g_mg_graph_clock;
  g_mg_graph_captures += 1;
  return exec;
}
#endif

extern "C" int sic_mg_solve(sic_mg_level_t* lv, int n_levels, const sic_mg_opts_t* o, sic_ksp_t* ksp,
                            const double* b_ext, double* x, double* work, void* stream) {
  if (int rc = mg_check_levels(lv, n_levels, o)) return rc;
  if (!ksp || !b_ext || !x || !work) return sic_fail("sic_mg_solve: null argument");
  if (int rc = mg_host_mirror()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  sic_mg_level_t& T = lv[n_levels - 1];
  const sic_problem_t* p = &T.prob;
  const uint8_t* fixed = T.fixed;
  const int nn = p->n_nodes, nd = 3 * nn, nc = p->n_cells;
  for (int l = 0; l < n_levels; ++l)
    if (!(lv[l].lambda_max > 0.0)) return sic_fail("sic_mg_solve: call sic_mg_setup first (lambda_max not set)");
  MgWork W = mg_work(work);
  MgScal* S = W.S;
  double* vec = W.partials + mg_partial_slots(nc, nn);
  double *r = vec, *pp = vec + nd, *q = vec + 2 * (size_t)nd;
  double* z = T.x;
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_MG_HEADER + SIC_MG_COUNTERS), st),
                              "memset mg header"))
    return rc;
  if (int rc = sic_check_cuda(cudaMemsetAsync(vec, 0, sizeof(double) * 3 * (size_t)nd, st), "memset mg-cg vectors"))
    return rc;
  for (int l = 0; l < n_levels; ++l)
    if (int rc = sic_check_cuda(cudaMemsetAsync(lv[l].t, 0, sizeof(double) * 3 * lv[l].prob.n_nodes, st), "memset t"))
      return rc;
  const int nb = mg_blocks(nn, SIC_VEC_THREADS), db = mg_blocks(nd, SIC_VEC_THREADS), cb = mg_blocks(nc, SIC_TILE_CELLS);
  const int rb = mg_reduce_blocks(nd);
  (void)nb;
  const int check = ksp->check_every > 0 ? ksp->check_every : 4;
  const int guess = ksp->guess_nonzero ? 1 : 0;
  const double rtol = ksp->rtol, atol = ksp->atol;
  const sic_halo_t* halo = mg_halo(T);
  const int multi = halo ? 1 : 0;
  const double* ow = halo ? halo->owner_w : nullptr;
  auto fin = [&](int op) { return MgFin{S, op, rtol, atol, guess, multi}; };
  // several GPUs: add the partial sums of the last reducing kernel over the ranks (optionally together with the halo
  // sum of `vec`: one exchange kernel), then run the scalar recurrence in a one-thread kernel -- still no host sync
  auto reduce = [&](int op, double* vec3, int skip_if_done) -> int {
    if (!multi) return 0;
    if (int rc = sic_exchange(halo, vec3, vec3 ? 3 : 0, S->sum, 1, stream)) return rc;
    k_mg_scal<<<1, 1, 0, st>>>(fin(op), skip_if_done);
    return sic_check_launch("k_mg_scal");
  };
  ksp->op_samples = 0;
  ksp->op_ms = 0.0;
  ksp->op_dot_samples = 0;
  ksp->op_dot_ms = 0.0;
  ksp->xchg_samples = 0;
  ksp->xchg_ms = 0.0;
  cudaEvent_t* ev = nullptr;
  if (ksp->time_operator) {
    if (!g_mg_ev[0]) for (int k = 0; k < 5; ++k) cudaEventCreate(&g_mg_ev[k]);
    ev = g_mg_ev;
  }

  if (guess) {   // reference norm of rtol: the residual of the zero guess (prescribed values only), as PETSc's ||b||
    k_mg_zero_free<<<db, SIC_VEC_THREADS, 0, st>>>(nd, pp, x, fixed);
    if (int rc = sic_residual0(p, b_ext, pp, r, fixed, halo, stream)) return rc;
    k_mg_dot<<<rb, SIC_VEC_THREADS, 0, st>>>(nd, r, r, ow, fin(MG_OP_REF), 0, W.partials, W.counter);
    if (int rc = reduce(MG_OP_REF, nullptr, 0)) return rc;
  }
  if (int rc = sic_residual0(p, b_ext, x, r, fixed, halo, stream)) return rc;
  k_mg_dot<<<rb, SIC_VEC_THREADS, 0, st>>>(nd, r, r, ow, fin(MG_OP_INIT_RR), 0, W.partials, W.counter);
  if (int rc = reduce(MG_OP_INIT_RR, nullptr, 0)) return rc;
  // first cycle of the solve: z = M^-1 r0, rz = r0.z, p = z ; q = 0 (its own graph when use_graph, below)
  auto first_cycle = [&](cudaEvent_t*) -> int {
    if (int rc = mg_vcycle(lv, n_levels, o, r, &S->done, st, nullptr)) return rc;
    k_mg_dot<<<rb, SIC_VEC_THREADS, 0, st>>>(nd, r, z, ow, fin(MG_OP_INIT_RZ), 1, W.partials, W.counter);
    if (int rc = reduce(MG_OP_INIT_RZ, nullptr, 1)) return rc;
    k_mg_cg_p<<<db, SIC_VEC_THREADS, 0, st>>>(nd, pp, z, q, fixed, S, 1);
    return 0;
  };

  // one CG iteration: q = K p with p.Kp summed per cell (cells are partitioned: no owner weights needed); several
  // GPUs: the halo sum of q and the sum of p.Kp over the ranks travel in ONE exchange.  Every argument is either a
  // device pointer or a value fixed until the next setup, so the iteration can be replayed from a CUDA graph.
  auto iteration = [&](cudaEvent_t* time_ev) -> int {
    if (time_ev) cudaEventRecord(time_ev[2], st);
    if (cb > 0) k_mg_ebe_dot<<<cb, SIC_TILE_CELLS, 0, st>>>(*p, pp, q, W.partials, &S->done);
    if (time_ev) cudaEventRecord(time_ev[3], st);
    k_mg_sum_pq<<<1, 1024, 0, st>>>(W.partials, cb, fin(MG_OP_PQ));
    if (int rc = reduce(MG_OP_PQ, q, 1)) return rc;
    k_mg_cg_update<<<rb, SIC_VEC_THREADS, 0, st>>>(nd, x, r, pp, q, fixed, ow, fin(MG_OP_RR), W.partials, W.counter);
    if (int rc = reduce(MG_OP_RR, nullptr, 1)) return rc;
    if (int rc = mg_vcycle(lv, n_levels, o, r, &S->done, st, time_ev)) return rc;
    k_mg_dot<<<rb, SIC_VEC_THREADS, 0, st>>>(nd, r, z, ow, fin(MG_OP_RZ), 1, W.partials, W.counter);
    if (int rc = reduce(MG_OP_RZ, nullptr, 1)) return rc;
    k_mg_cg_p<<<db, SIC_VEC_THREADS, 0, st>>>(nd, pp, z, q, fixed, S, 0);
    return 0;
  };
  ksp->graph_launches = 0;
  ksp->direct_iterations = 0;
#ifndef SIC_HOSTEMU
  // Captured on a stream of the library's own (the caller's is normally the legacy default stream, which cannot be
  // captured) and replayed on the caller's stream.  The lambdas above launch on `st` / `stream`, whatever they hold.
  cudaGraphExec_t gexec = nullptr, gfirst = nullptr;
  if (ksp->use_graph) {
    if (cudaStream_t cs = mg_capture_stream()) {
      void* const user = stream;
      st = cs; stream = (void*)cs;
      gexec = mg_iteration_graph(lv, n_levels, o, b_ext, x, work, rtol, atol, guess, cs, iteration);
      if (gexec) gfirst = mg_iteration_graph(lv, n_levels, o, b_ext, x, work, rtol, atol, guess, cs, first_cycle, 1);
      st = (cudaStream_t)user; stream = user;
    }
  }
  if (gfirst) {
    if (int rc = sic_check_cuda(cudaGraphLaunch(gfirst, st), "cudaGraphLaunch (mg-cg first cycle)")) return rc;
    ksp->graph_launches += 1;
  } else
#endif
  if (int rc = first_cycle(nullptr)) return rc;

  int launched = 0;
  bool timed_batch = false;
  while (true) {
    cudaMemcpyAsync(g_mg_host, S, sizeof(MgScal), cudaMemcpyDeviceToHost, st);
    if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "mg sync")) return rc;
    if (timed_batch) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) { ksp->op_ms += ms; ksp->op_samples += 1; }
      if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) { ksp->op_dot_ms += ms; ksp->op_dot_samples += 1; }
      if (multi && cudaEventElapsedTime(&ms, ev[1], ev[4]) == cudaSuccess) { ksp->xchg_ms += ms; ksp->xchg_samples += 1; }
      timed_batch = false;
    }
    if (g_mg_host->done || launched >= ksp->max_it) break;
    const int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
    for (int k = 0; k < batch; ++k) {
#ifndef SIC_HOSTEMU
      // the timed iteration is launched kernel by kernel (its events bracket one operator launch)
      const bool timed = ev && k == 0 && (launched == 0 || !gexec);
      if (timed) timed_batch = true;
      if (gexec && !timed) {
        if (int rc = sic_check_cuda(cudaGraphLaunch(gexec, st), "cudaGraphLaunch (mg-cg iteration)")) return rc;
        ksp->graph_launches += 1;
        continue;
      }
#endif
#ifdef SIC_HOSTEMU
      const bool timed = ev && k == 0;
      if (timed) timed_batch = true;
#endif
      if (int rc = iteration(timed ? ev : nullptr)) return rc;
      ksp->direct_iterations += 1;
    }
    launched += batch;
    if (int rc = sic_check_launch("mg-cg batch")) return rc;
  }
  if (multi && halo->p2p && sic_p2p_error(halo->p2p)) return sic_fail("P2P exchange timed out waiting for a peer");
  ksp->iterations = g_mg_host->iters;
  ksp->rnorm = sqrt(g_mg_host->rr);
  ksp->rnorm0 = sqrt(g_mg_host->rr0);
  if (g_mg_host->nanflag) ksp->reason = -9;
  else if (g_mg_host->done) ksp->reason = g_mg_host->reason ? g_mg_host->reason : 2;
  else ksp->reason = -3;
  return 0;
}
