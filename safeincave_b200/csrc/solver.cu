// solver.cu — hot-path part (3): block-Jacobi preconditioned CG / BiCGStab on the matrix-free operator.
//
// Replaces `self.solver.solve(b, X)` of MomentumEquation.py:1023-1025 (and :920-922), which in the
// reference is a user-configured PETSc KSP (cg / bicg / bcgs / gmres + asm).
// B200 design: the whole iteration lives on the device.  Scalars (alpha, beta, rho, omega, norms)
// stay in device memory, every dot product is fused into the kernel that produces its operand
// (p.Kp is accumulated per CELL inside the operator kernel at no extra memory traffic -- and, cells
// being partitioned, it needs no owner weights on several GPUs) and is finished by the last block to
// arrive, the convergence test sets a device flag that turns the remaining launches of a batch into
// no-ops, and the host only looks every `check_every` iterations.  Dirichlet rows/columns are handled
// by keeping search directions zero on constrained dofs (equivalent to assemble_matrix(bcs) +
// apply_lifting + set_bc, MomentumEquation.py:1010-1020).
//
// Several GPUs (halo != NULL): vectors are kept CONSISTENT on interface nodes; the operator result is
// halo-summed (ncclSend/ncclRecv), node dot products use owner weights, and the per-rank sums go through
// ncclAllReduce on the device buffer S->sum followed by a one-thread scalar kernel -- still no host sync.
#include <math.h>
#include <string.h>

#include "ebe_tma.cuh"

namespace sic {

struct Scal {
  double rz, pq, rr, rr0, alpha, beta, tol2;
  double rho, rhv, omega, ts, tt;
  double sum[4];   // per-rank partial sums, all-reduced in place on several GPUs
  double rr_ref;   // ||r||^2 of the zero initial guess (reference for rtol when warm-starting)
  int done, iters, nanflag, reason;
};
static_assert(sizeof(Scal) <= 64 * sizeof(double), "Scal must fit the reserved workspace header");

#define SIC_WS_HEADER 64   /* doubles reserved for Scal */
#define SIC_WS_COUNTERS 8  /* doubles reserved for ticket counters */

enum { OP_NONE = -1, OP_REF = 0, OP_CG_INIT, OP_CG_ALPHA, OP_CG_BETA, OP_BI_INIT, OP_BI_ALPHA, OP_BI_OMEGA, OP_BI_BETA,
       OP_CGCG_INIT, OP_CGCG_STEP };

__device__ __forceinline__ void check_convergence(Scal* S, double rr) {
  S->rr = rr;
  S->iters += 1;
  if (!(rr == rr) || isinf(rr)) { S->nanflag = 1; S->done = 1; S->reason = -9; }
  else if (rr <= S->tol2) { S->done = 1; }
}

__device__ __forceinline__ void init_tolerance(Scal* S, double rr, double rtol, double atol, int guess) {
  S->rr = rr;
  S->rr0 = rr;
  const double ref = guess ? S->rr_ref : rr;
  const double t = rtol * rtol * ref, a2 = atol * atol;
  S->tol2 = (t > a2) ? t : a2;
  S->iters = 0; S->nanflag = 0; S->reason = 0; S->done = 0;
  if (!(rr == rr) || isinf(rr)) { S->nanflag = 1; S->done = 1; S->reason = -9; }
  else if (rr <= S->tol2 || rr == 0.0) { S->done = 1; S->reason = (rr <= a2) ? 3 : 2; }
}

// The scalar recurrences.  Runs in the last block of the producing kernel (one GPU) or in k_scal after
// the allreduce (several GPUs).  Input: S->sum[].
__device__ __forceinline__ void scal_step(Scal* S, int op, double rtol, double atol, int guess) {
  switch (op) {
    case OP_REF: S->rr_ref = S->sum[0]; break;
    case OP_CG_INIT: S->rz = S->sum[0]; init_tolerance(S, S->sum[1], rtol, atol, guess); break;
    case OP_CG_ALPHA: S->pq = S->sum[0]; S->alpha = S->rz / S->sum[0]; break;
    case OP_CG_BETA: S->beta = S->sum[0] / S->rz; S->rz = S->sum[0]; check_convergence(S, S->sum[1]); break;
    case OP_BI_INIT:
      S->rho = S->sum[0]; S->alpha = 1.0; S->omega = 1.0;
      init_tolerance(S, S->sum[0], rtol, atol, guess);
      break;
    case OP_BI_ALPHA: S->rhv = S->sum[0]; S->alpha = S->rho / S->sum[0]; break;
    case OP_BI_OMEGA: S->ts = S->sum[0]; S->tt = S->sum[1]; S->omega = S->sum[0] / S->sum[1]; break;
    case OP_BI_BETA:
      S->beta = (S->sum[0] / S->rho) * (S->alpha / S->omega);
      S->rho = S->sum[0];
      check_convergence(S, S->sum[1]);
      break;
    // Chronopoulos-Gear: sum[0] = gamma = r.u, sum[1] = r.r, sum[2] = delta = u.Ku
    case OP_CGCG_INIT:
      S->rz = S->sum[0];
      init_tolerance(S, S->sum[1], rtol, atol, guess);
      S->alpha = S->sum[0] / S->sum[2];
      S->beta = 0.0;
      break;
    case OP_CGCG_STEP: {
      const double gamma_new = S->sum[0];
      S->beta = gamma_new / S->rz;
      S->alpha = gamma_new / (S->sum[2] - S->beta * gamma_new / S->alpha);
      S->rz = gamma_new;
      check_convergence(S, S->sum[1]);
      break;
    }
  }
}

__global__ void k_scal(Scal* S, int op, double rtol, double atol, int guess, int skip_if_done) {
  if (skip_if_done && S->done) return;
  scal_step(S, op, rtol, atol, guess);
}

struct Fin {  // what the last block does with the grid totals: store them at S->sum[slot..], then (one GPU) step
  Scal* S; int op; double rtol, atol; int guess; int multi; int slot;
  template <int NV>
  __device__ __forceinline__ void run(const double* tot) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) S->sum[slot + k] = tot[k];
    if (!multi && op != OP_NONE) scal_step(S, op, rtol, atol, guess);
  }
};

#ifndef SIC_EBE_IMPL
#define SIC_EBE_IMPL 3   /* 3: one thread per cell, front-batched loads (fem.cuh); 2: TMA-staged tiles (ebe_tma.cuh) */
#endif

#if SIC_EBE_IMPL == 3
#define SIC_EBE_BLOCK 128
#ifndef SIC_EBE_MINBLOCKS
#define SIC_EBE_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(SIC_EBE_BLOCK, SIC_EBE_MINBLOCKS) k_ebe_dot(sic_problem_t P, const double* __restrict__ x,
                                                             double* __restrict__ y, Fin fin,
                                                             double* __restrict__ partials, unsigned* counter) {
  __shared__ TileScratch sc;
  double v[1] = {ebe_tile_scatter<0>(P, x, y, sc, &fin.S->done)};
  block_partials<1, SIC_EBE_BLOCK>(v, partials);     // summed by k_sum_partials (no per-CTA ticket wait)
}

// one block: sum n block partials in order, then run the scalar recurrence
__global__ void __launch_bounds__(1024) k_sum_partials(const double* __restrict__ partials, int n, Fin fin) {
  if (fin.S->done) return;
  __shared__ double sh[32];
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 1024) a += partials[k];
  a = warp_sum(a);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[1] = {0.0};
    for (int k = 0; k < 32; ++k) tot[0] += sh[k];
    fin.run<1>(tot);
  }
}
__global__ void __launch_bounds__(SIC_EBE_BLOCK, SIC_EBE_MINBLOCKS) k_ebe_plain(sic_problem_t P, const double* __restrict__ x,
                                                               double* __restrict__ y, const Scal* S) {
  __shared__ TileScratch sc;
  ebe_tile_scatter<0>(P, x, y, sc, &S->done);
}
#else
// ---- operator kernel (TMA-staged persistent tiles, ebe_tma.cuh) with the fused p.Kp reduction ----
__global__ void __launch_bounds__(SIC_TILE, 2) k_ebe_dot(sic_problem_t P, const double* __restrict__ x,
                                                        double* __restrict__ y, Fin fin,
                                                        double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<0>& sm = *reinterpret_cast<TileSmem<0>*>(smem_raw);
  double v[1] = {ebe_tiles<0>(P, x, y, sm)};
  grid_reduce<1, SIC_TILE>(v, partials, counter, [&](const double* tot) { fin.run<1>(tot); });
}

// plain operator (no dot), skipping when converged
__global__ void __launch_bounds__(SIC_TILE, 2) k_ebe_plain(sic_problem_t P, const double* __restrict__ x,
                                                          double* __restrict__ y, const Scal* S) {
  if (S->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<0>& sm = *reinterpret_cast<TileSmem<0>*>(smem_raw);
  ebe_tiles<0>(P, x, y, sm);
}
#endif

__device__ __forceinline__ void precond3(const double* __restrict__ dinv, size_t n, const double r[3], double z[3]) {
  const double* d = dinv + 9 * n;
#pragma unroll
  for (int j = 0; j < 3; ++j) z[j] = __ldg(d + 3 * j) * r[0] + __ldg(d + 3 * j + 1) * r[1] + __ldg(d + 3 * j + 2) * r[2];
}

// z = fixed ? x : 0   (the zero initial guess that still carries the prescribed values)
__global__ void k_zero_free(int nd, double* __restrict__ z, const double* __restrict__ x,
                            const uint8_t* __restrict__ fixed) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < nd) z[d] = fixed[d] ? x[d] : 0.0;
}

// sum_owned r.r
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_norm2(int n_nodes, const double* __restrict__ r,
                                                          const double* __restrict__ w, Fin fin,
                                                          double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (n < n_nodes) {
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) { const double a = r[3 * (size_t)n + j]; v[0] += wn * a * a; }
  }
  grid_reduce<1, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run<1>(tot); });
}

// ---- PCG ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_init(int n_nodes, const double* __restrict__ r,
                                                            double* __restrict__ z, double* __restrict__ p,
                                                            double* __restrict__ q, const double* __restrict__ dinv,
                                                            const double* __restrict__ w, Fin fin,
                                                            double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (n < n_nodes) {
    double rn[3], zn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) rn[j] = r[3 * (size_t)n + j];
    precond3(dinv, n, rn, zn);
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      z[3 * (size_t)n + j] = zn[j];
      p[3 * (size_t)n + j] = zn[j];
      q[3 * (size_t)n + j] = 0.0;
      v[0] += wn * rn[j] * zn[j];
      v[1] += wn * rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_update(int n_nodes, double* __restrict__ x,
                                                              double* __restrict__ r, double* __restrict__ z,
                                                              const double* __restrict__ p,
                                                              const double* __restrict__ q,
                                                              const double* __restrict__ dinv,
                                                              const uint8_t* __restrict__ fixed,
                                                              const double* __restrict__ w, Fin fin,
                                                              double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = fin.S->alpha;
  double v[2] = {0.0, 0.0};
  if (n < n_nodes) {
    double rn[3], zn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      if (fixed[d]) { rn[j] = 0.0; }
      else {
        x[d] += alpha * p[d];
        rn[j] = r[d] - alpha * q[d];
      }
      r[d] = rn[j];
    }
    precond3(dinv, n, rn, zn);
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      z[3 * (size_t)n + j] = zn[j];
      v[0] += wn * rn[j] * zn[j];
      v[1] += wn * rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_p(int nd, double* __restrict__ p, const double* __restrict__ z,
                                                         double* __restrict__ q, const Scal* S) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  p[d] = z[d] + S->beta * p[d];
  q[d] = 0.0;
}

// ---- Chronopoulos-Gear CG -----------------------------------------------------------------------------
// u = M^-1 r ; p = s = w = 0 ; sums gamma = r.u, rr = r.r
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cgcg_init(int n_nodes, const double* __restrict__ r,
                                                              double* __restrict__ u, double* __restrict__ p,
                                                              double* __restrict__ s, double* __restrict__ w,
                                                              const double* __restrict__ dinv,
                                                              const double* __restrict__ ow, Fin fin,
                                                              double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (n < n_nodes) {
    double rn[3], un[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) rn[j] = r[3 * (size_t)n + j];
    precond3(dinv, n, rn, un);
    const double wn = ow ? ow[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      u[d] = un[j]; p[d] = 0.0; s[d] = 0.0; w[d] = 0.0;
      v[0] += wn * rn[j] * un[j];
      v[1] += wn * rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

// p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; w = 0 ; sums gamma, rr
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cgcg_vec(int n_nodes, double* __restrict__ x, double* __restrict__ r,
                                                             double* __restrict__ u, double* __restrict__ p,
                                                             double* __restrict__ s, double* __restrict__ w,
                                                             const double* __restrict__ dinv,
                                                             const uint8_t* __restrict__ fixed,
                                                             const double* __restrict__ ow, Fin fin,
                                                             double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = fin.S->alpha, beta = fin.S->beta;
  double v[2] = {0.0, 0.0};
  if (n < n_nodes) {
    double rn[3], un[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      if (fixed[d]) {
        rn[j] = 0.0; p[d] = 0.0; s[d] = 0.0;
      } else {
        const double pn = u[d] + beta * p[d];
        const double sn = w[d] + beta * s[d];
        p[d] = pn; s[d] = sn;
        x[d] += alpha * pn;
        rn[j] = r[d] - alpha * sn;
      }
      r[d] = rn[j];
      w[d] = 0.0;
    }
    precond3(dinv, n, rn, un);
    const double wn = ow ? ow[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      u[3 * (size_t)n + j] = un[j];
      v[0] += wn * rn[j] * un[j];
      v[1] += wn * rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

// ---- BiCGStab (right-preconditioned) -------------------------------------------------------------
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_init(int n_nodes, const double* __restrict__ r,
                                                            double* __restrict__ rh, double* __restrict__ p,
                                                            double* __restrict__ y, double* __restrict__ v,
                                                            double* __restrict__ t, const double* __restrict__ dinv,
                                                            const double* __restrict__ w, Fin fin,
                                                            double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[1] = {0.0};
  if (n < n_nodes) {
    double rn[3], yn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) rn[j] = r[3 * (size_t)n + j];
    precond3(dinv, n, rn, yn);
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      rh[d] = rn[j]; p[d] = rn[j]; y[d] = yn[j]; v[d] = 0.0; t[d] = 0.0;
      acc[0] += wn * rn[j] * rn[j];
    }
  }
  grid_reduce<1, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) { fin.run<1>(tot); });
}

// rhv = rh . v (free dofs)  ->  alpha = rho / rhv
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_dot1(int n_nodes, const double* __restrict__ rh,
                                                            const double* __restrict__ v,
                                                            const uint8_t* __restrict__ fixed,
                                                            const double* __restrict__ w, Fin fin,
                                                            double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[1] = {0.0};
  if (n < n_nodes) {
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      if (!fixed[d]) acc[0] += wn * rh[d] * v[d];
    }
  }
  grid_reduce<1, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) { fin.run<1>(tot); });
}

// s = r - alpha v ; z = M^-1 s ; t = 0
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_s(int n_nodes, const double* __restrict__ r,
                                                         const double* __restrict__ v, double* __restrict__ s,
                                                         double* __restrict__ z, double* __restrict__ t,
                                                         const double* __restrict__ dinv,
                                                         const uint8_t* __restrict__ fixed, const Scal* S) {
  if (S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const double alpha = S->alpha;
  double sn[3], zn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t d = 3 * (size_t)n + j;
    sn[j] = fixed[d] ? 0.0 : r[d] - alpha * v[d];
    s[d] = sn[j];
    t[d] = 0.0;
  }
  precond3(dinv, n, sn, zn);
#pragma unroll
  for (int j = 0; j < 3; ++j) z[3 * (size_t)n + j] = zn[j];
}

// ts = t.s, tt = t.t (free dofs) -> omega
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_dot2(int n_nodes, const double* __restrict__ t,
                                                            const double* __restrict__ s,
                                                            const uint8_t* __restrict__ fixed,
                                                            const double* __restrict__ w, Fin fin,
                                                            double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (n < n_nodes) {
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      if (!fixed[d]) { acc[0] += wn * t[d] * s[d]; acc[1] += wn * t[d] * t[d]; }
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

// x += alpha y + omega z ; r = s - omega t ; rho_new = rh.r ; rr = r.r -> beta
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_update(int n_nodes, double* __restrict__ x,
                                                              double* __restrict__ r, const double* __restrict__ s,
                                                              const double* __restrict__ t,
                                                              const double* __restrict__ y,
                                                              const double* __restrict__ z,
                                                              const double* __restrict__ rh,
                                                              const uint8_t* __restrict__ fixed,
                                                              const double* __restrict__ w, Fin fin,
                                                              double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = fin.S->alpha, omega = fin.S->omega;
  double acc[2] = {0.0, 0.0};
  if (n < n_nodes) {
    const double wn = w ? w[n] : 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      double rn = 0.0;
      if (!fixed[d]) {
        x[d] += alpha * y[d] + omega * z[d];
        rn = s[d] - omega * t[d];
      }
      r[d] = rn;
      acc[0] += wn * rh[d] * rn;
      acc[1] += wn * rn * rn;
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) { fin.run<2>(tot); });
}

// p = r + beta (p - omega v) ; y = M^-1 p ; v = 0
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_p(int n_nodes, double* __restrict__ p,
                                                         const double* __restrict__ r, double* __restrict__ v,
                                                         double* __restrict__ y, const double* __restrict__ dinv,
                                                         const uint8_t* __restrict__ fixed, const Scal* S) {
  if (S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const double beta = S->beta, omega = S->omega;
  double pn[3], yn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const size_t d = 3 * (size_t)n + j;
    pn[j] = fixed[d] ? 0.0 : r[d] + beta * (p[d] - omega * v[d]);
    p[d] = pn[j];
    v[d] = 0.0;
  }
  precond3(dinv, n, pn, yn);
#pragma unroll
  for (int j = 0; j < 3; ++j) y[3 * (size_t)n + j] = yn[j];
}

}  // namespace sic

using namespace sic;

static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

static int64_t partial_slots(int n_nodes) {
  return 2 * ((int64_t)n_nodes * 8 / SIC_EBE_THREADS + 3 * (int64_t)n_nodes / SIC_VEC_THREADS + 4);
}

extern "C" int64_t sic_ksp_workspace_doubles(int n_nodes, int method) {
  // header + counters + partials (sized for up to 8 cells per node) + vectors
  const int64_t nd = 3 * (int64_t)n_nodes;
  const int64_t nvec = (method == SIC_KSP_BICGSTAB) ? 8 : (method == SIC_KSP_CGCG ? 5 : 4);
  return SIC_WS_HEADER + SIC_WS_COUNTERS + partial_slots(n_nodes) + nvec * nd;
}

static Scal* g_host_scal = nullptr;  // pinned mirror of the device scalars
static cudaEvent_t g_ev[2] = {nullptr, nullptr};

// Brackets one operator launch per batch with CUDA events (measurement only).
struct OpTimer {
  sic_ksp_t* ksp;
  cudaStream_t st;
  bool armed = false;
  OpTimer(sic_ksp_t* k, cudaStream_t s) : ksp(k), st(s) {
    ksp->op_samples = 0;
    ksp->op_ms = 0.0;
    if (ksp->time_operator && !g_ev[0]) { cudaEventCreate(&g_ev[0]); cudaEventCreate(&g_ev[1]); }
  }
  void begin(int k) { if (ksp->time_operator && k == 0) { cudaEventRecord(g_ev[0], st); armed = true; } }
  void end(int k) { if (armed && k == 0) cudaEventRecord(g_ev[1], st); }
  void collect() {  // call after the stream has been synchronised
    if (!armed) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev[0], g_ev[1]) == cudaSuccess) { ksp->op_ms += ms; ksp->op_samples += 1; }
    armed = false;
  }
};

extern "C" int sic_ksp_solve(const sic_problem_t* p, sic_ksp_t* ksp, const double* b_ext, double* x,
                             const uint8_t* fixed, const double* dinv, double* work, const sic_halo_t* halo,
                             void* stream) {
  if (!p || !ksp || !b_ext || !x || !fixed || !dinv || !work) return sic_fail("sic_ksp_solve: null argument");
  if (ksp->method != SIC_KSP_CG && ksp->method != SIC_KSP_BICGSTAB && ksp->method != SIC_KSP_CGCG)
    return sic_fail("sic_ksp_solve: unknown method");
  cudaStream_t st = (cudaStream_t)stream;
  const int nn = p->n_nodes, nd = 3 * nn, nc = p->n_cells;
  const int64_t part = partial_slots(nn);
  if (!g_host_scal) {
    if (int rc = sic_check_cuda(cudaMallocHost((void**)&g_host_scal, sizeof(Scal)), "cudaMallocHost")) return rc;
  }
  const int multi = (halo && halo->n_ranks > 1) ? 1 : 0;
  if (multi && ((!halo->comm && !halo->p2p) || !halo->owner_w))
    return sic_fail("sic_ksp_solve: halo without communicator / owner weights");
  const double* w = multi ? halo->owner_w : nullptr;
  const double* ow_ptr = w;
  Scal* S = (Scal*)work;
  unsigned* counter = (unsigned*)(work + SIC_WS_HEADER);
  double* partials = work + SIC_WS_HEADER + SIC_WS_COUNTERS;
  double* vec = partials + part;
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_WS_HEADER + SIC_WS_COUNTERS), st),
                              "memset ksp header"))
    return rc;
  const int nb = blocks_for(nn, SIC_VEC_THREADS), db = blocks_for(nd, SIC_VEC_THREADS);
#if SIC_EBE_IMPL == 3
  const int cb = blocks_for(nc, SIC_EBE_BLOCK);
  const size_t esm = 0;
  const int ebt = SIC_EBE_BLOCK;
  static bool smem_ok = true;
#else
  const int cb = ebe_grid(nc);
  const size_t esm = sizeof(TileSmem<0>);
  const int ebt = SIC_TILE;
  static bool smem_ok = false;
#endif
  if (!smem_ok) {
    if (int rc = sic_check_cuda(ebe_allow_smem(k_ebe_dot, esm), "smem attr k_ebe_dot")) return rc;
    if (int rc = sic_check_cuda(ebe_allow_smem(k_ebe_plain, esm), "smem attr k_ebe_plain")) return rc;
    smem_ok = true;
  }
  const int check = ksp->check_every > 0 ? ksp->check_every : 25;
  const int guess = ksp->guess_nonzero ? 1 : 0;
  const double rtol = ksp->rtol, atol = ksp->atol;
  auto fin = [&](int op, int slot = 0) { return Fin{S, op, rtol, atol, guess, multi, slot}; };
  // several GPUs: all-reduce S->sum and run the scalar recurrence in a one-thread kernel
  auto reduce = [&](int op, int count, int skip_if_done) -> int {
    if (!multi) return 0;
    if (int rc = sic_exchange(halo, nullptr, 0, S->sum, count, stream)) return rc;
    k_scal<<<1, 1, 0, st>>>(S, op, rtol, atol, guess, skip_if_done);
    return sic_check_launch("k_scal");
  };
  int launched = 0;
  OpTimer timer(ksp, st);

  double *r = vec, *v1 = vec + nd, *v2 = vec + 2 * (size_t)nd, *v3 = vec + 3 * (size_t)nd;
  if (guess) {  // reference norm: residual of the zero guess (prescribed values only)
    k_zero_free<<<db, SIC_VEC_THREADS, 0, st>>>(nd, v1, x, fixed);
    if (int rc = sic_residual0(p, b_ext, v1, r, fixed, halo, stream)) return rc;
    k_norm2<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, w, fin(OP_REF), partials, counter);
    if (int rc = reduce(OP_REF, 1, 0)) return rc;
  }
  if (int rc = sic_residual0(p, b_ext, x, r, fixed, halo, stream)) return rc;

  if (ksp->method == SIC_KSP_CGCG) {
    double *u = v1, *pp = v2, *sv = v3, *w = vec + 4 * (size_t)nd;
    // one operator application + ONE reduction / exchange per iteration
    auto apply_and_reduce = [&](int op, int k) -> int {
      timer.begin(k);
      k_ebe_dot<<<cb, ebt, esm, st>>>(*p, u, w, fin(OP_NONE, 2), partials, counter);
      timer.end(k);
      k_sum_partials<<<1, 1024, 0, st>>>(partials, cb, fin(multi ? OP_NONE : op, 2));
      if (multi) {
        if (int rc = sic_exchange(halo, w, 3, S->sum, 3, stream)) return rc;
        k_scal<<<1, 1, 0, st>>>(S, op, rtol, atol, guess, op == OP_CGCG_STEP ? 1 : 0);
      }
      return sic_check_launch("cgcg apply");
    };
    k_cgcg_init<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, u, pp, sv, w, dinv, ow_ptr, fin(OP_NONE, 0), partials, counter + 1);
    if (int rc = apply_and_reduce(OP_CGCG_INIT, 1)) return rc;
    while (true) {
      cudaMemcpyAsync(g_host_scal, S, sizeof(Scal), cudaMemcpyDeviceToHost, st);
      if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "ksp sync")) return rc;
      timer.collect();
      if (g_host_scal->done || launched >= ksp->max_it) break;
      const int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
      for (int k = 0; k < batch; ++k) {
        k_cgcg_vec<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, x, r, u, pp, sv, w, dinv, fixed, ow_ptr, fin(OP_NONE, 0), partials,
                                                   counter + 1);
        if (int rc = apply_and_reduce(OP_CGCG_STEP, k)) return rc;
      }
      launched += batch;
    }
  } else if (ksp->method == SIC_KSP_CG) {
    double *z = v1, *pp = v2, *q = v3;
    k_cg_init<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, z, pp, q, dinv, w, fin(OP_CG_INIT), partials, counter);
    if (int rc = reduce(OP_CG_INIT, 2, 0)) return rc;
    while (true) {
      cudaMemcpyAsync(g_host_scal, S, sizeof(Scal), cudaMemcpyDeviceToHost, st);
      if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "ksp sync")) return rc;
      timer.collect();
      if (g_host_scal->done || launched >= ksp->max_it) break;
      const int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
      for (int k = 0; k < batch; ++k) {
        timer.begin(k);
        k_ebe_dot<<<cb, ebt, esm, st>>>(*p, pp, q, fin(OP_CG_ALPHA), partials, counter);
        timer.end(k);
#if SIC_EBE_IMPL == 3
        k_sum_partials<<<1, 1024, 0, st>>>(partials, cb, fin(OP_CG_ALPHA));
#endif
        if (multi) {   // halo sum of q and the sum over ranks of p.Kp travel in ONE exchange
          if (int rc = sic_exchange(halo, q, 3, S->sum, 1, stream)) return rc;
          k_scal<<<1, 1, 0, st>>>(S, OP_CG_ALPHA, rtol, atol, guess, 1);
        }
        k_cg_update<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, x, r, z, pp, q, dinv, fixed, w, fin(OP_CG_BETA), partials,
                                                    counter + 1);
        if (int rc = reduce(OP_CG_BETA, 2, 1)) return rc;
        k_cg_p<<<db, SIC_VEC_THREADS, 0, st>>>(nd, pp, z, q, S);
      }
      launched += batch;
      if (int rc = sic_check_launch("cg batch")) return rc;
    }
  } else {
    double *rh = v1, *pp = v2, *v = v3, *s = vec + 4 * (size_t)nd, *t = vec + 5 * (size_t)nd,
           *y = vec + 6 * (size_t)nd, *z = vec + 7 * (size_t)nd;
    k_bi_init<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, rh, pp, y, v, t, dinv, w, fin(OP_BI_INIT), partials, counter);
    if (int rc = reduce(OP_BI_INIT, 1, 0)) return rc;
    while (true) {
      cudaMemcpyAsync(g_host_scal, S, sizeof(Scal), cudaMemcpyDeviceToHost, st);
      if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "ksp sync")) return rc;
      timer.collect();
      if (g_host_scal->done || launched >= ksp->max_it) break;
      const int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
      for (int k = 0; k < batch; ++k) {
        timer.begin(k);
        k_ebe_plain<<<cb, ebt, esm, st>>>(*p, y, v, S);
        timer.end(k);
        if (int rc = sic_exchange(halo, v, 3, nullptr, 0, stream)) return rc;
        k_bi_dot1<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, rh, v, fixed, w, fin(OP_BI_ALPHA), partials, counter);
        if (int rc = reduce(OP_BI_ALPHA, 1, 1)) return rc;
        k_bi_s<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, v, s, z, t, dinv, fixed, S);
        k_ebe_plain<<<cb, ebt, esm, st>>>(*p, z, t, S);
        if (int rc = sic_exchange(halo, t, 3, nullptr, 0, stream)) return rc;
        k_bi_dot2<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, t, s, fixed, w, fin(OP_BI_OMEGA), partials, counter + 1);
        if (int rc = reduce(OP_BI_OMEGA, 2, 1)) return rc;
        k_bi_update<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, x, r, s, t, y, z, rh, fixed, w, fin(OP_BI_BETA), partials,
                                                    counter + 2);
        if (int rc = reduce(OP_BI_BETA, 2, 1)) return rc;
        k_bi_p<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, pp, r, v, y, dinv, fixed, S);
      }
      launched += batch;
      if (int rc = sic_check_launch("bicgstab batch")) return rc;
    }
  }
  if (multi && halo->p2p && sic_p2p_error(halo->p2p)) return sic_fail("P2P exchange timed out waiting for a peer");
  ksp->iterations = g_host_scal->iters;
  ksp->rnorm = sqrt(g_host_scal->rr);
  ksp->rnorm0 = sqrt(g_host_scal->rr0);
  if (g_host_scal->nanflag) ksp->reason = -9;
  else if (g_host_scal->done) ksp->reason = g_host_scal->reason ? g_host_scal->reason : 2;
  else ksp->reason = -3;
  return 0;
}

// ---- Krylov initial guess by extrapolation of the Newton iterates ------------------------------------------------
// The reference starts every KSP solve from zero (MomentumEquation.py:1023-1025).  Its Newton loop
// (Simulators.py:399-436) converges linearly, so the displacement iterates u_1, u_2, ... of a time step are close to a
// short linear recurrence; the next one is predicted from the last three differences
//     d1 = u0 - u1, d0 = u1 - u2, dm = u2 - u3      (u0 = latest solution)
//     d1 ~ a d0 + b dm  (least squares)    =>    guess = u0 + a d1 + b d0
// (one-term version a = <d1,d0>/<d0,d0>, b = 0 when the two-term fit is poor).  Only the STARTING POINT of the Krylov
// solve changes: the solve still runs to rtol relative to the zero-guess residual, so the result is the reference's
// to that tolerance; a model that did not explain the last increment, or a prediction that is not a contraction
// (|a d1 + b d0| >= |d1|), is discarded (guess_coefficients).  Measured on
// cavern_regular (oracle, 4 steps): initial residual 10-100x below the plain warm start from the 4th Newton
// iteration on, below rtol = 1e-10 itself from about the 12th.
namespace sic {

struct GuessScal { double sum[8]; double a, b, fit1, fit2; int used; int pad; };
#define SIC_GUESS_HEADER 16   /* doubles reserved for GuessScal */
#define SIC_GUESS_FIT_MAX 0.1 /* a model is trusted if it explains all but 10 % (squared norm) of the last increment */

// A prediction is only used if its model has just been seen to work on the LAST increment d1:
//   two terms: relative misfit of the least-squares fit d1 ~ a d0 + b dm;
//   one term : the ratio fitted on the PREVIOUS pair, a' = <d0,dm>/<dm,dm>, must have predicted d1 ~ a' d0.
// A Newton sequence that is not a linear recurrence (e.g. one that lands on the fixed point after two iterations,
// the uniform triaxial cube) fails both and keeps the plain warm start.
__device__ __forceinline__ void guess_coefficients(GuessScal* G, int n_iter) {
  const double g11 = G->sum[0], g00 = G->sum[1], g10 = G->sum[2], gmm = G->sum[3], g0m = G->sum[4], g1m = G->sum[5];
  double a = 0.0, b = 0.0, fit1 = 1.0, fit2 = 1.0;
  int used = 0;
  if (n_iter >= 4 && g00 > 0.0 && g11 > 0.0 && gmm > 0.0 && g11 == g11 && g00 == g00 && gmm == gmm) {
    const double det = g00 * gmm - g0m * g0m;
    if (det > 1e-10 * g00 * gmm) {
      const double a2 = (g10 * gmm - g1m * g0m) / det, b2 = (g1m * g00 - g10 * g0m) / det;
      fit2 = (g11 - 2.0 * a2 * g10 - 2.0 * b2 * g1m + a2 * a2 * g00 + 2.0 * a2 * b2 * g0m + b2 * b2 * gmm) / g11;
      const double pred2 = a2 * a2 * g11 + 2.0 * a2 * b2 * g10 + b2 * b2 * g00;   // |a d1 + b d0|^2
      if (fit2 == fit2 && fit2 < SIC_GUESS_FIT_MAX && pred2 == pred2 && pred2 < g11) { a = a2; b = b2; used = 2; }
    }
    if (!used) {
      const double ap = g0m / gmm;
      fit1 = (g11 - 2.0 * ap * g10 + ap * ap * g00) / g11;
      double a1 = g10 / g00;
      if (fit1 == fit1 && fit1 < SIC_GUESS_FIT_MAX && a1 > 0.0) { a = a1 < 0.95 ? a1 : 0.95; b = 0.0; used = 1; }
    }
  }
  G->a = a; G->b = b; G->fit1 = fit1; G->fit2 = fit2; G->used = used;
}

struct GuessFin {
  GuessScal* G; int n_iter; int multi;
  __device__ __forceinline__ void run(const double* tot) const {
#pragma unroll
    for (int k = 0; k < 6; ++k) G->sum[k] = tot[k];
    if (!multi) guess_coefficients(G, n_iter);
  }
};

// one thread per NODE (owner weights are per node): the six inner products of the last three differences
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_guess_dots(int nn, const double* __restrict__ u0, const double* __restrict__ u1,
                                                               const double* __restrict__ u2, const double* __restrict__ u3,
                                                               const double* __restrict__ node_w, GuessFin fin,
                                                               double* __restrict__ partials, unsigned* counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < nn) {
    const double w = node_w ? node_w[i] : 1.0;
    if (w != 0.0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t k = 3 * (size_t)i + c;
        const double a0 = u0[k], a1 = u1[k], a2 = u2[k];
        const double d1 = a0 - a1, d0 = a1 - a2, dm = u3 ? a2 - u3[k] : 0.0;
        v[0] += d1 * d1; v[1] += d0 * d0; v[2] += d1 * d0; v[3] += dm * dm; v[4] += d0 * dm; v[5] += d1 * dm;
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] *= w;
    }
  }
  grid_reduce<6, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot); });
}

__global__ void k_guess_scal(GuessScal* G, int n_iter) { guess_coefficients(G, n_iter); }

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_guess_apply(int nd, const double* __restrict__ u0, const double* __restrict__ u1,
                                                                const double* __restrict__ u2, double* __restrict__ x,
                                                                const GuessScal* __restrict__ G) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  const double a = G->a, b = G->b;
  const double a0 = u0[k], a1 = u1[k];
  x[k] = a0 + (a * (a0 - a1) + b * (a1 - u2[k]));
}

}  // namespace sic

extern "C" int64_t sic_guess_workspace_doubles(int n_nodes) {
  return SIC_GUESS_HEADER + SIC_WS_COUNTERS + 6 * ((int64_t)n_nodes / SIC_VEC_THREADS + 4);
}

extern "C" int sic_guess_extrapolate(int n_nodes, int n_iterates, const double* u0, const double* u1, const double* u2,
                                     const double* u3, double* x, const sic_halo_t* halo, double* work, double* coef_out,
                                     void* stream) {
  if (n_nodes < 0 || !u0 || !u1 || !u2 || !x || !work) return sic_fail("sic_guess_extrapolate: null argument");
  if (n_iterates < 3 || n_iterates > 4 || (n_iterates == 4 && !u3))
    return sic_fail("sic_guess_extrapolate: needs the last 3 or 4 iterates");
  cudaStream_t st = (cudaStream_t)stream;
  const int multi = (halo && halo->n_ranks > 1) ? 1 : 0;
  GuessScal* G = (GuessScal*)work;
  unsigned* counter = (unsigned*)(work + SIC_GUESS_HEADER);
  double* partials = work + SIC_GUESS_HEADER + SIC_WS_COUNTERS;
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_GUESS_HEADER + SIC_WS_COUNTERS), st), "guess memset"))
    return rc;
  if (n_nodes == 0 && !multi) return 0;
  const int nb = blocks_for(n_nodes > 0 ? n_nodes : 1, SIC_VEC_THREADS), db = blocks_for(3 * n_nodes, SIC_VEC_THREADS);
  k_guess_dots<<<nb, SIC_VEC_THREADS, 0, st>>>(n_nodes, u0, u1, u2, n_iterates == 4 ? u3 : nullptr,
                                               multi ? halo->owner_w : nullptr, GuessFin{G, n_iterates, multi}, partials, counter);
  if (int rc = sic_check_launch("k_guess_dots")) return rc;
  if (multi) {      // the same coefficients on every rank: the vectors stay consistent on interface nodes
    if (int rc = sic_exchange(halo, nullptr, 0, G->sum, 6, stream)) return rc;
    k_guess_scal<<<1, 1, 0, st>>>(G, n_iterates);
  }
  if (n_nodes > 0) k_guess_apply<<<db, SIC_VEC_THREADS, 0, st>>>(3 * n_nodes, u0, u1, u2, x, G);
  if (int rc = sic_check_launch("k_guess_apply")) return rc;
  if (coef_out) {   // diagnostics: {a, b, terms used, misfit of the one-term model, of the two-term model}; synchronises
    GuessScal h;
    if (int rc = sic_check_cuda(cudaMemcpyAsync(&h, G, sizeof(GuessScal), cudaMemcpyDeviceToHost, st), "guess copy")) return rc;
    if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "guess sync")) return rc;
    coef_out[0] = h.a; coef_out[1] = h.b; coef_out[2] = (double)h.used; coef_out[3] = h.fit1; coef_out[4] = h.fit2;
  }
  return 0;
}

// ---- FP64 peak micro-benchmark ---------------------------------------------------------------------
namespace sic {
__global__ void __launch_bounds__(256) k_fp64_fma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int k = 0; k < iters; ++k) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace sic

extern "C" int sic_fp64_peak(double* flops, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int sm = 0;
  if (int rc = sic_device_info(&sm, nullptr, nullptr)) return rc;
  const int blocks = sm * 8, threads = 256, iters = 1 << 15;
  double* out = nullptr;
  if (int rc = sic_check_cuda(cudaMalloc(&out, sizeof(double) * blocks * threads), "cudaMalloc")) return rc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fp64_fma<<<blocks, threads, 0, st>>>(out, 1024);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    k_fp64_fma<<<blocks, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (int rc = sic_check_launch("k_fp64_fma")) return rc;
  if (flops) *flops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3);
  return 0;
}
