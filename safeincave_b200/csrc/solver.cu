// solver.cu — hot-path part (3): block-Jacobi preconditioned CG / BiCGStab on the matrix-free operator.
//
// Replaces `self.solver.solve(b, X)` of MomentumEquation.py:1023-1025 (and :920-922), which in the
// reference is a user-configured PETSc KSP (cg / bicg / bcgs / gmres + asm).
// B200 design: the whole iteration lives on the device.  Scalars (alpha, beta, rho, omega, norms)
// stay in device memory, every dot product is fused into the kernel that produces its operand
// (p.Kp is accumulated per CELL inside the operator kernel at no extra memory traffic) and is
// finished by the last block to arrive, the convergence test sets a device flag that turns the
// remaining launches of a batch into no-ops, and the host only looks every `check_every`
// iterations.  Dirichlet rows/columns are handled by keeping search directions zero on constrained
// dofs (equivalent to assemble_matrix(bcs) + apply_lifting + set_bc, MomentumEquation.py:1010-1020).
#include <math.h>
#include <string.h>

#include "fem.cuh"

namespace sic {

struct Scal {
  double rz, pq, rr, rr0, alpha, beta, tol2;
  double rho, rhv, omega, ts, tt;
  int done, iters, nanflag, reason;
};
static_assert(sizeof(Scal) <= 64 * sizeof(double), "Scal must fit the reserved workspace header");

#define SIC_WS_HEADER 64   /* doubles reserved for Scal */
#define SIC_WS_COUNTERS 8  /* doubles reserved for ticket counters */

__device__ __forceinline__ void check_convergence(Scal* S, double rr) {
  S->rr = rr;
  S->iters += 1;
  if (!(rr == rr) || isinf(rr)) { S->nanflag = 1; S->done = 1; S->reason = -9; }
  else if (rr <= S->tol2) { S->done = 1; }
}

// ---- operator kernel with the fused p.Kp reduction --------------------------------------------
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_ebe_dot(sic_problem_t P, const double* __restrict__ x,
                                                            double* __restrict__ y, Scal* S,
                                                            double* __restrict__ partials, unsigned* counter) {
  if (S->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (i < P.n_cells) {
    CellGeom c;
    load_geom(P, i, c);
    double ua[12];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int j = 0; j < 3; ++j) ua[3 * a + j] = __ldg(x + 3 * (size_t)c.node[a] + j);
    }
    double eps[6], sig[6], f[12];
    strain_from_nodal(c, ua, eps);
    stress_from_CT(P, i, eps, sig);
    forces(c, sig, f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(y + 3 * (size_t)c.node[a] + j, f[3 * a + j]);
    }
    // x_e^T K_e x_e = V eps : sigma  (shear terms counted twice)
    v[0] = c.vol * ((eps[0] * sig[0] + eps[1] * sig[1] + eps[2] * sig[2]) +
                    2.0 * (eps[3] * sig[3] + eps[4] * sig[4] + eps[5] * sig[5]));
  }
  grid_reduce<1, SIC_EBE_THREADS>(v, partials, counter, [&](const double* tot) {
    S->pq = tot[0];
    S->alpha = S->rz / tot[0];
  });
}

// plain operator (no dot), skipping when converged
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_ebe_plain(sic_problem_t P, const double* __restrict__ x,
                                                              double* __restrict__ y, const Scal* S) {
  if (S->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  CellGeom c;
  load_geom(P, i, c);
  double ua[12];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int j = 0; j < 3; ++j) ua[3 * a + j] = __ldg(x + 3 * (size_t)c.node[a] + j);
  }
  double eps[6], sig[6], f[12];
  strain_from_nodal(c, ua, eps);
  stress_from_CT(P, i, eps, sig);
  forces(c, sig, f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(y + 3 * (size_t)c.node[a] + j, f[3 * a + j]);
  }
}

__device__ __forceinline__ void precond3(const double* __restrict__ dinv, size_t n, const double r[3], double z[3]) {
  const double* d = dinv + 9 * n;
#pragma unroll
  for (int j = 0; j < 3; ++j) z[j] = __ldg(d + 3 * j) * r[0] + __ldg(d + 3 * j + 1) * r[1] + __ldg(d + 3 * j + 2) * r[2];
}

// ---- PCG ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_init(int n_nodes, const double* __restrict__ r,
                                                            double* __restrict__ z, double* __restrict__ p,
                                                            double* __restrict__ q, const double* __restrict__ dinv,
                                                            Scal* S, double rtol, double atol,
                                                            double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (n < n_nodes) {
    double rn[3], zn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) rn[j] = r[3 * (size_t)n + j];
    precond3(dinv, n, rn, zn);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      z[3 * (size_t)n + j] = zn[j];
      p[3 * (size_t)n + j] = zn[j];
      q[3 * (size_t)n + j] = 0.0;
      v[0] += rn[j] * zn[j];
      v[1] += rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) {
    S->rz = tot[0];
    S->rr = tot[1];
    S->rr0 = tot[1];
    double t = rtol * rtol * tot[1];
    double a2 = atol * atol;
    S->tol2 = (t > a2) ? t : a2;
    S->iters = 0;
    S->nanflag = 0;
    S->reason = 0;
    S->done = 0;
    if (!(tot[1] == tot[1]) || isinf(tot[1])) { S->nanflag = 1; S->done = 1; S->reason = -9; }
    else if (tot[1] <= a2 || tot[1] == 0.0) { S->done = 1; S->reason = 3; }
  });
}

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_update(int n_nodes, double* __restrict__ x,
                                                              double* __restrict__ r, double* __restrict__ z,
                                                              const double* __restrict__ p,
                                                              const double* __restrict__ q,
                                                              const double* __restrict__ dinv,
                                                              const uint8_t* __restrict__ fixed, Scal* S,
                                                              double* __restrict__ partials, unsigned* counter) {
  if (S->done) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = S->alpha;
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
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      z[3 * (size_t)n + j] = zn[j];
      v[0] += rn[j] * zn[j];
      v[1] += rn[j] * rn[j];
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) {
    S->beta = tot[0] / S->rz;
    S->rz = tot[0];
    check_convergence(S, tot[1]);
  });
}

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_cg_p(int nd, double* __restrict__ p, const double* __restrict__ z,
                                                         double* __restrict__ q, const Scal* S) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  p[d] = z[d] + S->beta * p[d];
  q[d] = 0.0;
}

// ---- BiCGStab (right-preconditioned) -------------------------------------------------------------
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_init(int n_nodes, const double* __restrict__ r,
                                                            double* __restrict__ rh, double* __restrict__ p,
                                                            double* __restrict__ y, double* __restrict__ v,
                                                            double* __restrict__ t, const double* __restrict__ dinv,
                                                            Scal* S, double rtol, double atol,
                                                            double* __restrict__ partials, unsigned* counter) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[1] = {0.0};
  if (n < n_nodes) {
    double rn[3], yn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) rn[j] = r[3 * (size_t)n + j];
    precond3(dinv, n, rn, yn);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const size_t d = 3 * (size_t)n + j;
      rh[d] = rn[j]; p[d] = rn[j]; y[d] = yn[j]; v[d] = 0.0; t[d] = 0.0;
      acc[0] += rn[j] * rn[j];
    }
  }
  grid_reduce<1, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) {
    S->rho = tot[0];
    S->rr = tot[0];
    S->rr0 = tot[0];
    double tt = rtol * rtol * tot[0];
    double a2 = atol * atol;
    S->tol2 = (tt > a2) ? tt : a2;
    S->iters = 0; S->nanflag = 0; S->reason = 0; S->done = 0;
    S->alpha = 1.0; S->omega = 1.0;
    if (!(tot[0] == tot[0]) || isinf(tot[0])) { S->nanflag = 1; S->done = 1; S->reason = -9; }
    else if (tot[0] <= a2 || tot[0] == 0.0) { S->done = 1; S->reason = 3; }
  });
}

// rhv = rh . v (free dofs)  ->  alpha = rho / rhv
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_dot1(int nd, const double* __restrict__ rh,
                                                            const double* __restrict__ v,
                                                            const uint8_t* __restrict__ fixed, Scal* S,
                                                            double* __restrict__ partials, unsigned* counter) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[1] = {0.0};
  if (d < nd && !fixed[d]) acc[0] = rh[d] * v[d];
  grid_reduce<1, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) {
    S->rhv = tot[0];
    S->alpha = S->rho / tot[0];
  });
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
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_dot2(int nd, const double* __restrict__ t,
                                                            const double* __restrict__ s,
                                                            const uint8_t* __restrict__ fixed, Scal* S,
                                                            double* __restrict__ partials, unsigned* counter) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (d < nd && !fixed[d]) { acc[0] = t[d] * s[d]; acc[1] = t[d] * t[d]; }
  grid_reduce<2, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) {
    S->ts = tot[0];
    S->tt = tot[1];
    S->omega = tot[0] / tot[1];
  });
}

// x += alpha y + omega z ; r = s - omega t ; rho_new = rh.r ; rr = r.r -> beta
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_bi_update(int nd, double* __restrict__ x, double* __restrict__ r,
                                                              const double* __restrict__ s,
                                                              const double* __restrict__ t,
                                                              const double* __restrict__ y,
                                                              const double* __restrict__ z,
                                                              const double* __restrict__ rh,
                                                              const uint8_t* __restrict__ fixed, Scal* S,
                                                              double* __restrict__ partials, unsigned* counter) {
  if (S->done) return;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = S->alpha, omega = S->omega;
  double acc[2] = {0.0, 0.0};
  if (d < nd) {
    double rn = 0.0;
    if (!fixed[d]) {
      x[d] += alpha * y[d] + omega * z[d];
      rn = s[d] - omega * t[d];
    }
    r[d] = rn;
    acc[0] = rh[d] * rn;
    acc[1] = rn * rn;
  }
  grid_reduce<2, SIC_VEC_THREADS>(acc, partials, counter, [&](const double* tot) {
    S->beta = (tot[0] / S->rho) * (S->alpha / S->omega);
    S->rho = tot[0];
    check_convergence(S, tot[1]);
  });
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

static int64_t partial_doubles(int n_cells, int n_nodes) {
  int nb = blocks_for(n_cells, SIC_EBE_THREADS);
  int nv = blocks_for(3 * n_nodes, SIC_VEC_THREADS);
  int m = nb > nv ? nb : nv;
  return 2 * (int64_t)(m + 1);
}

extern "C" int64_t sic_ksp_workspace_doubles(int n_nodes, int method) {
  // header + counters + partials (sized for the worst case of 8 cells per node) + vectors
  int64_t nd = 3 * (int64_t)n_nodes;
  int64_t nvec = (method == SIC_KSP_BICGSTAB) ? 8 : 4;
  int64_t part = 2 * ((int64_t)n_nodes * 8 / SIC_EBE_THREADS + nd / SIC_VEC_THREADS + 4);
  return SIC_WS_HEADER + SIC_WS_COUNTERS + part + nvec * nd;
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
                             const uint8_t* fixed, const double* dinv, double* work, void* stream) {
  if (!p || !ksp || !b_ext || !x || !fixed || !dinv || !work) return sic_fail("sic_ksp_solve: null argument");
  if (ksp->method != SIC_KSP_CG && ksp->method != SIC_KSP_BICGSTAB) return sic_fail("sic_ksp_solve: unknown method");
  cudaStream_t st = (cudaStream_t)stream;
  const int nn = p->n_nodes, nd = 3 * nn, nc = p->n_cells;
  const int64_t part = 2 * ((int64_t)nn * 8 / SIC_EBE_THREADS + (int64_t)nd / SIC_VEC_THREADS + 4);
  if (partial_doubles(nc, nn) > part) return sic_fail("sic_ksp_solve: more than 8 cells per node on average");
  if (!g_host_scal) {
    if (int rc = sic_check_cuda(cudaMallocHost((void**)&g_host_scal, sizeof(Scal)), "cudaMallocHost")) return rc;
  }
  Scal* S = (Scal*)work;
  unsigned* counter = (unsigned*)(work + SIC_WS_HEADER);
  double* partials = work + SIC_WS_HEADER + SIC_WS_COUNTERS;
  double* vec = partials + part;
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_WS_HEADER + SIC_WS_COUNTERS), st),
                              "memset ksp header"))
    return rc;
  const int cb = blocks_for(nc, SIC_EBE_THREADS), nb = blocks_for(nn, SIC_VEC_THREADS),
            db = blocks_for(nd, SIC_VEC_THREADS);
  const int check = ksp->check_every > 0 ? ksp->check_every : 25;
  int launched = 0;
  OpTimer timer(ksp, st);
  if (ksp->method == SIC_KSP_CG) {
    double *r = vec, *z = vec + nd, *pp = vec + 2 * (size_t)nd, *q = vec + 3 * (size_t)nd;
    if (int rc = sic_residual0(p, b_ext, x, r, fixed, stream)) return rc;
    k_cg_init<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, z, pp, q, dinv, S, ksp->rtol, ksp->atol, partials, counter);
    while (true) {
      cudaMemcpyAsync(g_host_scal, S, sizeof(Scal), cudaMemcpyDeviceToHost, st);
      if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "ksp sync")) return rc;
      timer.collect();
      if (g_host_scal->done || launched >= ksp->max_it) break;
      int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
      for (int k = 0; k < batch; ++k) {
        timer.begin(k);
        k_ebe_dot<<<cb, SIC_EBE_THREADS, 0, st>>>(*p, pp, q, S, partials, counter);
        timer.end(k);
        k_cg_update<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, x, r, z, pp, q, dinv, fixed, S, partials, counter + 1);
        k_cg_p<<<db, SIC_VEC_THREADS, 0, st>>>(nd, pp, z, q, S);
      }
      launched += batch;
      if (int rc = sic_check_launch("cg batch")) return rc;
    }
  } else {
    double *r = vec, *rh = vec + nd, *pp = vec + 2 * (size_t)nd, *v = vec + 3 * (size_t)nd, *s = vec + 4 * (size_t)nd,
           *t = vec + 5 * (size_t)nd, *y = vec + 6 * (size_t)nd, *z = vec + 7 * (size_t)nd;
    if (int rc = sic_residual0(p, b_ext, x, r, fixed, stream)) return rc;
    k_bi_init<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, rh, pp, y, v, t, dinv, S, ksp->rtol, ksp->atol, partials, counter);
    while (true) {
      cudaMemcpyAsync(g_host_scal, S, sizeof(Scal), cudaMemcpyDeviceToHost, st);
      if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "ksp sync")) return rc;
      timer.collect();
      if (g_host_scal->done || launched >= ksp->max_it) break;
      int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
      for (int k = 0; k < batch; ++k) {
        timer.begin(k);
        k_ebe_plain<<<cb, SIC_EBE_THREADS, 0, st>>>(*p, y, v, S);
        timer.end(k);
        k_bi_dot1<<<db, SIC_VEC_THREADS, 0, st>>>(nd, rh, v, fixed, S, partials, counter);
        k_bi_s<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, r, v, s, z, t, dinv, fixed, S);
        k_ebe_plain<<<cb, SIC_EBE_THREADS, 0, st>>>(*p, z, t, S);
        k_bi_dot2<<<db, SIC_VEC_THREADS, 0, st>>>(nd, t, s, fixed, S, partials, counter + 1);
        k_bi_update<<<db, SIC_VEC_THREADS, 0, st>>>(nd, x, r, s, t, y, z, rh, fixed, S, partials, counter + 2);
        k_bi_p<<<nb, SIC_VEC_THREADS, 0, st>>>(nn, pp, r, v, y, dinv, fixed, S);
      }
      launched += batch;
      if (int rc = sic_check_launch("bicgstab batch")) return rc;
    }
  }
  ksp->iterations = g_host_scal->iters;
  ksp->rnorm = sqrt(g_host_scal->rr);
  ksp->rnorm0 = sqrt(g_host_scal->rr0);
  if (g_host_scal->nanflag) ksp->reason = -9;
  else if (g_host_scal->done) ksp->reason = g_host_scal->reason ? g_host_scal->reason : 2;
  else ksp->reason = -3;
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
