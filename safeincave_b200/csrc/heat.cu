// heat.cu — SURVEY 8f row 2: the heat equation of the thermo-mechanical path.
//
// Replaces HeatDiffusion.solve (HeatEquation.py:304-343: do.fem.form + assemble_matrix + assemble_vector +
// apply_lifting + set_bc + KSP.solve every step) and get_T_elems (:286-302).  Backward Euler on P1 tets:
//   (M/dt + K + R) T = (M/dt) T_old + q
// evaluated matrix-free with the exact element matrices (consistent mass V/20 (1 + delta_ab), conductivity
// V k g_a . g_b, boundary mass A/12 (1 + delta_ab)); Jacobi-preconditioned CG whose scalars live on the device
// (same pattern as solver.cu / mg.cu: dot products fused into the producing kernel, finished by the last
// block, a `done` flag turns the remaining launches of a batch into no-ops).  The system is mass-dominated at
// the reference's time steps (rho cp h^2 / (k dt) >> 1), so a handful of iterations suffice; this path is
// ~1 % of a thermo-mechanical step and is kept simple (one thread per cell, one FP64 atomic per cell node).
//
// Several GPUs (sic_heat_t.halo, the reference's heat solve is MPI-parallel: HeatEquation.py:344-364 ghostUpdate /
// scatter_forward): cells and boundary triangles are partitioned exactly as for the momentum equation, nodal vectors
// are kept consistent on the interface nodes, every cell / facet sum (operator result, right-hand side, diagonal) is
// completed by the halo sum of sic_exchange (one component), dot products use the owner weights and their partial
// sums travel through the same exchange kernel; the scalar recurrence then runs in a one-thread kernel.
#include <math.h>

#include "fem.cuh"

namespace sic {

struct HeatScal {
  double rz, pq, rr, rr0, rr_ref, alpha, beta, tol2;
  double sum[2];        // several GPUs: this rank's partial sums of the running reduction
  int done, iters, nanflag, reason;
};
static_assert(sizeof(HeatScal) <= 64 * sizeof(double), "HeatScal must fit the reserved workspace header");
#define SIC_HEAT_HEADER 64
#define SIC_HEAT_COUNTERS 8

enum { HT_REF = 0, HT_INIT, HT_PQ, HT_UPDATE };

struct HeatFin {
  HeatScal* S; int op; double rtol, atol; int multi;
  // one GPU: run the scalar recurrence at once; several: park the partial sums, k_heat_scal runs it after the exchange
  __device__ __forceinline__ void run(const double* tot) const {
    if (multi) { S->sum[0] = tot[0]; S->sum[1] = tot[1]; return; }
    step(tot);
  }
  __device__ __forceinline__ void step(const double* tot) const {
    switch (op) {
      case HT_REF: S->rr_ref = tot[0]; break;
      case HT_INIT: {   // tot = {r.z, r.r}
        S->rz = tot[0]; S->rr = tot[1]; S->rr0 = tot[1];
        const double t = rtol * rtol * S->rr_ref, a2 = atol * atol;
        S->tol2 = (t > a2) ? t : a2;
        S->iters = 0; S->nanflag = 0; S->reason = 0; S->done = 0;
        if (!(tot[1] == tot[1]) || isinf(tot[1])) { S->nanflag = 1; S->done = 1; S->reason = -9; }
        else if (tot[1] <= S->tol2 || tot[1] == 0.0) { S->done = 1; S->reason = (tot[1] <= a2) ? 3 : 2; }
        break;
      }
      case HT_PQ: S->pq = tot[0]; S->alpha = S->rz / tot[0]; break;
      case HT_UPDATE:   // tot = {r.z, r.r}
        S->beta = tot[0] / S->rz; S->rz = tot[0]; S->rr = tot[1]; S->iters += 1;
        if (!(tot[1] == tot[1]) || isinf(tot[1])) { S->nanflag = 1; S->done = 1; S->reason = -9; }
        else if (tot[1] <= S->tol2) { S->done = 1; }
        break;
    }
  }
};

__global__ void k_heat_scal(HeatFin fin, int skip_if_done) {
  if (skip_if_done && fin.S->done) return;
  fin.step(fin.S->sum);
}

// y += cm * M x + ck * K x over the cells (one thread per cell)
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_heat_cells(sic_heat_t H, double cm, double ck,
                                                               const double* __restrict__ x, double* __restrict__ y,
                                                               const int* done) {
  if (done && *done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_cells) return;
  const size_t ns = (size_t)H.cell_stride;
  int nd[4];
  double g[12], xa[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) nd[a] = __ldg(H.conn + a * ns + i);
#pragma unroll
  for (int k = 0; k < 12; ++k) g[k] = __ldg(H.grad + k * ns + i);
  const double vol = __ldg(H.vol + i);
#pragma unroll
  for (int a = 0; a < 4; ++a) xa[a] = x[nd[a]];
  const double m = cm * vol * __ldg(H.rho_cp + i) / 20.0;
  const double kk = ck * vol * __ldg(H.k + i);
  const double sum = (xa[0] + xa[1]) + (xa[2] + xa[3]);
  double gx = 0.0, gy = 0.0, gz = 0.0;     // grad T = sum_b x_b g_b
#pragma unroll
  for (int b = 0; b < 4; ++b) { gx += xa[b] * g[3 * b]; gy += xa[b] * g[3 * b + 1]; gz += xa[b] * g[3 * b + 2]; }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double v = m * (sum + xa[a]) + kk * (g[3 * a] * gx + g[3 * a + 1] * gy + g[3 * a + 2] * gz);
    atomicAdd(y + nd[a], v);
  }
}

// y += R x (Robin boundary mass) and/or y += q-load, over the boundary triangles
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_tris(sic_heat_t H, int with_robin, int with_load,
                                                              const double* __restrict__ x, double* __restrict__ y,
                                                              const int* done) {
  if (done && *done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H.n_tri) return;
  const int n0 = H.tri[t], n1 = H.tri[(size_t)H.n_tri + t], n2 = H.tri[2 * (size_t)H.n_tri + t];
  const double A = H.tri_area[t];
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  if (with_robin) {
    const double h = H.tri_h[t];
    if (h != 0.0) {
      const double w = h * A / 12.0, x0 = x[n0], x1 = x[n1], x2 = x[n2], s = (x0 + x1) + x2;
      v0 += w * (s + x0); v1 += w * (s + x1); v2 += w * (s + x2);
    }
  }
  if (with_load) {
    const double q = H.tri_q[t] * A / 3.0;
    v0 += q; v1 += q; v2 += q;
  }
  if (v0 != 0.0 || v1 != 0.0 || v2 != 0.0) {
    atomicAdd(y + n0, v0); atomicAdd(y + n1, v1); atomicAdd(y + n2, v2);
  }
}

// diagonal of M/dt + K: per cell node 2 (cm V rho cp / 20) + V k |g_a|^2 ; the Robin boundary mass adds 2 (h A / 12) per
// facet node (k_heat_diag_tris)
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_heat_diag_cells(sic_heat_t H, double cm, double* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_cells) return;
  const size_t ns = (size_t)H.cell_stride;
  const double vol = __ldg(H.vol + i);
  const double m = cm * vol * __ldg(H.rho_cp + i) / 10.0;     // 2 m_ab(offdiag) = V rho cp / 10
  const double kk = vol * __ldg(H.k + i);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double gx = __ldg(H.grad + (3 * a) * ns + i), gy = __ldg(H.grad + (3 * a + 1) * ns + i),
                 gz = __ldg(H.grad + (3 * a + 2) * ns + i);
    atomicAdd(d + __ldg(H.conn + a * ns + i), m + kk * (gx * gx + gy * gy + gz * gz));
  }
}
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_diag_tris(sic_heat_t H, double* __restrict__ d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H.n_tri) return;
  const double h = H.tri_h[t];
  if (h == 0.0) return;
  const double w = h * H.tri_area[t] / 6.0;
  atomicAdd(d + H.tri[t], w);
  atomicAdd(d + H.tri[(size_t)H.n_tri + t], w);
  atomicAdd(d + H.tri[2 * (size_t)H.n_tri + t], w);
}

// r = fixed ? 0 : b - t ; (guess pass) z = Dinv r ; p = z ; q = 0 ; sums r.z, r.r      [t = A x0]
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_init(int n, const double* __restrict__ b, const double* __restrict__ t,
                                                              const double* __restrict__ diag, const uint8_t* __restrict__ fixed,
                                                              double* __restrict__ r, double* __restrict__ p,
                                                              double* __restrict__ q, const double* __restrict__ w,
                                                              HeatFin fin, int only_norm,
                                                              double* __restrict__ partials, unsigned* counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (i < n) {
    const double wn = w ? w[i] : 1.0;       // several GPUs: every node is counted by its owner only
    const double rn = fixed[i] ? 0.0 : b[i] - t[i];
    if (only_norm) { v[0] = wn * rn * rn; }
    else {
      const double zn = fixed[i] ? 0.0 : rn / diag[i];
      r[i] = rn; p[i] = zn; q[i] = 0.0;
      v[0] = wn * rn * zn; v[1] = wn * rn * rn;
    }
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot); });
}

// p.q over the free nodes
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_pq(int n, const double* __restrict__ p, const double* __restrict__ q,
                                                            const uint8_t* __restrict__ fixed, const double* __restrict__ w,
                                                            HeatFin fin, double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (i < n && !fixed[i]) v[0] = (w ? w[i] : 1.0) * p[i] * q[i];
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot); });
}

// x += alpha p ; r -= alpha q ; z = Dinv r ; sums r.z, r.r
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_update(int n, double* __restrict__ x, double* __restrict__ r,
                                                                double* __restrict__ z, const double* __restrict__ p,
                                                                const double* __restrict__ q, const double* __restrict__ diag,
                                                                const uint8_t* __restrict__ fixed, const double* __restrict__ w,
                                                                HeatFin fin, double* __restrict__ partials, unsigned* counter) {
  if (fin.S->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double alpha = fin.S->alpha;
  double v[2] = {0.0, 0.0};
  if (i < n) {
    const double wn = w ? w[i] : 1.0;
    double rn = 0.0, zn = 0.0;
    if (!fixed[i]) {
      x[i] += alpha * p[i];
      rn = r[i] - alpha * q[i];
      zn = rn / diag[i];
    }
    r[i] = rn; z[i] = zn;
    v[0] = wn * rn * zn; v[1] = wn * rn * rn;
  }
  grid_reduce<2, SIC_VEC_THREADS>(v, partials, counter, [&](const double* tot) { fin.run(tot); });
}

// p = z + beta p ; q = 0
__global__ void __launch_bounds__(SIC_VEC_THREADS) k_heat_p(int n, double* __restrict__ p, const double* __restrict__ z,
                                                           double* __restrict__ q, const HeatScal* S) {
  if (S->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = z[i] + S->beta * p[i];
  q[i] = 0.0;
}

// z = fixed ? x : 0
__global__ void k_heat_zero_free(int n, double* __restrict__ z, const double* __restrict__ x, const uint8_t* __restrict__ fixed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = fixed[i] ? x[i] : 0.0;
}

__global__ void __launch_bounds__(SIC_EBE_THREADS) k_heat_cell_mean(sic_heat_t H, const double* __restrict__ Tn,
                                                                   double* __restrict__ Tc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.n_cells) return;
  const size_t ns = (size_t)H.cell_stride;
  const double a = Tn[__ldg(H.conn + i)], b = Tn[__ldg(H.conn + ns + i)], c = Tn[__ldg(H.conn + 2 * ns + i)],
               d = Tn[__ldg(H.conn + 3 * ns + i)];
  Tc[i] = 0.25 * ((a + b) + (c + d));
}

}  // namespace sic

using namespace sic;

static inline int ht_blocks(int n, int t) { return (n + t - 1) / t; }
static int64_t ht_partials(int n_nodes) { return 2 * ((int64_t)n_nodes / SIC_VEC_THREADS + 4); }

extern "C" int64_t sic_heat_workspace_doubles(int n_nodes) {
  return SIC_HEAT_HEADER + SIC_HEAT_COUNTERS + ht_partials(n_nodes) + 7 * (int64_t)n_nodes;
}

static HeatScal* g_heat_host = nullptr;

static int heat_check(const sic_heat_t* h) {
  if (!h) return sic_fail("heat: null problem");
  if (h->n_cells < 0 || h->cell_stride < h->n_cells || h->n_nodes < 0 || h->n_tri < 0) return sic_fail("heat: bad sizes");
  if (!h->conn || !h->grad || !h->vol || !h->rho_cp || !h->k || !h->fixed) return sic_fail("heat: null mesh / material array");
  if (h->n_tri > 0 && (!h->tri || !h->tri_area || !h->tri_h || !h->tri_q)) return sic_fail("heat: null boundary array");
  return 0;
}

static inline const sic_halo_t* heat_halo(const sic_heat_t* h) { return (h->halo && h->halo->n_ranks > 1) ? h->halo : nullptr; }

// y = A x = (M/dt + K + R) x   (y zeroed here; several GPUs: completed on the interface nodes by the halo sum)
static int heat_apply(const sic_heat_t* h, double inv_dt, const double* x, double* y, const int* done, cudaStream_t st) {
  if (int rc = sic_check_cuda(cudaMemsetAsync(y, 0, sizeof(double) * h->n_nodes, st), "heat memset")) return rc;
  if (h->n_cells > 0)
    k_heat_cells<<<ht_blocks(h->n_cells, SIC_EBE_THREADS), SIC_EBE_THREADS, 0, st>>>(*h, inv_dt, 1.0, x, y, done);
  if (h->n_tri > 0) k_heat_tris<<<ht_blocks(h->n_tri, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(*h, 1, 0, x, y, done);
  if (int rc = sic_check_launch("heat apply")) return rc;
  if (const sic_halo_t* halo = heat_halo(h)) return sic_exchange(halo, y, 1, nullptr, 0, (void*)st);
  return 0;
}

extern "C" int sic_heat_step(const sic_heat_t* h, double dt, const double* T_old, double* T, sic_ksp_t* ksp, double* work,
                             void* stream) {
  if (int rc = heat_check(h)) return rc;
  if (!T_old || !T || !ksp || !work) return sic_fail("sic_heat_step: null argument");
  if (!(dt > 0.0)) return sic_fail("sic_heat_step: dt must be positive");
  if (!g_heat_host)
    if (int rc = sic_check_cuda(cudaMallocHost((void**)&g_heat_host, sizeof(HeatScal)), "cudaMallocHost")) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = h->n_nodes;
  const double inv_dt = 1.0 / dt;
  HeatScal* S = (HeatScal*)work;
  unsigned* counter = (unsigned*)(work + SIC_HEAT_HEADER);
  double* partials = work + SIC_HEAT_HEADER + SIC_HEAT_COUNTERS;
  double* vec = partials + ht_partials(n);
  double *b = vec, *r = vec + n, *z = vec + 2 * (size_t)n, *p = vec + 3 * (size_t)n, *q = vec + 4 * (size_t)n,
         *diag = vec + 5 * (size_t)n, *tmp = vec + 6 * (size_t)n;
  if (int rc = sic_check_cuda(cudaMemsetAsync(work, 0, sizeof(double) * (SIC_HEAT_HEADER + SIC_HEAT_COUNTERS), st), "heat memset"))
    return rc;
  if (int rc = sic_check_cuda(cudaMemsetAsync(vec, 0, sizeof(double) * 7 * (size_t)n, st), "heat memset")) return rc;
  const int nb = ht_blocks(n, SIC_VEC_THREADS), cb = ht_blocks(h->n_cells, SIC_EBE_THREADS),
            tb = ht_blocks(h->n_tri, SIC_VEC_THREADS);
  const double rtol = ksp->rtol, atol = ksp->atol;
  const sic_halo_t* halo = heat_halo(h);
  const int multi = halo ? 1 : 0;
  const double* ow = halo ? halo->owner_w : nullptr;
  if (halo && !ow) return sic_fail("sic_heat_step: halo without owner weights");
  auto fin = [&](int op) { return HeatFin{S, op, rtol, atol, multi}; };
  // several GPUs: add the partial sums of the last reducing kernel over the ranks, then run the scalar recurrence
  auto reduce = [&](int op, int skip_if_done) -> int {
    if (!multi) return 0;
    if (int rc = sic_exchange(halo, nullptr, 0, S->sum, 2, stream)) return rc;
    k_heat_scal<<<1, 1, 0, st>>>(fin(op), skip_if_done);
    return sic_check_launch("k_heat_scal");
  };
  // b = (M/dt) T_old + q ; diag of A
  if (h->n_cells > 0) {
    k_heat_cells<<<cb, SIC_EBE_THREADS, 0, st>>>(*h, inv_dt, 0.0, T_old, b, nullptr);
    k_heat_diag_cells<<<cb, SIC_EBE_THREADS, 0, st>>>(*h, inv_dt, diag);
  }
  if (h->n_tri > 0) {
    k_heat_tris<<<tb, SIC_VEC_THREADS, 0, st>>>(*h, 0, 1, T_old, b, nullptr);
    k_heat_diag_tris<<<tb, SIC_VEC_THREADS, 0, st>>>(*h, diag);
  }
  if (int rc = sic_check_launch("heat rhs")) return rc;
  if (multi) {
    if (int rc = sic_exchange(halo, b, 1, nullptr, 0, stream)) return rc;
    if (int rc = sic_exchange(halo, diag, 1, nullptr, 0, stream)) return rc;
  }
  // reference norm of rtol: the residual of the zero guess (prescribed values only), PETSc's ||b|| after lifting
  k_heat_zero_free<<<nb, SIC_VEC_THREADS, 0, st>>>(n, tmp, T, h->fixed);
  if (int rc = heat_apply(h, inv_dt, tmp, q, nullptr, st)) return rc;
  k_heat_init<<<nb, SIC_VEC_THREADS, 0, st>>>(n, b, q, diag, h->fixed, r, p, z, ow, fin(HT_REF), 1, partials, counter);
  if (int rc = reduce(HT_REF, 0)) return rc;
  // r0 = b - A T(guess) on the free nodes
  if (int rc = heat_apply(h, inv_dt, T, tmp, nullptr, st)) return rc;
  k_heat_init<<<nb, SIC_VEC_THREADS, 0, st>>>(n, b, tmp, diag, h->fixed, r, p, q, ow, fin(HT_INIT), 0, partials, counter);
  if (int rc = reduce(HT_INIT, 0)) return rc;
  const int check = ksp->check_every > 0 ? ksp->check_every : 10;
  int launched = 0;
  while (true) {
    cudaMemcpyAsync(g_heat_host, S, sizeof(HeatScal), cudaMemcpyDeviceToHost, st);
    if (int rc = sic_check_cuda(cudaStreamSynchronize(st), "heat sync")) return rc;
    if (g_heat_host->done || launched >= ksp->max_it) break;
    const int batch = ksp->max_it - launched < check ? ksp->max_it - launched : check;
    for (int k = 0; k < batch; ++k) {
      // q = A p (q is zero on entry: k_heat_init / k_heat_p leave it so)
      if (h->n_cells > 0) k_heat_cells<<<cb, SIC_EBE_THREADS, 0, st>>>(*h, inv_dt, 1.0, p, q, &S->done);
      if (h->n_tri > 0) k_heat_tris<<<tb, SIC_VEC_THREADS, 0, st>>>(*h, 1, 0, p, q, &S->done);
      if (multi) if (int rc = sic_exchange(halo, q, 1, nullptr, 0, stream)) return rc;
      k_heat_pq<<<nb, SIC_VEC_THREADS, 0, st>>>(n, p, q, h->fixed, ow, fin(HT_PQ), partials, counter);
      if (int rc = reduce(HT_PQ, 1)) return rc;
      k_heat_update<<<nb, SIC_VEC_THREADS, 0, st>>>(n, T, r, z, p, q, diag, h->fixed, ow, fin(HT_UPDATE), partials, counter);
      if (int rc = reduce(HT_UPDATE, 1)) return rc;
      k_heat_p<<<nb, SIC_VEC_THREADS, 0, st>>>(n, p, z, q, S);
    }
    launched += batch;
    if (int rc = sic_check_launch("heat cg batch")) return rc;
  }
  if (multi && halo->p2p && sic_p2p_error(halo->p2p)) return sic_fail("P2P exchange timed out waiting for a peer");
  ksp->iterations = g_heat_host->iters;
  ksp->rnorm = sqrt(g_heat_host->rr);
  ksp->rnorm0 = sqrt(g_heat_host->rr0);
  if (g_heat_host->nanflag) ksp->reason = -9;
  else if (g_heat_host->done) ksp->reason = g_heat_host->reason ? g_heat_host->reason : 2;
  else ksp->reason = -3;
  return 0;
}

extern "C" int sic_heat_cell_mean(const sic_heat_t* h, const double* T_nodes, double* T_cells, void* stream) {
  if (int rc = heat_check(h)) return rc;
  if (!T_nodes || !T_cells) return sic_fail("sic_heat_cell_mean: null argument");
  if (h->n_cells == 0) return 0;
  k_heat_cell_mean<<<ht_blocks(h->n_cells, SIC_EBE_THREADS), SIC_EBE_THREADS, 0, (cudaStream_t)stream>>>(*h, T_nodes, T_cells);
  return sic_check_launch("k_heat_cell_mean");
}
