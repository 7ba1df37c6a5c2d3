// fields.cu — output fields of a converged step: mean stress p and von Mises stress q per cell, smoothed to the nodes
// (volume-weighted average over the incident cells) and back to the cells (mean of the four node values).
// Replaces compute_p_elems / q_elems / p_nodes / q_nodes (MomentumEquation.py:287-324, 944-976) and the scipy CSR
// smoother of Grid.py:198-242.  Runs once per SAVED step, outside the timed hot path.
#include "fem.cuh"

namespace sic {

__global__ void __launch_bounds__(SIC_EBE_THREADS) k_node_volumes(sic_problem_t P, double* __restrict__ node_vol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const size_t ns = (size_t)P.cell_stride;
  const double v = __ldg(P.vol + i);
#pragma unroll
  for (int a = 0; a < 4; ++a) atomicAdd(node_vol + __ldg(P.conn + a * ns + i), v);
}

// accumulate V p and V q of every cell on its four nodes
__global__ void __launch_bounds__(SIC_EBE_THREADS) k_pq_scatter(sic_problem_t P, double* __restrict__ p_nodes,
                                                               double* __restrict__ q_nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const size_t ns = (size_t)P.cell_stride;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) s[c] = P.sig[c * ns + i];
  const double I1 = s[0] + s[1] + s[2];
  const double I2 = s[0] * s[1] + s[1] * s[2] + s[0] * s[2] - s[3] * s[3] - s[4] * s[4] - s[5] * s[5];
  const double J2 = (1.0 / 3.0) * I1 * I1 - I2;
  const double v = __ldg(P.vol + i);
  const double vp = v * (I1 / 3.0), vq = v * sqrt(3.0 * J2);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = __ldg(P.conn + a * ns + i);
    atomicAdd(p_nodes + n, vp);
    atomicAdd(q_nodes + n, vq);
  }
}

__global__ void __launch_bounds__(SIC_VEC_THREADS) k_pq_divide(int n_nodes, const double* __restrict__ node_vol,
                                                              double* __restrict__ p_nodes, double* __restrict__ q_nodes) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const double v = node_vol[n];
  p_nodes[n] = p_nodes[n] / v;
  q_nodes[n] = q_nodes[n] / v;
}

__global__ void __launch_bounds__(SIC_EBE_THREADS) k_pq_cells(sic_problem_t P, const double* __restrict__ p_nodes,
                                                             const double* __restrict__ q_nodes, double* __restrict__ p_elems,
                                                             double* __restrict__ q_elems) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const size_t ns = (size_t)P.cell_stride;
  int nd[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) nd[a] = __ldg(P.conn + a * ns + i);
  if (p_elems) p_elems[i] = 0.25 * ((p_nodes[nd[0]] + p_nodes[nd[1]]) + (p_nodes[nd[2]] + p_nodes[nd[3]]));
  if (q_elems) q_elems[i] = 0.25 * ((q_nodes[nd[0]] + q_nodes[nd[1]]) + (q_nodes[nd[2]] + q_nodes[nd[3]]));
}

}  // namespace sic

using namespace sic;

static inline int fl_blocks(int n, int t) { return (n + t - 1) / t; }

extern "C" int sic_node_volumes(const sic_problem_t* p, double* node_vol, const sic_halo_t* halo, void* stream) {
  if (!p || !node_vol) return sic_fail("sic_node_volumes: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = sic_check_cuda(cudaMemsetAsync(node_vol, 0, sizeof(double) * p->n_nodes, st), "memset node_vol")) return rc;
  if (p->n_cells > 0) k_node_volumes<<<fl_blocks(p->n_cells, SIC_EBE_THREADS), SIC_EBE_THREADS, 0, st>>>(*p, node_vol);
  if (int rc = sic_check_launch("k_node_volumes")) return rc;
  return sic_exchange(halo, node_vol, 1, nullptr, 0, stream);
}

extern "C" int sic_pq_fields(const sic_problem_t* p, const double* node_vol, double* p_nodes, double* q_nodes,
                             double* p_elems, double* q_elems, const sic_halo_t* halo, void* stream) {
  if (!p || !node_vol || !p_nodes || !q_nodes || !p->sig) return sic_fail("sic_pq_fields: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = sic_check_cuda(cudaMemsetAsync(p_nodes, 0, sizeof(double) * p->n_nodes, st), "memset p_nodes")) return rc;
  if (int rc = sic_check_cuda(cudaMemsetAsync(q_nodes, 0, sizeof(double) * p->n_nodes, st), "memset q_nodes")) return rc;
  const int cb = fl_blocks(p->n_cells, SIC_EBE_THREADS);
  if (p->n_cells > 0) k_pq_scatter<<<cb, SIC_EBE_THREADS, 0, st>>>(*p, p_nodes, q_nodes);
  if (int rc = sic_check_launch("k_pq_scatter")) return rc;
  if (int rc = sic_exchange(halo, p_nodes, 1, nullptr, 0, stream)) return rc;
  if (int rc = sic_exchange(halo, q_nodes, 1, nullptr, 0, stream)) return rc;
  if (p->n_nodes > 0) k_pq_divide<<<fl_blocks(p->n_nodes, SIC_VEC_THREADS), SIC_VEC_THREADS, 0, st>>>(p->n_nodes, node_vol, p_nodes, q_nodes);
  if ((p_elems || q_elems) && p->n_cells > 0) k_pq_cells<<<cb, SIC_EBE_THREADS, 0, st>>>(*p, p_nodes, q_nodes, p_elems, q_elems);
  return sic_check_launch("sic_pq_fields");
}
