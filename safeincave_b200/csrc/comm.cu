// comm.cu — multi-GPU plumbing of the hot path: NCCL communicator (dlopen'ed), interface-node halo sum,
// scalar allreduce.  One process per GPU; traffic goes over NVLink 5 / NVSwitch.
//
// Replaces, for this path, the MPI traffic the reference delegates to DOLFINx/PETSc: the RHS ghost
// accumulate/scatter (MomentumEquation.py:915-917, 1018-1020), the solution scatter (:922, :1025), and the
// MatMult halos + dot-product reductions inside KSP.solve (:1024).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include "../../include/safeincave_cuda.h"
#include "common.cuh"

namespace {
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  // torch has normally loaded its bundled libnccl.so.2 already; the loader reuses it by SONAME
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return sic_fail("cannot dlopen libnccl.so.2 (needed only for multi-GPU runs)");
#define SIC_SYM(field, name)                                                     \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                     \
  if (!g_nccl.field) return sic_fail("libnccl lacks symbol " name);
  SIC_SYM(GetUniqueId, "ncclGetUniqueId")
  SIC_SYM(CommInitRank, "ncclCommInitRank")
  SIC_SYM(CommDestroy, "ncclCommDestroy")
  SIC_SYM(AllReduce, "ncclAllReduce")
  SIC_SYM(Send, "ncclSend")
  SIC_SYM(Recv, "ncclRecv")
  SIC_SYM(GroupStart, "ncclGroupStart")
  SIC_SYM(GroupEnd, "ncclGroupEnd")
  SIC_SYM(GetErrorString, "ncclGetErrorString")
#undef SIC_SYM
  g_nccl.handle = h;
  return 0;
}

int nccl_check(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return 0;
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
  return sic_fail(buf);
}
}  // namespace

namespace sic {
// buf[k*ncomp + c] = vec[idx[k]*ncomp + c]
__global__ void k_halo_pack(int n, int ncomp, const int32_t* __restrict__ idx, const double* __restrict__ vec,
                            double* __restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ncomp) return;
  const int k = t / ncomp, c = t - k * ncomp;
  buf[t] = vec[(size_t)idx[k] * ncomp + c];
}
// vec[idx[k]*ncomp + c] += buf[k*ncomp + c]; a node shared with several neighbours appears once per
// neighbour, hence the atomic (different k may hit the same node).
__global__ void k_halo_unpack_add(int n, int ncomp, const int32_t* __restrict__ idx, double* __restrict__ vec,
                                  const double* __restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ncomp) return;
  const int k = t / ncomp, c = t - k * ncomp;
  atomicAdd(vec + (size_t)idx[k] * ncomp + c, buf[t]);
}
}  // namespace sic

extern "C" int sic_comm_unique_id(uint8_t* id128) {
  if (!id128) return sic_fail("sic_comm_unique_id: null");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  if (int rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int sic_comm_init(const uint8_t* id128, int rank, int n_ranks, void** comm) {
  if (!id128 || !comm) return sic_fail("sic_comm_init: null");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  if (int rc = nccl_check(g_nccl.CommInitRank(&c, n_ranks, id, rank), "ncclCommInitRank")) return rc;
  *comm = (void*)c;
  return 0;
}

extern "C" int sic_comm_destroy(void* comm) {
  if (!comm) return 0;
  if (int rc = load_nccl()) return rc;
  return nccl_check(g_nccl.CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
}

extern "C" int sic_allreduce_sum(void* comm, double* dev_buf, int count, void* stream) {
  if (!comm || !dev_buf) return sic_fail("sic_allreduce_sum: null");
  if (int rc = load_nccl()) return rc;
  return nccl_check(g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)comm,
                                     (cudaStream_t)stream), "ncclAllReduce");
}

extern "C" int sic_halo_sum(const sic_halo_t* h, double* vec, int ncomp, void* stream) {
  if (!h || h->n_ranks <= 1 || h->n_shared_total == 0) return 0;
  if (!vec || !h->idx || !h->send_buf || !h->recv_buf || !h->comm) return sic_fail("sic_halo_sum: null argument");
  if (ncomp < 1 || ncomp > 9) return sic_fail("sic_halo_sum: ncomp must be 1..9");
  if (int rc = load_nccl()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = h->n_shared_total * ncomp;
  const int threads = 256, blocks = (n + threads - 1) / threads;
  sic::k_halo_pack<<<blocks, threads, 0, st>>>(h->n_shared_total, ncomp, h->idx, vec, h->send_buf);
  if (int rc = sic_check_launch("k_halo_pack")) return rc;
  if (int rc = nccl_check(g_nccl.GroupStart(), "ncclGroupStart")) return rc;
  for (int p = 0; p < h->n_peers; ++p) {
    const size_t off = (size_t)h->peer_off[p] * ncomp, cnt = (size_t)(h->peer_off[p + 1] - h->peer_off[p]) * ncomp;
    if (int rc = nccl_check(g_nccl.Send(h->send_buf + off, cnt, ncclDouble, h->peer[p], (ncclComm_t)h->comm, st), "ncclSend"))
      return rc;
    if (int rc = nccl_check(g_nccl.Recv(h->recv_buf + off, cnt, ncclDouble, h->peer[p], (ncclComm_t)h->comm, st), "ncclRecv"))
      return rc;
  }
  if (int rc = nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd")) return rc;
  sic::k_halo_unpack_add<<<blocks, threads, 0, st>>>(h->n_shared_total, ncomp, h->idx, vec, h->recv_buf);
  return sic_check_launch("k_halo_unpack_add");
}
