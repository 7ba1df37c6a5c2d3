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
#include <stdlib.h>
#include <string.h>

#include "../../include/safeincave_cuda.h"
#include "common.cuh"
#include "comm.cuh"

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

// =====================================================================================================
// Peer-to-peer exchange over NVLink (no NCCL): halo sum + scalar all-reduce in ONE kernel
// =====================================================================================================
namespace sic {

// grid = n_peers * bpp halo blocks (+ 1 scalar block when n_scal > 0); all of them must be resident at once (they wait
// for each other and for the peers): <= 16 * 32 + 1 blocks of 256 threads, a B200 holds 148 * 8.
//   halo block (p, c): chunk c of the nodal data exchanged with neighbour p (only ranks that share nodes);
//   scalar block     : thread r sends this rank's n_scal partial sums to rank r (EVERY rank, neighbour or not),
//                      waits for rank r's, then the sums are formed in rank order (identical on all ranks).
// Slot layout per (source rank, parity): [9*cap nodal values | NSCAL scalars | halo flag | scalar flag].
// every block has read the epochs before it takes its ticket, so the last block of the launch may advance them
__device__ __forceinline__ void p2p_block_done(const P2P& ctx, int ncomp, int n_scal, unsigned long long epoch,
                                               unsigned long long epoch_s) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ctx.epochs + 2, 1ull);
    if (t == (unsigned long long)gridDim.x - 1ull) {
      ctx.epochs[2] = 0ull;
      if (ncomp > 0) ctx.epochs[0] = epoch + 1ull;
      if (n_scal > 0) ctx.epochs[1] = epoch_s + 1ull;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(SIC_P2P_THREADS) k_p2p_exchange(sic_halo_t H, P2P ctx, double* __restrict__ vec, int ncomp,
                                                                 double* __restrict__ scal, int n_scal) {
  const unsigned long long epoch = ((volatile unsigned long long*)ctx.epochs)[0];
  const unsigned long long epoch_s = ((volatile unsigned long long*)ctx.epochs)[1];
  // once a wait has timed out the run is lost (sic_p2p_error): later launches do not wait again
  const long long timeout = (*(volatile int*)ctx.error) ? 0ll : SIC_P2P_TIMEOUT_CYCLES;
  const size_t scal_off = 9 * (size_t)ctx.cap;
  const int bpp = ctx.bpp;
  const int n_halo_blocks = (ncomp > 0) ? H.n_peers * bpp : 0;
  __shared__ double mine[SIC_P2P_NSCAL];
  if ((int)blockIdx.x >= n_halo_blocks) {
    // ---------------- scalar block -----------------------------------------------------------------
    const int r = threadIdx.x;
    const unsigned long long epoch_h = epoch;
    const int parity = (int)(epoch_s & 1ull);          // shadows the halo parity
    const unsigned long long epoch = epoch_s;          // and the halo epoch
    if (r < n_scal) mine[r] = scal[r];
    __syncthreads();
    if (r < ctx.n_ranks && r != ctx.rank) {
      double* out = p2p_slot(ctx.remote[r], ctx.slot_doubles, ctx.rank, parity) + scal_off;
      for (int j = 0; j < n_scal; ++j) out[j] = mine[j];
      __threadfence_system();
      *(volatile unsigned long long*)(out + SIC_P2P_NSCAL + 1) = epoch + 1;
      volatile unsigned long long* flag =
          (volatile unsigned long long*)(p2p_slot(ctx.local, ctx.slot_doubles, r, parity) + scal_off + SIC_P2P_NSCAL + 1);
      const long long t0 = clock64();
      while (*flag != epoch + 1) {
        if (clock64() - t0 > timeout) { atomicExch(ctx.error, 1); break; }   // ~2 min
      }
      __threadfence_system();
    }
    __syncthreads();
    if (r < n_scal) {
      double acc = 0.0;
      for (int q = 0; q < ctx.n_ranks; ++q)
        acc += (q == ctx.rank) ? mine[r] : __ldcv(p2p_slot(ctx.local, ctx.slot_doubles, q, parity) + scal_off + r);
      scal[r] = acc;
    }
    p2p_block_done(ctx, ncomp, n_scal, epoch_h, epoch_s);
    return;
  }
  // ---------------- halo block ---------------------------------------------------------------------
  p2p_halo_block(H, ctx, vec, ncomp, blockIdx.x / bpp, blockIdx.x % bpp, epoch, timeout);
  p2p_block_done(ctx, ncomp, n_scal, epoch, epoch_s);
}
}  // namespace sic

extern "C" int sic_p2p_create(int rank, int n_ranks, int cap_nodes, void** p2p, uint8_t* handle64) {
  if (!p2p || !handle64) return sic_fail("sic_p2p_create: null");
  if (n_ranks < 2 || n_ranks > SIC_P2P_MAX_RANKS) return sic_fail("sic_p2p_create: 2..16 ranks");
  sic::P2P* c = new sic::P2P();
  c->rank = rank; c->n_ranks = n_ranks; c->cap = cap_nodes > 0 ? cap_nodes : 1;
  {   // 3-component exchanges are the frequent ones: <= 16 values per thread
    const long long per_block = 16ll * SIC_P2P_THREADS;
    long long b = (3ll * c->cap + per_block - 1) / per_block;
    c->bpp = (int)(b < SIC_P2P_BPP_MIN ? SIC_P2P_BPP_MIN : (b > SIC_P2P_BPP_MAX ? SIC_P2P_BPP_MAX : b));
  }
  c->slot_doubles = 9 * (size_t)c->cap + SIC_P2P_NSCAL + 2;   // + halo flag + scalar flag
  const size_t bytes = sizeof(double) * c->slot_doubles * 2 * n_ranks;
  if (int rc = sic_check_cuda(cudaMalloc((void**)&c->local, bytes), "cudaMalloc mailbox")) return rc;
  if (int rc = sic_check_cuda(cudaMemset(c->local, 0, bytes), "memset mailbox")) return rc;
  const size_t cbytes = sizeof(unsigned) * SIC_P2P_MAX_RANKS + 7 * sizeof(unsigned long long);
  if (int rc = sic_check_cuda(cudaMalloc((void**)&c->counters, cbytes), "cudaMalloc")) return rc;
  cudaMemset(c->counters, 0, cbytes);
  c->gbar = (unsigned long long*)(c->counters + SIC_P2P_MAX_RANKS);
  c->error = (int*)(c->gbar + 1);
  c->epochs = c->gbar + 2;
  cudaIpcMemHandle_t h;
  if (int rc = sic_check_cuda(cudaIpcGetMemHandle(&h, c->local), "cudaIpcGetMemHandle")) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  memcpy(handle64, &h, 64);
  for (int r = 0; r < SIC_P2P_MAX_RANKS; ++r) c->remote[r] = nullptr;
  c->remote[rank] = c->local;
  *p2p = c;
  return 0;
}

extern "C" int sic_p2p_connect(void* p2p, const uint8_t* all_handles) {
  if (!p2p || !all_handles) return sic_fail("sic_p2p_connect: null");
  sic::P2P* c = (sic::P2P*)p2p;
  for (int r = 0; r < c->n_ranks; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    if (int rc = sic_check_cuda(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle"))
      return rc;
    c->remote[r] = (double*)ptr;
  }
  return 0;
}

extern "C" int sic_p2p_error(void* p2p) {   // synchronous: 1 if a wait on a peer's flag timed out
  if (!p2p) return 0;
  int e = 0;
  cudaMemcpy(&e, ((sic::P2P*)p2p)->error, sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}

extern "C" int sic_p2p_destroy(void* p2p) {
  if (!p2p) return 0;
  sic::P2P* c = (sic::P2P*)p2p;
  for (int r = 0; r < c->n_ranks; ++r)
    if (r != c->rank && c->remote[r]) cudaIpcCloseMemHandle(c->remote[r]);
  cudaFree(c->local);
  cudaFree(c->counters);
  delete c;
  return 0;
}

extern "C" int sic_exchange(const sic_halo_t* h, double* vec, int ncomp, double* scal, int n_scal, void* stream) {
  if (!h || h->n_ranks <= 1) return 0;
#ifdef SIC_DEBUG_SWITCHES      // never defined by build.py: timing experiment only (results are wrong)
  static int dbg_skip = -1;
  if (dbg_skip < 0) { const char* e = getenv("SIC_DBG_NOXCHG"); dbg_skip = (e && e[0] == '1') ? 1 : 0; }
  if (dbg_skip) return 0;
#endif
  if (n_scal < 0 || n_scal > SIC_P2P_NSCAL) return sic_fail("sic_exchange: n_scal must be 0..8");
  if (!h->p2p) {   // NCCL path
    if (ncomp > 0) { if (int rc = sic_halo_sum(h, vec, ncomp, stream)) return rc; }
    if (n_scal > 0) return sic_allreduce_sum(h->comm, scal, n_scal, stream);
    return 0;
  }
  sic::P2P* c = (sic::P2P*)h->p2p;
  cudaStream_t st = (cudaStream_t)stream;
  int cap_needed = 0;
  for (int p = 0; p < h->n_peers; ++p) {
    int cnt = h->peer_off[p + 1] - h->peer_off[p];
    if (cnt > cap_needed) cap_needed = cnt;
  }
  if (ncomp > 0 && cap_needed > c->cap) return sic_fail("sic_exchange: mailbox too small for this halo plan");
  if (h->n_ranks > SIC_P2P_THREADS) return sic_fail("sic_exchange: too many ranks");
  const int blocks = (ncomp > 0 ? h->n_peers * c->bpp : 0) + (n_scal > 0 ? 1 : 0);
  // every rank takes this path for every call (nodal data to the neighbours, scalars to everybody), so the
  // (device-resident) epoch counters of all ranks advance in lock step; a rank without neighbours on this halo plan
  // launches nothing for a nodal-only exchange and its nodal epoch is then never looked at by anybody
  if (blocks == 0) return 0;
  sic::k_p2p_exchange<<<blocks, SIC_P2P_THREADS, 0, st>>>(*h, *c, vec, ncomp, scal, n_scal);
  return sic_check_launch("k_p2p_exchange");
}
