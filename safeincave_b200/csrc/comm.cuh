// comm.cuh -- device side of the peer-to-peer exchange over NVLink (csrc/comm.cu), shared with the kernels that FUSE the
// halo exchange into the operator launch (csrc/mg.cu: k_mg_ebe_pc_x).  Not compiled by the host emulation.
#ifndef SIC_COMM_CUH_
#define SIC_COMM_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/safeincave_cuda.h"

namespace sic {

#define SIC_P2P_MAX_RANKS 16
#define SIC_P2P_BPP_MIN 4      /* blocks per peer: at least this many, ... */
#define SIC_P2P_BPP_MAX 32     /* ... at most this many, sized so that a thread moves <= 16 values (P2P.bpp) */
#define SIC_P2P_THREADS 256
// a rank may be late by as much as its host needs between two launches (rank 0 writing output files, uneven set-up
// work): wait ~2 minutes of SM clocks before declaring the peer dead (sic_p2p_error)
#define SIC_P2P_TIMEOUT_CYCLES 240000000000ll
#define SIC_P2P_NSCAL 8

struct P2P {
  int rank, n_ranks, cap;                 // cap: interface nodes per peer the mailbox can hold
  int bpp;                                // blocks per peer of every launch on this context (from cap: the interface of
                                          // a 7 M-cell rank is ~50 k nodes per neighbour = 1.2 MB per exchange, which 4
                                          // blocks moved 8 bytes at a time in ~150 dependent trips to HBM / NVLink)
  size_t slot_doubles;                    // doubles per (source rank, parity) slot: 9*cap data + NSCAL scalars + 1 flag
  double* local;                          // this rank's mailbox (cudaMalloc)
  double* remote[SIC_P2P_MAX_RANKS];      // peers' mailboxes mapped into this process (remote[rank] == local)
  unsigned* counters;                     // [n_ranks] blocks-done counters (device)
  unsigned long long* gbar;               // device: halo blocks that have finished READING vec, over all exchanges
  int* error;                             // device flag: a wait timed out
  unsigned long long* epochs;             // DEVICE counters (so that a launch can be replayed from a CUDA graph):
                                          // [0] nodal (halo) exchanges done so far (identical on every rank),
                                          // [1] scalar exchanges done so far (own counter so that two consecutive scalar
                                          //     exchanges always alternate the mailbox parity),
                                          // [2] blocks of the running launch that have finished (the last one advances
                                          //     [0] / [1] for the next launch),
                                          // [3] fused operator + exchange launches done so far, [4] interface tiles that
                                          //     have finished, over all fused launches (k_mg_ebe_pc_x waits for
                                          //     ([3] + 1) * n_iface_tiles before it sends)
};

__device__ __forceinline__ double* p2p_slot(double* mailbox, size_t slot_doubles, int src, int parity) {
  return mailbox + ((size_t)src * 2 + parity) * slot_doubles;
}


// One halo block: chunk `chunk` of `bpp` of the nodal data exchanged with neighbour p.  Sends this rank's partial sums
// straight into the peer's mailbox, publishes the flag with the last chunk, waits for the peer's flag and for every halo
// block of this rank to have finished READING vec (gbar), then adds what arrived.  Any block size.
__device__ __forceinline__ void p2p_halo_block(const sic_halo_t& H, const P2P& ctx, double* __restrict__ vec, int ncomp, int p,
                                               int chunk, unsigned long long epoch, long long timeout) {
  __shared__ int ok;
  const int bpp = ctx.bpp, nthr = (int)blockDim.x;
  const int n_halo_blocks = H.n_peers * bpp;
  const int parity = (int)(epoch & 1ull);
  const size_t scal_off = 9 * (size_t)ctx.cap;
  const int peer = H.peer[p];
  const int off = H.peer_off[p], cnt = H.peer_off[p + 1] - off;
  const int n = cnt * ncomp;
  const int per = (n + bpp - 1) / bpp;
  const int lo = chunk * per, hi = min(n, lo + per);
  // send: my partial sums for the nodes shared with `peer` go straight into ITS mailbox
  double* out = p2p_slot(ctx.remote[peer], ctx.slot_doubles, ctx.rank, parity);
  // four independent gathers in flight per thread, then the four remote stores
  for (int t0 = lo + (int)threadIdx.x; t0 < hi; t0 += 4 * nthr) {
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j * nthr;
      if (t < hi) { const int k = t / ncomp, c = t - k * ncomp; v[j] = __ldcg(vec + (size_t)H.idx[off + k] * ncomp + c); }   // L2: written by other CTAs
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j * nthr;
      if (t < hi) out[t] = v[j];
    }
  }
  __syncthreads();                           // CTA-scope ordering of everybody's stores before thread 0's fence
  if (threadIdx.x == 0) {
    __threadfence_system();                  // cumulative: covers the whole CTA's remote stores
    atomicAdd(ctx.gbar, 1ull);               // this block no longer reads vec
    const unsigned done = atomicAdd(ctx.counters + p, 1u);
    if (done == (unsigned)bpp - 1u) {        // last chunk for this peer: publish
      ctx.counters[p] = 0;
      __threadfence_system();
      *(volatile unsigned long long*)(out + scal_off + SIC_P2P_NSCAL) = epoch + 1;
    }
  }
  // receive: wait for the peer's flag in MY mailbox, then add its partial sums
  double* in = p2p_slot(ctx.local, ctx.slot_doubles, peer, parity);
  if (threadIdx.x == 0) {
    volatile unsigned long long* flag = (volatile unsigned long long*)(in + scal_off + SIC_P2P_NSCAL);
    const long long t0 = clock64();
    int good = 1;
    while (*flag != epoch + 1) {
      if (clock64() - t0 > timeout) { good = 0; atomicExch(ctx.error, 1); break; }   // ~2 min
    }
    // A node shared by three ranks sits in two neighbours' lists: nobody may ADD to vec before every halo
    // block of this launch has finished READING its part of vec (grid-wide barrier; all blocks are resident).
    const unsigned long long target = (epoch + 1) * (unsigned long long)n_halo_blocks;
    while (*(volatile unsigned long long*)ctx.gbar < target) {
      if (clock64() - t0 > timeout) { good = 0; atomicExch(ctx.error, 1); break; }
    }
    __threadfence_system();
    ok = good;
  }
  __syncthreads();
  if (ok) {
    for (int t0 = lo + (int)threadIdx.x; t0 < hi; t0 += 4 * nthr) {
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + j * nthr;
        if (t < hi) v[j] = __ldcv(in + t);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + j * nthr;
        if (t < hi) { const int k = t / ncomp, c = t - k * ncomp; atomicAdd(vec + (size_t)H.idx[off + k] * ncomp + c, v[j]); }
      }
    }
  }
}

}  // namespace sic
#endif
