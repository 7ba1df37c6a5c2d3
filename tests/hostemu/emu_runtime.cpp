// emu_runtime.cpp — block/thread scheduler of the host emulation (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include <stdio.h>

#include <vector>

#include "cuda_emu.h"
#include "safeincave_cuda.h"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

extern "C" void sic_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl sic_emu_switch
.type sic_emu_switch,@function
sic_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size sic_emu_switch,.-sic_emu_switch
)");

namespace sic_emu {

static const size_t kStack = 512 * 1024;
struct Fiber { char* stack = nullptr; void* sp = nullptr; bool done = true; int ncoll = 0; };
static thread_local std::vector<Fiber> g_fib;
static thread_local std::vector<double> g_slots[2];
static thread_local void* g_sched_sp = nullptr;
static thread_local int g_cur = -1;              // fiber index
static thread_local bool g_fiber_mode = false;
static thread_local const std::function<void()>* g_body = nullptr;

static void trampoline() {
  (*g_body)();
  Fiber& f = g_fib[g_cur];
  f.done = true;
  sic_emu_switch(&f.sp, g_sched_sp);
  abort();  // a finished fiber is never resumed
}

static void prepare(Fiber& f) {
  if (!f.stack) f.stack = (char*)aligned_alloc(64, kStack);
  uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;                   // keeps rsp = 8 (mod 16) at the trampoline's entry
  *--sp = (void*)&trampoline;        // return address popped by sic_emu_switch's ret
  for (int k = 0; k < 6; ++k) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
  f.sp = sp;
  f.done = false;
  f.ncoll = 0;
}

void collective() {
  Fiber& f = g_fib[g_cur];
  f.ncoll += 1;
  sic_emu_switch(&f.sp, g_sched_sp);
}

int collective_count() {
  return g_fib[g_cur].ncoll;
}

double* shfl_slot(int parity) {
  return g_slots[parity].data();
}

static void run_block_fibers(unsigned block) {
  if (g_fib.size() < block) g_fib.resize(block);
  for (int p = 0; p < 2; ++p) if (g_slots[p].size() < block) g_slots[p].resize(block);
  for (unsigned t = 0; t < block; ++t) prepare(g_fib[t]);
  g_fiber_mode = true;
  unsigned live = block;
  while (live) {
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = g_fib[t];
      if (f.done) continue;
      threadIdx.x = t;
      g_cur = (int)t;
      sic_emu_switch(&g_sched_sp, f.sp);
      if (f.done) --live;
    }
  }
  g_fiber_mode = false;
  g_cur = -1;
}

void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  if (g_body) { fprintf(stderr, "sic_emu: nested launch\n"); abort(); }
  g_body = &body;
  gridDim.x = grid; blockDim.x = block;
  // Every thread runs on its own fiber, always: running threads directly and switching to fibers only
  // when a collective is reached would re-execute the work done before it (k_post increments state).
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    run_block_fibers(block);
  }
  g_body = nullptr;
}

}  // namespace sic_emu

// ---- multi-GPU entry points: every emulated rank is a host thread; they meet here ------------------------
// sic_exchange / sic_halo_sum / sic_allreduce_sum with the semantics of csrc/comm.cu: halo sum of the interface
// nodes (every rank first PUBLISHES what it sends to each neighbour, then everybody adds what it received) and
// sums of scalars / vectors over the ranks in rank order (identical bits on every rank).
#include <chrono>
#include <condition_variable>
#include <mutex>

namespace {
struct RankSlot {
  std::vector<std::vector<double>> send;   // per neighbour (in the order of sic_halo_t.peer)
  std::vector<int> peer;
  std::vector<double> scal;
  const double* big = nullptr;             // sic_allreduce_sum operand
};
std::mutex g_mu;
std::condition_variable g_cv;
int g_waiting = 0;
unsigned long long g_generation = 0;
RankSlot g_rank[SIC_MAX_PEERS];

int g_timeout_s = 120;
// all `n` ranks arrive, or -1 after g_timeout_s (a rank died: its thread raised in Python)
int rank_barrier(int n) {
  std::unique_lock<std::mutex> lk(g_mu);
  const unsigned long long gen = g_generation;
  if (++g_waiting == n) { g_waiting = 0; ++g_generation; g_cv.notify_all(); return 0; }
  if (!g_cv.wait_for(lk, std::chrono::seconds(g_timeout_s), [&] { return g_generation != gen; })) { --g_waiting; return -1; }
  return 0;
}
struct EmuComm { int rank, n_ranks; };
}  // namespace

extern int sic_fail(const char* msg);

extern "C" {
void* sic_emu_make_comm(int rank, int n_ranks) { return new EmuComm{rank, n_ranks}; }
void sic_emu_set_timeout(int seconds) { g_timeout_s = seconds > 0 ? seconds : 1; }

int sic_exchange(const sic_halo_t* h, double* vec, int ncomp, double* scal, int n_scal, void*) {
  if (!h || h->n_ranks <= 1) return 0;
  if (h->n_ranks > SIC_MAX_PEERS) return sic_fail("emulated sic_exchange: too many ranks");
  RankSlot& me = g_rank[h->rank];
  me.send.assign(h->n_peers, {});
  me.peer.assign(h->peer, h->peer + h->n_peers);
  if (ncomp > 0) {
    for (int p = 0; p < h->n_peers; ++p) {
      const int off = h->peer_off[p], cnt = h->peer_off[p + 1] - off;
      me.send[p].resize((size_t)cnt * ncomp);
      for (int k = 0; k < cnt; ++k)
        for (int c = 0; c < ncomp; ++c) me.send[p][(size_t)k * ncomp + c] = vec[(size_t)h->idx[off + k] * ncomp + c];
    }
  }
  me.scal.assign(scal, scal + (n_scal > 0 ? n_scal : 0));
  if (rank_barrier(h->n_ranks)) return sic_fail("emulated sic_exchange: a rank did not arrive");
  if (ncomp > 0) {
    for (int p = 0; p < h->n_peers; ++p) {
      const RankSlot& q = g_rank[h->peer[p]];
      int back = -1;
      for (size_t j = 0; j < q.peer.size(); ++j) if (q.peer[j] == h->rank) back = (int)j;
      const int off = h->peer_off[p], cnt = h->peer_off[p + 1] - off;
      if (back < 0 || q.send[back].size() != (size_t)cnt * ncomp) return sic_fail("emulated sic_exchange: halo plans of two ranks disagree");
      for (int k = 0; k < cnt; ++k)
        for (int c = 0; c < ncomp; ++c) vec[(size_t)h->idx[off + k] * ncomp + c] += q.send[back][(size_t)k * ncomp + c];
    }
  }
  for (int j = 0; j < n_scal; ++j) {
    double acc = 0.0;
    for (int r = 0; r < h->n_ranks; ++r) acc += g_rank[r].scal[j];
    scal[j] = acc;
  }
  if (rank_barrier(h->n_ranks)) return sic_fail("emulated sic_exchange: a rank did not arrive");
  return 0;
}

int sic_halo_sum(const sic_halo_t* h, double* vec, int ncomp, void* st) { return sic_exchange(h, vec, ncomp, nullptr, 0, st); }

int sic_allreduce_sum(void* comm, double* buf, int count, void*) {
  if (!comm || !buf) return sic_fail("emulated sic_allreduce_sum: null");
  const EmuComm* c = (const EmuComm*)comm;
  if (c->n_ranks <= 1) return 0;
  g_rank[c->rank].big = buf;
  if (rank_barrier(c->n_ranks)) return sic_fail("emulated sic_allreduce_sum: a rank did not arrive");
  std::vector<double> tot((size_t)count, 0.0);
  for (int r = 0; r < c->n_ranks; ++r)
    for (int k = 0; k < count; ++k) tot[k] += g_rank[r].big[k];
  if (rank_barrier(c->n_ranks)) return sic_fail("emulated sic_allreduce_sum: a rank did not arrive");
  for (int k = 0; k < count; ++k) buf[k] = tot[k];
  if (rank_barrier(c->n_ranks)) return sic_fail("emulated sic_allreduce_sum: a rank did not arrive");
  return 0;
}

int sic_p2p_error(void*) { return 0; }
}
