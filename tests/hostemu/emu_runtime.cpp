// emu_runtime.cpp — block/thread scheduler of the host emulation (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include <stdio.h>

#include <vector>

#include "cuda_emu.h"
#include "safeincave_cuda.h"

dim3 threadIdx, blockIdx, blockDim, gridDim;

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
static std::vector<Fiber> g_fib;
static std::vector<double> g_slots[2];
static void* g_sched_sp = nullptr;
static int g_cur = -1;              // fiber index, -1: direct mode
static bool g_fiber_mode = false;
static const std::function<void()>* g_body = nullptr;

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

// ---- multi-GPU entry points: the emulation is single-"device" ----------------------------------------
extern "C" {
int sic_exchange(const sic_halo_t* h, double*, int, double*, int, void*) { return h ? -1 : 0; }
int sic_halo_sum(const sic_halo_t* h, double*, int, void*) { return h ? -1 : 0; }
int sic_p2p_error(void*) { return 0; }
}
