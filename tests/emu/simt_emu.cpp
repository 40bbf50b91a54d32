/* simt_emu.cpp -- TEST INFRASTRUCTURE ONLY (see simt_emu.h). */
#include "simt_emu.h"

#include <mutex>
#include <vector>

namespace simt {

Thread* cur = nullptr;
Dim3 g_blockIdx, g_blockDim, g_gridDim;
unsigned char* g_dynsmem = nullptr;
unsigned long long g_collectives = 0;

static void* g_sched_sp = nullptr;
static const std::function<void()>* g_body = nullptr;
static const size_t kStack = 256 * 1024;

/* Minimal x86-64 SysV context switch: save callee-saved registers on the current
 * stack, store its sp, load the other stack, restore, return into it. */
extern "C" void simt_switch(void** save_sp, void* load_sp);
__asm__(
    ".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch,.-simt_switch\n");

static void to_scheduler() { simt_switch(&cur->sp, g_sched_sp); }

static void fiber_main() {
  (*g_body)();
  cur->st = DONE;
  to_scheduler();
  fprintf(stderr, "simt: resumed a finished thread\n");
  abort();
}

uint64_t warp_collective(Op op, unsigned mask, uint64_t val, int arg) {
  Thread* t = cur;
  t->op = op; t->mask = mask; t->val = val; t->arg = arg; t->st = WAIT_WARP;
  to_scheduler();
  return t->res;
}

void cta_barrier() {
  cur->st = WAIT_CTA;
  to_scheduler();
}

/* named barriers (PTX bar.sync / bar.arrive with an explicit thread count): arrivals are counted per
 * barrier id; when `count` threads have arrived the waiting ones are released and the barrier resets. */
static int g_bar_arrived[16];
static int g_bar_expect[16];
void named_barrier(int id, int count, bool wait) {
  if (id < 1 || id > 15) { fprintf(stderr, "simt: bad barrier id %d\n", id); abort(); }
  if (g_bar_expect[id] && g_bar_expect[id] != count) { fprintf(stderr, "simt: barrier %d used with counts %d and %d\n", id, g_bar_expect[id], count); abort(); }
  g_bar_expect[id] = count;
  g_bar_arrived[id]++;
  if (wait) {
    cur->st = WAIT_BAR;
    cur->arg = id;
    to_scheduler();
  }
}

static void resolve_warp(Thread* th, unsigned base, unsigned nlanes, bool* did) {
  /* find a waiting lane; its mask names the participants */
  for (unsigned l0 = 0; l0 < nlanes; l0++) {
    Thread& a = th[base + l0];
    if (a.st != WAIT_WARP) continue;
    unsigned mask = a.mask;
    if (!(mask & (1u << l0))) { fprintf(stderr, "simt: lane %u not in its own mask %08x\n", l0, mask); abort(); }
    bool ready = true;
    for (unsigned l = 0; l < 32 && ready; l++) {
      if (!(mask & (1u << l))) continue;
      if (l >= nlanes) continue;             /* partial last warp: lanes that do not exist */
      Thread& b = th[base + l];
      if (b.st == DONE) { fprintf(stderr, "simt: lane %u exited but is named in mask %08x (op %d)\n", l, mask, a.op); abort(); }
      if (b.st != WAIT_WARP) ready = false;
      else if (b.op != a.op || b.mask != mask) {
        fprintf(stderr, "simt: divergent collective: lane %u op %d mask %08x vs lane %u op %d mask %08x\n",
                l0, a.op, mask, l, b.op, b.mask);
        abort();
      }
    }
    if (!ready) continue;
    g_collectives++;
    unsigned ballot = 0;
    for (unsigned l = 0; l < nlanes; l++)
      if ((mask & (1u << l)) && th[base + l].val) ballot |= 1u << l;
    for (unsigned l = 0; l < nlanes; l++) {
      if (!(mask & (1u << l))) continue;
      Thread& b = th[base + l];
      int src = (int)l;
      switch (a.op) {
        case OP_SYNCWARP: b.res = 0; break;
        case OP_BALLOT: b.res = ballot; break;
        case OP_SHFL: src = b.arg & 31; break;
        case OP_SHFL_UP: src = (int)l - b.arg; if (src < 0) src = (int)l; break;
        case OP_SHFL_DOWN: src = (int)l + b.arg; if (src > 31) src = (int)l; break;
        case OP_SHFL_XOR: src = (int)l ^ b.arg; break;
        case OP_MATCH_ANY: {
          unsigned m = 0;
          for (unsigned k = 0; k < nlanes; k++)
            if ((mask & (1u << k)) && th[base + k].val == b.val) m |= 1u << k;
          b.res = m;
          break;
        }
        default: abort();
      }
      if (a.op == OP_SHFL || a.op == OP_SHFL_UP || a.op == OP_SHFL_DOWN || a.op == OP_SHFL_XOR) {
        if ((unsigned)src < nlanes && (mask & (1u << src))) b.res = th[base + src].val;
        else b.res = b.val;   /* reading an inactive lane is undefined on hardware; keep own */
      }
    }
    for (unsigned l = 0; l < nlanes; l++)
      if (mask & (1u << l)) th[base + l].st = RUNNABLE;
    *did = true;
  }
}

static std::mutex g_launch_mu;   /* the scheduler state is global: one emulated kernel at a time */

void launch(Dim3 grid, Dim3 block, size_t dynsmem, const std::function<void()>& body) {
  std::lock_guard<std::mutex> guard(g_launch_mu);
  unsigned nthreads = block.x * block.y * block.z;
  std::vector<Thread> th(nthreads);
  std::vector<unsigned char> smem(dynsmem + 64);
  for (unsigned i = 0; i < nthreads; i++) th[i].stack = (char*)malloc(kStack);
  g_body = &body;
  g_blockDim = block;
  g_gridDim = grid;
  unsigned long long phase = 0;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = Dim3(bx, by, bz);
        memset(g_bar_arrived, 0, sizeof g_bar_arrived); memset(g_bar_expect, 0, sizeof g_bar_expect);
        memset(smem.data(), 0xCD, smem.size());
        g_dynsmem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
        for (unsigned i = 0; i < nthreads; i++) {
          Thread& t = th[i];
          t.lin = i;
          t.tidx = Dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
          t.st = RUNNABLE;
          t.op = OP_NONE;
          uintptr_t top = ((uintptr_t)t.stack + kStack) & ~(uintptr_t)15;
          void** sp = (void**)top;
          *--sp = nullptr;                 /* fake return address for fiber_main */
          *--sp = (void*)&fiber_main;      /* simt_switch 'ret's into this */
          for (int r = 0; r < 6; r++) *--sp = nullptr;
          t.sp = (void*)sp;
        }
        unsigned alive = nthreads;
        while (alive) {
          bool progress = false;
          bool rev = (phase++ & 1) != 0;
          for (unsigned k = 0; k < nthreads; k++) {
            unsigned i = rev ? nthreads - 1 - k : k;
            if (th[i].st != RUNNABLE) continue;
            cur = &th[i];
            simt_switch(&g_sched_sp, th[i].sp);
            progress = true;
            if (th[i].st == DONE) alive--;
          }
          bool did = false;
          for (unsigned base = 0; base < nthreads; base += 32)
            resolve_warp(th.data(), base, nthreads - base < 32 ? nthreads - base : 32, &did);
          for (int b = 1; b < 16; b++) {
            if (g_bar_expect[b] && g_bar_arrived[b] >= g_bar_expect[b]) {
              if (g_bar_arrived[b] > g_bar_expect[b]) { fprintf(stderr, "simt: barrier %d over-subscribed (%d of %d)\n", b, g_bar_arrived[b], g_bar_expect[b]); abort(); }
              for (unsigned i = 0; i < nthreads; i++)
                if (th[i].st == WAIT_BAR && th[i].arg == b) th[i].st = RUNNABLE;
              g_bar_arrived[b] = 0; g_bar_expect[b] = 0;
              did = true;
            }
          }
          unsigned waiting_cta = 0;
          for (unsigned i = 0; i < nthreads; i++) waiting_cta += th[i].st == WAIT_CTA;
          if (alive && waiting_cta == alive) {
            for (unsigned i = 0; i < nthreads; i++) if (th[i].st == WAIT_CTA) th[i].st = RUNNABLE;
            did = true;
          }
          if (!progress && !did && alive) {
            fprintf(stderr, "simt: deadlock in block (%u,%u,%u): %u threads alive\n", bx, by, bz, alive);
            for (unsigned i = 0; i < nthreads; i++)
              if (th[i].st != DONE) fprintf(stderr, "  thread %u state %d op %d mask %08x\n", i, th[i].st, th[i].op, th[i].mask);
            abort();
          }
        }
      }
  cur = nullptr;
  for (unsigned i = 0; i < nthreads; i++) free(th[i].stack);
}

}  // namespace simt
