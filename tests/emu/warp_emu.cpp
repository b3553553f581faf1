// TEST TOOL: fiber scheduler behind warp_emu.h (x86-64 SysV only).
#include "warp_emu.h"

void (*psd_emu_ring_drain)(const StorePool* sp) = nullptr;

namespace psd_emu {
Warp* g_warp = nullptr;

// Save callee-saved registers on the current stack, store its sp, switch to load_sp and restore.
asm(R"(
.text
.globl psd_emu_switch
.type psd_emu_switch,@function
psd_emu_switch:
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
.size psd_emu_switch,.-psd_emu_switch
)");

// named barriers of the block being run: arrival count and generation per id
static int g_named_count[16], g_named_need[16];
static unsigned g_named_gen[16];
static Fiber* g_named_waiters[16][1024];
static int g_named_nwait[16];

void named_arrive(int id, int n_threads, bool wait, int site) {
  Warp* w = g_warp;
  g_named_need[id] = n_threads;
  g_named_count[id]++;
  if (g_named_count[id] == n_threads) {          // the barrier completes: release the waiters, start a new generation
    for (int k = 0; k < g_named_nwait[id]; k++) g_named_waiters[id][k]->state = RUNNABLE;
    g_named_nwait[id] = 0; g_named_count[id] = 0; g_named_gen[id]++;
    return;
  }
  if (!wait) return;
  Fiber& me = w->f[w->cur];
  g_named_waiters[id][g_named_nwait[id]++] = &me;
  me.state = WAIT_NAMED; me.site = site;
  psd_emu_switch(&me.sp, w->sched_sp);
}

static void trampoline() {
  Warp* w = g_warp;
  w->entry(w->arg);
  Fiber& me = w->f[w->cur];
  me.state = DONE;
  psd_emu_switch(&me.sp, w->sched_sp);
  abort();  // a finished fiber is never resumed
}

static void die(Warp& w, const char* what) {
  fprintf(stderr, "warp_emu: %s\n", what);
  for (int i = 0; i < 32; i++) fprintf(stderr, "  lane %2d state %d line %d\n", i, w.f[i].state, w.f[i].site);
  abort();
}

static void init_warp(Warp& w, void (*entry)(void*), void* arg, int descending, int warp_id, int n_warps) {
  static const size_t kStack = 1 << 20;
  memset(&w, 0, sizeof w);
  w.entry = entry; w.arg = arg; w.descending = descending; w.warp_id = warp_id; w.n_warps = n_warps;
  for (int i = 0; i < 32; i++) {
    w.f[i].stack = (char*)aligned_alloc(64, kStack);
    uintptr_t top = ((uintptr_t)w.f[i].stack + kStack) & ~(uintptr_t)63;
    // layout consumed by psd_emu_switch: 6 saved registers, then the return address.  After the
    // `ret`, rsp must be 8 mod 16 (as if trampoline had been called).
    uint64_t* sp = (uint64_t*)(top - 8);
    *--sp = (uint64_t)(uintptr_t)&trampoline;
    for (int k = 0; k < 6; k++) *--sp = 0;
    w.f[i].sp = sp;
    w.f[i].state = RUNNABLE;
    w.f[i].site = -1;
  }
}

// One scheduling step of a warp: run every runnable lane to its next collective, then release the
// groups / the warp whose lanes all wait at the same call site.  Returns true on progress.
static bool step_warp(Warp& w) {
  bool progress = false;
  g_warp = &w;
  for (int k = 0; k < 32; k++) {
    const int i = w.descending ? 31 - k : k;
    if (w.f[i].state != RUNNABLE) continue;
    w.cur = i;
    psd_emu_switch(&w.sched_sp, w.f[i].sp);
    progress = true;
  }
  // release a 16-lane group whose lanes all wait at the same group collective
  for (int g = 0; g < 2; g++) {
    bool all = true;
    const int site = w.f[16 * g].site;
    for (int i = 16 * g; i < 16 * g + 16; i++) all = all && w.f[i].state == WAIT_GROUP && w.f[i].site == site;
    if (all) {
      for (int i = 16 * g; i < 16 * g + 16; i++) w.f[i].state = RUNNABLE;
      w.phase_g[g]++;
      progress = true;
    } else {
      int n_wait = 0;
      for (int i = 16 * g; i < 16 * g + 16; i++) n_wait += w.f[i].state == WAIT_GROUP;
      if (n_wait == 16) die(w, "divergent group collective (lanes of one group wait at different lines)");
    }
  }
  // release the warp when all lanes wait at the same whole-warp collective
  {
    bool all = true;
    const int site = w.f[0].site;
    for (int i = 0; i < 32; i++) all = all && w.f[i].state == WAIT_FULL && w.f[i].site == site;
    if (all) {
      for (int i = 0; i < 32; i++) w.f[i].state = RUNNABLE;
      w.phase_full++;
      progress = true;
    }
  }
  return progress;
}

void run_block(void (*entry)(void*), void* arg, int n_warps, int descending) {
  memset(g_named_count, 0, sizeof g_named_count); memset(g_named_nwait, 0, sizeof g_named_nwait);
  Warp* ws = new Warp[n_warps];
  for (int k = 0; k < n_warps; k++) init_warp(ws[k], entry, arg, descending, k, n_warps);
  Warp* outer = g_warp;
  for (;;) {
    bool progress = false;
    // warps advance in ascending or descending order too: a missing block barrier between a write
    // by one warp and a read by another shows up in one of the two orders
    for (int k = 0; k < n_warps; k++) progress = step_warp(ws[descending ? n_warps - 1 - k : k]) || progress;
    int n_done = 0, n_bar = 0, site = -2;
    bool same_site = true;
    for (int k = 0; k < n_warps; k++)
      for (int i = 0; i < 32; i++) {
        const Fiber& f = ws[k].f[i];
        n_done += f.state == DONE;
        if (f.state == WAIT_BLOCK) { n_bar++; if (site == -2) site = f.site; else same_site = same_site && f.site == site; }
      }
    if (n_done == 32 * n_warps) break;
    if (n_bar > 0 && n_bar + n_done == 32 * n_warps) {
      if (!same_site || n_done > 0) die(ws[0], "divergent block barrier (lanes wait at different lines, or some lanes have exited)");
      for (int k = 0; k < n_warps; k++)
        for (int i = 0; i < 32; i++) ws[k].f[i].state = RUNNABLE;
      progress = true;
    }
    if (!progress) die(ws[0], "deadlock: no lane can run and no collective is complete");
  }
  g_warp = outer;
  for (int k = 0; k < n_warps; k++)
    for (int i = 0; i < 32; i++) free(ws[k].f[i].stack);
  delete[] ws;
}

void run_warp(void (*entry)(void*), void* arg, int descending) { run_block(entry, arg, 1, descending); }
}  // namespace psd_emu
