// TEST TOOL: fiber scheduler behind warp_emu.h (x86-64 SysV only).
#include "warp_emu.h"

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

void run_warp(void (*entry)(void*), void* arg, int descending) {
  static const size_t kStack = 1 << 20;
  Warp w;
  memset(&w, 0, sizeof w);
  w.entry = entry; w.arg = arg; w.descending = descending;
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
  Warp* outer = g_warp;
  g_warp = &w;
  for (;;) {
    bool progress = false;
    for (int k = 0; k < 32; k++) {
      const int i = descending ? 31 - k : k;
      if (w.f[i].state != RUNNABLE) continue;
      w.cur = i;
      psd_emu_switch(&w.sched_sp, w.f[i].sp);
      progress = true;
    }
    int n_done = 0;
    for (int i = 0; i < 32; i++) n_done += w.f[i].state == DONE;
    if (n_done == 32) break;
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
    if (!progress) die(w, "deadlock: no lane can run and no collective is complete");
  }
  g_warp = outer;
  for (int i = 0; i < 32; i++) free(w.f[i].stack);
}
}  // namespace psd_emu
