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
  me.done = true;
  psd_emu_switch(&me.sp, w->sched_sp);
  abort();  // a finished fiber is never resumed
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
    w.f[i].done = false;
    w.f[i].site = -1;
  }
  Warp* outer = g_warp;
  g_warp = &w;
  for (;;) {
    int n_done = 0, site = -2;
    for (int k = 0; k < 32; k++) {
      int i = descending ? 31 - k : k;
      if (w.f[i].done) { n_done++; continue; }
      w.cur = i;
      psd_emu_switch(&w.sched_sp, w.f[i].sp);
      if (w.f[i].done) { n_done++; continue; }
      if (site == -2) site = w.f[i].site;
      else if (site != w.f[i].site) {
        fprintf(stderr, "warp_emu: divergent collective: lane %d at line %d, others at line %d\n", i, w.f[i].site, site);
        abort();
      }
    }
    if (n_done == 32) break;
    if (n_done != 0) {
      fprintf(stderr, "warp_emu: %d lanes exited while others wait at line %d\n", n_done, site);
      abort();
    }
    w.phase++;
  }
  g_warp = outer;
  for (int i = 0; i < 32; i++) free(w.f[i].stack);
}
}  // namespace psd_emu
