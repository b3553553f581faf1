// warp_emu.h -- TEST TOOL: runs the product's warp-cooperative device source (fpop_warp.cuh) on the
// CPU as 32 cooperatively scheduled fibers, one per lane, so its control flow and arithmetic can
// be diffed against the oracle in the GPU-less build container.  Never linked into the product
// library; the product has no CPU execution path.
//
// Model: every warp collective (shuffle, ballot, syncwarp) is a barrier.  The scheduler resumes
// lanes in ascending or descending order (PSD_EMU_ORDER) until each reaches its next collective;
// all 32 lanes must arrive at the same call site, otherwise the run aborts (that would be a
// divergent full-mask collective on the GPU).  Running both orders exposes missing __syncwarp()s
// between a shared-memory write and a cross-lane read.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace psd_emu {

struct Fiber { void* sp; char* stack; bool done; int site; };
struct Warp {
  Fiber f[32];
  void* sched_sp;
  int cur;
  int phase;
  uint64_t xchg[2][32];
  void (*entry)(void*);
  void* arg;
  int descending;
};
extern Warp* g_warp;

extern "C" void psd_emu_switch(void** save_sp, void* load_sp);

inline int lane() { return g_warp->cur; }

inline void barrier(int site) {
  Warp* w = g_warp;
  Fiber& me = w->f[w->cur];
  me.site = site;
  psd_emu_switch(&me.sp, w->sched_sp);
}

inline uint64_t exchange(uint64_t v, int src_lane, int site) {
  Warp* w = g_warp;
  const int p = w->phase & 1;
  const int me = w->cur;
  w->xchg[p][me] = v;
  barrier(site);
  return g_warp->xchg[p][src_lane & 31];
}

inline uint32_t ballot(int pred, int site) {
  Warp* w = g_warp;
  const int p = w->phase & 1;
  w->xchg[p][w->cur] = pred ? 1 : 0;
  barrier(site);
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) m |= (uint32_t)(g_warp->xchg[p][i] & 1) << i;
  return m;
}

void run_warp(void (*entry)(void*), void* arg, int descending);

}  // namespace psd_emu

// ---- the primitive set fpop_warp.cuh is written against ------------------------------------------
#define PSD_DEV static inline
#define PSD_DEVNI static __attribute__((noinline))
#define PSD_SITE __LINE__

static inline int psd_lane() { return psd_emu::lane(); }
static inline uint64_t psd_bits_(double v) { uint64_t u; memcpy(&u, &v, 8); return u; }
static inline double psd_dbl_(uint64_t u) { double v; memcpy(&v, &u, 8); return v; }

#define psd_shfl_d(v, src) psd_dbl_(psd_emu::exchange(psd_bits_(v), (src), PSD_SITE))
#define psd_shfl_i(v, src) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (src), PSD_SITE))
#define psd_shfl_u64(v, src) (psd_emu::exchange((uint64_t)(v), (src), PSD_SITE))
// up/down: lanes whose source falls outside the warp keep their own value
#define psd_shfl_up_d(v, d) psd_dbl_(psd_emu::exchange(psd_bits_(v), (psd_lane() - (d) < 0 ? psd_lane() : psd_lane() - (d)), PSD_SITE))
#define psd_shfl_down_d(v, d) psd_dbl_(psd_emu::exchange(psd_bits_(v), (psd_lane() + (d) > 31 ? psd_lane() : psd_lane() + (d)), PSD_SITE))
#define psd_shfl_up_i(v, d) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (psd_lane() - (d) < 0 ? psd_lane() : psd_lane() - (d)), PSD_SITE))
#define psd_shfl_xor_d(v, m) psd_dbl_(psd_emu::exchange(psd_bits_(v), psd_lane() ^ (m), PSD_SITE))
#define psd_shfl_xor_i(v, m) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), psd_lane() ^ (m), PSD_SITE))
#define psd_ballot(p) psd_emu::ballot((p), PSD_SITE)
#define psd_syncwarp() psd_emu::barrier(PSD_SITE)

static inline int psd_ffs(uint32_t m) { return __builtin_ffs((int)m); }
static inline int psd_clz(uint32_t m) { return m ? __builtin_clz(m) : 32; }
static inline int psd_popc(uint32_t m) { return __builtin_popcount(m); }
static inline unsigned long long psd_atomic_add_ull(unsigned long long* p, unsigned long long v) {
  unsigned long long old = *p; *p = old + v; return old;
}
static inline int psd_atomic_add_int(int* p, int v) { int old = *p; *p = old + v; return old; }
// streaming (evict-first) stores of the cost-function store: plain stores here
static inline void psd_st_cs_d2(double* p, double x, double y) { p[0] = x; p[1] = y; }
static inline void psd_st_cs_i(int* p, int v) { *p = v; }
static inline void psd_st_cs_u64(unsigned long long* p, unsigned long long v) { *p = v; }
static inline void psd_st_cs_u4(unsigned* p, unsigned a, unsigned b, unsigned c, unsigned d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
