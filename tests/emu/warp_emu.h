// warp_emu.h -- TEST TOOL: runs the product's warp-cooperative device source (fpop_warp.cuh) on the
// CPU as 32 cooperatively scheduled fibers, one per lane, so its control flow and arithmetic can
// be diffed against the oracle in the GPU-less build container.  Never linked into the product
// library; the product has no CPU execution path.
//
// Model: every warp collective (shuffle, ballot, sync) is a barrier, either over the whole warp or
// over one 16-lane group (the two half-warps run different operator chains and only meet at
// whole-warp collectives).  The scheduler resumes runnable lanes in ascending or descending order
// until each blocks at its next collective; a group (or the warp) is released when all its lanes
// wait at the SAME call site, otherwise the run aborts (that would be a divergent collective on the
// GPU).  Running both orders exposes missing syncs between a shared-memory write and a
// cross-lane read.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace psd_emu {

enum { RUNNABLE = 0, WAIT_GROUP = 1, WAIT_FULL = 2, DONE = 3, WAIT_BLOCK = 4, WAIT_NAMED = 5 };
struct Fiber { void* sp; char* stack; int state; int site; };
struct Warp {
  Fiber f[32];
  void* sched_sp;
  int cur;
  int phase_full;
  int phase_g[2];
  uint64_t xchg_full[2][32];
  uint64_t xchg_g[2][2][16];
  void (*entry)(void*);
  void* arg;
  int descending;
  int warp_id;     // index of this warp in its block (run_block); 0 for run_warp
  int n_warps;
};
extern Warp* g_warp;

extern "C" void psd_emu_switch(void** save_sp, void* load_sp);

inline int lane() { return g_warp->cur; }
inline int warp_id() { return g_warp->warp_id; }
inline int n_warps() { return g_warp->n_warps; }

inline void block(int kind, int site) {
  Warp* w = g_warp;
  Fiber& me = w->f[w->cur];
  me.state = kind; me.site = site;
  psd_emu_switch(&me.sp, w->sched_sp);
}

inline uint64_t exchange(uint64_t v, int src_lane, int site) {
  Warp* w = g_warp;
  const int p = w->phase_full & 1;
  w->xchg_full[p][w->cur] = v;
  block(WAIT_FULL, site);
  return g_warp->xchg_full[p][src_lane & 31];
}
inline uint32_t ballot(int pred, int site) {
  Warp* w = g_warp;
  const int p = w->phase_full & 1;
  w->xchg_full[p][w->cur] = pred ? 1 : 0;
  block(WAIT_FULL, site);
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) m |= (uint32_t)(g_warp->xchg_full[p][i] & 1) << i;
  return m;
}
// 16-lane group versions: src is a lane index within the group
inline uint64_t g_exchange(uint64_t v, int src_local, int site) {
  Warp* w = g_warp;
  const int g = w->cur >> 4, gl = w->cur & 15;
  const int p = w->phase_g[g] & 1;
  w->xchg_g[g][p][gl] = v;
  block(WAIT_GROUP, site);
  return g_warp->xchg_g[g][p][src_local & 15];
}
inline uint32_t g_ballot(int pred, int site) {
  Warp* w = g_warp;
  const int g = w->cur >> 4, gl = w->cur & 15;
  const int p = w->phase_g[g] & 1;
  w->xchg_g[g][p][gl] = pred ? 1 : 0;
  block(WAIT_GROUP, site);
  uint32_t m = 0;
  for (int i = 0; i < 16; i++) m |= (uint32_t)(g_warp->xchg_g[g][p][i] & 1) << i;
  return m;
}

// Named barriers (PTX bar.sync / bar.arrive with an id and a thread count): arrivals are counted per
// id; a fiber that syncs blocks until the count reaches n_threads, a fiber that only arrives goes on.
void named_arrive(int id, int n_threads, bool wait, int site);
void run_warp(void (*entry)(void*), void* arg, int descending);
// A thread block of n_warps warps (the latency kernel: the warps of one problem).  Block barriers
// (block(WAIT_BLOCK, site)) release when every lane of every warp waits at the same call site.
void run_block(void (*entry)(void*), void* arg, int n_warps, int descending);

}  // namespace psd_emu

// ---- the primitive set fpop_warp.cuh is written against ------------------------------------------
#define PSD_DEV static inline
#define PSD_DEVNI static __attribute__((noinline))
#define PSD_SITE __LINE__

static inline int psd_lane() { return psd_emu::lane(); }
#if defined(PSD_G32)
static inline int psd_glane() { return psd_emu::lane(); }
#else
static inline int psd_glane() { return psd_emu::lane() & 15; }
#endif
static inline uint64_t psd_bits_(double v) { uint64_t u; memcpy(&u, &v, 8); return u; }
static inline double psd_dbl_(uint64_t u) { double v; memcpy(&v, &u, 8); return v; }

// block barrier (multi-warp blocks only; with one warp it degenerates to a warp barrier)
#define psd_cta_sync() psd_emu::block(psd_emu::WAIT_BLOCK, PSD_SITE)
#define psd_bar_sync(id, n) psd_emu::named_arrive((id), (n), true, PSD_SITE)
#define psd_bar_arrive(id, n) psd_emu::named_arrive((id), (n), false, PSD_SITE)
static inline int psd_warp_in_block() { return psd_emu::warp_id(); }
// whole-warp collectives
#define psd_shfl_d(v, src) psd_dbl_(psd_emu::exchange(psd_bits_(v), (src), PSD_SITE))
#define psd_shfl_i(v, src) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (src), PSD_SITE))
#define psd_shfl_u64(v, src) (psd_emu::exchange((uint64_t)(v), (src), PSD_SITE))
#define psd_shfl_xor_d(v, m) psd_dbl_(psd_emu::exchange(psd_bits_(v), psd_lane() ^ (m), PSD_SITE))
#define psd_shfl_xor_i(v, m) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), psd_lane() ^ (m), PSD_SITE))
#define psd_shfl_up_i(v, d) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (psd_lane() - (d) < 0 ? psd_lane() : psd_lane() - (d)), PSD_SITE))
#define psd_ballot(p) psd_emu::ballot((p), PSD_SITE)
#define psd_syncwarp() psd_emu::block(psd_emu::WAIT_FULL, PSD_SITE)
#if defined(PSD_G32)
// one operator group per warp (latency kernel): group collectives are whole-warp collectives
static inline int psd_glane32_() { return psd_emu::lane(); }
#define psd_g_shfl_d(v, src) psd_dbl_(psd_emu::exchange(psd_bits_(v), (src), PSD_SITE))
#define psd_g_shfl_i(v, src) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (src), PSD_SITE))
#define psd_g_shfl_up_d(v, d) psd_dbl_(psd_emu::exchange(psd_bits_(v), (psd_lane() - (d) < 0 ? psd_lane() : psd_lane() - (d)), PSD_SITE))
#define psd_g_shfl_down_d(v, d) psd_dbl_(psd_emu::exchange(psd_bits_(v), (psd_lane() + (d) > 31 ? psd_lane() : psd_lane() + (d)), PSD_SITE))
#define psd_g_shfl_up_i(v, d) ((int)(int64_t)psd_emu::exchange((uint64_t)(int64_t)(v), (psd_lane() - (d) < 0 ? psd_lane() : psd_lane() - (d)), PSD_SITE))
#define psd_g_ballot(p) psd_emu::ballot((p), PSD_SITE)
#define psd_g_sync() psd_emu::block(psd_emu::WAIT_FULL, PSD_SITE)
#else
// 16-lane group collectives; up/down: lanes whose source falls outside the group keep their value
#define psd_g_shfl_d(v, src) psd_dbl_(psd_emu::g_exchange(psd_bits_(v), (src), PSD_SITE))
#define psd_g_shfl_i(v, src) ((int)(int64_t)psd_emu::g_exchange((uint64_t)(int64_t)(v), (src), PSD_SITE))
#define psd_g_shfl_up_d(v, d) psd_dbl_(psd_emu::g_exchange(psd_bits_(v), (psd_glane() - (d) < 0 ? psd_glane() : psd_glane() - (d)), PSD_SITE))
#define psd_g_shfl_down_d(v, d) psd_dbl_(psd_emu::g_exchange(psd_bits_(v), (psd_glane() + (d) > 15 ? psd_glane() : psd_glane() + (d)), PSD_SITE))
#define psd_g_shfl_up_i(v, d) ((int)(int64_t)psd_emu::g_exchange((uint64_t)(int64_t)(v), (psd_glane() - (d) < 0 ? psd_glane() : psd_glane() - (d)), PSD_SITE))
#define psd_g_ballot(p) psd_emu::g_ballot((p), PSD_SITE)
#define psd_g_sync() psd_emu::block(psd_emu::WAIT_GROUP, PSD_SITE)
#endif

static inline int psd_ffs(uint32_t m) { return __builtin_ffs((int)m); }
static inline int psd_clz(uint32_t m) { return m ? __builtin_clz(m) : 32; }
static inline int psd_popc(uint32_t m) { return __builtin_popcount(m); }
static inline unsigned long long psd_atomic_add_ull(unsigned long long* p, unsigned long long v) {
  unsigned long long old = *p; *p = old + v; return old;
}
static inline int psd_atomic_add_int(int* p, int v) { int old = *p; *p = old + v; return old; }
static inline void psd_fence_system() {}
static inline void psd_fence_device() {}
// bulk shared -> global copies of the record store: immediate copies here
static inline void psd_bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
static inline void psd_bulk_fence() {}
static inline void psd_bulk_commit() {}
static inline void psd_bulk_wait_read_1() {}
static inline void psd_bulk_wait_read_0() {}
static inline void psd_bulk_wait_all() {}
// ring drain (store spill): the emulator has no concurrent host thread, so a warp that finds the
// free queue empty runs the host's drain step itself (emu_fpop.cpp)
struct StorePool;
extern void (*psd_emu_ring_drain)(const StorePool* sp);
#define PSD_RING_WAIT(sp, pos) do { if (*(sp).ring.free_tail <= (pos)) psd_emu_ring_drain(&(sp)); if (*(sp).ring.free_tail <= (pos)) { fprintf(stderr, "warp_emu: ring drain freed nothing\n"); abort(); } } while (0)
// streaming (evict-first) stores of the cost-function store: plain stores here
static inline void psd_st_cs_i(int* p, int v) { *p = v; }
static inline void psd_st_cs_u64(unsigned long long* p, unsigned long long v) { *p = v; }
static inline void psd_st_cs_u4(unsigned* p, unsigned a, unsigned b, unsigned c, unsigned d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
