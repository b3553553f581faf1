// TEST TOOL: runs the product's device source (peaksegdisk_b200/csrc/fpop_warp.cuh) for one problem
// under the CPU warp emulator and exposes it with the oracle's in-memory signature, so tests can
// diff it against the oracle row by row without a GPU.  Not part of the product.
#define PSD_EMU 1
#include <vector>
#include <cmath>
#include "warp_emu.h"
#include "../../peaksegdisk_b200/csrc/fpop_warp.cuh"

namespace {
struct Job {
  DpProblem pb; WarpWs ws; WarpWs ws_g; StorePool sp; DpResult* res;
  psd_trace_fn trace; void* trace_user;
  int* seg_row; double* seg_x;
  int order0 = 0; int cursor = 1;
#if defined(PSD_G32)
  LatShared lat;
  bool helpers = true;     // PSD_EMU_NO_HELPERS=1: two warps only, each chain solves both Newton roots itself
#endif
};
// The host side of the store's DMA drain (fpop_gpu.cu: RingDrain), done synchronously: copy every
// published ring slot to its place in the "host" region and hand the slot back.
struct EmuRing { unsigned long long next = 0, ftail = 0, drained = 0; } g_ring;
void emu_ring_drain(const StorePool* sp) {
  const StoreRing& r = sp->ring;
  for (;;) {
    volatile unsigned long long* e = r.done_q + 2ull * (g_ring.next & r.q_mask);
    if (e[0] != g_ring.next + 1ull) break;
    const unsigned long long host_chunk = e[1] >> 32, slot = e[1] & 0xffffffffull;
    memcpy(sp->host_base + host_chunk * sp->chunk_bytes, r.base + slot * sp->chunk_bytes, sp->chunk_bytes);
    ((unsigned int*)r.free_q)[g_ring.ftail & r.q_mask] = (unsigned int)slot;
    g_ring.ftail++; g_ring.next++; g_ring.drained++;
    *(unsigned long long*)r.free_tail = g_ring.ftail;
  }
}
#if defined(PSD_G32)
// latency kernel: a block of two warps owns the problem (one chain per warp); the backtrack is one warp
void lane_main(void* arg) {
  Job* J = (Job*)arg;
  if (psd_warp_in_block() >= 2) lat_helper_loop(&J->lat.help[psd_warp_in_block() - 2], psd_warp_in_block() - 2);   // helper warps
  else dp_run_latency(J->ws, J->ws_g, J->pb, J->res, J->sp, &J->lat, J->helpers, J->trace, J->trace_user);
  psd_cta_sync();
  if (J->sp.ring.n_slots && psd_warp_in_block() == 0 && psd_lane() == 0) emu_ring_drain(&J->sp);
  psd_cta_sync();
  if (psd_warp_in_block() == 0) backtrack_problem(J->sp, J->pb.index, J->pb.n_rows, J->res, J->seg_row, J->seg_x);
}
#define PSD_EMU_WARPS (J.helpers ? 4 : 2)
#else
void lane_main(void* arg) {
  Job* J = (Job*)arg;
  DpQueue Q;
  Q.problems = &J->pb; Q.order = &J->order0; Q.n_order = 1; Q.cursor = &J->cursor; Q.results = J->res; Q.first_slot = 0;
  dp_run_queue(J->ws, J->ws_g, Q, J->sp, J->trace, J->trace_user);
  psd_syncwarp();
  if (J->sp.ring.n_slots && psd_lane() == 0) emu_ring_drain(&J->sp);
  psd_syncwarp();
  backtrack_problem(J->sp, J->pb.index, J->pb.n_rows, J->res, J->seg_row, J->seg_x);
}
#define PSD_EMU_WARPS 1
#endif
}  // namespace

#if defined(PSD_EMU_STATS)
unsigned long long psd_emu_stats[8][65];
#endif

extern "C" {
#if defined(PSD_EMU_STATS)
void emu_stats_read(int slot, unsigned long long* out65, int reset) {
  for (int i = 0; i < 65; i++) { out65[i] = psd_emu_stats[slot][i]; if (reset) psd_emu_stats[slot][i] = 0; }
}
#endif


// Same outputs as oracle_fpop_rows (non-trivial problems only).  cap = list capacity to emulate.
// Returns the DpResult status (0 ok, 101 piece overflow, ...).
int emu_fpop_rows(int n_rows, const int* chrom_start, const int* chrom_end, const int* coverage,
                  double penalty, int cap, int spill_cap, int descending, double* out_summary,
                  int* seg_start, int* seg_end, int* seg_peak, double* seg_mean,
                  psd_trace_fn trace, void* trace_user, int* n_spills) {
  std::vector<int> w(n_rows);
  double W = 0, dmin = INFINITY, dmax = -INFINITY;
  for (int t = 0; t < n_rows; t++) {
    w[t] = chrom_end[t] - chrom_start[t]; W += w[t];
    double lx = psd_log((double)coverage[t], psd_log_tab_host);
    if (lx < dmin) dmin = lx;
    if (dmax < lx) dmax = lx;
  }
  Job J;
  std::vector<unsigned long long> index(n_rows);
  J.pb.weight = w.data(); J.pb.coverage = coverage; J.pb.n_rows = n_rows; J.pb.penalty = penalty;
  J.pb.dmin = dmin; J.pb.dmax = dmax; J.pb.index = index.data();
  // shared-memory-tier stand-in (capacity cap) and the global workspace the warp can move to
  const int ccap = 3 * cap;
  std::vector<double> wsmem(PSD_WS_BYTES(cap, ccap) / 8 + 16);
  J.ws.base = (unsigned char*)wsmem.data(); J.ws.scratch = nullptr; J.ws.flags = (int*)J.ws.base; J.ws.cap = cap; J.ws.ccap = ccap; J.ws.help = nullptr;
#if defined(PSD_G32)
  J.helpers = getenv("PSD_EMU_NO_HELPERS") == nullptr;
#endif
  std::vector<double> wsmem_g(spill_cap > 0 ? PSD_WS_BYTES(spill_cap, 3 * spill_cap) / 8 + 16 : 1);
  J.ws_g.base = spill_cap > 0 ? (unsigned char*)wsmem_g.data() : nullptr; J.ws_g.scratch = nullptr; J.ws_g.flags = J.ws.flags;
  J.ws_g.cap = spill_cap; J.ws_g.ccap = 3 * spill_cap; J.ws_g.help = nullptr;
  const unsigned long long chunk = 1 << 16;
  std::vector<unsigned char> pool;
  unsigned long long cursor = 0;
  // generous pool: header + 20 bytes per piece, pieces <= cap per function
  unsigned long long pool_bytes = (unsigned long long)n_rows * (32ull + 40ull * (unsigned)(cap > spill_cap ? cap : spill_cap) + 64ull) + chunk;
  if (pool_bytes > (6ull << 30)) pool_bytes = 6ull << 30;
  pool.resize((pool_bytes / chunk + 1) * chunk);
  // PSD_EMU_HBM_CHUNKS (test knob): only that many chunks count as "HBM", the rest of the buffer
  // plays the pinned-host spill region
  unsigned long long hbm_chunks = pool.size() / chunk, host_cursor = 0;
  if (const char* e = getenv("PSD_EMU_HBM_CHUNKS")) { const unsigned long long v = strtoull(e, 0, 10); if (v < hbm_chunks) hbm_chunks = v; }
  J.sp.base = pool.data(); J.sp.cursor = &cursor; J.sp.n_chunks = hbm_chunks; J.sp.chunk_bytes = chunk;
  J.sp.host_base = pool.data() + hbm_chunks * chunk; J.sp.host_cursor = &host_cursor; J.sp.host_chunks = pool.size() / chunk - hbm_chunks;
  // PSD_EMU_RING_SLOTS (test knob): spilled chunks go through a ring of that many "HBM" slots and the
  // drain protocol instead of being written in place
  memset(&J.sp.ring, 0, sizeof J.sp.ring);
  std::vector<unsigned char> ring_mem; std::vector<unsigned int> free_q; std::vector<unsigned long long> done_q;
  unsigned long long ring_head = 0, done_head = 0, free_tail = 0;
  if (const char* e = getenv("PSD_EMU_RING_SLOTS")) {
    const unsigned long long ns = strtoull(e, 0, 10);
    if (ns > 0) {
      unsigned q_len = 4; while (q_len < 4 * ns) q_len <<= 1;
      ring_mem.resize(ns * chunk); free_q.assign(q_len, 0); done_q.assign(2ull * q_len, 0);
      for (unsigned long long i = 0; i < ns; i++) free_q[i] = (unsigned int)i;
      free_tail = ns;
      J.sp.ring.base = ring_mem.data(); J.sp.ring.n_slots = ns; J.sp.ring.head = &ring_head; J.sp.ring.done_head = &done_head;
      J.sp.ring.free_tail = &free_tail; J.sp.ring.free_q = free_q.data(); J.sp.ring.done_q = done_q.data(); J.sp.ring.q_mask = q_len - 1;
      g_ring = EmuRing(); g_ring.ftail = ns;
      psd_emu_ring_drain = emu_ring_drain;
    }
  }
  DpResult res; res.status = -1;
  J.res = &res; J.trace = trace; J.trace_user = trace_user;
  std::vector<int> seg_row(n_rows + 1); std::vector<double> seg_x(n_rows + 1);
  J.seg_row = seg_row.data(); J.seg_x = seg_x.data();
  psd_emu::run_block(lane_main, &J, PSD_EMU_WARPS, descending);
  if (res.status != 0) return res.status;
  const int ns = res.n_segments, np = (ns - 1) / 2;
  out_summary[0] = penalty; out_summary[1] = ns; out_summary[2] = np; out_summary[3] = W; out_summary[4] = n_rows;
  out_summary[5] = res.best_cost; out_summary[6] = res.best_cost * W - penalty * np; out_summary[7] = res.n_equality;
  out_summary[8] = (double)res.total_intervals / (n_rows * 2); out_summary[9] = res.max_intervals;
  if (n_spills) *n_spills = res.pad_;
  if (getenv("PSD_EMU_RING_SLOTS") && getenv("PSD_EMU_RING_REPORT")) fprintf(stderr, "emu ring: %llu chunks drained\n", g_ring.drained);
  int prev_end = chrom_end[n_rows - 1];
  for (int s = 0; s < ns; s++) {
    const int st = (s < ns - 1) ? chrom_end[seg_row[s]] : chrom_start[0];
    seg_start[s] = st; seg_end[s] = prev_end; seg_peak[s] = s & 1; seg_mean[s] = psd_exp(seg_x[s], psd_exp_tab_host);
    prev_end = st;
  }
  return 0;
}

}  // extern "C"
