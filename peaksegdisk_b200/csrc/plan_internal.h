// plan_internal.h -- host-side types shared by fpop_gpu.cu (device plan) and host_api.cpp (C ABI,
// bedGraph text I/O).  Not installed; the public surface is include/peaksegdisk_b200.h.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <memory>
#include <thread>
#include <string>
#include <vector>
#include "../../include/peaksegdisk_b200.h"

// The rows of one bedGraph (or one psd_plan_add call) and their pass-1 totals.  Problems that solve
// the same file at different penalties share one RowData: it is parsed, summed, packed and copied to
// the device once.
struct RowData {
  std::vector<int32_t> chrom_start, chrom_end, coverage, weight;
  double bases = 0, sum_wz = 0;   // pass-1 totals (src/PeakSegFPOPLog.cpp:187-189)
  double dmin = 0, dmax = 0;      // log(min coverage), log(max coverage)
};

struct HostProblem {
  int status = 0;                 // reference-style input status (0 ok)
  bool trivial = false;           // one-segment model: penalty Inf or constant coverage
  bool penalty_is_inf = false;
  double penalty = 0;
  int64_t n_rows = 0;
  std::shared_ptr<RowData> rows;  // row problems (null for count-vector problems)
  // count-vector problems (psd_plan_add_counts): no row arrays on the host, the device run-length
  // encodes `counts`; positions are 0..n_pos and seg_row then holds coordinates, not row numbers
  bool from_counts = false;
  int64_t n_pos = 0;
  std::vector<int32_t> counts;
  int64_t raw_off = 0;            // offset into the packed device count buffer
  double bases = 0, sum_wz = 0;   // pass-1 totals (src/PeakSegFPOPLog.cpp:187-189)
  double dmin = 0, dmax = 0;      // log(min coverage), log(max coverage)
  int64_t row_off = 0;            // offset into the packed device row arrays (shared by the problems of one RowData)
  int64_t index_off = 0;          // offset into the device record-index array (one entry per row per PROBLEM)
  // results
  int result_status = -1;         // -1 not solved yet
  int n_segments = 0, n_equality = 0;
  double best_cost = 0, total_intervals = 0, max_intervals = 0;
  std::vector<int> seg_row;       // last row of the previous segment, last segment first (-1 for the first)
  std::vector<double> seg_x;      // log-mean per segment
};

inline int hp_first_start(const HostProblem& h) { return h.from_counts ? 0 : h.rows->chrom_start[0]; }
inline int hp_last_end(const HostProblem& h) { return h.from_counts ? (int)h.n_pos : h.rows->chrom_end[h.n_rows - 1]; }
// chromStart of segment s (segments are stored last first; s < n_segments - 1)
inline int hp_seg_start(const HostProblem& h, int s) { return h.from_counts ? h.seg_row[s] : h.rows->chrom_end[h.seg_row[s]]; }

// PSD_TRACE=1: wall-clock stage marks on stderr (where does a small solve's latency go?)
struct Trace {
  bool on; std::chrono::steady_clock::time_point t;
  Trace() : on(getenv("PSD_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[psd trace] %-34s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Runs fn(i) for i in [0,n) on up to hardware_concurrency host threads.
template <class F> void parallel_for(int n, F fn) {
  int nt = (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > n) nt = n;
  if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; t++) pool.emplace_back([&]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i); } });
  for (auto& th : pool) th.join();
}

struct psd_plan;
psd_plan* psd_plan_create_impl(int device);
void psd_plan_destroy_impl(psd_plan* p);
psd_plan* psd_plan_acquire_parked();          // file entry points: reuse the previous call's buffers
void psd_plan_release_parked(psd_plan* p);
void psd_plan_drop_parked();
std::vector<HostProblem>& psd_plan_problems(psd_plan* p);
const std::vector<HostProblem>& psd_plan_problems_c(const psd_plan* p);
void psd_plan_invalidate(psd_plan* p);
void psd_plan_mark_penalty_changed(psd_plan* p);
const psd_stats& psd_plan_stats_ref(const psd_plan* p);
int psd_plan_upload_impl(psd_plan* p, void* stream);
int psd_plan_solve_impl(psd_plan* p, void* stream);
int psd_plan_download_impl(psd_plan* p, void* stream);
int psd_plan_store_function_impl(psd_plan* p, int id, int row, int which, int cap, int* n_out, double* hi, int* back_i, double* back_x);
int psd_device_count_impl();
int psd_set_option_impl(const char* name, double value);
int psd_option_devices();                      // option "devices" / env PSD_DEVICES: GPUs one batched file call may use
void psd_set_last_error(const std::string& s);
const std::string& psd_get_last_error();
