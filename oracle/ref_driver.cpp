// TEST INFRASTRUCTURE ONLY.  Thin C-linkage driver around the UNMODIFIED reference solver
// (compiled from /root/reference/src by oracle/Makefile into oracle/_ref/).
//   ref_fpop <bedGraph> <penalty> <db>          -> exit status = PeakSegFPOP_disk status
//   libref_fpop.so: ref_fpop_disk(), ref_fpop_batch() (a pool of host threads, one problem each;
//   the reference keeps no global state, SURVEY.md 8b "Threading").
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include <cstdio>
#include "PeakSegFPOPLog.h"

extern "C" int ref_fpop_disk(const char* bedGraph, const char* penalty, const char* db) {
  return PeakSegFPOP_disk((char*)bedGraph, (char*)penalty, (char*)db);
}

// Runs problems [0,n) on n_threads workers; returns wall seconds. status_out[i] = solver status.
extern "C" double ref_fpop_batch(int n, const char* const* bedGraph, const char* const* penalty,
                                 const char* const* db, int n_threads, int* status_out) {
  std::atomic<int> next(0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; t++) {
    pool.emplace_back([&]() {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        status_out[i] = PeakSegFPOP_disk((char*)bedGraph[i], (char*)penalty[i], (char*)db[i]);
        std::remove(db[i]);
      }
    });
  }
  for (auto& th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

#ifdef REF_FPOP_MAIN
int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s bedGraph penalty db\n", argv[0]); return 100; }
  return PeakSegFPOP_disk(argv[1], argv[2], argv[3]);
}
#endif
