#!/usr/bin/env python3
"""Experiment (PSD_TIMING build): cycle breakdown per row."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PSD_LIB"] = os.path.join(ROOT, "peaksegdisk_b200", "libpsd_timing.so")
os.environ["PSD_LATENCY_MODE"] = "2"      # this tool reads the throughput kernel's counters
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth, _lib
npb = int(sys.argv[1]); rows = int(sys.argv[2]); pen = float(sys.argv[3]) if len(sys.argv) > 3 else 1000.0
plan = psd.Plan(0)
k = 0; tot = 0
while len(plan) < npb:
    s, e, c = synth.poisson_problem(k, int(rows * 1.6)); k += 1
    if len(c) < rows: continue
    plan.add(s[:rows], e[:rows], c[:rows], pen); tot += rows
plan.upload(); plan.solve()
buf = (C.c_ulonglong * 32)()
_lib.lib.psd_debug_read(buf, 32, 1)
hist = (C.c_ulonglong * 256)()
_lib.lib.psd_debug_hist(hist, 1)
plan.solve(); st = plan.stats()
_lib.lib.psd_debug_hist(hist, 0)
_lib.lib.psd_debug_read(buf, 32, 0)
names = {0: "min_less (g0)", 1: "min_more (g1)", 2: "min_env up (g0)", 3: "min_env down (g1)", 4: "epilogue (both lanes counted)", 5: "wait at barrier 1, after min_less/min_more (both lanes counted)", 6: "wait at barrier 2, after min_env (both lanes counted)",
         8: "  env: enumerate", 9: "  env: pair rule loop", 10: "  env: merge/emit", 16: "    pair: loads + eq flags", 17: "    pair: exp,exp,log,exp (dmid)", 18: "    pair: dl,dr,log,exp (two_roots)", 19: "    pair: post-Newton", 20: "    pair-loop chunks (count/row)", 21: "    intervals (count/row)", 12: "    pair: root_left", 13: "    pair: root_right"}
print("problems=%d rows=%d dp_ms=%.2f -> %.0f cycles/row wall (1965 MHz)" % (npb, tot, st["dp_ms"], st["dp_ms"] * 1e-3 * 1.965e9 / rows / max(1, (npb + 147) // 148 / 14 if npb > 148 else 1)))
for i in sorted(names):
    print("%-66s %8.0f cycles/row" % (names[i], buf[i] / tot))

# distribution of the phase durations over (warp, row): what the other warps of a block wait for
for slot, name in enumerate(["min_less/min_more", "min_env, small", "min_env, large (> 36 pieces in its four lists)"]):
    h = [hist[64 * slot + b] for b in range(64)]
    n = sum(h)
    if not n:
        continue
    cum, marks = 0, {}
    for b, v in enumerate(h):
        cum += v
        for q in (0.1, 0.5, 0.9, 0.99, 0.999):
            if q not in marks and cum >= q * n:
                marks[q] = (b + 1) * 2048
    mean = sum((b + 0.5) * 2048 * v for b, v in enumerate(h)) / n
    print("%-48s n=%9d mean=%6.0f  p10=%6d p50=%6d p90=%6d p99=%6d p99.9=%6d cycles" % (name, n, mean, marks[0.1], marks[0.5], marks[0.9], marks[0.99], marks[0.999]))
