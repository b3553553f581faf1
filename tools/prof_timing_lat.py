#!/usr/bin/env python3
"""Experiment (PSD_TIMING build, `make -C peaksegdisk_b200/csrc timing`): cycle breakdown per row of the
LATENCY kernel (one problem per block, one chain per warp) on single problems.
usage: python tools/prof_timing_lat.py [mono|<positions>] [penalty]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PSD_LIB"] = os.path.join(ROOT, "peaksegdisk_b200", "libpsd_timing.so")
os.environ["PSD_LATENCY_MODE"] = "1"
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth, _lib
what = sys.argv[1] if len(sys.argv) > 1 else "mono"
pen = float(sys.argv[2]) if len(sys.argv) > 2 else 10.5
if what == "mono":
    _, s, e, c = synth.read_bedgraph(os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph"))
else:
    s, e, c = synth.poisson_problem(2024, int(what))
plan = psd.Plan(0)
plan.add(s, e, c, pen)
plan.upload(); plan.solve()
buf = (C.c_ulonglong * 32)()
_lib.lib.psd_debug_read_lat(buf, 32, 1)
plan.solve(); st = plan.stats()
_lib.lib.psd_debug_read_lat(buf, 32, 0)
rows = len(c)
names = {0: "min_less (warp 0)", 1: "min_more (warp 1)", 2: "min_env up (warp 0)", 3: "min_env down (warp 1)", 4: "after the barrier: counters, store (both warps summed)",
         5: "waiting at the row barrier (both warps summed)", 8: "  env: enumerate (both)", 9: "  env: pair rule loop (both)", 10: "  env: merge/emit (both)",
         16: "    pair: loads + eq flags", 17: "    pair: exp,exp,log,exp (dmid)", 18: "    pair: dl,dr,log,exp (two_roots)", 12: "    pair: root_left", 13: "    pair: root_right", 19: "    pair: post-Newton"}
print("%s rows=%d pen=%g dp_ms=%.2f -> %.0f cycles/row wall (1965 MHz), %.2f us/row" % (what, rows, pen, st["dp_ms"], st["dp_ms"] * 1e-3 * 1.965e9 / rows, 1e3 * st["dp_ms"] / rows))
for i in sorted(names):
    print("%-58s %8.0f cycles/row" % (names[i], buf[i] / rows))
