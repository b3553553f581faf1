#!/usr/bin/env python3
"""Statistics (PSD_TIMING build): how many Newton rounds does min_env run per call, and how many
would a job list compacted across the 16-interval chunks need?  Config-2 mix of penalties.
usage: python tools/prof_newton_rounds.py [n_vectors] [n_positions]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PSD_LIB"] = os.path.join(ROOT, "peaksegdisk_b200", "libpsd_timing.so")
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth, _lib
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
for pens in ([1e2], [1e3], [1e4], [1e5], [1e6], synth.C2_PENALTIES):
    plan = psd.Plan(0); rows = 0
    for k in range(nv):
        s, e, c = synth.poisson_problem(k, n)
        for pen in pens:
            plan.add(s, e, c, pen); rows += len(c)
    plan.upload()
    buf = (C.c_ulonglong * 32)()
    _lib.lib.psd_debug_read(buf, 32, 1)
    plan.solve()
    _lib.lib.psd_debug_read(buf, 32, 0)
    calls = buf[24]
    print("penalties=%s: min_env calls/row=%.2f  intervals/call=%.1f  chunks/call=%.2f  calls with >16 intervals=%.1f%%  "
          "Newton jobs/call=%.2f  rounds/call=%.3f  compacted rounds/call=%.3f  calls with Newton=%.1f%%" % (
              pens, calls / rows, buf[21] / calls, buf[20] / calls, 100.0 * buf[27] / calls, buf[22] / calls, buf[23] / calls,
              buf[26] / calls, 100.0 * buf[25] / calls), flush=True)
