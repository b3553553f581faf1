#!/usr/bin/env python3
"""Experiment (PSD_TIMING build): when does each SM's block run out of work in the config-2 bench batch?"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PSD_LIB"] = os.path.join(ROOT, "peaksegdisk_b200", "libpsd_timing.so")
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth, _lib
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plan = psd.Plan(0)
for k in range(nv):
    s, e, c = synth.poisson_problem(k)
    for pen in synth.C2_PENALTIES:
        plan.add(s, e, c, pen)
plan.upload(); plan.solve(); plan.solve()
st = plan.stats()
buf = (C.c_ulonglong * 160)()
_lib.lib.psd_debug_block_ends(buf)
t = np.array(buf[:148], dtype=np.float64)
t = (t - t.max()) / 1e6   # ms before the last block ended
print("dp_ms %.1f; blocks ended (ms before the last one): min %.1f p10 %.1f median %.1f p90 %.1f" % (
    st["dp_ms"], t.min(), np.percentile(t, 10), np.median(t), np.percentile(t, 90)))
print("mean idle fraction of the SMs at the tail: %.3f" % (float(np.mean(-t)) / st["dp_ms"]))
