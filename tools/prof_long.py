#!/usr/bin/env python3
"""Per-row latency of ONE long problem per penalty (config-3 shape).
usage: python tools/prof_long.py [n_positions] [penalty ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
pens = [float(x) for x in sys.argv[2:]] or [1e3, 31622.7766016838, 1e6]
s, e, c = synth.poisson_problem(12345, n)
for pen in pens:
    plan = psd.Plan(0)
    pid = plan.add(s, e, c, pen)
    plan.run()
    st = plan.stats(); r = plan.loss_row(pid)
    print("penalty=%g rows=%d dp_ms=%.1f us/row=%.2f peaks=%d mean.intervals=%.2f max.intervals=%d overflow_tier=%d" % (
        pen, len(c), st["dp_ms"], 1e3 * st["dp_ms"] / len(c), r["peaks"], r["mean.intervals"], r["max.intervals"], st["n_overflow_tier"]), flush=True)
