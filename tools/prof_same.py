#!/usr/bin/env python3
"""Experiment: every warp solves the SAME problem (zero phase imbalance) vs distinct problems."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth
npb = int(sys.argv[1]); rows = int(sys.argv[2]); same = int(sys.argv[3])
plan = psd.Plan(0)
tot = 0
s0, e0, c0 = synth.poisson_problem(7, int(rows * 1.6))
k = 0
while len(plan) < npb:
    if same: s, e, c = s0, e0, c0
    else:
        s, e, c = synth.poisson_problem(k, int(rows * 1.6)); k += 1
        if len(c) < rows: continue
    plan.add(s[:rows], e[:rows], c[:rows], 1000.0); tot += rows
plan.upload()
for _ in range(2):
    plan.solve(); st = plan.stats()
print("same=%d problems=%d rows=%d dp_ms=%.2f rows/s=%.3e" % (same, len(plan), tot, st["dp_ms"], tot / (st["dp_ms"] / 1e3)), flush=True)
