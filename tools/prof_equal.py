#!/usr/bin/env python3
"""Experiment driver: n_problems equal-length problems (rows truncated to `rows`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth
npb = int(sys.argv[1]); rows = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
plan = psd.Plan(0)
tot = 0
k = 0
while len(plan) < npb:
    s, e, c = synth.poisson_problem(k, int(rows * 1.6)); k += 1
    if len(c) < rows: continue
    for pen in synth.C2_PENALTIES:
        if len(plan) < npb:
            plan.add(s[:rows], e[:rows], c[:rows], pen); tot += rows
plan.upload()
for _ in range(reps):
    plan.solve(); st = plan.stats()
    print("problems=%d rows=%d dp_ms=%.2f rows/s=%.3e warps/sm=%d" % (len(plan), tot, st["dp_ms"], tot / (st["dp_ms"] / 1e3), st["warps_per_sm"]), flush=True)
plan.download()
assert all(plan.result(i).status == 0 for i in range(len(plan)))
