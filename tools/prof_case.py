#!/usr/bin/env python3
"""Small fixed workload for ncu: one batched solve (DP kernel + backtrack kernel).
usage: python tools/prof_case.py [n_vectors] [n_per_vector] [n_solves]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 480
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
plan = psd.Plan(0)
rows = 0
for k in range(nv):
    s, e, c = synth.poisson_problem(k, n if n > 0 else None)     # 0: config 2's own vector lengths (1e4..1e5)
    for pen in ([float(os.environ['PSD_PEN'])] * 5 if os.environ.get('PSD_PEN') else synth.C2_PENALTIES):
        plan.add(s, e, c, pen); rows += len(c)
plan.upload()
for _ in range(reps):
    t0 = time.time(); plan.solve(); dt = time.time() - t0
    st = plan.stats()
    print("problems=%d rows=%d dp_ms=%.2f bt_ms=%.3f rows/s=%.3e alg_GB/s=%.2f warps/sm=%d" % (
        len(plan), rows, st["dp_ms"], st["backtrack_ms"], rows / (st["dp_ms"] / 1e3), st["store_bytes_algorithmic"] / st["dp_ms"] / 1e6, st["warps_per_sm"]), flush=True)
plan.download()
assert all(plan.result(i).status == 0 for i in range(len(plan)))
