#!/usr/bin/env python3
"""Latency kernel (one problem per block, one chain per warp) vs throughput kernel (one problem per
warp) on the same problems: per-row time of single problems and rows/s of small waves.  Results of
the two kernels are compared bit for bit on every case.
usage: python tools/prof_modes.py [quick]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
lib = psd._lib.lib


def run(probs, mode, max_blocks=2):
    lib.psd_set_option(b"latency_mode", float(mode))
    lib.psd_set_option(b"latency_max_blocks", float(max_blocks))
    plan = psd.Plan(0)
    for (s, e, c, pen) in probs:
        plan.add(s, e, c, pen)
    plan.upload(); plan.solve()          # warm-up (allocations, first launch)
    plan.solve(); plan.download()
    st = plan.stats()
    res = [(plan.loss_row(i), plan.segments(i)) for i in range(len(plan))]
    plan.close()
    lib.psd_set_option(b"latency_mode", 0.0); lib.psd_set_option(b"latency_max_blocks", 16.0)
    return st, res


def same(a, b):
    return all(x[0] == y[0] and all(np.array_equal(p, q) for p, q in zip(x[1], y[1])) for x, y in zip(a, b))


out = []
mono = synth.read_bedgraph(os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph"))[1:]
cases = [("mono27ac x1 pen 10.5", [mono + (10.5,)]),
         ("mono27ac x1 pen 1952.6", [mono + (1952.6,)])]
long_rows = synth.poisson_problem(2024, 100000 if quick else 400000)
for pen in (0.0, 1787.15394555632, 6037.57092280531):
    cases.append(("c3-like %d rows x1 pen %g" % (len(long_rows[2]), pen), [long_rows + (pen,)]))
inc = synth.increasing_problem(1000 if quick else 3000)
cases.append(("c5 increasing(%d) pens 1e2,1e4,1e6" % len(inc[2]), [inc + (p,) for p in (1e2, 1e4, 1e6)]))
wave_rows = [synth.poisson_problem(7000 + k, 8000 if quick else 20000) for k in range(148)]
for n_prob in (74, 148, 296, 592, 1184):
    cases.append(("wave %d problems x %d positions" % (n_prob, 8000 if quick else 20000),
                  [wave_rows[k % 148] + (synth.C2_PENALTIES[(k // 148) % 5],) for k in range(n_prob)]))
for name, probs in cases:
    rows = sum(len(p[2]) for p in probs)
    longest = max(len(p[2]) for p in probs)
    st_t, res_t = run(probs, 2)
    st_l, res_l = run(probs, 1, 16)
    rec = {"case": name, "problems": len(probs), "rows": rows,
           "throughput_kernel_ms": round(st_t["dp_ms"], 3), "latency_kernel_ms": round(st_l["dp_ms"], 3),
           "throughput_us_per_row_longest": round(1e3 * st_t["dp_ms"] / longest, 3), "latency_us_per_row_longest": round(1e3 * st_l["dp_ms"] / longest, 3),
           "throughput_rows_per_s": rows / st_t["dp_ms"] * 1e3, "latency_rows_per_s": rows / st_l["dp_ms"] * 1e3,
           "latency_piece_cap": st_l["piece_cap"], "latency_warps_per_sm": st_l["warps_per_sm"], "identical": bool(same(res_t, res_l)),
           "speedup": round(st_t["dp_ms"] / st_l["dp_ms"], 3)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
assert all(r["identical"] for r in out)
