#!/usr/bin/env python3
"""Config 4 at reduced scale: hg19-shaped problems (24 chromosomes x S samples, rows proportional to
chromosome length, Mono27ac-like weights), one penalty each, with the HBM pool optionally capped so
that part of the cost-function store spills to pinned host memory.
usage: tools/run_config4_lite.py [samples=2] [row_scale=0.05] [store_gb=0 (auto)] [check=1]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
store_gb = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
check = int(sys.argv[4]) if len(sys.argv) > 4 else 1
chrom, s0, e0, c0 = synth.read_bedgraph(os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph"))
w0 = (e0 - s0).astype(np.int64)
w0 = np.minimum(w0, 2000)          # keep coordinates below 2^31 when tiled
if store_gb > 0:
    psd._lib.lib.psd_set_option(b"store_gb", store_gb)
probs = []
t0 = time.time()
for ci in range(24):
    for si in range(samples):
        probs.append(synth.hg19_problem(ci, si, w0, c0.astype(np.float64), scale_rows=scale))
print("generated %d problems, %d rows total, max %d rows, in %.1f s" % (len(probs), sum(len(p[2]) for p in probs), max(len(p[2]) for p in probs), time.time() - t0), flush=True)
plan = psd.Plan(0)
ids = [plan.add(s, e, c, pen) for (s, e, c, pen) in probs]
t0 = time.time(); plan.run(); wall = time.time() - t0
st = plan.stats()
rows = st["rows_solved"]
print("solved: wall %.2f s, dp %.0f ms, backtrack %.2f ms, %.3e rows/s, store written %.2f GB (algorithmic %.2f GB), spilled to host %.2f GB, waves %d, tier-switched problems %d" % (
    wall, st["dp_ms"], st["backtrack_ms"], rows / (st["dp_ms"] / 1e3), st["store_bytes_written"] / 1e9, st["store_bytes_algorithmic"] / 1e9,
    st["store_bytes_spilled_host"] / 1e9, st["n_waves"], st["n_overflow_tier"]), flush=True)
assert all(plan.result(i).status == 0 for i in ids)
if check:
    import oracle_bind
    k = int(np.argmin([len(p[2]) for p in probs]))
    s, e, c, pen = probs[k]
    t0 = time.time()
    ost, summ, oseg = oracle_bind.solve_rows(s, e, c, pen)
    got = plan.loss_row(ids[k]); seg = plan.segments(ids[k])
    ok = ost == 0 and got["segments"] == int(summ[1]) and got["total.loss"] == summ[6] and np.array_equal(seg[0], oseg[0]) and np.array_equal(seg[1], oseg[1])
    print("oracle check on the smallest problem (%d rows, %.1f s on the CPU): %s, peaks=%d" % (len(c), time.time() - t0, "identical" if ok else "MISMATCH", got["peaks"]))
    assert ok
