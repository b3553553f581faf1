#!/usr/bin/env python3
"""Device run-length encoding (rle_gpu.cuh) on config-2 count vectors: rate of the four RLE kernels
against the HBM roofline, and the host-side alternative (numpy RLE + row upload) beside it.
usage: python tools/prof_rle.py [n_vectors]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
vecs = [synth.poisson_counts(s % 1024).astype(np.int32) for s in range(nv)]
peak = 6554.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
plan = psd.Plan(0)
t0 = time.time()
for v in vecs:
    plan.add_counts(v, 1000.0)
t_add = time.time() - t0
best = None
for rep in range(5):
    t0 = time.time(); plan.upload(); t_up = time.time() - t0
    st = plan.stats()
    if best is None or st["rle_ms"] < best["rle_ms"]:
        best = dict(st, upload_wall_ms=1e3 * t_up)
pos = best["rle_positions"]; rows = (best["rle_bytes_algorithmic"] - 4 * pos) // 12
gbps = best["rle_bytes_algorithmic"] / best["rle_ms"] / 1e6
t0 = time.time(); rows_host = [synth.rle_rows(v) for v in vecs]; t_np = time.time() - t0
plan2 = psd.Plan(0)
t0 = time.time()
for s, e, c in rows_host:
    plan2.add(s, e, c, 1000.0)
t_add2 = time.time() - t0
t0 = time.time(); plan2.upload(); t_up2 = time.time() - t0
print(json.dumps({"vectors": nv, "positions": pos, "rows": int(rows), "rle_ms": best["rle_ms"], "rle_launches": best["n_rle_launches"],
                  "algorithmic_bytes": best["rle_bytes_algorithmic"], "achieved_gbps": gbps, "peak_gbps": peak, "frac": gbps / peak,
                  "h2d_ms": best["h2d_ms"], "h2d_bytes": best["h2d_bytes"], "upload_wall_ms_counts": best["upload_wall_ms"],
                  "add_counts_wall_s": t_add, "host_numpy_rle_s": t_np, "add_rows_wall_s": t_add2,
                  "upload_wall_ms_rows_first": 1e3 * t_up2, "h2d_bytes_rows": plan2.stats()["h2d_bytes"]}))
