#!/usr/bin/env python3
"""Device RLE on single very long count vectors (thousands of tiles in one look-back chain)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth
rng = np.random.default_rng(1)
for n in (20_000_000, 100_000_000):
    # long runs (many tiles without a head) mixed with noisy stretches
    parts = []
    left = n
    while left > 0:
        k = int(min(left, rng.integers(1, 200_000)))
        parts.append(np.full(k, rng.integers(0, 5), np.int32) if rng.random() < 0.5 else rng.poisson(2.0, k).astype(np.int32))
        left -= k
    v = np.concatenate(parts)
    plan = psd.Plan(0)
    t0 = time.time(); pid = plan.add_counts(v, 1000.0); t_add = time.time() - t0
    plan.upload()
    st = plan.stats()
    s, e, c = synth.rle_rows(v)
    print("n=%d rows=%d add=%.2fs rle_ms=%.3f GB/s=%.0f  (device run count verified against the host's at upload)" % (
        n, len(c), t_add, st["rle_ms"], st["rle_bytes_algorithmic"] / st["rle_ms"] / 1e6), flush=True)
    assert plan.result(pid).n_rows == len(c)
    del plan
