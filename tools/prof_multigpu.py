#!/usr/bin/env python3
"""One batched file call spread over the GPUs of the box (option "devices"): Mono27ac at 2,000
penalties, wall time with 1 and with all GPUs.  usage: python tools/prof_multigpu.py"""
import os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import peaksegdisk_b200 as psd
from peaksegdisk_b200.api import r_paste

mono = os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph")
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
f = os.path.join(tmp, "m.bedGraph"); shutil.copy(mono, f)
pens = [r_paste(p) for p in np.exp(np.linspace(np.log(1.0), np.log(1e6), 2000))]
ndev = psd._lib.lib.psd_device_count()
for k in sorted({1, 2, ndev}):
    if k > ndev:
        continue
    psd._lib.lib.psd_set_option(b"devices", float(k))
    for rep in range(2):
        t0 = time.time(); st = psd.PeakSegFPOP_file_batch([f] * len(pens), pens); dt = time.time() - t0
        assert st == [0] * len(pens)
    print("devices=%d: %.3f s for %d problems (%.2f M rows*penalties/s)" % (k, dt, len(pens), 6921 * len(pens) / dt / 1e6), flush=True)
shutil.rmtree(tmp)
