#!/usr/bin/env python3
"""Where does the wall time of ONE small solve go?  Stage timings of the plan API and of the file
entry point on Mono27ac (6,921 rows, penalty 10.5).  usage: python tools/prof_latency.py"""
import os, sys, time, shutil, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth

mono = os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph")
_, s, e, c = synth.read_bedgraph(mono)
for rep in range(3):
    t = [time.time()]
    plan = psd.Plan(0); t.append(time.time())
    plan.add(s, e, c, 10.5); t.append(time.time())
    plan.upload(); t.append(time.time())
    plan.solve(); t.append(time.time())
    plan.download(); t.append(time.time())
    st = plan.stats()
    del plan; t.append(time.time())
    names = ["create", "add", "upload", "solve", "download", "destroy"]
    print("plan rep %d: " % rep + "  ".join("%s=%.1fms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])) +
          "  [dp kernel %.1f ms, backtrack %.3f ms]" % (st["dp_ms"], st["backtrack_ms"]), flush=True)
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
f = os.path.join(tmp, "m.bedGraph"); shutil.copy(mono, f)
for rep in range(3):
    t0 = time.time(); psd.PeakSegFPOP_file(f, "10.5"); print("file entry rep %d: %.1f ms" % (rep, 1e3 * (time.time() - t0)), flush=True)
shutil.rmtree(tmp)
