#!/usr/bin/env python3
"""CPU-only statistics from the warp emulator (PSD_EMU_STATS build of the device source): how large
are the quantities that decide the number of 16-/32-lane passes per operator call?  Those passes are
what the warps of a block wait for at the phase barriers (profiles/README.md).
usage: python tools/emu_stats.py [n_positions=30000] [n_vectors=4]"""
import ctypes as C, multiprocessing as mp, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
LIB = os.path.join(ROOT, "tests", "_build", "libpsd_emu_stats.so")
NAMES = ["min_mono: input pieces", "min_mono: flat-stretch windows (Newton rounds)", "min_env: pieces of the previous function",
         "min_env: overlap intervals of one chain", "min_env: intervals of both chains (pooled stage)", "min_env: candidates of one chain",
         "min_env: intervals of both chains that need exp/log", "min_env: intervals of both chains that run Newton"]

def build():
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-DPSD_EMU_STATS", "-Wno-unused-variable",
                           "-Wno-unused-function", "-I" + emu, "-x", "c++", "-shared", "-o", LIB, os.path.join(emu, "emu_fpop.cpp"), os.path.join(emu, "warp_emu.cpp")])

def one(args):
    seed, n, pen = args
    from peaksegdisk_b200 import synth
    import emu_compare as ec
    emu = C.CDLL(LIB)
    emu.emu_fpop_rows.restype = C.c_int
    s, e, c = synth.poisson_problem(seed, n)
    dummy = (C.c_ulonglong * 65)()
    for k in range(8):
        emu.emu_stats_read(k, dummy, 1)          # a pool worker runs several jobs: start from zero
    st, summ, seg, _ = ec.run_emu(emu, s, e, c, pen, cap=48, spill_cap=512)
    out = np.zeros((8, 65), np.uint64)
    for k in range(8):
        emu.emu_stats_read(k, out[k].ctypes.data_as(C.POINTER(C.c_ulonglong)), 0)
    return pen, len(c), st, out

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    build()
    from peaksegdisk_b200 import synth
    jobs = [(seed, n, pen) for seed in range(nv) for pen in synth.C2_PENALTIES]
    with mp.Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(one, jobs)
    assert all(r[2] == 0 for r in res)
    for pens in [[p] for p in synth.C2_PENALTIES] + [synth.C2_PENALTIES]:
        tot = sum(r[3] for r in res if r[0] in pens)
        rows = sum(r[1] for r in res if r[0] in pens)
        print("penalties %s (%d rows):" % (pens, rows))
        for k, name in enumerate(NAMES):
            h = tot[k].astype(float); cnt = h.sum()
            mean = (h * np.arange(65)).sum() / cnt
            print("  %-52s mean %5.1f   >16: %5.1f %%   >32: %5.2f %%   >48: %5.2f %%" % (
                name, mean, 100 * h[17:].sum() / cnt, 100 * h[33:].sum() / cnt, 100 * h[49:].sum() / cnt))
