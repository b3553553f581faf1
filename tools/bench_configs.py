#!/usr/bin/env python3
"""GPU library vs the unmodified reference (oracle/_ref/libref_fpop.so) on the box's host cores for
the BASELINE.json configurations that are NOT the bench line (config 2 is bench.py):

  c1  Mono27ac (6,921 rows), penalty "10.5", one problem, through the file entry point
  c3  one long problem (default 1e6 rows), a few penalties of a search chain (per-solve latency)
  c5  increasing counts (default N=1e4: cost functions of thousands of pieces, global tier)
  c1x Mono27ac x 1,000 penalties in one batched call (what the plan API is for)
  c2f 128 config-2 vectors as bedGraph FILES x 5 penalties, one batched file call (text parse and
      result files included on both sides)

Prints one JSON object; both sides produce the same files and the outputs are compared byte for byte.
usage: python tools/bench_configs.py [--c3-rows N] [--c5-rows N] [--skip c3,c5]"""
import argparse, ctypes as C, json, os, shutil, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth
from peaksegdisk_b200.api import r_paste
import oracle_bind


def ref_batch(files, pens, dbs, threads):
    lib = C.CDLL(oracle_bind.REF_SO)
    lib.ref_fpop_batch.restype = C.c_double
    n = len(files)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    st = (C.c_int * n)()
    secs = lib.ref_fpop_batch(n, arr(files), arr(pens), arr(dbs), threads, st)
    assert all(s == 0 for s in st), list(st)
    return secs


def outputs(path, pen):
    return [open("%s_penalty=%s_%s" % (path, pen, suffix), "rb").read() for suffix in ("segments.bed", "loss.tsv")]


def both(tag, files, pens, tmp, threads, rows):
    """Runs the (file, penalty) problems through the reference and through our batched file entry."""
    gdir, rdir = os.path.join(tmp, tag + "_gpu"), os.path.join(tmp, tag + "_ref")
    os.makedirs(gdir); os.makedirs(rdir)
    gfiles, rfiles = [], []
    for f in files:
        for d, lst in ((gdir, gfiles), (rdir, rfiles)):
            dst = os.path.join(d, os.path.basename(f))
            if not os.path.exists(dst):
                shutil.copy(f, dst)
            lst.append(dst)
    dbs_g = [os.path.join(gdir, "db%d" % i) for i in range(len(files))]
    dbs_r = [os.path.join(rdir, "db%d" % i) for i in range(len(files))]
    t0 = time.time(); st = psd.PeakSegFPOP_file_batch(gfiles, pens, dbs_g); cold = time.time() - t0
    t0 = time.time(); st = psd.PeakSegFPOP_file_batch(gfiles, pens, dbs_g); warm = time.time() - t0
    ref_s = ref_batch(rfiles, pens, dbs_r, threads)
    same = all(outputs(g, p) == outputs(r, p) for g, r, p in zip(gfiles, rfiles, pens))
    return {"problems": len(files), "rows_x_penalties": rows, "gpu_seconds_first_call": round(cold, 4),
            "gpu_seconds": round(warm, 4), "reference_seconds": round(ref_s, 4), "reference_threads": threads,
            "gpu_rows_per_s": rows / warm, "reference_rows_per_s": rows / ref_s,
            "files_identical": bool(same)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-rows", type=int, default=1000000)
    ap.add_argument("--c5-rows", type=int, default=10000)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    skip = set(args.skip.split(","))
    assert os.path.exists(oracle_bind.REF_SO), "needs oracle/_ref/libref_fpop.so (make -C oracle ref)"
    threads = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(prefix="psdcfg", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {"host_threads": threads}
    try:
        mono = os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph")
        if "c1" not in skip:
            out["c1_mono27ac_one_problem"] = both("c1", [mono], ["10.5"], tmp, 1, 6921)
            pens = [r_paste(p) for p in np.exp(np.linspace(np.log(1.0), np.log(1e6), 1000))]
            out["c1x_mono27ac_1000_penalties"] = both("c1x", [mono] * 1000, pens, tmp, threads, 6921 * 1000)
        if "c2f" not in skip:
            # config 2 through FILES: 128 of the count vectors as bedGraph text, 5 penalties each, one batched call
            files, pens, rows = [], [], 0
            for seed in range(128):
                s_, e_, c_ = synth.poisson_problem(seed)
                f = os.path.join(tmp, "v%d.bedGraph" % seed); synth.write_bedgraph(f, s_, e_, c_)
                for pen in synth.C2_PENALTIES:
                    files.append(f); pens.append(r_paste(pen)); rows += len(c_)
            out["c2_files_128_vectors_x_5_penalties"] = both("c2f", files, pens, tmp, threads, rows)
        if "c3" not in skip:
            n_raw = int(args.c3_rows / 0.75)
            s, e, c = synth.poisson_problem(12345, n_raw)
            f = os.path.join(tmp, "c3.bedGraph"); synth.write_bedgraph(f, s, e, c)
            pens = ["1000", "31622.7766016838", "1000000"]        # three steps of a chain, solved together
            out["c3_one_long_problem_3_penalties"] = both("c3", [f] * 3, pens, tmp, min(3, threads), 3 * len(c))
            out["c3_one_long_problem_3_penalties"]["rows"] = int(len(c))
        if "c5" not in skip:
            s, e, c = synth.increasing_problem(args.c5_rows)
            f = os.path.join(tmp, "c5.bedGraph"); synth.write_bedgraph(f, s, e, c)
            out["c5_increasing_counts"] = both("c5", [f, f], ["0", "1000000"], tmp, min(2, threads), 2 * len(c))
            out["c5_increasing_counts"]["rows"] = int(len(c))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
