#!/usr/bin/env python3
"""BASELINE config 4: hg19-shaped coverage, 24 chromosomes x S samples, rows proportional to the
chromosome length (chr1 = 1e7 x row_scale), one penalty per problem -- a FIXED problem list solved on
1, 2, 4, 8 GPUs of the box (strong scaling).  Sharding is by problem, longest-processing-time first
(peaksegdisk_b200/shard.py), one Plan and one host thread per GPU, no exchange between GPUs.

  python tools/run_config4.py --scale 0.05 --samples 64 --devices 1,2,4,8 --out profiles/r02_config4.json
  python tools/run_config4.py --big --store-gb 2            # one 1e7-row chromosome, HBM pool capped so that
                                                            # most of its store goes through the DMA-drain spill;
                                                            # the unmodified reference solves it on a host core meanwhile

Every multi-GPU result is compared bit for bit with the 1-GPU result; the two problems pinned in
tests/golden/golden_fullsize.json (section c4) are compared with the reference's loss line."""
import argparse, hashlib, json, os, subprocess, sys, tempfile, threading, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import peaksegdisk_b200 as psd
from peaksegdisk_b200 import synth, shard
import helpers


def solve_on(devices, probs):
    """LPT-shard probs over `devices` GPUs, solve concurrently.  Returns (wall_s, per-device stats, results by problem)."""
    shards = shard.lpt_assign([len(p[2]) for p in probs], len(devices))
    plans, ids = [], []
    for d, idx in zip(devices, shards):
        plan = psd.Plan(d)
        ids.append([plan.add(*probs[i]) for i in idx])
        plans.append(plan)
    errs = [None] * len(plans)

    def run(k):
        try:
            plans[k].run()
        except Exception as exc:      # surfaced below
            errs[k] = exc
    t0 = time.time()
    threads = [threading.Thread(target=run, args=(k,)) for k in range(len(plans))]
    [t.start() for t in threads]; [t.join() for t in threads]
    wall = time.time() - t0
    if any(errs):
        raise RuntimeError(errs)
    stats = [pl.stats() for pl in plans]
    res = [None] * len(probs)
    for pl, idx, pid in zip(plans, shards, ids):
        for i, j in zip(idx, pid):
            row = pl.loss_row(j)
            seg = pl.segments(j)
            res[i] = (row, hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in seg)).hexdigest())
    for pl in plans:
        pl.close()
    return wall, stats, shards, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--devices", default="1")
    ap.add_argument("--store-gb", type=float, default=0.0)
    ap.add_argument("--host-spill-gb", type=float, default=-1.0)
    ap.add_argument("--big", action="store_true", help="one full-scale chr1 problem (1e7 rows) instead of the batch")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    lib = psd._lib.lib
    if args.store_gb > 0:
        lib.psd_set_option(b"store_gb", args.store_gb)
    lib.psd_set_option(b"host_spill_gb", args.host_spill_gb)
    n_dev = lib.psd_device_count()
    out = {"host_threads": len(os.sched_getaffinity(0)), "gpus_visible": n_dev}
    if args.big:
        t0 = time.time()
        s, e, c, pen = helpers.c4_lite_problem(0, 1, 1.0)          # chr1, sample 0, full scale: 1e7 tiled rows before RLE
        out["generate_s"] = round(time.time() - t0, 1)
        tmp = tempfile.mkdtemp(prefix="psdc4big", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        bg = os.path.join(tmp, "chr1.bedGraph")
        i32p = psd._lib.C.POINTER(psd._lib.C.c_int32)
        lib.psd_write_bedgraph(bg.encode(), b"chr1", len(c), s.ctypes.data_as(i32p), e.ctypes.data_as(i32p), c.ctypes.data_as(i32p))
        pen_s = psd.r_paste(pen)
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_fpop")
        ref = {}

        def run_ref():       # the unmodified reference on one host core, db on tmpfs, while the GPU works
            t = time.time()
            ref["rc"] = subprocess.call([ref_bin, bg, pen_s, bg + ".refdb"], stdout=subprocess.DEVNULL)
            ref["seconds"] = time.time() - t
            ref["db_bytes"] = os.path.getsize(bg + ".refdb") if os.path.exists(bg + ".refdb") else None
            pre = "%s_penalty=%s" % (bg, pen_s)
            ref["loss"] = open(pre + "_loss.tsv").read()
            ref["seg_sha"] = hashlib.sha256(open(pre + "_segments.bed", "rb").read()).hexdigest()
            for f in (bg + ".refdb", pre + "_loss.tsv", pre + "_segments.bed"):
                os.path.exists(f) and os.unlink(f)
        th = threading.Thread(target=run_ref) if os.path.exists(ref_bin) else None
        if th:
            th.start()
        t0 = time.time()
        status = psd.PeakSegFPOP_file_batch([bg], [pen_s])
        wall = time.time() - t0
        bs = psd._lib.last_batch_stats()
        pre = "%s_penalty=%s" % (bg, pen_s)
        loss = open(pre + "_loss.tsv").read()
        seg_sha = hashlib.sha256(open(pre + "_segments.bed", "rb").read()).hexdigest()
        os.unlink(pre + "_loss.tsv"); os.unlink(pre + "_segments.bed")
        if th:
            th.join()
        out["big"] = {"rows": int(len(c)), "penalty": pen_s, "status": status, "gpu_wall_s": round(wall, 2), "dp_ms": bs["dp_ms"],
                      "us_per_row": 1e3 * bs["dp_ms"] / len(c), "stages": bs, "loss": loss, "reference": ref,
                      "identical_to_reference": bool(ref) and ref.get("rc") == 0 and ref["loss"] == loss and ref["seg_sha"] == seg_sha}
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
        print(json.dumps(out, indent=1), flush=True)
        if args.out:
            json.dump(out, open(args.out, "w"), indent=1)
        assert status == [0] and (not ref or out["big"]["identical_to_reference"])
        return
    t0 = time.time()
    n_prob = 24 * args.samples
    helpers._mono_rows()      # fill the cache before the threads start
    with ThreadPoolExecutor(min(32, out["host_threads"])) as ex:
        probs = list(ex.map(lambda k: helpers.c4_lite_problem(k, args.samples, args.scale), range(n_prob)))
    # the reference (and R) see the penalty as its 15-digit string (R/PeakSegFPOP_dir.R:64): solve exactly that value
    probs = [(s, e, c, float(psd.r_paste(pen))) for (s, e, c, pen) in probs]
    rows = sum(len(p[2]) for p in probs)
    out.update(problems=n_prob, rows=int(rows), longest=int(max(len(p[2]) for p in probs)), row_scale=args.scale, samples=args.samples,
               generate_s=round(time.time() - t0, 1))
    print("generated %d problems, %d rows, longest %d, in %.1f s" % (n_prob, rows, out["longest"], out["generate_s"]), flush=True)
    gold = {g["key"][2]: g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_fullsize.json")))["c4"]} if args.scale == 0.05 else {}
    runs, base = [], None
    for k in [int(x) for x in args.devices.split(",")]:
        if k > n_dev:
            print("skipping %d GPUs: only %d visible" % (k, n_dev)); continue
        wall, stats, shards, res = solve_on(list(range(k)), probs)
        if base is None:
            base = res
        same = all(a == b for a, b in zip(base, res))
        dp = [st["dp_ms"] for st in stats]
        rec = {"gpus": k, "wall_s": round(wall, 3), "dp_ms_per_gpu": [round(x, 1) for x in dp], "makespan_dp_ms": round(max(dp), 1),
               "imbalance_max_over_mean": round(max(dp) / (sum(dp) / len(dp)), 3), "rows_per_s_device_time": rows / (max(dp) / 1e3),
               "rows_per_gpu": [int(sum(len(probs[i][2]) for i in idx)) for idx in shards],
               "latency_kernel_waves": [st["n_latency_waves"] for st in stats], "store_gb_written": round(sum(st["store_bytes_written"] for st in stats) / 1e9, 2),
               "spilled_gb": round(sum(st["store_bytes_spilled_host"] for st in stats) / 1e9, 2),
               "drained_dma_gb": round(sum(st["store_bytes_drained_dma"] for st in stats) / 1e9, 2),
               "identical_to_first_run": bool(same)}
        runs.append(rec)
        print(json.dumps(rec), flush=True)
        assert same
    # the two problems the reference solved (tools/make_golden_fullsize.py): (chromosome 1, sample 0), (chromosome 21, sample 0)
    checked = 0
    for gk, g in gold.items():
        idx = (gk // 2) * args.samples
        row = base[idx][0]
        f = g["loss"].rstrip("\n").split("\t")
        assert (row["segments"], row["peaks"], row["bedGraph.lines"]) == (int(f[1]), int(f[2]), int(f[4])) and row["total.loss"] == float(f[6]), (gk, row, f)
        checked += 1
    out["runs"] = runs
    out["reference_checked_problems"] = checked
    if runs:
        t1 = runs[0]["makespan_dp_ms"]
        out["strong_scaling"] = [{"gpus": r["gpus"], "speedup_vs_first": round(t1 / r["makespan_dp_ms"], 3)} for r in runs]
    print(json.dumps(out, indent=1), flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
