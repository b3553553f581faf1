#!/usr/bin/env python3
"""Generate tests/golden/golden_fullsize.json: BASELINE.json's configurations at their STATED sizes,
solved by the UNMODIFIED reference (oracle/_ref/ref_fpop, compiled from /root/reference/src by
`make -C oracle ref`).  Run in the build container (the reference sources do not exist on the GPU box);
the fixture is committed.  Per (input, penalty): the whole _loss.tsv line, sha256 of _segments.bed,
its first/last lines and the size of the reference's db file.

  c2   config 2 at full size: 8 vectors of N = 1e5 positions (seeds 2000..2007) x {1e2..1e6}, plus the
       bench's own vectors (seeds 0..3, the longest (588) and the shortest (1022) of rank 0's batch)
       x {1e2..1e6}: bench.py compares these problems of its timed batch with this fixture
  c3   config 3: seed 2024, 1,333,333 positions (1.0 M bedGraph rows): penalties 0, Inf and the
       whole sequentialSearch_dir chain to 100 peaks (R/sequentialSearch_dir.R:31-98 emulated over the
       reference binary, as tools/make_golden.py does for Mono27ac)
  c3s  the same search on a 1e5-position vector (seed 2025), target 30 peaks: the second pinned chain,
       small enough for the default GPU test run
  c5   config 5 (vignettes/Worst_case.Rmd:19-41): increasing(10000) at 1e2 / 1e4 / 1e6
  c4   config 4 lite (tools/run_config4_lite.py 2 0.05): the smallest and the largest of the 48
       hg19-shaped problems

usage: tools/make_golden_fullsize.py [jobs=8] [only=c2,c3,c3s,c5,c4]
"""
import hashlib, json, os, subprocess, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from peaksegdisk_b200 import synth
from peaksegdisk_b200.api import r_paste
import helpers
REF = os.path.join(ROOT, "oracle", "_ref", "ref_fpop")
OUT = os.path.join(ROOT, "tests", "golden", "golden_fullsize.json")
TMP = tempfile.mkdtemp(prefix="psdgold", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run_ref(path, pen, tag):
    db = "%s.%s.db" % (path, tag)
    t0 = time.time()
    p = subprocess.run([REF, path, pen, db], capture_output=True, text=True)
    secs = time.time() - t0
    pre = "%s_penalty=%s" % (path, pen)
    seg = open(pre + "_segments.bed").read()
    loss = open(pre + "_loss.tsv").read()
    db_bytes = os.path.getsize(db) if os.path.isfile(db) else None
    for f in (db, pre + "_segments.bed", pre + "_loss.tsv"):
        if os.path.isfile(f):
            os.unlink(f)
    lines = seg.splitlines()
    return {"penalty": pen, "status": p.returncode, "loss": loss, "segments_sha256": hashlib.sha256(seg.encode()).hexdigest(),
            "segments_head": lines[:2], "segments_tail": lines[-1:], "db_bytes": db_bytes, "reference_seconds": round(secs, 2)}


def write_rows(name, rows):
    path = os.path.join(TMP, name)
    synth.write_bedgraph(path, *rows)
    return path


def search_chain(path, target, pool_tag):
    """R/sequentialSearch_dir.R:31-98 over the reference binary: returns the list of solves in order."""
    def solve(pen_str, it):
        r = run_ref(path, pen_str, pool_tag)
        f = r["loss"].split("\t")
        return dict(r, penalty_str=pen_str, peaks=int(f[2]), total_loss=float(f[6]), iteration=it)
    with ThreadPoolExecutor(2) as ex:
        fo, fu = ex.submit(solve, "0", 1), ex.submit(solve, "Inf", 1)
        over, under = fo.result(), fu.result()
    chain = [over, under]
    it = 1
    while True:
        if target in (under["peaks"], over["peaks"]):
            break
        nxt = (over["total_loss"] - under["total_loss"]) / (under["peaks"] - over["peaks"])
        if nxt < 0:
            break
        it += 1
        m = solve(r_paste(nxt), it)
        chain.append(m)
        print("  chain", pool_tag, it, m["penalty_str"], m["peaks"], flush=True)
        if m["peaks"] in (under["peaks"], over["peaks"]):
            break
        if m["peaks"] < target:
            under = m
        else:
            over = m
    return chain


def main():
    jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else {"c2", "c3", "c3s", "c5", "c4"}
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    ex = ThreadPoolExecutor(jobs)
    futs = []   # (section, record-without-results, future)

    def submit(section, meta, path, pen):
        futs.append((section, meta, ex.submit(run_ref, path, pen, "%s%d" % (section, len(futs)))))

    chains = {}
    if "c3" in only:
        rows = synth.poisson_problem(2024, 1333333)
        p3 = write_rows("c3.bedGraph", rows)
        chains["c3"] = (ex.submit(search_chain, p3, 100, "c3"), {"seed": 2024, "positions": 1333333, "n_rows": len(rows[2]), "target_peaks": 100})
    if "c3s" in only:
        rows = synth.poisson_problem(2025, 100000)
        p3s = write_rows("c3s.bedGraph", rows)
        chains["c3s"] = (ex.submit(search_chain, p3s, 30, "c3s"), {"seed": 2025, "positions": 100000, "n_rows": len(rows[2]), "target_peaks": 30})
    if "c5" in only:
        rows = synth.increasing_problem(10000)
        p5 = write_rows("c5.bedGraph", rows)
        for pen in ("1e+06", "10000", "100"):
            submit("c5", {"kind": "increasing", "key": [10000], "n_rows": 10000}, p5, pen)
    if "c4" in only:
        probs = helpers.c4_lite_problems(2, 0.05)
        sizes = [len(p[2]) for p in probs]
        for k in (sizes.index(max(sizes)), sizes.index(min(sizes))):
            s, e, c, pen = probs[k]
            path = write_rows("c4_%d.bedGraph" % k, (s, e, c))
            submit("c4", {"kind": "c4lite", "key": [2, 0.05, k], "n_rows": len(c), "penalty_value": pen}, path, r_paste(pen))
    if "c2" in only:
        for seed, n in [(s, 100000) for s in range(2000, 2008)] + [(s, None) for s in (588, 0, 1, 2, 3, 1022)]:
            rows = synth.poisson_problem(seed, n)
            path = write_rows("c2_%d.bedGraph" % seed, rows)
            for pen in synth.C2_PENALTIES:
                submit("c2", {"kind": "poisson", "key": [seed, n], "n_rows": len(rows[2])}, path, r_paste(pen))
    for section in only:
        if section not in chains:
            gold[section] = []
    for section, meta, f in futs:
        r = f.result()
        assert r["status"] == 0, (section, meta, r)
        gold[section].append(dict(meta, **r))
        print(section, meta.get("key"), r["penalty"], r["loss"].split("\t")[1:3], r["reference_seconds"], "s", flush=True)
    for name, (f, meta) in chains.items():
        gold[name] = dict(meta, chain=f.result())
    json.dump(gold, open(OUT, "w"), indent=1)
    print("wrote", OUT, {k: (len(v["chain"]) if isinstance(v, dict) else len(v)) for k, v in gold.items()})


if __name__ == "__main__":
    main()
