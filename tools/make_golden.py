#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference solver (oracle/_ref/ref_fpop, built by
`make -C oracle ref` from /root/reference/src).  Run in the build container; the fixtures are committed
because /root/reference does not exist on the GPU box.

Fixtures:
  golden_small.json   the reference's own test vectors (tests/testthat/*.R):
                      full segments.bed / loss.tsv text and status per (input, penalty)
  golden_errors.json  every input-error status code (tests/testthat/test-CRAN-cpp-errors.R)
  golden_mono27ac.json  Mono27ac chr11 coverage at 5 penalties + the sequentialSearch chain (19 peaks)
  golden_synth.json   seeded synthetic problems (peaksegdisk_b200/synth.py): loss line + sha256(segments)
"""
import hashlib, json, math, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peaksegdisk_b200 import synth
from peaksegdisk_b200.api import r_paste
REF = os.path.join(ROOT, "oracle", "_ref", "ref_fpop")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_ref(text, pen, workdir, name="cov.bedGraph", db=None):
    path = os.path.join(workdir, name)
    if text is not None:
        with open(path, "w") as f:
            f.write(text)
    for suf in ("_segments.bed", "_loss.tsv"):
        try: os.unlink("%s_penalty=%s%s" % (path, pen, suf))
        except OSError: pass
    db = db or path + ".db"
    p = subprocess.run([REF, path, pen, db], capture_output=True, text=True)
    out = {"status": p.returncode, "stdout": p.stdout}
    for key, suf in (("segments", "_segments.bed"), ("loss", "_loss.tsv")):
        fn = "%s_penalty=%s%s" % (path, pen, suf)
        out[key] = open(fn).read() if os.path.exists(fn) else None
    out["db_bytes"] = os.path.getsize(db) if os.path.isfile(db) else None
    if os.path.isfile(db): os.unlink(db)
    return out


def rows_text(rows, sep="\t", chrom="chr1"):
    return "".join(sep.join([chrom] + [str(v) for v in r]) + "\n" for r in rows)


def main():
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp()
    small = []
    def add(name, text, pens, source):
        for pen in pens:
            r = run_ref(text, pen, tmp)
            small.append({"name": name, "source": source, "input": text, "penalty": pen, **r})
    four = rows_text([(0, 10, 2), (10, 20, 10), (20, 30, 14), (30, 40, 13)])
    add("four", four, ["10.5", "0", "Inf", "1e6", "3"], "tests/testthat/test-CRAN-PeakSegFPOP_file.R:19-23")
    hap3 = ("chr6_dbb_hap3 3491790 3491834 2\nchr6_dbb_hap3 3491834 3491836 1\nchr6_dbb_hap3 3491836 3697362 0\n"
            "chr6_dbb_hap3 3697362 3697408 1\nchr6_dbb_hap3 3697408 3701587 0\nchr6_dbb_hap3 3701587 3701633 1\n"
            "chr6_dbb_hap3 3701633 3736386 0\n")
    add("hap3", hap3, ["8.66939314852865e+17", "866939314852865280", "10", "5", "300", "0", "Inf"],
        "tests/testthat/test-CRAN-PeakSegFPOP_dir.R:6-14")
    add("zeros", rows_text([(1, 2, 0), (2, 3, 0), (3, 4, 0)]), ["0", "5", "Inf"], "test-CRAN-PeakSegFPOP_dir.R:87-103")
    add("fives", rows_text([(1, 2, 5), (2, 3, 5), (3, 4, 5)]), ["0", "5", "Inf"], "test-CRAN-PeakSegFPOP_dir.R:113-129")
    add("zero-zero-five", rows_text([(1, 2, 0), (2, 3, 0), (3, 4, 5)]), ["0", "10000", "Inf", "0.5"], "test-CRAN-PeakSegFPOP_dir.R:139-160")
    vec = [1, 3, 0, 4, 2]
    add("vec5", rows_text([(i, i + 1, v) for i, v in enumerate(vec)], chrom="chrUnknown"), ["0", "Inf", "1", "0.1"],
        "test-CRAN-PeakSegFPOP_vec.R:5-15")
    supp = [3, 9, 18, 15, 20, 2]
    add("supp", rows_text([(i, i + 1, v) for i, v in enumerate(supp)]), ["0", "Inf", "2.5"], "test-CRAN-sequentialSearch.R:8-15")
    add("two-rows-no-newline", "chr1 0 1 5\nchr1 1 3 3", ["0.1", "Inf", "0"], "test-CRAN-cpp-errors.R:7")
    add("one-row", "chr1 0 1 5\n", ["300", "0"], "test-CRAN-PeakSegFPOP_dir.R:66-70")
    add("crlf", "chr1\t0\t10\t2\r\nchr1\t10\t20\t10\r\nchr1\t20\t30\t14\r\n", ["1"], "format edge: CRLF line ends")
    add("spaces", "  chr1   0   10   2  \n chr1 10 20 10\nchr1 20 30 14\n", ["1"], "format edge: extra blanks")
    add("negative-count", rows_text([(0, 1, 3), (1, 2, -2), (2, 3, 5), (3, 4, 1)]), ["0", "1"], "format edge: negative coverage")
    add("plus-sign", "chr1 +0 +10 +2\nchr1 10 20 +7\n", ["0"], "format edge: explicit plus signs")
    json.dump(small, open(os.path.join(GOLD, "golden_small.json"), "w"), indent=1)

    errors = []
    def adderr(name, text, pen, source, db=None, missing=False):
        r = run_ref(None if missing else text, pen, tmp, name="err_%s.bedGraph" % name, db=db)
        errors.append({"name": name, "source": source, "input": text, "penalty": pen, "missing": missing,
                       "db": "dir" if db else None, **r})
    ok2 = "chr1 0 1 5\nchr1 1 3 3"
    src = "tests/testthat/test-CRAN-cpp-errors.R"
    adderr("pen-not-numeric", ok2, "foo", src + ":28-37")
    adderr("pen-nan", ok2, "NaN", src + ":39-48")
    adderr("pen-negative", ok2, "-1", src + ":50-59")
    adderr("pen-inf-lower", ok2, "inf", "SURVEY 8b: only the exact string Inf selects the no-peaks branch")
    adderr("pen-infinity", ok2, "Infinity", "SURVEY 8b")
    adderr("pen-prefix", ok2, "10abc", "SURVEY 8b: stod accepts a numeric prefix")
    adderr("pen-neg-zero", ok2, "-0", "SURVEY 8b")
    adderr("pen-hex", ok2, "0x10", "SURVEY 8b")
    adderr("pen-space", ok2, " 1.5", "SURVEY 8b: leading whitespace")
    adderr("pen-empty", ok2, "", "stod on empty string")
    adderr("missing-file", None, "1", src + ":61-70", missing=True)
    adderr("three-columns", "chr1 0 1\nchr1 1 3 3\n", "1", src + ":72-83")
    adderr("non-integer", "chr1 0 1 5.5\nchr1 1 3 3\n", "1", src + ":85-96")
    adderr("gap", "chr1 0 1 5\nchr1 2 3 3\n", "1", src + ":98-109")
    adderr("empty", "", "1", src + ":111-122")
    adderr("blank-line", "chr1 0 1 5\n\nchr1 1 3 3\n", "1", "empty line inside the file")
    adderr("text-count", "chr1 0 1 x\n", "1", "non-numeric fourth column")
    adderr("five-columns", "chr1 0 1 5 extra\n", "1", "trailing column")
    adderr("reversed", "chr1 10 5 1\nchr1 0 10 2\n", "1", "test-CRAN-PeakSegFPOP_dir.R:180-184")
    adderr("error-order-pen-first", "chr1 0 1\n", "-3", "penalty errors precede file errors (:152-159)")
    adderr("error-second-line", "chr1 0 1 5\nchr1 1 2\n", "Inf", "errors are reported even on the Inf branch")
    dbdir = os.path.join(tmp, "dbdir"); os.makedirs(dbdir, exist_ok=True)
    adderr("db-is-directory", ok2 + "\nchr1 3 4 9\n", "0.1", src + ":156-170", db=dbdir)
    adderr("db-is-directory-inf", ok2, "Inf", "the trivial branch never touches the db", db=dbdir)
    json.dump(errors, open(os.path.join(GOLD, "golden_errors.json"), "w"), indent=1)

    mono_path = os.path.join(GOLD, "Mono27ac_coverage.bedGraph")
    mono_txt = open(mono_path).read()
    mono = {"penalties": {}, "search19": []}
    for pen in ["0", "10.5", "1952.6", "1e6", "Inf", "0.1", "100"]:
        r = run_ref(mono_txt, pen, tmp, name="mono.bedGraph")
        seg = r["segments"]
        mono["penalties"][pen] = {"status": r["status"], "loss": r["loss"], "segments_sha256": hashlib.sha256(seg.encode()).hexdigest(),
                                  "segments_head": seg.splitlines()[:3], "segments_tail": seg.splitlines()[-1:], "db_bytes": r["db_bytes"]}
    # sequential search, target 19 peaks (test-TRAVIS-sequentialSearch.R:25-29), emulating R/sequentialSearch_dir.R
    def solve(pen_str):
        r = run_ref(mono_txt, pen_str, tmp, name="mono.bedGraph")
        f = r["loss"].split("\t")
        return {"penalty_str": pen_str, "peaks": int(f[2]), "total_loss": float(f[6]), "loss": r["loss"]}
    target = 19
    under, over = solve("Inf"), solve("0")
    chain = [dict(over, iteration=1), dict(under, iteration=1)]
    it = 1
    while True:
        if target in (under["peaks"], over["peaks"]): break
        nxt = (over["total_loss"] - under["total_loss"]) / (under["peaks"] - over["peaks"])
        if nxt < 0: break
        it += 1
        m = solve(r_paste(nxt))
        chain.append(dict(m, iteration=it))
        if m["peaks"] in (under["peaks"], over["peaks"]): break
        if m["peaks"] < target: under = m
        else: over = m
    mono["search19"] = chain
    json.dump(mono, open(os.path.join(GOLD, "golden_mono27ac.json"), "w"), indent=1)

    syn = []
    def addsyn(kind, key, rows, pens):
        s, e, c = rows
        txt = "".join("chrUnknown\t%d\t%d\t%d\n" % t for t in zip(s.tolist(), e.tolist(), c.tolist()))
        for pen in pens:
            r = run_ref(txt, pen, tmp, name="syn.bedGraph")
            syn.append({"kind": kind, "key": key, "penalty": pen, "status": r["status"], "loss": r["loss"], "n_rows": len(c),
                        "segments_sha256": hashlib.sha256(r["segments"].encode()).hexdigest(), "db_bytes": r["db_bytes"]})
    for seed, n in [(0, 3000), (1, 10000), (2, 20000), (3, 1000), (4, 1500), (5, 800), (6, 2500), (7, 5000)]:
        addsyn("poisson", [seed, n], synth.poisson_problem(seed, n), ["0", "100", "1000", "10000", "1e+05", "1e+06"])
    for n in [10, 100, 1000, 3000]:
        addsyn("increasing", [n], synth.increasing_problem(n), ["0", "100", "10000", "1e+06"])
    json.dump(syn, open(os.path.join(GOLD, "golden_synth.json"), "w"), indent=1)
    print("wrote", len(small), "small,", len(errors), "error,", len(mono["penalties"]), "mono,", len(chain), "chain,", len(syn), "synthetic cases")
    for c in chain: print(c["iteration"], c["penalty_str"], c["peaks"])


if __name__ == "__main__":
    main()
