#!/usr/bin/env python3
"""Differential fuzz of the drop-in boundary's INPUT handling (no GPU needed): random well- and
ill-formed bedGraph text and penalty strings go through psd_fpop_disk and through the unmodified
reference (oracle/_ref/ref_fpop); the status codes must agree, and so must the output files
whenever the reference stops before its DP (input errors, penalty Inf, constant coverage).
Inputs whose reference status is 0 on the DP branch need a GPU on our side and are skipped here
(the GPU parity tests cover that branch).  usage: tools/fuzz_parse_vs_reference.py [n=400] [seed=0]"""
import os, random, shutil, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from peaksegdisk_b200 import _lib
import oracle_bind
from helpers import outputs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
assert oracle_bind.ref_available(), "needs oracle/_ref/ref_fpop"

SEPS = ["\t", " ", "  ", "\t ", " \t"]
def number(good=True):
    if good:
        return str(rng.choice([0, 1, 2, 5, 7, 13, 100, 99999]))
    return rng.choice(["1.5", "abc", "3x", "+4", "-2", "", "1e3", "0x10", "7.0", "2147483648", "99999999999", "1,5", ".5", "NA", "-0", "  8"])

def make_text():
    rows, pos = [], rng.choice([0, 0, 10, 12345])
    k = rng.choice([1, 1, 2, 3, 5, 8])
    const = rng.random() < 0.35
    cval = number()
    for _ in range(k):
        w = rng.choice([1, 2, 10, 500])
        rows.append(["chr%s" % rng.choice(["1", "X", "Unknown", "1_gl000191_random"]) if rng.random() < 0.2 else "chr1",
                     str(pos), str(pos + w), cval if const else number()])
        pos += w
    kind = rng.random()
    if kind < 0.12 and rows:        # gap or overlap between consecutive rows
        r = rng.randrange(len(rows)); rows[r][1] = str(int(rows[r][1]) + rng.choice([-1, 1, 5]))
    elif kind < 0.24 and rows:      # bad 4th column
        rows[rng.randrange(len(rows))][3] = number(False)
    elif kind < 0.34 and rows:      # too few / too many columns
        r = rng.randrange(len(rows)); rows[r] = rows[r][:rng.choice([1, 2, 3])] if rng.random() < 0.6 else rows[r] + ["extra"]
    elif kind < 0.40 and rows:      # bad coordinate
        r = rng.randrange(len(rows)); rows[r][rng.choice([1, 2])] = number(False)
    lines = [rng.choice(SEPS).join(r) for r in rows]
    eol = rng.choice(["\n", "\n", "\n", "\r\n"])
    text = eol.join(lines) + (eol if rng.random() < 0.85 else "")
    if rng.random() < 0.08:
        text = text.replace(eol, eol + eol, 1)   # blank line
    if rng.random() < 0.05:
        text = ""
    if rng.random() < 0.05:
        text = "track type=bedGraph\n" + text
    return text

PENS = ["Inf", "Inf", "Inf", "inf", "0", "10.5", "1e3", "-1", "-Inf", "NaN", "nan", "abc", "", "1e400", "0x1p3", "  5", "5  ", "1,5", "Infinity", "+Inf", "1e-400"]
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
bad = skipped = compared = aborted = 0
by_status = {}
try:
    for k in range(n):
        text, pen = make_text(), rng.choice(PENS)
        gp, rp = os.path.join(tmp, "g%d.bedGraph" % k), os.path.join(tmp, "r%d.bedGraph" % k)
        for p in (gp, rp):
            with open(p, "w", newline="") as f:
                f.write(text)
        rs = oracle_bind.ref_disk(rp, pen, rp + ".db")
        ref_out = outputs(rp, pen)
        on_dp_branch = rs in (0, 7) and os.path.exists(rp + ".db")
        if rs == 0 and on_dp_branch:
            skipped += 1
            continue
        if rs < 0:      # the reference itself aborts (std::stod throws out_of_range for "1e400"); ours returns status 1
            aborted += 1
            continue
        gs = _lib.lib.psd_fpop_disk(gp.encode(), pen.encode(), (gp + ".db").encode())
        compared += 1
        by_status[rs] = by_status.get(rs, 0) + 1
        if gs != rs or outputs(gp, pen) != ref_out or os.path.exists(gp + ".db") != os.path.exists(rp + ".db"):
            bad += 1
            if bad <= 10:
                print("MISMATCH status ours=%d ref=%d penalty=%r text=%r" % (gs, rs, pen, text[:200]))
                print("   ours:", outputs(gp, pen), "\n   ref: ", ref_out)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
print("%d cases: %d compared (reference statuses %s), %d need the GPU branch (skipped), %d where the reference aborts, %d mismatches" % (
    n, compared, dict(sorted(by_status.items())), skipped, aborted, bad))
sys.exit(1 if bad else 0)
