"""The parity oracle (oracle/fpop_oracle.cpp, our CPU restatement of the reference) is pinned here
against golden vectors generated from the UNMODIFIED reference (tools/make_golden.py ->
tests/golden/*.json) and, when oracle/_ref exists, against the reference binary itself, byte for
byte including the per-row cost-function db."""
import filecmp
import os
import pytest
import numpy as np
import oracle_bind
from helpers import golden, sha, outputs, synth_rows, rows_text

MATH_MODES = [1] + ([0] if oracle_bind.libm_matches_golden() else [])


def _run_oracle(tmp_path, text, pen, mode, name="cov.bedGraph", db=None, missing=False):
    path = str(tmp_path / name)
    if not missing:
        with open(path, "w") as f:
            f.write(text)
    db = db or path + ".db"
    if os.path.isfile(db):
        os.unlink(db)
    st = oracle_bind.oracle_disk(path, pen, db, mode)
    seg, loss = outputs(path, pen)
    size = os.path.getsize(db) if os.path.isfile(db) else None
    return st, seg, loss, size


@pytest.mark.parametrize("mode", MATH_MODES)
def test_small_vectors(tmp_path, mode):
    for k, case in enumerate(golden("golden_small.json")):
        st, seg, loss, size = _run_oracle(tmp_path, case["input"], case["penalty"], mode, name="c%d.bedGraph" % k)
        assert st == case["status"], case["name"]
        assert seg == case["segments"], (case["name"], case["penalty"])
        assert loss == case["loss"], (case["name"], case["penalty"])
        assert size == case["db_bytes"], (case["name"], case["penalty"])


def test_error_codes(tmp_path, capfd):
    dbdir = tmp_path / "dbdir"
    dbdir.mkdir()
    for k, case in enumerate(golden("golden_errors.json")):
        st, seg, loss, _ = _run_oracle(tmp_path, case["input"], case["penalty"], 1, name="e%d.bedGraph" % k,
                                       db=str(dbdir) if case["db"] else None, missing=case["missing"])
        assert st == case["status"], case["name"]
        assert seg == case["segments"] and loss == case["loss"], case["name"]


@pytest.mark.parametrize("mode", MATH_MODES)
def test_mono27ac(tmp_path, mode):
    g = golden("golden_mono27ac.json")
    text = open(os.path.join(os.path.dirname(__file__), "golden", "Mono27ac_coverage.bedGraph")).read()
    for pen, want in g["penalties"].items():
        st, seg, loss, size = _run_oracle(tmp_path, text, pen, mode, name="mono.bedGraph")
        assert st == 0 and loss == want["loss"], pen
        assert sha(seg) == want["segments_sha256"], pen
        assert size == want["db_bytes"], pen


@pytest.mark.parametrize("mode", MATH_MODES)
def test_synthetic(tmp_path, mode):
    for case in golden("golden_synth.json"):
        if case["n_rows"] > 8000:
            continue   # keep the CPU suite short; the big ones are covered by the GPU parity tests
        s, e, c = synth_rows(case["kind"], case["key"])
        st, seg, loss, size = _run_oracle(tmp_path, rows_text(s, e, c), case["penalty"], mode, name="syn.bedGraph")
        assert st == case["status"] and loss == case["loss"], case
        assert sha(seg) == case["segments_sha256"] and size == case["db_bytes"], case


@pytest.mark.skipif(not oracle_bind.ref_available() or not oracle_bind.libm_matches_golden(),
                    reason="needs oracle/_ref and a glibc-2.39-FMA libm")
def test_against_reference_binary_including_db(tmp_path):
    """Live comparison with the compiled reference on fresh seeds: outputs AND the per-row db."""
    from peaksegdisk_b200 import synth
    for seed, n, pen in [(11, 2000, "0"), (12, 3000, "7.5"), (13, 2500, "250"), (14, 1500, "1e5")]:
        s, e, c = synth.poisson_problem(seed, n)
        text = rows_text(s, e, c)
        a, b = str(tmp_path / "a.bedGraph"), str(tmp_path / "b.bedGraph")
        for p in (a, b):
            open(p, "w").write(text)
        assert oracle_bind.ref_disk(a, pen, a + ".db") == 0
        assert oracle_bind.oracle_disk(b, pen, b + ".db", 0) == 0
        for suf in ("_penalty=%s_loss.tsv" % pen, "_penalty=%s_segments.bed" % pen, ".db"):
            assert filecmp.cmp(a + suf, b + suf, shallow=False), (seed, pen, suf)


def test_in_memory_matches_files(tmp_path):
    from peaksegdisk_b200 import synth
    s, e, c = synth.poisson_problem(21, 1500)
    st, summ, seg = oracle_bind.solve_rows(s, e, c, 50.0)
    path = str(tmp_path / "x.bedGraph")
    open(path, "w").write(rows_text(s, e, c))
    assert oracle_bind.oracle_disk(path, "50", path + ".db", 1) == 0
    seg_txt, loss_txt = outputs(path, "50")
    f = loss_txt.split("\t")
    assert st == 0 and int(f[1]) == int(summ[1]) and int(f[2]) == int(summ[2])
    assert float(f[6]) == summ[6]
    lines = seg_txt.splitlines()
    assert [int(l.split("\t")[1]) for l in lines] == seg[0].tolist()
    assert [int(l.split("\t")[2]) for l in lines] == seg[1].tolist()
