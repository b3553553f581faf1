"""Host-side logic that needs no GPU: the C ABI exports, the bedGraph/penalty front end's status codes
and created files, the trivial (one-segment) branch which the reference also solves in closed form
on the host, and R-compatible number formatting."""
import ctypes as C
import os
import re
import pytest
from helpers import ROOT, golden, outputs


def test_library_exports_every_declared_symbol():
    from peaksegdisk_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "peaksegdisk_b200.h")).read()
    declared = set(re.findall(r"\b(psd_[a-z_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared <= set(_lib.C_ABI_SYMBOLS), declared - set(_lib.C_ABI_SYMBOLS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in _lib.C_ABI_SYMBOLS:
        assert hasattr(lib, name), name


def test_status_messages_match_interface_cpp():
    from peaksegdisk_b200 import _lib
    assert _lib.status_text(2, penalty="-1") == "penalty=-1 must be non-negative"
    assert _lib.status_text(1, penalty="NaN") == "penalty=NaN but must be finite"
    assert _lib.status_text(3, bedgraph="f") == "unable to open input file for reading f"
    assert _lib.status_text(4, bedgraph="f") == "each line of input data file f should have exactly four columns"
    assert _lib.status_text(5, bedgraph="f") == "fourth column of input data file f should be integer"
    assert _lib.status_text(6, bedgraph="f") == "there should be no gaps (columns 2-3) in input data file f"
    assert _lib.status_text(7, db="d") == "unable to write to cost function database file d"
    assert _lib.status_text(9, bedgraph="f") == "input file f contains no data"
    assert _lib.status_text(10, penalty="foo") == "penalty string 'foo' is not numeric; it should be convertible to double"


def test_input_errors_and_trivial_branch_without_gpu(tmp_path, capfd):
    """Every golden case that ends before the DP (input errors, Inf penalty, constant coverage)."""
    from peaksegdisk_b200 import _lib
    dbdir = tmp_path / "dbdir"
    dbdir.mkdir()
    n = 0
    for k, case in enumerate(golden("golden_errors.json")):
        if case["status"] in (0, 7) and case["penalty"] != "Inf":
            continue   # reaches the DP: GPU test
        path = str(tmp_path / ("e%d.bedGraph" % k))
        if not case["missing"]:
            open(path, "w").write(case["input"])
        db = str(dbdir) if case["db"] else path + ".db"
        st = _lib.lib.psd_fpop_disk(path.encode(), case["penalty"].encode(), db.encode())
        assert st == case["status"], case["name"]
        assert outputs(path, case["penalty"]) == (case["segments"], case["loss"]), case["name"]
        assert not os.path.isfile(path + ".db")
        n += 1
    for k, case in enumerate(golden("golden_small.json")):
        trivial = case["penalty"] == "Inf" or len(set(l.split()[3] for l in case["input"].splitlines())) == 1
        if not trivial:
            continue
        path = str(tmp_path / ("t%d.bedGraph" % k))
        open(path, "w").write(case["input"])
        st = _lib.lib.psd_fpop_disk(path.encode(), case["penalty"].encode(), (path + ".db").encode())
        assert st == 0, case["name"]
        assert outputs(path, case["penalty"]) == (case["segments"], case["loss"]), (case["name"], case["penalty"])
        assert not os.path.exists(path + ".db"), "the trivial branch must not touch the db"
        n += 1
    assert n > 25


def test_count_vector_front_end_trivial_models_without_gpu():
    """psd_plan_add_counts (R/PeakSegFPOP_vec.R:18-25): the one-segment models (penalty Inf, constant
    vector) are closed-form on the host and equal the oracle on the rle() rows; argument errors."""
    import numpy as np
    import oracle_bind
    import peaksegdisk_b200 as psd
    from peaksegdisk_b200 import synth
    rng = np.random.default_rng(5)
    cases = [(rng.poisson(3.0, size=1000), float("inf")), (np.full(77, 4), 10.0), (np.zeros(5, int), 0.0),
             (np.array([1, 3, 0, 4, 2]), float("inf")), (np.array([9]), 1.0)]
    plan = psd.Plan()
    ids = [plan.add_counts(v, pen) for v, pen in cases]
    plan.run()                       # nothing to send to a device
    for pid, (v, pen) in zip(ids, cases):
        s, e, c = synth.rle_rows(v)
        st, summ, oseg = oracle_bind.solve_rows(s, e, c, pen)
        got = plan.loss_row(pid)
        assert st == 0 and (got["segments"], got["peaks"], got["bases"], got["bedGraph.lines"]) == (1, 0, len(v), len(c))
        assert got["total.loss"] == summ[6] and got["mean.pen.cost"] == summ[5]
        seg = plan.segments(pid)
        assert (int(seg[0][0]), int(seg[1][0]), int(seg[2][0])) == (0, len(v), 0) and seg[3][0] == oseg[3][0]
    assert plan.stats()["n_rle_launches"] == 0
    with pytest.raises(ValueError):
        plan.add_counts(np.array([1, -1]), 1.0)
    with pytest.raises(ValueError):
        plan.add_counts(np.array([1, 2]), -1.0)
    with pytest.raises(ValueError):
        plan.add_counts(np.array([0.5, 2.0]), 1.0)
    with pytest.raises(ValueError):
        psd.PeakSegFPOP_vec_batch([np.array([1, 2])], [float("nan")])
    with pytest.raises(ValueError):
        plan.add_counts(np.array([1, 2 ** 40]), 1.0)                 # would wrap in int32
    with pytest.raises(ValueError):
        plan.add(np.array([0, 1]), np.array([1, 2 ** 33]), np.array([1, 2]), 1.0)


def test_input_handling_fuzz_against_the_compiled_reference():
    """Random well- and ill-formed bedGraph text / penalty strings: same status codes and, on every
    branch that ends before the DP, byte-identical files (including the sign of NaN in degenerate
    loss lines) as the unmodified reference binary.  tools/fuzz_parse_vs_reference.py, 300 cases."""
    import subprocess, sys
    import oracle_bind
    if not oracle_bind.ref_available():
        pytest.skip("oracle/_ref/ref_fpop not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parse_vs_reference.py"), "300", "11"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert " 0 mismatches" in out.stdout


def test_not_enough_columns_message(tmp_path, capfd):
    from peaksegdisk_b200 import _lib
    path = str(tmp_path / "x.bedGraph")
    open(path, "w").write("chr1 0 1 5\n\nchr1 1 3 3\n")
    assert _lib.lib.psd_fpop_disk(path.encode(), b"1", (path + ".db").encode()) == 4


def test_r_paste():
    from peaksegdisk_b200 import r_paste
    cases = {0.0: "0", 10.5: "10.5", 1e6: "1e+06", 1e5: "1e+05", 10000.0: "10000", 100000.0: "1e+05", 123456.0: "123456",
             1715.8495636069199918: "1715.84956360692", 157.99473732931699: "157.994737329317",
             866939314852865280.0: "866939314852865280", 0.1: "0.1", 1e-4: "1e-04", 0.00012345: "0.00012345",
             float("inf"): "Inf", 1952.6: "1952.6", 3.0: "3", 1e15: "1e+15", 123456789012345680.0: "123456789012345680", 2.0**53: "9007199254740992", 1e22: "1e+22"}
    for x, want in cases.items():
        assert r_paste(x) == want, (x, r_paste(x), want)


def test_api_argument_errors(tmp_path):
    import peaksegdisk_b200 as psd
    with pytest.raises(ValueError, match="must be the name of a data file to segment"):
        psd.PeakSegFPOP_file(str(tmp_path / "missing.bedGraph"), "1")
    f = tmp_path / "a.bedGraph"
    f.write_text("chr1 0 1 5\nchr1 1 3 3\n")
    with pytest.raises(ValueError, match="pen.str must be a character string"):
        psd.PeakSegFPOP_file(str(f), 10.5)
    with pytest.raises(ValueError, match="must be a non-negative numeric scalar"):
        psd.PeakSegFPOP_file(str(f), "-1")
    with pytest.raises(ValueError, match="must be the name of a directory"):
        psd.PeakSegFPOP_dir(str(tmp_path / "nodir"), "1")
    with pytest.raises(ValueError, match="pen.num must be non-negative numeric scalar"):
        psd.PeakSegFPOP_vec([1, 2, 3], -1)
    with pytest.raises(ValueError, match="count.vec must be integer"):
        psd.PeakSegFPOP_vec([1.5, 2.0], 1)
    # Inf penalty: closed-form model, identical to the reference's files (test-CRAN-PeakSegFPOP_vec.R:9-15)
    fit = psd.PeakSegFPOP_vec([1, 3, 0, 4, 2], float("inf"))
    assert len(fit["segments"]) == 1 and int(fit["loss"]["peaks"][0]) == 0


def test_reference_style_caller_links_against_the_library(tmp_path):
    """A caller written like the reference's interface.cpp (C++-linkage declaration of
    PeakSegFPOP_disk, src/PeakSegFPOPLog.h:15) links against the library and runs.  Without a GPU
    only the closed-form branch (penalty Inf) can be exercised here."""
    import subprocess
    from peaksegdisk_b200 import _lib
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-O1", "-o", exe, os.path.join(ROOT, "tests", "native", "dropin_link.cpp"),
                           "-L" + libdir, "-lpeaksegdisk_b200", "-Wl,-rpath," + libdir])
    bg = str(tmp_path / "two.bedGraph")
    open(bg, "w").write("chr1 0 1 5\nchr1 1 3 3")
    out = subprocess.run([exe, bg, "Inf", bg + ".db"], capture_output=True, text=True)
    assert out.returncode == 0 and "status=0" in out.stdout
    case = [c for c in golden("golden_small.json") if c["name"] == "two-rows-no-newline" and c["penalty"] == "Inf"][0]
    assert outputs(bg, "Inf") == (case["segments"], case["loss"])
    out = subprocess.run([exe, bg, "-1", bg + ".db"], capture_output=True, text=True)
    assert out.returncode == 2


@pytest.mark.parametrize("fixture,key,target", [("golden_fullsize.json", "c3", 100), ("golden_fullsize.json", "c3s", 30),
                                                ("golden_mono27ac.json", "search19", 19)])
def test_search_state_machine_follows_the_reference_chain(tmp_path, fixture, key, target):
    """R/sequentialSearch_dir.R:31-98 on the host: fed with the reference's own _loss.tsv lines (the
    1.0 M-row config-3 problem, the 75,892-row one, Mono27ac) the search must ask for exactly the
    reference's next penalty string at every step -- the %.20g fields are parsed with correct
    rounding, the quotient is formatted like R's paste()."""
    import json
    from peaksegdisk_b200 import api
    g = json.load(open(os.path.join(ROOT, "tests", "golden", fixture)))[key]
    chain = g["chain"] if isinstance(g, dict) else g
    by_pen = {c["penalty_str"]: c for c in chain}
    st = api._Search("unused", target)
    asked = []
    while st.next_pen:
        fits = []
        for _, pen in st.requests():
            asked.append(pen)
            assert pen in by_pen, "the search asked for %s, which the reference never solved" % pen
            f = tmp_path / "loss.tsv"
            f.write_text(by_pen[pen]["loss"])
            fits.append({"loss": api._read_loss(str(f))})
        st.update(fits)
    assert sorted(asked) == sorted(by_pen)
    assert int(st.candidate["peaks"]) == chain[-1]["peaks"] and api.r_paste(float(st.candidate["penalty"])) == chain[-1]["penalty_str"]


IFACE_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_interface_b200")


@pytest.mark.skipif(not os.path.exists(IFACE_BIN), reason="oracle/_ref/ref_interface_b200 not built (make -C oracle interface; needs /root/reference)")
def test_reference_interface_cpp_linked_against_the_library(tmp_path):
    """The reference's OWN src/interface.cpp (compiled unmodified by oracle/Makefile with an R-header
    shim) linked against libpeaksegdisk_b200.so: R_init_PeakSegDisk registers the .C routine, the
    driver calls PeakSegFPOP_interface(char**, char**, char**) as R's .C() does, and every input error
    surfaces as the Rf_error text tests/testthat/test-CRAN-cpp-errors.R expects.  (Branches that end
    before the DP need no GPU; the DP branch is covered by the -m gpu test of the same binary.)"""
    import subprocess
    from peaksegdisk_b200 import _lib
    dbdir = tmp_path / "dbdir"
    dbdir.mkdir()
    n = 0
    for k, case in enumerate(golden("golden_errors.json")):
        if case["status"] == 0 and case["penalty"] != "Inf":
            continue          # DP branch: needs the GPU
        path = str(tmp_path / ("e%d.bedGraph" % k))
        if not case["missing"]:
            open(path, "w").write(case["input"])
        db = str(dbdir) if case["db"] else path + ".db"
        out = subprocess.run([IFACE_BIN, path, case["penalty"], db], capture_output=True, text=True)
        if case["status"] == 0:
            assert out.returncode == 0 and out.stdout.endswith("ok\n"), case["name"]
        else:
            assert out.returncode == 1, (case["name"], out.stderr)
            assert out.stderr == "Error: " + _lib.status_text(case["status"], path, case["penalty"], db) + "\n", case["name"]
        assert outputs(path, case["penalty"]) == (case["segments"], case["loss"]), case["name"]
        n += 1
    assert n >= 15


def test_write_bedgraph_and_in_memory_validation(tmp_path):
    """Host-only entry points: psd_write_bedgraph writes the text R/writeBedGraph.R:35-37 produces;
    psd_plan_add enforces the file path's contract on in-memory rows (contiguous: status 6; positive
    widths and non-negative coverage: PSD_ERR_ARG) -- no GPU is touched."""
    import numpy as np
    from peaksegdisk_b200 import _lib, synth, Plan
    i32p = C.POINTER(C.c_int32)
    s, e, c = synth.poisson_problem(3, 5000)
    c = c.copy(); c[7] = -12                      # the writer prints what it is given, sign included
    a, b = str(tmp_path / "a.bedGraph"), str(tmp_path / "b.bedGraph")
    synth.write_bedgraph(a, s, e, c, chrom="chr7_x")
    assert _lib.lib.psd_write_bedgraph(b.encode(), b"chr7_x", len(c), s.ctypes.data_as(i32p), e.ctypes.data_as(i32p), c.ctypes.data_as(i32p)) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    assert _lib.lib.psd_write_bedgraph(str(tmp_path / "no" / "dir.bedGraph").encode(), b"c", 1, s.ctypes.data_as(i32p), e.ctypes.data_as(i32p), c.ctypes.data_as(i32p)) == 111
    plan = Plan(0)
    ok = np.array([0, 5, 9], np.int32), np.array([5, 9, 20], np.int32), np.array([1, 7, 0], np.int32)
    assert plan.add(*ok, 1.0) == 0
    for rows in [(np.array([0, 6, 9], np.int32), ok[1], ok[2]),            # gap
                 (np.array([0, 5, 5], np.int32), np.array([5, 5, 20], np.int32), ok[2]),   # zero width
                 (ok[0], ok[1], np.array([1, -7, 0], np.int32))]:          # negative coverage
        with pytest.raises(ValueError):
            plan.add(*rows, 1.0)
    assert len(plan) == 1
    st = _lib.last_batch_stats()
    assert set(st) >= {"parse_ms", "run_ms", "dp_ms", "n_problems"}
