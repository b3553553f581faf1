"""The product's device source (peaksegdisk_b200/csrc/fpop_warp.cuh) executed under the CPU warp
emulator (tests/emu: 32 fibers = 32 lanes; a TEST TOOL, not a product path) must reproduce the oracle
bit for bit at every row: coefficients, breakpoints and back-pointers of both cost functions.
Lane scheduling order is run both ascending and descending to expose missing warp syncs."""
import os
import subprocess
import numpy as np
import pytest
from helpers import ROOT, golden, parse_rows
import oracle_bind


@pytest.fixture(scope="module")
def ec():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    oracle_bind.ensure_built()
    import emu_compare
    return emu_compare


def test_reference_vectors_row_by_row(ec):
    for case in golden("golden_small.json"):
        if case["status"] != 0 or case["penalty"] == "Inf":
            continue
        s, e, c = parse_rows(case["input"])
        if len(set(c.tolist())) < 2:
            continue   # constant coverage: trivial model, never reaches the kernel
        assert ec.compare(s, e, c, float(case["penalty"]), cap=16, descending=0), (case["name"], case["penalty"])
        assert ec.compare(s, e, c, float(case["penalty"]), cap=16, descending=1, trace=False), case["name"]


@pytest.mark.parametrize("seed,n,pen", [(0, 1500, 0.0), (1, 1500, 4.0), (2, 2000, 100.0), (3, 1200, 1e4), (4, 900, 1e6)])
def test_poisson_row_by_row(ec, seed, n, pen):
    from peaksegdisk_b200 import synth
    s, e, c = synth.poisson_problem(seed, n)
    assert ec.compare(s, e, c, pen, cap=64, descending=seed & 1)


@pytest.mark.parametrize("n,pen", [(100, 0.0), (200, 1e2), (300, 1e4), (300, 1e6)])
def test_many_pieces_chunked_lists(ec, n, pen):
    """increasing counts: > 32 pieces per function, exercising the chunked (multi-pass) loops"""
    from peaksegdisk_b200 import synth
    s, e, c = synth.increasing_problem(n)
    assert ec.compare(s, e, c, pen, cap=512)


def test_mono27ac_prefix(ec):
    from peaksegdisk_b200 import synth
    chrom, s, e, c = synth.read_bedgraph(os.path.join(ROOT, "tests", "golden", "Mono27ac_coverage.bedGraph"))
    assert ec.compare(s[:2500], e[:2500], c[:2500], 10.5, cap=64)
    assert ec.compare(s[:2500], e[:2500], c[:2500], 1952.6, cap=64, descending=1)


def test_piece_overflow_is_reported(ec):
    from peaksegdisk_b200 import synth
    s, e, c = synth.increasing_problem(300)
    orc, emu = ec.load()
    st, _, _, _ = ec.run_emu(emu, s, e, c, 1e4, cap=16)
    assert st == 101


@pytest.mark.parametrize("cap,pen", [(4, 50.0), (6, 1e3), (8, 1e4), (4, 0.0)])
def test_tier_switch_is_transparent(ec, cap, pen):
    """A tiny shared-memory tier forces the warp to move its functions to the global workspace,
    repeat the row there and move back later; every row must still match the oracle."""
    from peaksegdisk_b200 import synth
    s, e, c = synth.poisson_problem(31, 1200)
    info = {}
    assert ec.compare(s, e, c, pen, cap=cap, spill_cap=256, info=info)
    assert info["spills"] >= 1
    assert ec.compare(s, e, c, pen, cap=cap, spill_cap=256, descending=1, trace=False)


def test_spill_tier_overflow_is_reported(ec):
    from peaksegdisk_b200 import synth
    s, e, c = synth.increasing_problem(300)
    orc, emu = ec.load()
    st, _, _, _ = ec.run_emu(emu, s, e, c, 1e4, cap=16, spill_cap=64)
    assert st == 101


@pytest.mark.timeout(600)
def test_degenerate_rows_terminate(ec):
    """The C entry accepts what the reference accepts, including rows R's writeBedGraph would refuse
    (chromEnd <= chromStart, negative counts: costs become inf / NaN).  The reference's output is
    then garbage; the device code must still TERMINATE (a hung kernel takes the GPU with it) and
    report ok / backtrack-lost / internal, never loop.  Same generator as a 54-case sweep that was
    also run with a per-case watchdog."""
    import random
    orc, emu = ec.load()
    rng = random.Random(3)
    seen = set()
    for k in range(24):
        S, E, Cc, pos = [], [], [], rng.choice([0, 7])
        for _ in range(rng.choice([2, 3, 5, 9, 20, 60])):
            w = rng.choice([1, 2, 10, 0, -1, -5]) if rng.random() < 0.3 else rng.choice([1, 2, 10, 300])
            z = rng.choice([0, 1, 5, 40, -3, -100]) if rng.random() < 0.25 else rng.choice([0, 1, 2, 5, 40])
            S.append(pos); E.append(pos + w); Cc.append(z); pos += w
        if len(set(Cc)) < 2:
            continue
        st, summ, seg, _ = ec.run_emu(emu, np.array(S, np.int32), np.array(E, np.int32), np.array(Cc, np.int32),
                                      float(rng.choice([0, 1, 10.5, 1000, 1e6])), cap=48, spill_cap=512)
        assert st in (0, 101, 103, 104), st
        seen.add(st)
    assert 0 in seen


@pytest.fixture()
def lat(ec, monkeypatch):
    """The latency kernel's layout (fpop_lat.cu: -DPSD_G32, a block of two warps per problem, one chain
    per warp, one block barrier per row) under the multi-warp emulator."""
    monkeypatch.setenv("PSD_EMU_LIB", os.path.join(ROOT, "tests", "_build", "libpsd_emu_lat.so"))
    return ec


@pytest.mark.parametrize("seed,n,pen,cap", [(0, 1500, 0.0, 64), (1, 1500, 4.0, 64), (2, 2000, 100.0, 64), (3, 1200, 1e4, 64),
                                            (4, 900, 1e6, 64), (5, 3000, 3e4, 48)])
def test_latency_layout_poisson_row_by_row(lat, seed, n, pen, cap):
    from peaksegdisk_b200 import synth
    s, e, c = synth.poisson_problem(seed, n)
    assert lat.compare(s, e, c, pen, cap=cap, descending=0)
    assert lat.compare(s, e, c, pen, cap=cap, descending=1, trace=False)   # warps and lanes scheduled in the opposite order


def test_latency_layout_reference_vectors_and_many_pieces(lat):
    from peaksegdisk_b200 import synth
    for case in golden("golden_small.json"):
        if case["status"] != 0 or case["penalty"] == "Inf":
            continue
        s, e, c = parse_rows(case["input"])
        if len(set(c.tolist())) < 2:
            continue
        assert lat.compare(s, e, c, float(case["penalty"]), cap=16, descending=k_order(case)), (case["name"], case["penalty"])
    for n, pen in [(100, 0.0), (200, 1e2), (300, 1e4), (300, 1e6)]:   # > 32 pieces: the chunked loops at 32 lanes per chain
        s, e, c = synth.increasing_problem(n)
        assert lat.compare(s, e, c, pen, cap=512)


def k_order(case):
    return len(case["input"]) & 1


@pytest.mark.parametrize("cap,pen", [(4, 50.0), (6, 1e3), (8, 1e4), (4, 0.0)])
def test_latency_layout_tier_switch(lat, cap, pen):
    """both warps must take the tier switch together (flag words are per warp and per row parity)"""
    from peaksegdisk_b200 import synth
    s, e, c = synth.poisson_problem(31, 1200)
    info = {}
    assert lat.compare(s, e, c, pen, cap=cap, spill_cap=256, info=info)
    assert info["spills"] >= 1
    assert lat.compare(s, e, c, pen, cap=cap, spill_cap=256, descending=1, trace=False)
    orc, emu = lat.load()
    st, _, _, _ = lat.run_emu(emu, *synth.increasing_problem(300), 1e4, cap=16, spill_cap=64)
    assert st == 101


@pytest.mark.parametrize("libname", ["libpsd_emu.so", "libpsd_emu_lat.so"])
def test_store_ring_drain_protocol(ec, monkeypatch, libname):
    """The store spill's DMA-drain protocol (StoreRing in fpop_warp.cuh): with 2 "HBM" chunks and a
    ring of 3 slots every further chunk is written into a ring slot, published in the done-queue,
    copied to the "host" region by the (emulated, synchronous) drain and its slot handed back through
    the free-queue; the backtrack then reads the host region.  Both kernels' layouts."""
    from peaksegdisk_b200 import synth
    monkeypatch.setenv("PSD_EMU_LIB", os.path.join(ROOT, "tests", "_build", libname))
    monkeypatch.setenv("PSD_EMU_HBM_CHUNKS", "2")
    monkeypatch.setenv("PSD_EMU_RING_SLOTS", "3")
    for seed, n, pen in [(2, 2000, 100.0), (5, 3000, 3e4)]:
        s, e, c = synth.poisson_problem(seed, n)
        assert ec.compare(s, e, c, pen, cap=64, descending=0)
        assert ec.compare(s, e, c, pen, cap=64, descending=1, trace=False)
    s, e, c = synth.increasing_problem(300)
    assert ec.compare(s, e, c, 1e4, cap=512)
