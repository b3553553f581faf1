import hashlib
import json
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return json.load(open(os.path.join(GOLD, name)))


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


def read_or_none(path):
    return open(path).read() if os.path.exists(path) else None


def outputs(bedgraph, pen):
    return (read_or_none("%s_penalty=%s_segments.bed" % (bedgraph, pen)),
            read_or_none("%s_penalty=%s_loss.tsv" % (bedgraph, pen)))


def synth_rows(kind, key):
    from peaksegdisk_b200 import synth
    if kind == "poisson":
        return synth.poisson_problem(key[0], key[1])
    return synth.increasing_problem(key[0])


def rows_text(s, e, c, chrom="chrUnknown"):
    return "".join("%s\t%d\t%d\t%d\n" % (chrom, a, b, d) for a, b, d in zip(s.tolist(), e.tolist(), c.tolist()))


def parse_rows(text):
    """rows of a well-formed bedGraph text"""
    s, e, c = [], [], []
    for line in text.splitlines():
        f = line.split()
        if len(f) >= 4:
            s.append(int(f[1])); e.append(int(f[2])); c.append(int(f[3]))
    return np.array(s, np.int32), np.array(e, np.int32), np.array(c, np.int32)


def loss_fields(line):
    f = line.rstrip("\n").split("\t")
    return {"penalty": f[0], "segments": int(f[1]), "peaks": int(f[2]), "bases": int(f[3]), "lines": int(f[4]),
            "mean_pen_cost": float(f[5]), "total_loss": float(f[6]), "equality": int(f[7]),
            "mean_intervals": float(f[8]), "max_intervals": float(f[9])}


_MONO_ROWS = None


def _mono_rows():
    global _MONO_ROWS
    if _MONO_ROWS is None:
        from peaksegdisk_b200 import synth
        _, s0, e0, c0 = synth.read_bedgraph(os.path.join(GOLD, "Mono27ac_coverage.bedGraph"))
        _MONO_ROWS = (np.minimum((e0 - s0).astype(np.int64), 2000), c0.astype(np.float64))
    return _MONO_ROWS


def c4_lite_problem(k, samples=2, scale=0.05):
    """Problem k (= chromosome * samples + sample) of BASELINE config 4 at reduced scale (SURVEY.md 8d,
    C4): rows proportional to the hg19 chromosome length (chr1 = 1e7 * scale), Mono27ac-like weights
    (clipped to 2,000 bases so tiled coordinates stay below 2^31), its own penalty.
    Returns (chromStart, chromEnd, coverage, penalty)."""
    from peaksegdisk_b200 import synth
    w0, c0 = _mono_rows()
    return synth.hg19_problem(k // samples, k % samples, w0, c0, scale_rows=scale)


def c4_lite_problems(samples=2, scale=0.05):
    """All 24 x `samples` problems of config 4 lite, in (chromosome, sample) order."""
    return [c4_lite_problem(k, samples, scale) for k in range(24 * samples)]


def parse_reference_db(path, n_rows):
    """The reference's scratch db (src/PeakSegFPOPLog.cpp:12-34, 76-141): 2N std::streampos entries
    (16 bytes: offset + mbstate; up functions at [0,N), down at [N,2N); 0 = never written), then the
    appended records  int32 size | int32 n_pieces | int32 chromEnd | n x {f64 max_log_mean, i32 data_i,
    f64 prev_log_mean}.  Returns {(row, which): (chromEnd, max_log_mean[], data_i[], prev_log_mean[])}."""
    raw = open(path, "rb").read()
    pos = np.frombuffer(raw, dtype="<i8", count=4 * n_rows).reshape(2 * n_rows, 2)[:, 0]
    rec = np.dtype([("hi", "<f8"), ("bi", "<i4"), ("bx", "<f8")], align=False)
    assert rec.itemsize == 20
    out = {}
    for el in range(2 * n_rows):
        p = int(pos[el])
        if p == 0:
            continue
        size, n_pieces, chrom_end = np.frombuffer(raw, dtype="<i4", count=3, offset=p)
        assert size == 8 + 20 * n_pieces
        a = np.frombuffer(raw, dtype=rec, count=int(n_pieces), offset=p + 12)
        out[(el % n_rows, el // n_rows)] = (int(chrom_end), a["hi"].copy(), a["bi"].copy(), a["bx"].copy())
    return out
