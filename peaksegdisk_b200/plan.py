"""Batched in-memory solves: many (rows, penalty) problems in one launch, one warp per problem."""
import ctypes as C
import math
import numpy as np
from . import _lib

LOSS_COLUMNS = ["penalty", "segments", "peaks", "bases", "bedGraph.lines", "mean.pen.cost", "total.loss",
                "equality.constraints", "mean.intervals", "max.intervals"]   # R/col.name.list.R:12-15
SEGMENT_COLUMNS = ["chrom", "chromStart", "chromEnd", "status", "mean"]     # R/col.name.list.R:16


def _i32(a):
    a = np.asarray(a)
    if a.dtype != np.int32 and a.size and np.issubdtype(a.dtype, np.integer):
        if int(a.max()) > 2147483647 or int(a.min()) < -2147483648:   # R integers are 32-bit; do not wrap silently
            raise ValueError("values do not fit a 32-bit integer")
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


class Plan:
    """A batch of independent problems.  add() copies rows to the host side of the plan;
    upload()/solve()/download() are the three timed phases; run() does all three."""

    def __init__(self, device=-1):
        self._h = _lib.lib.psd_plan_create(int(device))
        if not self._h:
            raise RuntimeError("peaksegdisk_b200: cannot create a plan: " + _lib.lib.psd_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.psd_plan_destroy(self._h)
            self._h = None

    __del__ = close

    def __len__(self):
        return _lib.lib.psd_plan_size(self._h)

    def add(self, chrom_start, chrom_end, coverage, penalty):
        s, sp = _i32(chrom_start)
        e, ep = _i32(chrom_end)
        c, cp = _i32(coverage)
        if not (len(s) == len(e) == len(c)):
            raise ValueError("chromStart, chromEnd and coverage must have the same length")
        is_inf = 1 if (isinstance(penalty, float) and math.isinf(penalty) and penalty > 0) else 0
        pid = _lib.lib.psd_plan_add(self._h, len(c), sp, ep, cp, 0.0 if is_inf else float(penalty), is_inf)
        if pid < 0:
            raise ValueError(_lib.status_text(-pid, penalty=str(penalty)))
        return pid

    def add_counts(self, counts, penalty):
        """One problem from a count vector (counts[i] = coverage of base [i, i+1)); the run-length
        encoding into rows happens on the device at upload (R/PeakSegFPOP_vec.R:18-25 does it in R)."""
        v = np.asarray(counts)
        if not np.issubdtype(v.dtype, np.integer):
            raise ValueError("count.vec must be integer")
        c, cp = _i32(v)
        is_inf = 1 if (isinstance(penalty, float) and math.isinf(penalty) and penalty > 0) else 0
        pid = _lib.lib.psd_plan_add_counts(self._h, len(c), cp, 0.0 if is_inf else float(penalty), is_inf)
        if pid < 0:
            raise ValueError(_lib.status_text(-pid, penalty=str(penalty)))
        return pid

    def set_penalty(self, pid, penalty):
        is_inf = 1 if (math.isinf(penalty) and penalty > 0) else 0
        rc = _lib.lib.psd_plan_set_penalty(self._h, pid, 0.0 if is_inf else float(penalty), is_inf)
        if rc:
            raise ValueError(_lib.status_text(rc, penalty=str(penalty)))

    def _call(self, fn, stream):
        rc = fn(self._h, C.c_void_p(int(stream) if stream else 0))
        if rc:
            raise RuntimeError("peaksegdisk_b200: " + _lib.status_text(rc))

    def upload(self, stream=0):
        self._call(_lib.lib.psd_plan_upload, stream)

    def solve(self, stream=0):
        self._call(_lib.lib.psd_plan_solve, stream)

    def download(self, stream=0):
        self._call(_lib.lib.psd_plan_download, stream)

    def run(self, stream=0):
        self._call(_lib.lib.psd_plan_run, stream)

    def result(self, pid):
        r = _lib.PsdResult()
        rc = _lib.lib.psd_plan_result(self._h, pid, C.byref(r))
        if rc:
            raise RuntimeError(_lib.status_text(rc))
        return r

    def loss_row(self, pid):
        """The 10 fields of _loss.tsv as a dict (R/col.name.list.R:12-15)."""
        r = self.result(pid)
        if r.status:
            raise RuntimeError("peaksegdisk_b200: " + _lib.status_text(r.status))
        vals = [r.penalty, r.n_segments, r.n_peaks, int(r.bases), r.n_rows, r.mean_pen_cost, r.total_loss,
                r.n_equality, r.mean_intervals, r.max_intervals]
        return dict(zip(LOSS_COLUMNS, vals))

    def segments(self, pid):
        """(chromStart, chromEnd, is_peak, mean) arrays, last segment first as in _segments.bed."""
        r = self.result(pid)
        if r.status:
            raise RuntimeError("peaksegdisk_b200: " + _lib.status_text(r.status))
        n = r.n_segments
        s = np.zeros(n, np.int32); e = np.zeros(n, np.int32); k = np.zeros(n, np.int32); m = np.zeros(n)
        rc = _lib.lib.psd_plan_segments(self._h, pid, s.ctypes.data_as(C.POINTER(C.c_int32)),
                                        e.ctypes.data_as(C.POINTER(C.c_int32)), k.ctypes.data_as(C.POINTER(C.c_int32)),
                                        m.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise RuntimeError(_lib.status_text(rc))
        return s, e, k, m

    def store_function(self, pid, row, which, cap=65536):
        """The stored cost function `which` (0 up, 1 down) of row `row`: (max_log_mean, data_i,
        prev_log_mean) arrays, the fields of the reference's db record (src/PeakSegFPOPLog.cpp:18-34)."""
        n = C.c_int32(0)
        hi = np.zeros(cap); bi = np.zeros(cap, np.int32); bx = np.zeros(cap)
        rc = _lib.lib.psd_plan_store_function(self._h, pid, row, which, cap, C.byref(n), hi.ctypes.data_as(C.POINTER(C.c_double)),
                                              bi.ctypes.data_as(C.POINTER(C.c_int32)), bx.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise RuntimeError(_lib.status_text(rc))
        return hi[:n.value].copy(), bi[:n.value].copy(), bx[:n.value].copy()

    def stats(self):
        st = _lib.PsdStats()
        _lib.lib.psd_plan_get_stats(self._h, C.byref(st))
        return {name: getattr(st, name) for name, _ in st._fields_}


def solve_batch(problems, device=-1, stream=0):
    """problems: iterable of (chromStart, chromEnd, coverage, penalty).  Returns (plan, ids)."""
    plan = Plan(device)
    ids = [plan.add(s, e, c, pen) for (s, e, c, pen) in problems]
    plan.run(stream)
    return plan, ids


def solve_counts_batch(problems, device=-1, stream=0):
    """problems: iterable of (count_vector, penalty): the in-memory form of PeakSegFPOP_vec for many
    vectors at once; the vectors are run-length encoded on the device.  Returns (plan, ids)."""
    plan = Plan(device)
    ids = [plan.add_counts(v, pen) for (v, pen) in problems]
    plan.run(stream)
    return plan, ids
