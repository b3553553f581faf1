"""Test-side ctypes binding of the parity oracle (oracle/_build/liboracle_fpop.so) and of the compiled
reference (oracle/_ref).  Imported only by tests, smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_fpop.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_fpop.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_fpop")
# fingerprint of glibc 2.39's FMA exp/log (recorded when tests/golden was generated)
GOLDEN_LIBM_FINGERPRINT_FILE = os.path.join(ROOT, "tests", "golden", "libm_fingerprint.txt")

_orc = None


def ensure_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])


def oracle():
    global _orc
    if _orc is None:
        ensure_built()
        _orc = C.CDLL(ORACLE_SO)
        _orc.oracle_fpop_rows.restype = C.c_int
        _orc.oracle_fpop_disk.restype = C.c_int
        _orc.oracle_fpop_disk.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        _orc.oracle_libm_fingerprint.restype = C.c_uint64
    return _orc


def libm_matches_golden():
    """True when the host libm is bit-identical (on the probe set) to the one the goldens came from."""
    try:
        want = int(open(GOLDEN_LIBM_FINGERPRINT_FILE).read().strip())
    except OSError:
        return False
    return int(oracle().oracle_libm_fingerprint()) == want


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def solve_rows(s, e, c, penalty, math_mode=1):
    """Oracle on in-memory rows.  math_mode 1 = psd_math (bit-identical to glibc 2.39 FMA libm),
    0 = the host's libm.  Returns (status, summary[10], (start, end, is_peak, mean))."""
    o = oracle()
    s = np.ascontiguousarray(s, np.int32); e = np.ascontiguousarray(e, np.int32); c = np.ascontiguousarray(c, np.int32)
    n = len(c)
    o.oracle_set_math(math_mode)
    summ = np.zeros(10)
    ss = np.zeros(n, np.int32); se = np.zeros(n, np.int32); sp = np.zeros(n, np.int32); sm = np.zeros(n)
    is_inf = 1 if np.isinf(penalty) else 0
    st = o.oracle_fpop_rows(n, _ip(s), _ip(e), _ip(c), C.c_double(0.0 if is_inf else penalty), is_inf,
                            summ.ctypes.data_as(C.POINTER(C.c_double)), _ip(ss), _ip(se), _ip(sp),
                            sm.ctypes.data_as(C.POINTER(C.c_double)))
    k = int(summ[1])
    return st, summ, (ss[:k].copy(), se[:k].copy(), sp[:k].copy(), sm[:k].copy())


def oracle_disk(bedgraph, penalty_str, db, math_mode=1):
    o = oracle()
    o.oracle_set_math(math_mode)
    return o.oracle_fpop_disk(bedgraph.encode(), penalty_str.encode(), db.encode())


def ref_available():
    ensure_built()
    return os.path.exists(REF_BIN)


def ref_disk(bedgraph, penalty_str, db):
    """The unmodified reference solver (oracle/_ref/ref_fpop)."""
    return subprocess.call([REF_BIN, bedgraph, penalty_str, db], stdout=subprocess.DEVNULL)
