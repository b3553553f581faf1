"""Dev/test helper: run one problem through the oracle and through the device source under the
CPU warp emulator, compare per-row cost functions bit for bit, report the first difference."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ORACLE_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double))
EMU_TRACE = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double))

def load():
    orc = C.CDLL(os.path.join(ROOT, "oracle/_build/liboracle_fpop.so"))
    emu = C.CDLL(os.environ.get("PSD_EMU_LIB") or os.path.join(ROOT, "tests/_build/libpsd_emu.so"))
    orc.oracle_fpop_rows.restype = C.c_int
    emu.emu_fpop_rows.restype = C.c_int
    return orc, emu

def ip(a): return a.ctypes.data_as(C.POINTER(C.c_int))
def dp(a): return a.ctypes.data_as(C.POINTER(C.c_double))

def run_oracle(orc, s, e, c, pen, trace=False, math_mode=1):
    n = len(c)
    rows = {}
    def hook(user, row, which, npc, a, b, cc, hi, bi, bx):
        rows[(row, which)] = (np.array(a[:npc]), np.array(b[:npc]), np.array(cc[:npc]), np.array(hi[:npc]),
                              np.array(bi[:npc]), np.array(bx[:npc]))
    cb = ORACLE_HOOK(hook)
    orc.oracle_set_math(math_mode)
    orc.oracle_set_row_hook(cb if trace else C.cast(None, ORACLE_HOOK), None)
    summ = np.zeros(10); ss = np.zeros(n, np.int32); se = np.zeros(n, np.int32); sp = np.zeros(n, np.int32); sm = np.zeros(n)
    st = orc.oracle_fpop_rows(n, ip(s), ip(e), ip(c), C.c_double(pen), 0, dp(summ), ip(ss), ip(se), ip(sp), dp(sm))
    orc.oracle_set_row_hook(C.cast(None, ORACLE_HOOK), None)
    k = int(summ[1])
    return st, summ, (ss[:k].copy(), se[:k].copy(), sp[:k].copy(), sm[:k].copy()), rows

def run_emu(emu, s, e, c, pen, cap=64, descending=0, trace=False, spill_cap=0, info=None):
    n = len(c)
    rows = {}
    def tr(user, row, which, npc, cap_, base):
        arr = np.ctypeslib.as_array(base, shape=(cap_ * 6,))
        a = arr[0:npc].copy(); b = arr[cap_:cap_ + npc].copy(); cc = arr[2 * cap_:2 * cap_ + npc].copy()
        hi = arr[3 * cap_:3 * cap_ + npc].copy(); bx = arr[4 * cap_:4 * cap_ + npc].copy()
        bi = np.frombuffer(arr[5 * cap_:5 * cap_ + (cap_ + 1) // 2].tobytes(), dtype=np.int32)[:npc].copy()
        rows[(row, which)] = (a, b, cc, hi, bi, bx)
    cb = EMU_TRACE(tr)
    summ = np.zeros(10); ss = np.zeros(n + 1, np.int32); se = np.zeros(n + 1, np.int32); sp = np.zeros(n + 1, np.int32); sm = np.zeros(n + 1)
    nsp = C.c_int(0)
    st = emu.emu_fpop_rows(n, ip(s), ip(e), ip(c), C.c_double(pen), cap, spill_cap, descending, dp(summ), ip(ss), ip(se), ip(sp), dp(sm),
                           cb if trace else C.cast(None, EMU_TRACE), None, C.byref(nsp))
    if info is not None:
        info['spills'] = nsp.value
    k = int(summ[1])
    return st, summ, (ss[:k].copy(), se[:k].copy(), sp[:k].copy(), sm[:k].copy()), rows

def bits(a): return np.asarray(a, dtype=np.float64).view(np.uint64)

def first_row_diff(ro, re_, n):
    names = ["a", "b", "c", "hi", "back_i", "back_x"]
    for t in range(n):
        for which in (0, 1):
            fo, fe = ro.get((t, which)), re_.get((t, which))
            if fo is None and fe is None: continue
            if fo is None or fe is None: return "row %d which %d missing in %s" % (t, which, "oracle" if fo is None else "emu")
            if len(fo[0]) != len(fe[0]):
                return "row %d %s: n_pieces oracle=%d emu=%d\n oracle hi=%s\n emu    hi=%s\n oracle bx=%s\n emu bx=%s" % (
                    t, "up" if which == 0 else "down", len(fo[0]), len(fe[0]), fo[3], fe[3], fo[5], fe[5])
            for k, nm in enumerate(names):
                x, y = fo[k], fe[k]
                same = np.array_equal(x, y) if nm == "back_i" else np.array_equal(bits(x), bits(y))
                if not same:
                    return "row %d %s field %s differs\n oracle=%s\n emu   =%s" % (t, "up" if which == 0 else "down", nm, x, y)
    return None

def compare(s, e, c, pen, cap=64, descending=0, trace=True, verbose=True, spill_cap=0, info=None):
    orc, emu = load()
    so, summ_o, seg_o, rows_o = run_oracle(orc, s, e, c, pen, trace)
    se_, summ_e, seg_e, rows_e = run_emu(emu, s, e, c, pen, cap, descending, trace, spill_cap, info)
    ok = so == se_ == 0 and np.array_equal(bits(summ_o), bits(summ_e)) and all(
        np.array_equal(x, y) if x.dtype != np.float64 else np.array_equal(bits(x), bits(y)) for x, y in zip(seg_o, seg_e))
    msg = None
    if trace:
        msg = first_row_diff(rows_o, rows_e, len(c))
        ok = ok and msg is None
    if verbose and not ok:
        print("status oracle=%d emu=%d" % (so, se_)); print("oracle summary", summ_o); print("emu    summary", summ_e); print(msg)
    return ok

if __name__ == "__main__":
    from peaksegdisk_b200 import synth
    s = np.array([0, 10, 20, 30], np.int32); e = s + 10; c = np.array([2, 10, 14, 13], np.int32)
    print("four rows:", compare(s, e, c, 10.5))
