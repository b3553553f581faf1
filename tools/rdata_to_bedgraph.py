#!/usr/bin/env python3
"""Dev-time fixture generator: read Mono27ac$coverage out of the reference's data/Mono27ac.RData
(xz-compressed R serialization, format RDX2 / XDR) without R, and write it as the tab-separated
bedGraph text R's writeBedGraph would produce (R/writeBedGraph.R:35-37).

usage: tools/rdata_to_bedgraph.py /root/reference/data/Mono27ac.RData tests/golden/Mono27ac_coverage.bedGraph
"""
import lzma, struct, sys

class Reader:
    def __init__(self, b):
        self.b, self.p, self.refs = b, 0, []
    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]; self.p += 4; return v
    def raw(self, n):
        v = self.b[self.p:self.p + n]; self.p += n; return v

    def item(self):
        flags = self.i32()
        t = flags & 0xff
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254: return None                       # NILVALUE
        if t == 255: return self.refs[(flags >> 8) - 1]  # REFSXP
        if t == 1:                                      # SYMSXP
            name = self.item(); self.refs.append(name); return name
        if t == 2:                                      # LISTSXP (pairlist) -> list of (tag, value)
            out = []
            while True:
                if has_attr: self.item()
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.i32(); t = flags & 0xff
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                if t == 254: return out
                if t != 2: raise ValueError("pairlist tail type %d" % t)
        if t == 9:                                      # CHARSXP
            n = self.i32(); return None if n == -1 else self.raw(n).decode()
        if t in (10, 13):                               # LGLSXP / INTSXP
            n = self.i32(); v = list(struct.unpack_from(">%di" % n, self.b, self.p)); self.p += 4 * n
        elif t == 14:                                   # REALSXP
            n = self.i32(); v = list(struct.unpack_from(">%dd" % n, self.b, self.p)); self.p += 8 * n
        elif t == 16:                                   # STRSXP
            n = self.i32(); v = [self.item() for _ in range(n)]
        elif t == 19:                                   # VECSXP
            n = self.i32(); v = [self.item() for _ in range(n)]
        elif t == 22:                                   # EXTPTRSXP (data.table .internal.selfref)
            self.refs.append("extptr"); self.item(); self.item(); v = "extptr"
        else:
            raise ValueError("unsupported SEXP type %d at %d" % (t, self.p))
        attrs = dict(self.item()) if has_attr else {}
        return {"v": v, "attr": attrs} if attrs else v

def main(src, dst):
    b = lzma.open(src).read()
    assert b[:5] == b"RDX2\n" and b[5:7] == b"X\n", b[:8]
    r = Reader(b); r.p = 7
    r.i32(); r.i32(); r.i32()                         # format version, writer version, min reader
    top = dict(r.item())
    mono = top["Mono27ac"]
    names = mono["attr"]["names"]
    cov = mono["v"][names.index("coverage")]
    cols = cov["attr"]["names"]
    data = dict(zip(cols, cov["v"]))
    chrom = data["chrom"]
    if isinstance(chrom, dict):                        # factor
        lv = chrom["attr"]["levels"]; chrom = [lv[i - 1] for i in chrom["v"]]
    n = len(chrom)
    with open(dst, "w") as f:
        for i in range(n):
            f.write("%s\t%d\t%d\t%d\n" % (chrom[i], data["chromStart"][i], data["chromEnd"][i], data["count"][i]))
    print("wrote %d rows to %s" % (n, dst))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
