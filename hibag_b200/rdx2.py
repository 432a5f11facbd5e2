"""Minimal reader for R's RDX2 (XDR) workspace files (.rdata/.RData/.rda), no R needed.

Published HIBAG models and the reference's shipped fixtures (data/*.rdata, inst/extdata/*.RData)
are R workspaces holding an "hlaAttrBagObj" list (reference R/HIBAG.R:1041-1068, hlaModelToObj);
api.hlaModelFromRData turns one into an HLAModel. Also used by tools/make_golden.py.
Handles the SEXP types that occur there: NILSXP, SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP,
REALSXP, STRSXP, VECSXP, REFSXP, NILVALUE_SXP plus attributes.
"""
import gzip
import lzma
import struct

import numpy as np

NA_INTEGER = -2147483648


class RObj:
    """A vector / list value with its R attributes."""
    __slots__ = ("value", "attr")

    def __init__(self, value, attr=None):
        self.value = value
        self.attr = attr or {}

    def names(self):
        n = self.attr.get("names")
        return None if n is None else list(n.value)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.value[self.names().index(key)]
        return self.value[key]

    def __repr__(self):
        return "RObj(%r, attr=%s)" % (type(self.value).__name__, list(self.attr))


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 0xFE or t == 0:      # NILVALUE_SXP / NILSXP
            return None
        if t == 0xFF:                # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 0xFD:                # GLOBALENV_SXP
            return "<globalenv>"
        if t == 1:                   # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                   # LISTSXP (pairlist) -> list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                if t != 2:
                    self.p -= 4
                    tail = self.item()
                    assert tail is None, "dotted pairlist not supported"
                    return out
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
        if t == 9:                   # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("latin-1")
            self.p += n
            return s
        if t in (10, 13):            # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
        elif t == 14:                # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
        elif t == 16:                # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == 19:                # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError("SEXP type %d at offset %d" % (t, self.p))
        attr = {}
        if has_attr:
            for tag, val in self.item():
                attr[tag] = val
        return RObj(v, attr)


def load(path):
    """Return {name: object} for an RDX2 file compressed with gzip, xz or not at all."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    assert raw[:5] == b"RDX2\n" and raw[5:7] == b"X\n", "not an XDR RDX2 file"
    r = _Reader(raw)
    r.p = 7
    r.i32(); r.i32(); r.i32()        # format version, writer R version, min reader version
    top = r.item()
    return {tag: val for tag, val in top}
