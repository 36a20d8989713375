"""Minimal reader for R ``.rda`` files (bzip2/gzip + XDR serialisation v2/v3).

ORACLE / TEST INFRASTRUCTURE.  Used only by ``tests/golden/make_golden.py`` to
turn the reference's shipped datasets (``/root/reference/data/*.rda``) into
small ``.npz`` fixtures, because R is not installed here.
Only the SEXP types those two files contain are handled.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.o)
        self.o += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:      # NILVALUE
            return None
        if typ == 255:      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:        # LISTSXP (pairlist) -> list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                typ = flags & 0xFF
                if typ == 254:
                    return out
                if typ != 2:
                    raise ValueError("unexpected pairlist tail type %d" % typ)
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
        if typ == 9:        # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if typ in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
            return self._with_attr(v, has_attr)
        if typ == 14:       # REALSXP
            n = self.i32()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
            return self._with_attr(v, has_attr)
        if typ == 16:       # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        if typ == 19:       # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        raise ValueError("unsupported SEXP type %d at offset %d" % (typ, self.o))

    def _with_attr(self, v, has_attr):
        if not has_attr:
            return v
        attrs = dict(self.item() or [])
        if isinstance(v, list) and "names" in attrs:
            return {"__names__": attrs["names"], "__values__": v, "__attrs__": attrs}
        return v


def read_rda(path: str) -> dict:
    """Return ``{object_name: {column: ndarray}}`` for data.frames in ``path``."""
    with open(path, "rb") as fh:
        blob = fh.read()
    if blob[:3] == b"BZh":
        blob = bz2.decompress(blob)
    elif blob[:2] == b"\x1f\x8b":
        blob = gzip.decompress(blob)
    elif blob[:6] == b"\xfd7zXZ\x00":
        blob = lzma.decompress(blob)
    if blob[:5] != b"RDX2\n" and blob[:5] != b"RDX3\n":
        raise ValueError("not an RDX2/RDX3 file")
    if blob[5:7] != b"X\n":
        raise ValueError("only XDR serialisation supported")
    r = _Reader(blob)
    r.o = 7
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    out = {}
    for name, obj in top:
        if isinstance(obj, dict) and "__names__" in obj:
            out[name] = {k: v for k, v in zip(obj["__names__"], obj["__values__"])}
        else:
            out[name] = obj
    return out
