"""Reader for R `save()` files (RDX2: gzip + XDR serialization version 2).

TEST INFRASTRUCTURE ONLY: used by tests/golden/make_golden.py, in the build container, to pull
the reference's fixture data/EMresults.rda (man/TitanCNA-dataset.Rd) into a small .npz.  Written
from the public description of R's serialization format (R Internals, "Serialization Formats");
supports the node types that fixture contains.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

NA_INT = -2147483648


class RObj:
    """value + attributes (dict)."""

    def __init__(self, value, attrs=None):
        self.value = value
        self.attrs = attrs or {}

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attrs={list(self.attrs)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:                      # NILVALUE
            return None
        if typ in (253, 242, 241):          # global / empty / base env
            return RObj("<env>")
        if typ == 255:                      # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:                        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ in (2, 6):                   # LISTSXP / LANGSXP: pairlist -> list of (tag, value)
            out = []
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nxt = self.i32()
                t2 = nxt & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6):
                    self.o -= 4
                    out.append((None, self.item()))
                    break
                has_attr = bool(nxt & (1 << 9))
                has_tag = bool(nxt & (1 << 10))
            return out
        if typ == 9:                        # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", errors="replace")
        if typ == 22:                       # EXTPTRSXP
            obj = RObj("<extptr>")
            self.refs.append(obj)
            self.item()
            self.item()
            if has_attr:
                self.item()
            return obj
        if typ == 4:                        # ENVSXP
            obj = RObj("<env>")
            self.refs.append(obj)
            self.i32()
            for _ in range(4):
                self.item()
            return obj
        if typ in (10, 13):                 # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif typ == 14:                     # REALSXP
            n = self.i32()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif typ == 16:                     # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif typ in (19, 20):               # VECSXP / EXPRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"R type {typ} at offset {self.o}")
        attrs = {}
        if has_attr:
            for tag, val in self.item():
                attrs[tag] = val
        if typ == 10:
            v = v.astype(bool) if not np.any(v == NA_INT) else v
        return RObj(v, attrs)


def simplify(o):
    """RObj tree -> plain python: named lists -> dict, matrices reshaped (column-major)."""
    if not isinstance(o, RObj):
        return o
    v = o.value
    names = o.attrs.get("names")
    if isinstance(v, list):
        items = [simplify(x) for x in v]
        if names is not None and isinstance(names, RObj) and all(isinstance(x, RObj) or x is None or True for x in v):
            nm = names.value
            if len(nm) == len(items) and not all(isinstance(x, (str, type(None))) for x in items):
                return {k: it for k, it in zip(nm, items)}
        return items
    if isinstance(v, np.ndarray):
        dim = o.attrs.get("dim")
        if dim is not None:
            return v.reshape(tuple(int(d) for d in dim.value), order="F")
        return v
    return v


def load_rda(path: str) -> dict:
    with gzip.open(path, "rb") as fh:
        buf = fh.read()
    if not buf.startswith(b"RDX2\nX\n"):
        raise ValueError("not an RDX2/XDR file")
    r = _Reader(buf)
    r.o = 7
    ver, _, _ = r.i32(), r.i32(), r.i32()
    if ver != 2:
        raise ValueError(f"serialization version {ver} unsupported")
    top = r.item()
    return {tag: simplify(val) for tag, val in top}
