"""Minimal reader for R's RDX3 (XDR) serialization -- oracle/test infrastructure.

Only what is needed to read the reference's stored known-answer models in
data/precalculated_Xfgpm_objects.rda (documented in
R/8_precalculated_Xfgpm_objects.R:1-76).  Used by tests/golden/make_golden.py in
the build container; never at test/bench run time (the GPU box has no
/root/reference).
"""
import gzip
import struct
import numpy as np


class RObj:
    """An R object with attributes (S4 objects: slots live in attrs)."""

    def __init__(self, kind, value=None, attrs=None):
        self.kind = kind
        self.value = value
        self.attrs = attrs or {}

    def __getitem__(self, k):
        if isinstance(self.value, dict):
            return self.value[k]
        if isinstance(k, str):  # named list
            names = self.attrs.get("names")
            return self.value[list(names.value).index(k)]
        return self.value[k]

    def slot(self, name):
        return self.attrs[name]

    def __repr__(self):
        return f"RObj({self.kind}, attrs={list(self.attrs)})"


class _Reader:
    def __init__(self, data):
        self.d = data
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.d, self.p)[0]
        self.p += 4
        return v

    def f64s(self, n):
        a = np.frombuffer(self.d, dtype=">f8", count=n, offset=self.p).astype(np.float64)
        self.p += 8 * n
        return a

    def i32s(self, n):
        a = np.frombuffer(self.d, dtype=">i4", count=n, offset=self.p).astype(np.int32)
        self.p += 4 * n
        return a

    def length(self):
        n = self.i32()
        if n == -1:
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + lo
        return n

    def pairlist_to_dict(self, flags):
        """Read a LISTSXP chain starting with already-read flags; return ordered dict tag->value."""
        out = {}
        idx = 0
        while True:
            typ = flags & 0xFF
            if typ == 254:  # NILVALUE
                return out
            if typ != 2:  # dotted pair (e.g. ALTREP wrapper state = CONS(data, meta))
                out["__cdr__"] = self.item(flags)
                return out
            has_attr = bool(flags & (1 << 9))
            has_tag = bool(flags & (1 << 10))
            if has_attr:
                self.item()
            tag = None
            if has_tag:
                tag = self.item()
            car = self.item()
            out[tag if tag is not None else idx] = car
            idx += 1
            flags = self.i32()

    def item(self, flags=None):
        if flags is None:
            flags = self.i32()
        typ = flags & 0xFF
        is_obj = bool(flags & (1 << 8))
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:
            return None
        if typ == 253:  # global env
            return RObj("globalenv")
        if typ == 242 or typ == 241:  # empty env / base env
            return RObj("env")
        if typ == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.d[self.p:self.p + n].decode("utf-8", errors="replace")
            self.p += n
            return s
        if typ == 2:  # LISTSXP
            return RObj("pairlist", self.pairlist_to_dict(flags))
        if typ in (6, 3, 5):  # LANGSXP, CLOSXP, PROMSXP: parse structurally as pairlist-like
            attrs = None
            if has_attr:
                attrs = self.item()
            tag = self.item() if has_tag else None
            car = self.item()
            cdr = self.item()
            return RObj("lang", (tag, car, cdr))
        if typ == 4:  # ENVSXP
            self.i32()  # locked
            env = RObj("env")
            self.refs.append(env)
            self.item(); self.item(); self.item(); self.item()
            return env
        if typ in (10, 13):  # LGLSXP, INTSXP
            n = self.length()
            v = self.i32s(n)
            return self._with_attrs("int" if typ == 13 else "lgl", v, has_attr)
        if typ == 14:
            n = self.length()
            v = self.f64s(n)
            return self._with_attrs("real", v, has_attr)
        if typ == 16:
            n = self.length()
            v = [self.item() for _ in range(n)]
            return self._with_attrs("str", v, has_attr)
        if typ in (19, 20):
            n = self.length()
            v = [self.item() for _ in range(n)]
            return self._with_attrs("list", v, has_attr)
        if typ == 25:  # S4SXP
            return self._with_attrs("S4", None, has_attr)
        if typ == 238:  # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = list(info.value.values())[0] if isinstance(info, RObj) else None
            # wrap_real / wrap_integer / ...: state = pairlist(data, meta)
            data = list(state.value.values())[0] if isinstance(state, RObj) and state.kind == "pairlist" else state
            if isinstance(data, RObj):
                if isinstance(attr, RObj) and attr.kind == "pairlist":
                    data = RObj(data.kind, data.value, dict(data.attrs))
                    for k, v in attr.value.items():
                        data.attrs[k] = v
                return data
            raise ValueError(f"unsupported ALTREP class {cls}")
        raise ValueError(f"unsupported SEXP type {typ} at {self.p}")

    def _with_attrs(self, kind, v, has_attr):
        o = RObj(kind, v)
        if has_attr:
            a = self.item()
            if isinstance(a, RObj) and a.kind == "pairlist":
                o.attrs = dict(a.value)
        return o


def read_rda(path):
    """Return {name: RObj} for the top-level objects stored in an .rda (gzip RDX3)."""
    data = gzip.open(path).read()
    assert data[:5] == b"RDX3\n", data[:5]
    r = _Reader(data)
    r.p = 5
    assert data[r.p:r.p + 2] == b"X\n"
    r.p += 2
    r.i32(); r.i32(); r.i32()  # format version, writer R, min reader R
    n = r.i32()
    r.p += n  # native encoding
    top = r.item()
    return dict(top.value)


def as_array(o):
    """REAL/INT vector (+ optional dim attr, column-major) -> numpy array."""
    v = np.asarray(o.value)
    dim = o.attrs.get("dim")
    if dim is not None:
        return v.reshape(tuple(int(x) for x in dim.value), order="F").copy()
    return v
