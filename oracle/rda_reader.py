"""Minimal reader for R's RDX2 (XDR, serialization version 2) workspace files.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by
tests/golden/make_golden.py to pull the `joostTest` fixture out of the
reference's R/sysdata.rda without needing R.  Only the SEXP types that occur in
that file are handled.
"""
import bz2
import gzip
import lzma
import struct


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def ints(self, n):
        import numpy as np
        v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype("int32")
        self.o += 4 * n
        return v

    def reals(self, n):
        import numpy as np
        v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype("float64")
        self.o += 8 * n
        return v

    def length(self):
        n = self.i32()
        if n == -1:  # long vector
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_obj = bool(flags & (1 << 8))
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:  # NILVALUE
            return None
        if t in (253, 242, 241):  # global / empty / base env
            return {"_env": t}
        if t == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (249, 250):  # namespace / package reference
            info = self.item()
            self.refs.append(info)
            return info
        if t == 4:  # ENVSXP
            env = {"_envsxp": True}
            self.refs.append(env)
            self.i32()  # locked
            env["enclos"] = self.item()
            env["frame"] = self.item()
            env["hashtab"] = self.item()
            env["attrib"] = self.item()
            return env
        if t in (2, 6):  # LISTSXP / LANGSXP -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6):
                    raise ValueError("unexpected cdr type %d" % t2)
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("utf-8", "replace")
            self.o += n
            return s
        if t in (10, 13):
            val = self.ints(self.length())
        elif t == 14:
            val = self.reals(self.length())
        elif t == 16:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == 19:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == 25:  # S4SXP: slots are the attributes
            val = {"_s4": True}
        else:
            raise ValueError("unhandled SEXP type %d at offset %d" % (t, self.o))
        if has_attr:
            attrs = dict(self.item())
            if isinstance(val, dict):
                val.update(attrs)
            else:
                val = {"_value": val, "_attr": attrs}
        return val


def _decompress(raw):
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rda(path):
    """Return {name: object} for every top-level object stored in an .rda file."""
    with open(path, "rb") as f:
        buf = _decompress(f.read())
    if buf[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 file")
    if buf[5:7] != b"X\n":
        raise ValueError("only XDR format supported")
    r = _Reader(buf)
    r.o = 7
    r.i32(); r.i32(); r.i32()  # serialization version, writer, min reader
    top = r.item()
    return {k: v for k, v in top}


def strip(v):
    """Drop attribute wrappers: plain vector / list value."""
    if isinstance(v, dict) and "_value" in v:
        return v["_value"]
    return v


def names_of(v):
    if isinstance(v, dict) and "_attr" in v and "names" in v["_attr"]:
        return strip(v["_attr"]["names"])
    return None


def sparse_to_scipy(obj):
    """dgCMatrix / dsCMatrix S4 object -> scipy.sparse.csc_matrix (dsCMatrix mirrored)."""
    import numpy as np
    import scipy.sparse as sp
    i = strip(obj["i"])
    p = strip(obj["p"])
    x = strip(obj["x"])
    dim = strip(obj["Dim"])
    m = sp.csc_matrix((np.asarray(x, dtype=np.float64), np.asarray(i), np.asarray(p)),
                      shape=(int(dim[0]), int(dim[1])))
    if "uplo" in obj:  # symmetric class: only one triangle stored
        m = m + sp.triu(m, 1).T if strip(obj["uplo"])[0] == "U" else m + sp.tril(m, -1).T
    return m.tocsc()
