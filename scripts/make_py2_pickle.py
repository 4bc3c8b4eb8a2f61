"""Hand-assembles tests/golden/stitcher_py2.pkl: the bytes CPython 2.7 + numpy write for
``pickle.dump(stitcher, f, pickle.HIGHEST_PROTOCOL)`` of a calibrated reference ``Stitcher``
(PostScripts/Stitcher/StitcherClass.py:138-148) - protocol 2 with Python-2 ``str`` opcodes
(SHORT_BINSTRING / BINSTRING, no BINUNICODE), the labels as a numpy ``'S4'`` array and the
homographies as lists of float64 row arrays (``params_to_list`` does ``list(ndarray)``), both through
``numpy.core.multiarray._reconstruct`` with raw byte-string states.  No Python 2 exists in this image,
so the opcode stream is written out by hand below; ``pickletools.dis`` of the result is checked into
tests/golden/README.md's description.

The instance opcodes cover both class flavours Python 2 has: ``Stitcher`` is written the way an
OLD-style class instance is (MARK, class, OBJ, dict, BUILD), ``StitcherBase`` the way a new-style one
is (class, empty tuple, NEWOBJ, dict, BUILD) - the reference's ``Debugger`` base is not shipped, so
either may be what robots hold.

    python scripts/make_py2_pickle.py
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multicamera_stitching_b200 import synthetic  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "stitcher_py2.pkl")


class Py2Writer(object):
    def __init__(self):
        self.b = bytearray(b"\x80\x02")   # PROTO 2
        self.memo = 0

    def put(self):                          # BINPUT, like py2's memoize() after every container / string
        self.b += b"q" + bytes([self.memo]) if self.memo < 256 else b"r" + struct.pack("<i", self.memo)
        self.memo += 1

    def raw_str(self, data, memo=True):     # a Python-2 str: SHORT_BINSTRING / BINSTRING
        data = bytes(data)
        self.b += (b"U" + bytes([len(data)]) if len(data) < 256 else b"T" + struct.pack("<i", len(data))) + data
        if memo:
            self.put()

    def glob(self, module, name):
        self.b += b"c" + module.encode() + b"\n" + name.encode() + b"\n"
        self.put()

    def integer(self, v):
        v = int(v)
        if 0 <= v < 256:
            self.b += b"K" + bytes([v])
        elif 0 <= v < 65536:
            self.b += b"M" + struct.pack("<H", v)
        else:
            self.b += b"J" + struct.pack("<i", v)

    def value(self, v):
        if v is None:
            self.b += b"N"
        elif v is True:
            self.b += b"\x88"
        elif v is False:
            self.b += b"\x89"
        elif isinstance(v, (int, np.integer)):
            self.integer(v)
        elif isinstance(v, float):
            self.b += b"G" + struct.pack(">d", v)
        elif isinstance(v, str):
            self.raw_str(v.encode("ascii"))
        elif isinstance(v, np.ndarray):
            self.ndarray(v)
        elif isinstance(v, tuple):
            for x in v:
                self.value(x)
            self.b += {0: b")", 1: b"\x85", 2: b"\x86", 3: b"\x87"}[len(v)]
            if len(v):
                self.put()
        elif isinstance(v, list):
            self.b += b"]"
            self.put()
            if v:
                self.b += b"("
                for x in v:
                    self.value(x)
                self.b += b"e"
        else:
            raise TypeError(type(v))

    def ndarray(self, a):
        # numpy.core.multiarray._reconstruct(numpy.ndarray, (0,), 'b') + __setstate__((1, shape, dtype, False, raw))
        a = np.ascontiguousarray(a)
        self.glob("numpy.core.multiarray", "_reconstruct")
        self.glob("numpy", "ndarray")
        self.value((0,))
        self.raw_str(b"b")
        self.b += b"\x87"
        self.put()
        self.b += b"R"
        self.put()
        self.b += b"("                                   # state tuple via MARK ... TUPLE
        self.integer(1)
        self.value(tuple(int(s) for s in a.shape))
        self.glob("numpy", "dtype")                      # dtype('f8' | 'S4', 0, 1)
        self.raw_str(a.dtype.str.lstrip("<|=").encode())
        self.integer(0)
        self.integer(1)
        self.b += b"\x87"
        self.put()
        self.b += b"R"
        self.put()
        self.b += b"("                                   # dtype state (3, byteorder, None, None, None, size, align, flags)
        self.integer(3)
        self.raw_str(b"<" if a.dtype.kind == "f" else b"|")
        self.b += b"NNN"
        if a.dtype.kind == "f":
            self.b += b"J" + struct.pack("<i", -1) + b"J" + struct.pack("<i", -1)
        else:
            self.integer(a.dtype.itemsize)
            self.integer(1)
        self.integer(0)
        self.b += b"t"
        self.put()
        self.b += b"b"
        self.b += b"\x89"                                # not Fortran order
        self.raw_str(a.tobytes())
        self.b += b"t"
        self.put()
        self.b += b"b"

    def attrs(self, d):                                  # instance __dict__ + BUILD
        self.b += b"}"
        self.put()
        self.b += b"("
        for k, v in d.items():
            self.raw_str(k.encode("ascii"))
            self.value(v)
        self.b += b"ub"


def base_state(s):
    """__dict__ of a reference StitcherBase after params_to_list() (reference :190-209, :485-494)."""
    def rows(m):
        return None if m is None else [np.asarray(r, dtype=np.float64) for r in np.asarray(m)]
    return {
        "sid": str(s.sid), "super_mode": bool(s.super_mode),
        "cachedBH": rows(s.cachedBH), "cachedBINVH": rows(s.cachedBINVH),
        "Bpts": None if s.Bpts is None else [tuple(int(v) for v in p) for p in s.Bpts],
        "BimgSize": None if s.BimgSize is None else tuple(int(v) for v in s.BimgSize),
        "cachedAH": rows(s.cachedAH), "cachedAINVH": rows(s.cachedAINVH),
        "Apts": None if s.Apts is None else [[int(v) for v in p] for p in s.Apts],
        "AimgSize": None if s.AimgSize is None else tuple(int(v) for v in s.AimgSize),
        "matches": None, "status": None,
        "ABSize": None if s.ABSize is None else tuple(int(v) for v in s.ABSize),
        "x_limits": None if s.x_limits is None else [int(v) for v in s.x_limits],
        "y_limits": None if s.y_limits is None else [int(v) for v in s.y_limits],
    }


def main():
    st, _, labels, _ = synthetic.synthetic_stitcher(3, 72, 128, 3)
    w = Py2Writer()
    # old-style instance: MARK, class, OBJ
    w.b += b"("
    w.glob("StitcherClass", "Stitcher")
    w.b += b"o"
    w.put()
    w.b += b"}"
    w.put()
    w.b += b"("
    w.raw_str(b"img_labels")
    w.ndarray(np.array([l.encode("ascii") for l in labels], dtype="S4"))
    w.raw_str(b"stitcher_labels")
    w.value([str(l) for l in st.stitcher_labels])
    w.raw_str(b"stitchers")
    w.b += b"]"
    w.put()
    w.b += b"("
    for s in st.stitchers:
        # new-style instance: class, (), NEWOBJ
        w.b += b"cStitcherClass\nStitcherBase\n"
        w.put()
        w.b += b")\x81"
        w.put()
        w.attrs(base_state(s))
    w.b += b"e"
    w.b += b"ub."
    with open(OUT, "wb") as f:
        f.write(bytes(w.b))
    print("wrote", os.path.normpath(OUT), len(w.b), "bytes")


if __name__ == "__main__":
    main()
