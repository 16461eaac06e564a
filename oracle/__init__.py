"""oracle -- TEST INFRASTRUCTURE, never on the product path.

ctypes bindings for the two CPU checkers:

* ``Oracle``    : this repo's plain-C restatement of the reference's longest-match
                  search + greedy parse (oracle/sqz_oracle.c; cites
                  /root/reference/attic/map_experiment/squeeze.h:337-395).
* ``Reference`` : the UNMODIFIED reference codec compiled from /root/reference
                  into oracle/_ref/ by oracle/Makefile (ref_harness.c).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference) may import this package.  sqz_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

RULES_G1 = (3, 257)          # min_len, max_len; max_dist = window - 1   squeeze.h:13-15,342
RULES_HEAD = (2, 254)        # src/sqz.c:29-30,637-654 (compiled out)    max_dist = window - 1
RULES_BST = (2, 254)         # bst.c:3,230-252                           max_dist = window


def build(force: bool = False) -> None:
    """Compile liboracle.so, and oracle/_ref when /root/reference is mounted."""
    args = ["make", "-s", "-C", HERE, "all"]
    if force:
        subprocess.check_call(["make", "-s", "-C", HERE, "clean"])
    subprocess.check_call(args)


class _Rules(C.Structure):
    _fields_ = [("min_len", C.c_uint32), ("max_len", C.c_uint32), ("max_dist", C.c_uint32)]


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def fnv1a64(buf) -> int:
    return Oracle.get().fnv(buf)


class Oracle:
    _inst = None

    @classmethod
    def get(cls) -> "Oracle":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        u8p, u16p, u32p, u64 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.c_uint64
        rp = C.POINTER(_Rules)
        L.oracle_best.argtypes = [u8p, u64, u64, rp, u32p, u32p]
        L.oracle_best.restype = None
        L.oracle_match_table.argtypes = [u8p, u64, rp, u64, u64, u16p, u16p]
        L.oracle_match_table.restype = None
        L.oracle_fast_table.argtypes = [u8p, u64, rp, u64, u64, u16p, u16p]
        L.oracle_fast_table.restype = None
        L.oracle_tokens.argtypes = [u8p, u64, rp, u32p, u64]
        L.oracle_tokens.restype = u64
        L.oracle_tokens_from_table.argtypes = [u8p, u64, u16p, u16p, C.c_uint32, u64, u32p, u64, C.POINTER(u64)]
        L.oracle_tokens_from_table.restype = u64
        L.oracle_fnv1a64.argtypes = [C.c_void_p, u64]
        L.oracle_fnv1a64.restype = u64
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_get_threads.restype = C.c_int
        self.L = L

    @staticmethod
    def rules(window: int, min_len: int = 3, max_len: int = 257, max_dist: int | None = None) -> _Rules:
        return _Rules(min_len, max_len, window - 1 if max_dist is None else max_dist)

    def threads(self) -> int:
        return int(self.L.oracle_get_threads())

    def set_threads(self, n: int) -> None:
        self.L.oracle_set_threads(int(n))

    def fnv(self, buf) -> int:
        a = np.ascontiguousarray(buf)
        return int(self.L.oracle_fnv1a64(a.ctypes.data, a.nbytes))

    def best(self, data, i: int, window: int, **kw):
        d = _u8(data)
        r = self.rules(window, **kw)
        l, p = C.c_uint32(), C.c_uint32()
        self.L.oracle_best(_ptr(d, C.c_uint8), d.size, i, C.byref(r), C.byref(l), C.byref(p))
        return l.value, p.value

    def match_table(self, data, window: int, first: int = 0, count: int | None = None, fast: bool = False, **kw):
        d = _u8(data)
        if count is None:
            count = d.size - first
        r = self.rules(window, **kw)
        ln = np.zeros(count, dtype=np.uint16)
        ds = np.zeros(count, dtype=np.uint16)
        fn = self.L.oracle_fast_table if fast else self.L.oracle_match_table
        fn(_ptr(d, C.c_uint8), d.size, C.byref(r), first, count, _ptr(ln, C.c_uint16), _ptr(ds, C.c_uint16))
        return ln, ds

    def tokens(self, data, window: int, **kw) -> np.ndarray:
        d = _u8(data)
        r = self.rules(window, **kw)
        cap = max(d.size, 1)
        t = np.zeros(cap, dtype=np.uint32)
        n = self.L.oracle_tokens(_ptr(d, C.c_uint8), d.size, C.byref(r), _ptr(t, C.c_uint32), cap)
        return t[:n].copy()

    def tokens_from_table(self, data, ln, ds, min_len: int = 3, start: int = 0):
        d = _u8(data)
        ln = np.ascontiguousarray(ln, dtype=np.uint16)
        ds = np.ascontiguousarray(ds, dtype=np.uint16)
        cap = max(d.size, 1)
        t = np.zeros(cap, dtype=np.uint32)
        end = C.c_uint64()
        n = self.L.oracle_tokens_from_table(_ptr(d, C.c_uint8), d.size, _ptr(ln, C.c_uint16), _ptr(ds, C.c_uint16),
                                            min_len, start, _ptr(t, C.c_uint32), cap, C.byref(end))
        return t[:n].copy(), int(end.value)


class Reference:
    """The unmodified reference (oracle/_ref/libsqzref*.so)."""

    _inst = {}

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(os.path.join(REF_DIR, "libsqzref.so"))

    @classmethod
    def get(cls, release: bool = False) -> "Reference":
        if release not in cls._inst:
            cls._inst[release] = cls(release)
        return cls._inst[release]

    def __init__(self, release: bool = False):
        name = "libsqzref_rel.so" if release else "libsqzref.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference is mounted")
        L = C.CDLL(path)
        u8p, u32p, u64 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_uint64
        u64p, dp = C.POINTER(u64), C.POINTER(C.c_double)
        L.ref_compress.argtypes = [u8p, u64, C.c_int, u8p, u64, u64p, dp]
        L.ref_compress_cb.argtypes = [u8p, u64, C.c_int, u8p, u64, u64p]
        L.ref_read_header.argtypes = [u8p, u64, u64p, C.POINTER(C.c_int)]
        L.ref_decompress.argtypes = [u8p, u64, u8p, u64, u64p, dp]
        L.ref_tokens.argtypes = [u8p, u64, u32p, u64, u64p]
        L.ref_encode_tokens.argtypes = [u32p, u64, u64, C.c_int, C.c_int, u8p, u64, u64p, dp]
        self.L = L
        self.release = release
        self.last_seconds = 0.0
        bst = os.path.join(REF_DIR, "libsqzbst.so")
        self.B = None
        if os.path.exists(bst):
            B = C.CDLL(bst)
            B.ref_bst_table.argtypes = [u8p, u64, u64, u64, u64, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
            B.ref_bst_table.restype = None
            self.B = B

    @staticmethod
    def _cap(n: int) -> int:
        return (n * 9) // 8 + 4096

    def compress(self, data, win_bits: int = 15, file_mode: bool = False) -> bytes:
        d = _u8(data)
        out = np.zeros(self._cap(d.size), dtype=np.uint8)
        w = C.c_uint64()
        if file_mode:
            r = self.L.ref_compress_cb(_ptr(d, C.c_uint8), d.size, win_bits, _ptr(out, C.c_uint8), out.size, C.byref(w))
        else:
            s = C.c_double()
            r = self.L.ref_compress(_ptr(d, C.c_uint8), d.size, win_bits, _ptr(out, C.c_uint8), out.size,
                                    C.byref(w), C.byref(s))
            self.last_seconds = s.value
        if r != 0:
            raise OSError(r, "reference squeeze.compress failed")
        return out[: w.value].tobytes()

    def decompress(self, comp) -> bytes:
        c = _u8(comp)
        n, wb = C.c_uint64(), C.c_int()
        r = self.L.ref_read_header(_ptr(c, C.c_uint8), c.size, C.byref(n), C.byref(wb))
        if r != 0:
            raise OSError(r, "reference squeeze.read_header failed")
        out = np.zeros(max(n.value, 1), dtype=np.uint8)
        got, s = C.c_uint64(), C.c_double()
        r = self.L.ref_decompress(_ptr(c, C.c_uint8), c.size, _ptr(out, C.c_uint8), out.size, C.byref(got), C.byref(s))
        self.last_seconds = s.value
        if r != 0:
            raise OSError(r, "reference squeeze.decompress failed")
        return out[: got.value].tobytes()

    def tokens(self, comp) -> np.ndarray:
        c = _u8(comp)
        n, wb = C.c_uint64(), C.c_int()
        r = self.L.ref_read_header(_ptr(c, C.c_uint8), c.size, C.byref(n), C.byref(wb))
        if r != 0:
            raise OSError(r, "reference squeeze.read_header failed")
        t = np.zeros(max(n.value, 1), dtype=np.uint32)
        cnt = C.c_uint64()
        r = self.L.ref_tokens(_ptr(c, C.c_uint8), c.size, _ptr(t, C.c_uint32), t.size, C.byref(cnt))
        if r != 0:
            raise OSError(r, "reference token walk failed")
        return t[: cnt.value].copy()

    def encode_tokens(self, tokens, nbytes: int, win_bits: int = 15, file_mode: bool = False) -> bytes:
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        out = np.zeros(self._cap(nbytes) + 64, dtype=np.uint8)
        w, s = C.c_uint64(), C.c_double()
        r = self.L.ref_encode_tokens(_ptr(t, C.c_uint32), t.size, nbytes, win_bits, int(file_mode),
                                     _ptr(out, C.c_uint8), out.size, C.byref(w), C.byref(s))
        self.last_seconds = s.value
        if r != 0:
            raise OSError(r, "reference token encode failed")
        return out[: w.value].tobytes()

    def bst_table(self, data, window: int, first: int = 0, count: int | None = None):
        """bst.c:230-252 lz77_find for positions [first, first+count) (rule set iii)."""
        if self.B is None:
            raise FileNotFoundError("oracle/_ref/libsqzbst.so")
        d = _u8(data)
        if count is None:
            count = d.size - first
        ln = np.zeros(count, dtype=np.uint16)
        ds = np.zeros(count, dtype=np.uint16)
        self.B.ref_bst_table(_ptr(d, C.c_uint8), d.size, window, first, count, _ptr(ln, C.c_uint16), _ptr(ds, C.c_uint16))
        return ln, ds
