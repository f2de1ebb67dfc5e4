"""ctypes binding of oracle/libbr_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under br_b200/ does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "libbr_oracle.so"

ONE, TWO, GRAPH, GREEDY, GAP_SIZE = range(5)
METHOD_IDS = {"one": ONE, "two": TWO, "graph": GRAPH, "greedy": GREEDY, "gap_size": GAP_SIZE, "gap-size": GAP_SIZE}


def build(force=False):
    src_m = max((_DIR / f).stat().st_mtime for f in ("br_oracle.cpp", "br_oracle.h", "Makefile"))
    if force or not _SO.exists() or _SO.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def _load():
    if not _SO.exists():
        build()
    lib = C.CDLL(str(_SO))
    u8p, u64p, vp, sz = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)

    sig("bro_nuc2bit", C.c_uint64, C.c_uint8)
    sig("bro_bit2nuc", C.c_uint8, C.c_uint64)
    sig("bro_seq2bit", C.c_uint64, C.c_char_p, sz)
    sig("bro_revcomp", C.c_uint64, C.c_uint64, C.c_int)
    sig("bro_canonical", C.c_uint64, C.c_uint64, C.c_int)
    sig("bro_set_new", vp, C.c_int)
    sig("bro_set_from_bitfield", vp, C.c_int, vp, sz)
    sig("bro_hash_new", vp, C.c_int)
    sig("bro_hash_add_reads", None, vp, vp, vp, sz)
    sig("bro_hash_size", sz, vp)
    sig("bro_set_free", None, vp)
    sig("bro_set_k", C.c_int, vp)
    sig("bro_set_set", None, vp, C.c_uint64, C.c_int)
    sig("bro_set_get", C.c_int, vp, C.c_uint64)
    sig("bro_set_bits", vp, vp, C.POINTER(sz))
    sig("bro_set_get_batch", None, vp, vp, sz, vp)
    sig("bro_set_insert_all_kmers", None, vp, C.c_char_p, sz)
    sig("bro_counter_new", vp, C.c_int)
    sig("bro_counter_free", None, vp)
    sig("bro_counter_count", None, vp, vp, vp, sz, C.c_int)
    sig("bro_counter_raw", vp, vp, C.POINTER(sz))
    sig("bro_spectrum", None, vp, vp, C.c_int)
    sig("bro_first_minimum", C.c_int, vp)
    sig("bro_spectrum_threshold", C.c_int, vp, C.c_int, C.c_double)
    sig("bro_solid_from_count", vp, vp, C.c_int, C.c_int)
    sig("bro_alt_nucs", C.c_int, vp, C.c_uint64, vp)
    sig("bro_next_nucs", C.c_int, vp, C.c_uint64, vp)
    sig("bro_correct_error", C.c_long, vp, C.c_int, C.c_int, C.c_int, C.c_uint64, vp, sz, vp, sz, C.POINTER(sz))
    sig("bro_correct", sz, vp, C.c_int, C.c_int, C.c_int, vp, sz, vp, sz)
    sig("bro_run_correction", vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, vp, vp, sz, C.c_int)
    sig("bro_result_data", vp, vp)
    sig("bro_result_offsets", vp, vp)
    sig("bro_result_free", None, vp)
    sig("bro_bio_global", sz, vp, sz, vp, sz, vp, sz)
    sig("bro_match_alignement", C.c_int, vp, sz, vp, sz, vp, sz, C.POINTER(C.c_long))
    sig("bro_max_threads", C.c_int)
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def max_threads():
    return lib().bro_max_threads()


def seq2bit(s: bytes) -> int:
    return lib().bro_seq2bit(s, len(s))


def canonical(kmer: int, k: int) -> int:
    return lib().bro_canonical(kmer, k)


def revcomp(kmer: int, k: int) -> int:
    return lib().bro_revcomp(kmer, k)


class Solid:
    """pcon::solid::Solid as used behind set::Pcon (src/set/pcon.rs)."""

    def __init__(self, k=None, _handle=None):
        self._h = _handle if _handle is not None else lib().bro_set_new(k)
        assert self._h

    @classmethod
    def from_bitfield(cls, k, bits):
        b = _u8(bits)
        h = lib().bro_set_from_bitfield(k, _ptr(b), b.size)
        if not h:
            raise ValueError("bitfield length does not match k")
        return cls(_handle=h)

    @classmethod
    def from_solid_payload(cls, payload: bytes):
        """Solid::from_stream after gunzip: byte 0 = k, rest = bitfield (src/set/pcon.rs:18-25)."""
        return cls.from_bitfield(payload[0], payload[1:])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bro_set_free(self._h)
            self._h = None

    @property
    def k(self):
        return lib().bro_set_k(self._h)

    def set(self, kmer, value=True):
        lib().bro_set_set(self._h, kmer, int(value))

    def get(self, kmer):
        return bool(lib().bro_set_get(self._h, kmer))

    def insert_all_kmers(self, seq: bytes):
        lib().bro_set_insert_all_kmers(self._h, seq, len(seq))

    def bits(self):
        n = C.c_size_t()
        p = lib().bro_set_bits(self._h, C.byref(n))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()

    def get_batch(self, kmers):
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(km.size, dtype=np.uint8)
        lib().bro_set_get_batch(self._h, _ptr(km), km.size, _ptr(out))
        return out

    def alt_nucs(self, kmer):
        out = (C.c_uint64 * 4)()
        n = lib().bro_alt_nucs(self._h, kmer, out)
        return list(out[:n])

    def next_nucs(self, kmer):
        out = (C.c_uint64 * 4)()
        n = lib().bro_next_nucs(self._h, kmer, out)
        return list(out[:n])

    # --- Corrector surface -------------------------------------------------------------
    def correct(self, method, seq: bytes, confirm=5, max_search=7):
        s = _u8(seq)
        cap = 4 * s.size + 1024
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n = lib().bro_correct(self._h, method, confirm, max_search, _ptr(s), s.size, _ptr(out), cap)
            if n <= cap:
                return out[:n].tobytes()
            cap = n

    def correct_error(self, method, kmer, seq: bytes, confirm=5, max_search=7):
        s = _u8(seq)
        cap = 1 << 16
        out = np.empty(cap, dtype=np.uint8)
        off = C.c_size_t()
        n = lib().bro_correct_error(self._h, method, confirm, max_search, kmer, _ptr(s), s.size, _ptr(out), cap, C.byref(off))
        if n < 0:
            return None
        assert n <= cap
        return out[:n].tobytes(), off.value

    def run_correction(self, methods, seq, offsets, confirm=5, max_search=7, two_side=False, threads=1):
        """Batch form of src/lib.rs:21-69; returns (bytes ndarray, offsets ndarray)."""
        m = np.ascontiguousarray(methods, dtype=np.uint8)
        s = _u8(seq)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        h = lib().bro_run_correction(self._h, _ptr(m), m.size, confirm, max_search, int(two_side), _ptr(s), _ptr(off), n, threads)
        try:
            po = lib().bro_result_offsets(h)
            o = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
            tot = int(o[-1])
            if tot:
                pd = lib().bro_result_data(h)
                d = np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint8)), shape=(tot,)).copy()
            else:
                d = np.empty(0, dtype=np.uint8)
        finally:
            lib().bro_result_free(h)
        return d, o


class Hash(Solid):
    """set::Hash (src/set/hash.rs): exact set of canonical k-mers behind the same KmerSet::get, any
    k <= 31.  Everything `Solid` offers except the bitfield works on it (correct, run_correction, ...)."""

    def __init__(self, k):
        super().__init__(_handle=lib().bro_hash_new(k))

    @classmethod
    def from_reads(cls, k, seq, offsets):
        """Hash::from_fasta (src/set/hash.rs:41-60) over records given as (seq, offsets)."""
        h = cls(k)
        h.add_reads(seq, offsets)
        return h

    def add_reads(self, seq, offsets):
        s = _u8(seq)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        lib().bro_hash_add_reads(self._h, _ptr(s), _ptr(off), off.size - 1)

    def __len__(self):
        return lib().bro_hash_size(self._h)

    def bits(self):
        raise TypeError("set::Hash has no bitfield")


class Counter:
    """pcon::counter::Counter<u8> + the count2solid glue of src/main.rs:72-115."""

    def __init__(self, k):
        self.k = k
        self._h = lib().bro_counter_new(k)
        if not self._h:
            raise MemoryError("counter table")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bro_counter_free(self._h)
            self._h = None

    def count(self, seq, offsets, threads=1):
        s = _u8(seq)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        lib().bro_counter_count(self._h, _ptr(s), _ptr(off), off.size - 1, threads)

    def raw(self):
        n = C.c_size_t()
        p = lib().bro_counter_raw(self._h, C.byref(n))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,))

    def spectrum(self, threads=1):
        h = np.zeros(256, dtype=np.uint64)
        lib().bro_spectrum(self._h, _ptr(h), threads)
        return h

    @staticmethod
    def first_minimum(hist):
        h = np.ascontiguousarray(hist, dtype=np.uint64)
        r = lib().bro_first_minimum(_ptr(h))
        return None if r < 0 else r

    @staticmethod
    def spectrum_threshold(hist, method, percent):
        """method: "rarefaction" | "percent-most" | "percent-least" (src/cli.rs:227-241)."""
        h = np.ascontiguousarray(hist, dtype=np.uint64)
        code = {"rarefaction": 2, "percent-most": 3, "percent-least": 4}[method]
        r = lib().bro_spectrum_threshold(_ptr(h), code, float(percent))
        return None if r < 0 else r

    def to_solid(self, abundance, threads=1):
        return Solid(_handle=lib().bro_solid_from_count(self._h, abundance, threads))


OPS = "MXDI"  # Match, Subst, Del, Ins


def bio_global(x: bytes, y: bytes) -> str:
    xa, ya = _u8(x), _u8(y)
    cap = xa.size + ya.size + 2
    ops = np.empty(cap, dtype=np.uint8)
    n = lib().bro_bio_global(_ptr(xa), xa.size, _ptr(ya), ya.size, _ptr(ops), cap)
    return "".join(OPS[o] for o in ops[:n])


def match_alignement(before: bytes, read: bytes, corr: bytes):
    b, r, c = _u8(before), _u8(read), _u8(corr)
    off = C.c_long()
    ok = lib().bro_match_alignement(_ptr(b), b.size, _ptr(r), r.size, _ptr(c), c.size, C.byref(off))
    return off.value if ok else None
