"""Host-side mirror of br's `correct` module and batch driver over the C ABI.

    pub trait Corrector { fn valid_kmer(&self) -> &BoxKmerSet;
                          fn correct_error(&self, kmer, seq) -> Option<(Vec<u8>, usize)>;
                          fn k(&self) -> u8;
                          fn correct(&self, seq: &[u8]) -> Vec<u8>; }          src/correct/mod.rs:44-108

Each corrector object is only a (method, parameters, set) description: the work happens in
libbrgpu.so (speculative per-segment scan + per-read merge).  `correct(seq)` exists for the reference's unit tests (one read per
call); `run_correction` / `correct_batch` / `correct_reads` are the real entry points and take the
whole chunk (src/lib.rs:90-128) in one call.
"""
import ctypes as C

import numpy as np

from . import _lib, fasta
from ._lib import check, lib
from .runtime import Reads, _addr, _ptr, as_offsets, as_u8
from .set import KmerSet, Pcon

METHOD_IDS = {
    "one": _lib.ONE,
    "two": _lib.TWO,
    "graph": _lib.GRAPH,
    "greedy": _lib.GREEDY,
    "gap_size": _lib.GAP_SIZE,
    "gap-size": _lib.GAP_SIZE,
}
DEFAULT_CORRECTIONS = ["one", "two", "graph", "greedy", "gap_size"]  # src/cli.rs:121-131
DEFAULT_CONFIRM = 5  # src/cli.rs:135-137
DEFAULT_MAX_SEARCH = 7  # src/cli.rs:140-142
CHUNK_RECORDS = 8192  # hard-coded in src/lib.rs:90


class Corrector:
    method = None

    def __init__(self, valid_kmer: Pcon, confirm=DEFAULT_CONFIRM, max_search=DEFAULT_MAX_SEARCH):
        self._set = valid_kmer
        self.confirm = int(confirm)
        self.max_search = int(max_search)

    def valid_kmer(self) -> KmerSet:
        return self._set

    def k(self) -> int:
        return self._set.k()

    def correct(self, seq: bytes) -> bytes:
        """Corrector::correct (src/correct/mod.rs:53-107) for one read."""
        s = as_u8(seq)
        cap = 2 * s.size + 256
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n = C.c_uint64()
            st = lib.brgpu_correct_one(self._set.ctx._h, self._set._h, self.method, self.confirm, self.max_search,
                                       _ptr(s), s.size, _ptr(out), cap, C.byref(n))
            if st == _lib.E_OVERFLOW and int(n.value) > cap:  # buffer too small: *required tells how much
                cap = int(n.value)
                continue
            check(st, self._set.ctx._h)
            return out[: n.value].tobytes()


class One(Corrector):  # src/correct/exist/one.rs:74
    method = _lib.ONE

    def __init__(self, valid_kmer, c):
        super().__init__(valid_kmer, confirm=c)


class Two(Corrector):  # src/correct/exist/two.rs:328
    method = _lib.TWO

    def __init__(self, valid_kmer, c):
        super().__init__(valid_kmer, confirm=c)


class Graph(Corrector):  # src/correct/graph.rs:29-37
    method = _lib.GRAPH

    def __init__(self, valid_kmer):
        super().__init__(valid_kmer)


class Greedy(Corrector):  # src/correct/greedy.rs:41-54
    method = _lib.GREEDY

    def __init__(self, valid_kmer, max_search, nb_validate):
        super().__init__(valid_kmer, confirm=nb_validate, max_search=max_search)


class GapSize(Corrector):  # src/correct/gap_size.rs:29-42
    method = _lib.GAP_SIZE

    def __init__(self, valid_kmer, c):
        super().__init__(valid_kmer, confirm=c)


def build_methods(params, solid: Pcon, confirm=DEFAULT_CONFIRM, max_search=DEFAULT_MAX_SEARCH):
    """src/lib.rs:141-164 — same argument mapping: One/Two/GapSize get `confirm`, Greedy gets
    (max_search, confirm)."""
    methods = []
    for m in params:
        mid = METHOD_IDS[m] if isinstance(m, str) else int(m)
        if mid == _lib.ONE:
            methods.append(One(solid, confirm))
        elif mid == _lib.TWO:
            methods.append(Two(solid, confirm))
        elif mid == _lib.GRAPH:
            methods.append(Graph(solid))
        elif mid == _lib.GREEDY:
            methods.append(Greedy(solid, max_search, confirm))
        elif mid == _lib.GAP_SIZE:
            methods.append(GapSize(solid, confirm))
        else:
            raise ValueError(f"unknown correction method {m!r}")
    return methods


def _chain_params(methods):
    """The C ABI takes one (confirm, max_search) pair for the chain, as br's CLI does."""
    if not methods:
        return np.empty(0, dtype=np.uint8), DEFAULT_CONFIRM, DEFAULT_MAX_SEARCH, None
    solid = methods[0]._set
    ids = np.array([m.method for m in methods], dtype=np.uint8)
    confirm = {m.confirm for m in methods if not isinstance(m, Graph)}
    max_search = {m.max_search for m in methods if isinstance(m, Greedy)}
    if len(confirm) > 1 or len(max_search) > 1 or any(m._set is not solid for m in methods):
        raise ValueError("all methods of a chain must share the set, confirm and max_search (as in br's CLI)")
    return ids, (confirm.pop() if confirm else DEFAULT_CONFIRM), (max_search.pop() if max_search else DEFAULT_MAX_SEARCH), solid


def correct_reads(methods, reads: Reads, two_side=False, asynchronous=False) -> Reads:
    """Device-resident chunk in, device-resident corrected chunk out (input order kept).
    asynchronous: return once the chain is enqueued (brgpu_correct_reads_async); the set and `reads` must
    stay alive until the result's first use (download, `.bases`, `.wait()`)."""
    ids, confirm, max_search, solid = _chain_params(methods)
    if solid is None:
        raise ValueError("empty method list")
    h = C.c_void_p()
    f = lib.brgpu_correct_reads_async if asynchronous else lib.brgpu_correct_reads
    check(f(solid.ctx._h, solid._h, _ptr(ids), ids.size, confirm, max_search, int(bool(two_side)), reads._h, C.byref(h)),
          solid.ctx._h)
    out = Reads(solid.ctx, h)
    if asynchronous:
        out._keep = (solid, reads)
    return out


def correct_batch(methods, seq, offsets, two_side=False, out=None, out_offsets=None):
    """Host buffers in, host buffers out: the per-chunk body of run_correction (src/lib.rs:93-128)
    in one call.  Returns (uint8 array, uint64 offsets)."""
    ids, confirm, max_search, solid = _chain_params(methods)
    if solid is None:
        raise ValueError("empty method list")
    s, off = as_u8(seq), as_offsets(offsets)
    n = (off.numel() if hasattr(off, "numel") else off.size) - 1
    total = int(off[n]) - int(off[0])
    if out_offsets is None:
        out_offsets = np.empty(n + 1, dtype=np.uint64)
    if out is None:
        out = np.empty(total + total // 8 + 64 * n + 64, dtype=np.uint8)
    while True:
        cap = out.numel() if hasattr(out, "numel") else out.size
        req = C.c_uint64()
        st = lib.brgpu_correct_batch(solid.ctx._h, solid._h, _ptr(ids), ids.size, confirm, max_search,
                                     int(bool(two_side)), _addr(s), _addr(off), n, _addr(out), cap,
                                     _addr(out_offsets), C.byref(req))
        if st == _lib.E_OVERFLOW and req.value > cap:
            out = np.empty(int(req.value), dtype=np.uint8)
            continue
        check(st, solid.ctx._h)
        return out[: req.value], out_offsets


def run_correction(inputs, outputs, methods, two_side=False, record_buffer_len=CHUNK_RECORDS):
    """src/lib.rs:21-139: for each (input, output) pair read FASTA records, correct them chunk by
    chunk (8192 records, the reference's hard-coded chunk), write FASTA in input order.
    `record_buffer_len` is accepted and ignored exactly like the reference does (src/lib.rs:84,90)."""
    del record_buffer_len
    for inp, outp in zip(inputs, outputs):
        defs, seq, off = fasta.read_fasta(inp)
        close = False
        if not hasattr(outp, "write"):
            outp = open(outp, "wb")
            close = True
        try:
            for cdefs, cseq, coff in fasta.iter_chunks(defs, seq, off, CHUNK_RECORDS):
                d, o = correct_batch(methods, cseq, coff, two_side=two_side)
                fasta.write_fasta(outp, cdefs, d, o)
        finally:
            if close:
                outp.close()
