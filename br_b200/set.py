"""Host-side mirror of br's `set` module (src/set.rs, src/set/pcon.rs) over the C ABI.

    trait KmerSet: Sync { fn get(&self, kmer: u64) -> bool; fn k(&self) -> u8; }   src/set.rs:17-21

`Pcon` keeps the dense canonical bitfield in HBM.  `get` exists for interface parity and for the
reference's unit tests; it is a batch of one (a round trip per k-mer), so real callers use
`get_batch`.  Constructors follow src/set/pcon.rs and the count->solid glue of src/main.rs:72-115.
"""
import ctypes as C
import gzip

import numpy as np

from . import _lib
from ._lib import check, lib
from .runtime import Context, Reads, _addr, _ptr, as_offsets, as_u8


def nuc2bit(b: int) -> int:
    return (b >> 1) & 3


def seq2bit(seq: bytes) -> int:
    """cocktail::kmer::seq2bit — host helper for building k-mers to pass to get()/insert()."""
    kmer = 0
    for b in seq:
        kmer = (kmer << 2) | ((b >> 1) & 3)
    return kmer


SELECTIONS = {
    "first-minimum": _lib.ABUNDANCE_FIRST_MINIMUM, "first_minimum": _lib.ABUNDANCE_FIRST_MINIMUM,
    "rarefaction": _lib.ABUNDANCE_RAREFACTION,
    "percent-most": _lib.ABUNDANCE_PERCENT_AT_MOST, "percent_most": _lib.ABUNDANCE_PERCENT_AT_MOST,
    "percent-least": _lib.ABUNDANCE_PERCENT_AT_LEAST, "percent_least": _lib.ABUNDANCE_PERCENT_AT_LEAST,
}


def selection_code(abundance, abundance_selection):
    """(selection, abundance) for the ABI: an explicit abundance wins (src/main.rs:96); neither an
    abundance nor a method is Error::AbundanceThresholdOrAbundanceMethod (:109)."""
    if abundance is not None:
        return _lib.ABUNDANCE_EXPLICIT, int(abundance)
    if abundance_selection is None:
        return _lib.ABUNDANCE_EXPLICIT, -1
    if abundance_selection not in SELECTIONS:
        raise ValueError(f"unknown abundance selection {abundance_selection!r}")
    return SELECTIONS[abundance_selection], -1


def spectrum_threshold(hist, abundance_selection, percent=0.0):
    """pcon Spectrum::get_threshold on a 256-bin histogram; None when there is no threshold."""
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    r = lib.brgpu_spectrum_threshold(_ptr(h), SELECTIONS[abundance_selection], float(percent))
    return None if r < 0 else r


def csv_kmers(stream, k):
    """seq2bit of the first column of every data record of a CSV (the loop of from_csv, src/set/pcon.rs:34-42,
    src/set/hash.rs:27-36).  A field that is not k letters long cannot name a k-mer of the set and is refused."""
    from .fasta import read_csv_first_column

    fields = read_csv_first_column(stream)
    for i, f in enumerate(fields):
        if len(f) != k:
            raise ValueError(f"csv record {i + 2}: the first column must hold a {k}-mer")
    if not fields:
        return np.empty(0, dtype=np.uint64)
    codes = (np.frombuffer(b"".join(fields), dtype=np.uint8).reshape(len(fields), k).astype(np.uint64) >> np.uint64(1)) & np.uint64(3)
    shifts = (np.uint64(2) * np.arange(k - 1, -1, -1, dtype=np.uint64))
    return np.bitwise_or.reduce(codes << shifts, axis=1)


class KmerSet:
    """src/set.rs:17-21"""

    def get(self, kmer: int) -> bool:
        raise NotImplementedError

    def k(self) -> int:
        raise NotImplementedError


class Pcon(KmerSet):
    """set::Pcon (src/set/pcon.rs:13-196): dense bitfield of 2^(2k-1) canonical k-mers."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle
        ctx._adopt(self)

    # --- constructors ------------------------------------------------------------------------
    @classmethod
    def new(cls, ctx, k):
        """Pcon::new(pcon::solid::Solid::new(k)) — empty set (src/set/pcon.rs:183)."""
        h = C.c_void_p()
        check(lib.brgpu_set_new(ctx._h, k, C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_pcon_solid(cls, ctx, stream):
        """src/set/pcon.rs:18-25: a (gzip) `.solid` stream — byte 0 = k, rest = bitfield.
        Decompression is host I/O (niffler in the reference); the payload goes to the GPU."""
        data = stream if isinstance(stream, (bytes, bytearray)) else stream.read()
        if data[:2] == b"\x1f\x8b":
            data = gzip.decompress(data)
        buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        check(lib.brgpu_set_from_solid_payload(ctx._h, _ptr(buf), buf.size, C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_bitfield(cls, ctx, k, bits):
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        h = C.c_void_p()
        check(lib.brgpu_set_from_bitfield(ctx._h, k, _ptr(b), b.size, C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_reads(cls, ctx, reads, k, abundance=None, abundance_selection=None, percent=0.0):
        """The `fasta` sub-command (src/main.rs:72-115): count, spectrum, threshold, bitfield.
        `reads` is a device-resident Reads or a (seq, offsets) pair of host buffers.
        abundance: -a; abundance_selection: None, "first-minimum", or "rarefaction" /
        "percent-most" / "percent-least" with `percent` (src/cli.rs:227-241)."""
        k = k - (~(k & 1) & 1)  # Fasta::kmer_size forces k odd (src/cli.rs:277-279)
        sel, ab = selection_code(abundance, abundance_selection)
        h = C.c_void_p()
        if isinstance(reads, Reads):
            check(lib.brgpu_set_from_reads_ex(ctx._h, k, ab, sel, float(percent), reads._h, C.byref(h)), ctx._h)
        else:
            s, off = as_u8(reads[0]), as_offsets(reads[1])
            n = (off.numel() if hasattr(off, "numel") else off.size) - 1
            check(lib.brgpu_set_from_host_reads_ex(ctx._h, k, ab, sel, float(percent), _addr(s), _addr(off), n, C.byref(h)),
                  ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_chunks(cls, ctx, chunks, k, abundance=None, abundance_selection=None, percent=0.0):
        """The `fasta` sub-command over a stream of record chunks (count_fasta(inputs, 8192) reads chunk by
        chunk, src/main.rs:72-78): every chunk — a Reads or a (seq, offsets) pair — is partitioned on its
        own and dropped; the set is counted over all partitions at once (brgpu_set_from_kmers).  Same result
        as from_reads over the concatenation; peak device memory is 2 B per k-mer plus one chunk."""
        k = k - (~(k & 1) & 1)
        sel, ab = selection_code(abundance, abundance_selection)
        if k < 15:  # small tables: the literal counter accumulates chunk by chunk
            c = Counter(ctx, k)
            for ch in chunks:
                r = ch if isinstance(ch, Reads) else Reads.upload(ctx, *ch)
                c.count(r)
                if r is not ch:
                    r.free()
            if ab < 0:
                if sel == _lib.ABUNDANCE_EXPLICIT:
                    raise _lib.BrgpuError(_lib.E_NEED_ABUNDANCE)
                a = spectrum_threshold(c.spectrum(), abundance_selection, percent)
                if a is None:
                    raise _lib.BrgpuError(_lib.E_NO_THRESHOLD)
                ab = a
            s = c.to_set(ab)
            c.free()
            return s
        parts = []
        try:
            for ch in chunks:
                r = ch if isinstance(ch, Reads) else Reads.upload(ctx, *ch)
                h = C.c_void_p()
                check(lib.brgpu_kmers_create(ctx._h, k, r._h, C.byref(h)), ctx._h)
                parts.append(h)
                if r is not ch:
                    r.free()
            arr = (C.c_void_p * max(1, len(parts)))(*[p.value for p in parts])
            out = C.c_void_p()
            check(lib.brgpu_set_from_kmers(ctx._h, arr, len(parts), ab, sel, float(percent), C.byref(out)), ctx._h)
        finally:
            for p in parts:
                lib.brgpu_kmers_free(p)
        return cls(ctx, out)

    @classmethod
    def from_fasta(cls, ctx, stream, k):
        """Pcon::from_fasta (src/set/pcon.rs:47-112): presence of every canonical k-mer of every record with
        len >= k — the counting pass with the threshold `count > 0`."""
        from .fasta import read_fasta

        _, seq, off = read_fasta(stream)
        return cls._presence(ctx, seq, off, k)

    @classmethod
    def from_fastq(cls, ctx, stream, k):
        """Pcon::from_fastq (src/set/pcon.rs:114-181, cargo feature `fastq`): from_fasta over FASTQ records."""
        from .fasta import read_fastq

        _, seq, off = read_fastq(stream)
        return cls._presence(ctx, seq, off, k)

    @classmethod
    def _presence(cls, ctx, seq, off, k):
        h = C.c_void_p()
        s, o = as_u8(seq), as_offsets(off)
        check(lib.brgpu_set_from_host_reads(ctx._h, k, 0, _lib.ABUNDANCE_EXPLICIT, _addr(s), _addr(o), o.size - 1, C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_csv(cls, ctx, stream, k):
        """Pcon::from_csv (src/set/pcon.rs:27-45, cargo feature `csv`): Solid::new(k), then
        set.set(seq2bit(record[0]), true) for every data record (the first record is the header)."""
        out = cls.new(ctx, k)
        out.insert(csv_kmers(stream, k))
        return out

    # --- KmerSet --------------------------------------------------------------------------------
    def k(self):
        return lib.brgpu_set_k(self._h)

    def get(self, kmer: int) -> bool:
        """Pcon::get (src/set/pcon.rs:189-191); forward k-mers are canonicalised."""
        return bool(self.get_batch(np.array([kmer], dtype=np.uint64))[0])

    def get_batch(self, kmers):
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(km.size, dtype=np.uint8)
        check(lib.brgpu_set_get_batch(self._h, _ptr(km), km.size, _ptr(out)), self.ctx._h)
        return out

    def insert(self, kmers):
        """Solid::set(kmer, true) for a batch (canonicalises, like the reference's tests rely on)."""
        km = np.ascontiguousarray(np.atleast_1d(kmers), dtype=np.uint64)
        check(lib.brgpu_set_insert_batch(self._h, _ptr(km), km.size), self.ctx._h)

    def insert_all_kmers(self, seq: bytes):
        """`for kmer in Tokenizer::new(seq, k) { data.set(kmer, true) }` of the reference's tests."""
        k = self.k()
        if len(seq) < k:
            return
        mask = (1 << (2 * k)) - 1
        kmers, km = [], seq2bit(seq[: k - 1])
        for b in seq[k - 1 :]:
            km = ((km << 2) & mask) | nuc2bit(b)
            kmers.append(km)
        self.insert(np.array(kmers, dtype=np.uint64))

    # --- extras -----------------------------------------------------------------------------------
    @property
    def abundance(self):
        a = lib.brgpu_set_abundance(self._h)
        return None if a < 0 else a

    def spectrum(self):
        h = np.zeros(256, dtype=np.uint64)
        check(lib.brgpu_set_spectrum(self._h, _ptr(h)), self.ctx._h)
        return h

    def bitfield(self):
        n = lib.brgpu_set_bitfield_bytes(self._h)
        out = np.empty(n, dtype=np.uint8)
        check(lib.brgpu_set_export_bitfield(self._h, _ptr(out), n), self.ctx._h)
        return out

    def to_solid_payload(self) -> bytes:
        """The body of a `.solid` file before gzip: u8 k || bitfield."""
        return bytes([self.k()]) + self.bitfield().tobytes()

    def free(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib.brgpu_set_free(self._h)
        self._h = None

    def __del__(self):
        self.free()


class Hash(Pcon):
    """set::Hash (src/set/hash.rs:14-186): the set behind br's `large-kmer` sub-command — canonical
    k-mers of any k <= 31 in a device hash table.  Same KmerSet surface as Pcon (get, get_batch, k,
    insert, insert_all_kmers) and accepted by every corrector; it has no bitfield or spectrum."""

    @classmethod
    def new(cls, ctx, k, expected_kmers=0):
        h = C.c_void_p()
        check(lib.brgpu_set_hash_new(ctx._h, k, int(expected_kmers), C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def from_reads(cls, ctx, reads, k):
        """Hash::from_fasta (src/set/hash.rs:41-100): presence of every canonical k-mer of every record
        with len >= k.  `reads` is a device-resident Reads or a (seq, offsets) pair of host buffers."""
        h = C.c_void_p()
        if isinstance(reads, Reads):
            check(lib.brgpu_set_hash_from_reads(ctx._h, k, reads._h, C.byref(h)), ctx._h)
        else:
            s, off = as_u8(reads[0]), as_offsets(reads[1])
            n = (off.numel() if hasattr(off, "numel") else off.size) - 1
            check(lib.brgpu_set_hash_from_host_reads(ctx._h, k, _addr(s), _addr(off), n, C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def _presence(cls, ctx, seq, off, k):  # Hash::from_fasta / from_fastq over parsed records
        return cls.from_reads(ctx, (seq, off), k)

    @classmethod
    def from_csv(cls, ctx, stream, k):
        """Hash::from_csv (src/set/hash.rs:20-39): canonical(seq2bit(record[0]), k) of every data record."""
        km = csv_kmers(stream, k)
        out = cls.new(ctx, k, expected_kmers=int(km.size))
        out.insert(km)
        return out

    def add_reads(self, reads: Reads):
        """One more chunk of records (the 8192-record loop of src/set/hash.rs:76-97)."""
        check(lib.brgpu_set_hash_add_reads(self._h, reads._h), self.ctx._h)

    def __len__(self):
        return lib.brgpu_set_hash_size(self._h)

    @property
    def abundance(self):
        return None

    def spectrum(self):
        raise TypeError("set::Hash is presence-only: it has no spectrum")

    def bitfield(self):
        raise TypeError("set::Hash has no bitfield")


class Counter:
    """pcon::counter::Counter<u8> in HBM (src/main.rs:73-78)."""

    def __init__(self, ctx, k):
        self.ctx, self.k = ctx, k
        h = C.c_void_p()
        check(lib.brgpu_counts_create(ctx._h, k, C.byref(h)), ctx._h)
        self._h = h
        ctx._adopt(self)

    def count(self, reads: Reads):
        check(lib.brgpu_counts_add_reads(self._h, reads._h), self.ctx._h)

    def spectrum(self):
        h = np.zeros(256, dtype=np.uint64)
        check(lib.brgpu_counts_spectrum(self._h, _ptr(h)), self.ctx._h)
        return h

    @staticmethod
    def first_minimum(hist):
        h = np.ascontiguousarray(hist, dtype=np.uint64)
        r = lib.brgpu_spectrum_first_minimum(_ptr(h))
        return None if r < 0 else r

    def raw(self):
        n = lib.brgpu_counts_len(self._h)
        out = np.empty(n, dtype=np.uint8)
        check(lib.brgpu_counts_download(self._h, _ptr(out), n), self.ctx._h)
        return out

    def to_set(self, abundance):
        h = C.c_void_p()
        check(lib.brgpu_set_from_counts(self._h, int(abundance), C.byref(h)), self.ctx._h)
        return Pcon(self.ctx, h)

    def free(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib.brgpu_counts_free(self._h)
        self._h = None

    def __del__(self):
        self.free()
