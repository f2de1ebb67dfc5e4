"""Context and device-resident record chunks (host side of include/brgpu.h)."""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _addr(buf):
    """Host address of a numpy array or a (pinned) torch CPU tensor."""
    if buf is None:
        return None
    if hasattr(buf, "data_ptr"):
        return C.c_void_p(buf.data_ptr())
    return buf.ctypes.data_as(C.c_void_p)


def as_u8(seq):
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    if hasattr(seq, "data_ptr"):
        return seq
    return np.ascontiguousarray(seq, dtype=np.uint8)


def as_offsets(offsets):
    if hasattr(offsets, "data_ptr"):
        return offsets
    return np.ascontiguousarray(offsets, dtype=np.uint64)


_ACTG = np.frombuffer(b"ACTG", dtype=np.uint8)  # bit2nuc: code 0..3 -> letter


def pack_2bit(seq):
    """2-bit transport form of concatenated sequences (include/brgpu.h, "2-bit transport"): returns
    (packed uint8, exc_pos uint64, exc_byte uint8) — four bases per byte, first base in the two high
    bits, and the exception list of every byte that is not the upper-case letter of its own code."""
    s = np.ascontiguousarray(seq, dtype=np.uint8)
    codes = (s >> 1) & 3
    exc_pos = np.flatnonzero(s != _ACTG[codes]).astype(np.uint64)
    exc_byte = s[exc_pos.astype(np.int64)].copy()
    pad = (-s.size) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)])
    q = codes.reshape(-1, 4)
    packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)
    return packed, exc_pos, exc_byte


def unpack_2bit(packed, n_bases, exc_pos=None, exc_byte=None):
    """Inverse of pack_2bit: ASCII bytes of `n_bases` bases with the exceptions written back."""
    p = np.ascontiguousarray(packed, dtype=np.uint8)[: (int(n_bases) + 3) // 4]
    codes = np.empty((p.size, 4), dtype=np.uint8)
    for j in range(4):
        codes[:, j] = (p >> (6 - 2 * j)) & 3
    out = _ACTG[codes.reshape(-1)[: int(n_bases)]]
    if exc_pos is not None and len(exc_pos):
        out[np.asarray(exc_pos, dtype=np.int64)] = np.asarray(exc_byte, dtype=np.uint8)
    return out


def bind_to_gpu_numa_node(device=0):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host
    buffers it allocates afterwards (first touch) and its PCIe copies stay on that socket.  With
    one process per GPU this keeps 8 ranks from pushing their uploads and downloads across the
    inter-socket link.  Returns the node id, or None when the topology is not visible (containers
    often hide it); never raises."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        finally:
            pynvml.nvmlShutdown()
        if isinstance(bus, bytes):
            bus = bus.decode()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001 - topology probing is best effort by design
        return None


class Context:
    """One per process per GPU.  `stream` may be a torch.cuda.Stream (or a raw cudaStream_t
    integer) so that torch events and the library's kernels share one stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        raw = None
        if stream is not None:
            raw = C.c_void_p(getattr(stream, "cuda_stream", stream))
        check(lib.brgpu_ctx_create(int(device), raw, C.byref(h)))
        self._h = h
        self.stream_ptr = raw.value if raw is not None else None  # the caller's stream (None: the library made its own)
        self.device = int(device)
        self._children = weakref.WeakSet()  # live Reads / Pcon / Counter handles of this context

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        """Frees every handle still alive in this context, then the context itself (the C
        objects keep a raw pointer to their context, so the order matters)."""
        if getattr(self, "_h", None):
            for child in list(self._children):
                child.free()
            for p in self.__dict__.pop("_ipc_cache", {}).values():  # peer mappings kept by br_b200.dist
                lib.brgpu_ipc_close(self._h, p)
            lib.brgpu_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def synchronize(self):
        check(lib.brgpu_ctx_synchronize(self._h), self._h)

    def set_option(self, name, value):
        """Test / A-B switches of include/brgpu.h (brgpu_ctx_set_option): "no_compact", "one_level_partition",
        "scan_mode" (0 default, 1 or "warp", 2 or "groups").  None of them changes a result."""
        if name == "scan_mode" and isinstance(value, str):
            value = {"default": 0, "warp": 1, "groups": 2}[value]
        check(lib.brgpu_ctx_set_option(self._h, name.encode(), int(value)), self._h)

    # --- instrumentation -------------------------------------------------------------------
    def profile_enable(self, on=True):
        check(lib.brgpu_profile_enable(self._h, int(on)), self._h)

    def profile_reset(self):
        check(lib.brgpu_profile_reset(self._h), self._h)

    def profile(self):
        """{kernel name: {"ms", "launches", "algo_bytes", "lookups"}} accumulated since the last reset;
        "lookups" = KmerSet::get calls the kernel issued (scan / merge kernels)."""
        out = {}
        n = lib.brgpu_profile_count(self._h)
        name = C.create_string_buffer(64)
        ms, nb, ln, lk = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
        for i in range(n):
            check(lib.brgpu_profile_get(self._h, i, name, 64, C.byref(ms), C.byref(ln), C.byref(nb)), self._h)
            check(lib.brgpu_profile_get_lookups(self._h, i, C.byref(lk)), self._h)
            out[name.value.decode()] = {"ms": ms.value, "launches": ln.value, "algo_bytes": nb.value, "lookups": lk.value}
        return out

    def probe_random_gather(self, table_bytes):
        """Random 8-byte gathers per second over a table of that size (brgpu_probe_random_gather)."""
        v = C.c_double()
        check(lib.brgpu_probe_random_gather(self._h, int(table_bytes), C.byref(v)), self._h)
        return v.value

    @property
    def launch_count(self):
        return lib.brgpu_launch_count(self._h)

    @property
    def scan_lookups(self):
        """KmerSet::get calls issued by the correction scans while profiling was on (since the last reset)."""
        return lib.brgpu_scan_lookups(self._h)


class Reads:
    """A chunk of records resident in HBM (the 8192-record buffer of src/lib.rs:90, any size)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._h = handle
        ctx._adopt(self)

    @classmethod
    def upload(cls, ctx, seq, offsets):
        s, off = as_u8(seq), as_offsets(offsets)
        n = (off.numel() if hasattr(off, "numel") else off.size) - 1
        h = C.c_void_p()
        check(lib.brgpu_reads_upload(ctx._h, _addr(s), _addr(off), n, C.byref(h)), ctx._h)
        r = cls(ctx, h)
        r._keep = (s, off)
        return r

    @classmethod
    def upload_async(cls, ctx, seq, offsets):
        """Upload on the context's copy stream; returns at once.  Pass pinned buffers and leave them
        alone until a call that takes the reads has returned."""
        s, off = as_u8(seq), as_offsets(offsets)
        n = (off.numel() if hasattr(off, "numel") else off.size) - 1
        h = C.c_void_p()
        check(lib.brgpu_reads_upload_async(ctx._h, _addr(s), _addr(off), n, C.byref(h)), ctx._h)
        r = cls(ctx, h)
        r._keep = (s, off)
        return r

    @classmethod
    def upload_packed(cls, ctx, packed, offsets, exc_pos=None, exc_byte=None, asynchronous=False):
        """2-bit transport upload (brgpu_reads_upload_packed[_async]); offsets are in bases, from 0."""
        off = as_offsets(offsets)
        n = (off.numel() if hasattr(off, "numel") else off.size) - 1
        n_exc = 0 if exc_pos is None else int(exc_pos.numel() if hasattr(exc_pos, "numel") else exc_pos.size)
        h = C.c_void_p()
        f = lib.brgpu_reads_upload_packed_async if asynchronous else lib.brgpu_reads_upload_packed
        check(f(ctx._h, _addr(packed), _addr(off), n, _addr(exc_pos) if n_exc else None, _addr(exc_byte) if n_exc else None,
                n_exc, C.byref(h)), ctx._h)
        r = cls(ctx, h)
        r._keep = (packed, off, exc_pos, exc_byte)
        return r

    def download_packed(self, packed=None, out_offsets=None, exc_pos=None, exc_byte=None, counts=None, asynchronous=False):
        """2-bit transport download: returns (packed, offsets, exc_pos, exc_byte, counts) with counts[0] =
        bases and counts[1] = exceptions (valid after download_wait() when asynchronous).  Buffers are
        allocated when not given (exception capacity: 1/64 of the bases; pass buffers sized to the input's
        exception count to be exact)."""
        n = len(self)
        if out_offsets is None:
            out_offsets = np.empty(n + 1, dtype=np.uint64)
        if packed is None:
            packed = np.empty(self.bases // 4 + 8, dtype=np.uint8)
        if exc_pos is None:
            cap = max(16, (packed.numel() if hasattr(packed, "numel") else packed.size) // 16)
            exc_pos, exc_byte = np.empty(cap, dtype=np.uint64), np.empty(cap, dtype=np.uint8)
        if counts is None:
            counts = np.zeros(2, dtype=np.uint64)
        pcap = packed.numel() if hasattr(packed, "numel") else packed.size
        ecap = exc_pos.numel() if hasattr(exc_pos, "numel") else exc_pos.size
        f = lib.brgpu_reads_download_packed_async if asynchronous else lib.brgpu_reads_download_packed
        check(f(self._h, _addr(packed), pcap, _addr(out_offsets), _addr(exc_pos) if ecap else None,
                _addr(exc_byte) if ecap else None, ecap, _addr(counts)), self.ctx._h)
        self._dl = (packed, out_offsets, exc_pos, exc_byte, counts)
        return packed, out_offsets, exc_pos, exc_byte, counts

    @classmethod
    def synth(cls, ctx, genome_seed, read_seed, first_read_id, start, tlen, strand, thresholds):
        """Synthetic reads generated on the device (brgpu_reads_synth); br_b200.synth.host_reads is the
        numpy mirror that yields the same bytes."""
        start = np.ascontiguousarray(start, dtype=np.uint64)
        tlen = np.ascontiguousarray(tlen, dtype=np.uint32)
        strand = np.ascontiguousarray(strand, dtype=np.uint8)
        thr = np.ascontiguousarray(thresholds, dtype=np.uint32)
        h = C.c_void_p()
        check(lib.brgpu_reads_synth(ctx._h, int(genome_seed), int(read_seed), int(first_read_id), _ptr(start) or None,
                                    _ptr(tlen) or None, _ptr(strand) or None, tlen.size, _ptr(thr), C.byref(h)), ctx._h)
        return cls(ctx, h)

    def download_async(self, out, out_offsets):
        """Enqueue the copy back on the copy stream; returns the byte count.  `download_wait()`
        blocks until `out[:count]` and `out_offsets` are filled."""
        cap = out.numel() if hasattr(out, "numel") else out.size
        req = C.c_uint64()
        check(lib.brgpu_reads_download_async(self._h, _addr(out), cap, _addr(out_offsets), C.byref(req)), self.ctx._h)
        self._dl = (out, out_offsets)
        return int(req.value)

    def wait(self):
        """Completes an asynchronous correction (brgpu_reads_wait): the chain has run, overflow handled."""
        check(lib.brgpu_reads_wait(self._h), self.ctx._h)
        self._keep = None

    def download_wait(self):
        check(lib.brgpu_reads_download_wait(self._h), self.ctx._h)
        self._dl = None

    def free(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib.brgpu_reads_free(self._h)
        self._h = None

    def __del__(self):
        self.free()

    def __len__(self):
        return lib.brgpu_reads_count(self._h)

    @property
    def bases(self):
        return lib.brgpu_reads_bases(self._h)

    def download(self, out=None, out_offsets=None):
        """Returns (uint8 array of concatenated sequences, uint64 offsets)."""
        n = len(self)
        if out_offsets is None:
            out_offsets = np.empty(n + 1, dtype=np.uint64)
        if out is None:
            out = np.empty(max(1, self.bases), dtype=np.uint8)
        cap = out.numel() if hasattr(out, "numel") else out.size
        req = C.c_uint64()
        check(lib.brgpu_reads_download(self._h, _addr(out), cap, _addr(out_offsets), C.byref(req)), self.ctx._h)
        return out[: req.value], out_offsets
