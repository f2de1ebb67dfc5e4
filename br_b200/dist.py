"""Multi-GPU plumbing: one process per GPU, reads sharded by record, one exchange step.

Part 1 shards like this: every rank counts its own reads into a private full-size
table; the index space is cut into `world` slices; rank r merges slice r of all tables with a
saturating add, reading the peers' slices straight over NVLink (CUDA-IPC mapped peer memory, a
hand-written kernel — NCCL has no saturating u8 sum); the 256-bin spectra are all-reduced when the
threshold is data-derived; every rank thresholds its slice and the bitfield slices are
all-gathered with NCCL so each GPU ends up with the full replicated bitfield.  Part 2 needs no
communication: each rank corrects its own reads against its replica.

The protocol is written against a small `ops` object so that the same code runs on the GPUs
(`GpuOps`, C ABI + torch.distributed/NCCL) and in the CPU test-suite (a numpy fake over gloo).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def slice_bounds(n, world, rank):
    """Slice `rank` of [0, n) cut into `world` 1024-aligned pieces (last one takes the rest)."""
    per = (n // world) & ~1023
    if per == 0:
        return (0, n) if rank == 0 else (n, n)
    b = per * rank
    e = n if rank == world - 1 else per * (rank + 1)
    return b, e


def shard_records(offsets, world, rank):
    """Contiguous record range for `rank`, balanced by bases (prefix sum over lengths)."""
    off = np.asarray(offsets, dtype=np.uint64)
    n = off.size - 1
    total = int(off[-1]) - int(off[0])
    lo = int(np.searchsorted(off - off[0], np.uint64(total * rank // world), side="left"))
    hi = int(np.searchsorted(off - off[0], np.uint64(total * (rank + 1) // world), side="left")) if rank < world - 1 else n
    return min(lo, n), min(hi, n)


BUCKET_BITS = 15  # table indices per bucket = 2^15 (must match BUCKET_BITS in csrc/set_kernels.cu)


def bucket_bounds(n_buckets, world, rank):
    """Contiguous bucket range owned by `rank` (last rank takes the remainder)."""
    per = n_buckets // world
    if per == 0:
        return (0, n_buckets) if rank == 0 else (n_buckets, n_buckets)
    return per * rank, (n_buckets if rank == world - 1 else per * (rank + 1))


def _pick_abundance(ops, abundance, abundance_selection, spectrum_of_range, percent=0.0):
    if abundance is not None:
        return int(abundance)
    if abundance_selection is None:
        raise ValueError("need an abundance threshold or an abundance method")
    hist = ops.all_reduce_sum(spectrum_of_range())  # 2 KiB
    if abundance_selection in ("first-minimum", "first_minimum"):
        a = ops.first_minimum(hist)
    else:
        a = ops.spectrum_threshold(hist, abundance_selection, percent)
    if a is None:
        raise RuntimeError("can't compute the abundance threshold")
    return int(a)


def build_set_sharded(ops, k, abundance=None, abundance_selection=None, percent=0.0):
    """Runs the exchange protocol; returns whatever `ops.finish()` returns (the replicated set).

    k >= 15: no count tables at all.  Every rank partitions its own k-mers into buckets of 2^15
    table indices, exports the partition, and counts its own bucket range over all ranks'
    partitions (the peers' residues are read over NVLink inside the counting kernel).
    k < 15: the literal table protocol (private tables, saturating merge of the owned slice)."""
    if k >= 15 and getattr(ops, "supports_kmers", False):
        return _build_from_kmers(ops, k, abundance, abundance_selection, percent)
    return _build_from_tables(ops, k, abundance, abundance_selection, percent)


def _build_from_kmers(ops, k, abundance, abundance_selection, percent=0.0):
    world, rank = ops.world, ops.rank
    n_buckets = 1 << (2 * k - 1 - BUCKET_BITS)
    b0, b1 = bucket_bounds(n_buckets, world, rank)
    ops.partition_local(k)                   # two-level partition of the own shard
    handles = ops.exchange_kmer_handles()    # all-gather of the IPC handles + residue offsets at the rank boundaries
    ops.barrier("partitions complete")       # every partition is complete before anyone reads it
    ops.open_peers(handles)
    abundance = _pick_abundance(ops, abundance, abundance_selection, lambda: ops.count_range(b0, b1, None), percent)
    ops.count_range(b0, b1, abundance)       # count + threshold the owned range (peer data over NVLink)
    bounds = [tuple(x << BUCKET_BITS for x in bucket_bounds(n_buckets, world, r)) for r in range(world)]
    ops.all_gather_bitfield(b0 << BUCKET_BITS, b1 << BUCKET_BITS, 1 << (2 * k - 1), bounds)
    ops.barrier("peers done reading")        # peers are done reading this rank's partition
    return ops.finish(abundance)


def _build_from_tables(ops, k, abundance, abundance_selection, percent=0.0):
    world, rank = ops.world, ops.rank
    n = 1 << (2 * k - 1)
    begin, end = slice_bounds(n, world, rank)
    ops.count_local(k)                       # private table, own shard
    handles = ops.exchange_handles()         # all-gather of the 64-byte IPC handles
    ops.barrier("tables complete")           # every table is complete before anyone reads it
    ops.merge_slice(handles, begin, end)     # saturating reduce of slice `rank` over NVLink
    abundance = _pick_abundance(ops, abundance, abundance_selection, lambda: ops.spectrum_slice(begin, end), percent)
    ops.threshold_slice(abundance, begin, end)
    ops.all_gather_bitfield(begin, end, n, [slice_bounds(n, world, r) for r in range(world)])  # NCCL all-gather
    ops.barrier("peers done reading")        # peers are done reading this rank's table
    return ops.finish(abundance)


class _CudaArray:
    """Zero-copy view of library-owned device memory for torch (__cuda_array_interface__)."""

    def __init__(self, ptr, nbytes, typestr="|u1"):
        n = nbytes // int(typestr[2:])
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class GpuOps:
    """The protocol's steps on a real GPU: C ABI for the kernels, torch.distributed for the
    plumbing (handle exchange, the tiny all-reduce, the bitfield all-gather).

    Ordering without host barriers.  When the context runs on torch's current stream every NCCL
    collective is ordered, on every rank, after the library work enqueued before it.  The handle
    all-gather therefore cannot complete on a rank before every peer's partition kernels have — it is
    the "partitions complete" barrier — and the bitfield all-gather cannot complete before every peer's
    counting kernel has read this rank's residues — it is the "peers done reading" barrier.  So a step
    has no `dist.barrier()` and no stream synchronisation of its own; the one host round trip left is
    reading the gathered handles and offsets.  On any other stream the explicit barriers stay."""

    def __init__(self, ctx, reads, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        # `reads`: this rank's shard as one device-resident chunk or as a list of chunks (a shard larger than
        # 2^32 slot bytes is held in several; every rank must hold the same number of chunks)
        self.chunks = list(reads) if isinstance(reads, (list, tuple)) else [reads]
        self.ctx, self.reads, self.group = ctx, self.chunks[0], group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.counter = None
        self.kmers = None
        self.set = None
        self._peers = []
        self.supports_kmers = True
        self.dev = f"cuda:{ctx.device}"
        self.stream_ordered = ctx.stream_ptr is not None and torch.cuda.current_stream(ctx.device).cuda_stream == ctx.stream_ptr

    # ---- bucketed k-mer protocol (k >= 15) ----
    def partition_local(self, k):
        from .set import Pcon

        self.k = k
        self.parts = []
        for ch in self.chunks:
            h = C.c_void_p()
            check(lib.brgpu_kmers_create(self.ctx._h, k, ch._h, C.byref(h)), self.ctx._h)
            self.parts.append(h)
        self.kmers = self.parts[0]
        s = C.c_void_p()
        check(lib.brgpu_set_new_sliced(self.ctx._h, k, C.byref(s)), self.ctx._h)
        self.set = Pcon(self.ctx, s)

    def exchange_kmer_handles(self):
        """One all-gather of: the two CUDA-IPC handles (residues, bucket offsets) and this rank's residue
        offsets at every rank's bucket boundaries (so that the owner of a bucket range knows which
        contiguous piece of this rank's residues it needs).  The offsets are picked out of the
        library's offset array on the device: no host round trip before the collective."""
        torch = self.torch
        n_buckets = lib.brgpu_kmers_buckets(self.kmers)
        cuts = [bucket_bounds(n_buckets, self.world, r)[0] for r in range(self.world)] + [n_buckets]
        idx = self._cut_index(tuple(cuts))
        per = 16 + len(cuts)  # int64 words per partition: 128 B of handles + the offsets at the cuts
        mine = torch.empty(per * len(self.parts), dtype=torch.int64, device=self.dev)
        for j, part in enumerate(self.parts):
            h = (C.c_uint8 * 128)()
            check(lib.brgpu_kmers_ipc_export(part, h), self.ctx._h)
            base = torch.as_tensor(_CudaArray(lib.brgpu_kmers_offsets_ptr(part), (n_buckets + 1) * 8, "<i8"), device=self.dev)
            mine[j * per : j * per + 16].copy_(torch.frombuffer(bytearray(bytes(h)), dtype=torch.int64), non_blocking=True)
            torch.index_select(base, 0, idx, out=mine[j * per + 16 : (j + 1) * per])
        allh = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=self.dev)
        self.dist.all_gather_into_tensor(allh, mine, group=self.group)  # every rank holds the same number of chunks
        rows = allh.cpu().numpy().reshape(self.world, len(self.parts), per)  # the step's one host round trip
        return [[rows[r, j].tobytes() for j in range(len(self.parts))] for r in range(self.world)]

    def _cut_index(self, cuts):
        cache = self.ctx.__dict__.setdefault("_cut_index_cache", {})
        t = cache.get(cuts)
        if t is None:
            t = self.torch.tensor(list(cuts), dtype=self.torch.int64, device=self.dev)
            cache[cuts] = t
        return t

    def _ipc_open_cached(self, handle: bytes):
        """cudaIpcOpenMemHandle costs milliseconds (and so does closing); the library reuses its big
        buffers from call to call, so the peers' handles repeat — the mappings are kept per context and
        closed by Context.close().  The owner never frees an exported block before its context dies
        (brgpu.cu: pool_flush skips exported blocks), so a cached mapping cannot dangle."""
        cache = self.ctx.__dict__.setdefault("_ipc_cache", {})
        p = cache.get(handle)
        if p is None:
            p = C.c_void_p()
            check(lib.brgpu_ipc_open(self.ctx._h, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(p)), self.ctx._h)
            cache[handle] = p
        return p

    def open_peers(self, handles):
        self._peer_res, self._peer_off, self._peer_first, self._peer_last = [], [], [], []
        for r, parts in enumerate(handles):
            if r == self.rank:
                continue
            for hb in (parts if isinstance(parts, list) else [parts]):  # one entry per chunk of peer r
                self._peer_res.append(self._ipc_open_cached(hb[:64]))
                self._peer_off.append(self._ipc_open_cached(hb[64:128]))
                cuts = np.frombuffer(hb[128:], dtype=np.uint64)  # the partition's residue offsets at the rank boundaries
                self._peer_first.append(int(cuts[self.rank]))
                self._peer_last.append(int(cuts[self.rank + 1]))

    def count_range(self, b0, b1, abundance):
        n = len(self._peer_res)
        res = (C.c_void_p * max(1, n))(*[p.value for p in self._peer_res])
        off = (C.c_void_p * max(1, n))(*[p.value for p in self._peer_off])
        first = (C.c_uint64 * max(1, n))(*self._peer_first)
        last = (C.c_uint64 * max(1, n))(*self._peer_last)
        # the spectrum comes back to the host only when the threshold is derived from it
        hist = np.zeros(256, dtype=np.uint64) if abundance is None else None
        local = (C.c_void_p * len(self.parts))(*[p.value for p in self.parts])
        check(lib.brgpu_kmers_count_parts(self.ctx._h, local, len(self.parts), res, off, first, last, n, b0, b1,
                                          -1 if abundance is None else int(abundance),
                                          None if abundance is None else self.set._h,
                                          None if hist is None else hist.ctypes.data_as(C.c_void_p)), self.ctx._h)
        return hist

    def count_local(self, k):
        from .set import Counter, Pcon

        self.k = k
        self.counter = Counter(self.ctx, k)
        for ch in self.chunks:
            self.counter.count(ch)
        self.set = Pcon.new(self.ctx, k)

    def exchange_handles(self):
        h = (C.c_uint8 * 64)()
        check(lib.brgpu_counts_ipc_export(self.counter._h, h), self.ctx._h)
        mine = self.torch.tensor(list(h), dtype=self.torch.uint8, device=self.dev)
        allh = [self.torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allh, mine, group=self.group)
        return [bytes(t.cpu().tolist()) for t in allh]

    def barrier(self, why=""):
        del why
        if self.stream_ordered:
            return  # the neighbouring collective already orders the ranks (see the class docstring)
        self.ctx.synchronize()
        self.dist.barrier(group=self.group)

    def merge_slice(self, handles, begin, end):
        ptrs = [self._ipc_open_cached(hb) for r, hb in enumerate(handles) if r != self.rank]
        arr = (C.c_void_p * max(1, len(ptrs)))(*[p.value for p in ptrs])
        check(lib.brgpu_counts_merge_slice(self.counter._h, arr, len(ptrs), begin, end), self.ctx._h)

    def spectrum_slice(self, begin, end):
        h = np.zeros(256, dtype=np.uint64)
        check(lib.brgpu_counts_spectrum_slice(self.counter._h, begin, end, h.ctypes.data_as(C.c_void_p)), self.ctx._h)
        return h

    def all_reduce_sum(self, hist):
        t = self.torch.from_numpy(hist.astype(np.int64)).to(self.dev)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint64)

    @staticmethod
    def first_minimum(hist):
        from .set import Counter

        return Counter.first_minimum(hist)

    @staticmethod
    def spectrum_threshold(hist, abundance_selection, percent):
        from .set import spectrum_threshold

        return spectrum_threshold(hist, abundance_selection, percent)

    def threshold_slice(self, abundance, begin, end):
        check(lib.brgpu_set_threshold_slice(self.set._h, self.counter._h, abundance, begin, end), self.ctx._h)

    def all_gather_bitfield(self, begin, end, n_bits, bounds=None):
        """[begin, end) is this rank's bit range; `bounds` lists every rank's range in rank order.  The
        slices of the occupancy summary (one bit per 64 bitfield bits) travel with the bitfield's."""
        torch = self.torch
        n_bytes = lib.brgpu_set_bitfield_bytes(self.set._h)
        ptr = lib.brgpu_set_device_ptr(self.set._h)
        full = torch.as_tensor(_CudaArray(ptr, n_bytes), device=self.dev)
        sb = C.c_uint64()
        sptr = lib.brgpu_set_summary_ptr(self.set._h, C.byref(sb))
        summary = torch.as_tensor(_CudaArray(sptr, sb.value), device=self.dev) if sptr and self.kmers is not None else None
        self._summary_gathered = False
        if self.world == 1:
            self._summary_gathered = summary is not None
            return
        if bounds is None:
            mine = torch.tensor([begin, end], dtype=torch.int64, device=self.dev)
            allb = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(allb, mine, group=self.group)
            bounds = [tuple(int(x) for x in t.cpu().tolist()) for t in allb]
        if not self.stream_ordered:  # the library's stream produced the slice; NCCL runs on torch's current stream
            self.ctx.synchronize()
        if summary is not None and self._exchange_compacted(begin, end, bounds, n_bytes, summary):
            return
        per = (end - begin) // 8
        if all((e - b) // 8 == per for b, e in bounds) and per * self.world == n_bytes:
            # in place: rank r's slice already sits at offset r * per of the output
            self.dist.all_gather_into_tensor(full, full[begin // 8 : end // 8], group=self.group)
            if summary is not None and per % 64 == 0:
                self.dist.all_gather_into_tensor(summary, summary[begin // 512 : end // 512], group=self.group)
                self._summary_gathered = True
        else:  # ragged last slice: broadcast slice by slice
            for r, (b, e) in enumerate(bounds):
                src = self.dist.get_global_rank(self.group, r) if self.group else r
                self.dist.broadcast(full[b // 8 : e // 8], src=src, group=self.group)
                if summary is not None and b % 512 == 0 and e % 512 == 0:
                    self.dist.broadcast(summary[b // 512 : e // 512], src=src, group=self.group)
            self._summary_gathered = summary is not None and all(b % 512 == 0 and e % 512 == 0 for b, e in bounds)

    def _exchange_compacted(self, begin, end, bounds, n_bytes, summary):
        """A sparse set travels in rank-compacted form: every rank compacts its slice (occupied 64-bit blocks in
        index order); ONE small all-gather carries the slice's block count and CUDA-IPC handle (the step's second
        host round trip); if the blocks of all slices together take less than half the bitfield, every rank pulls
        all slices over NVLink into its block array with one kernel (`brgpu_set_compact_pull`: slices are
        contiguous index ranges, so their concatenation in rank order IS the compacted set) and the summary
        slices are all-gathered.  E. coli x 8 at N = 8: 0.3 GB instead of 1 GiB on the wire, no compaction pass
        over the gathered bitfield.  Returns False when the dense exchange should run instead.

        Ordering (context on torch's current stream): the small all-gather completes on a rank only after every
        peer's compaction kernel has (slices complete before anybody pulls); the summary all-gather AFTER the pull
        completes only after every peer's pull has (nobody reuses its slice buffer while a peer still reads it)."""
        torch = self.torch
        if any(b % 2048 or e % 2048 for b, e in bounds) or any(((e - b) // 8) != ((end - begin) // 8) for b, e in bounds):
            return False
        ptr, n_mine = C.c_void_p(), C.c_uint64()
        check(lib.brgpu_set_slice_compact(self.set._h, begin, end, C.byref(ptr), C.byref(n_mine)), self.ctx._h)
        words = [-1] + [0] * 8
        if ptr.value:
            h = (C.c_uint8 * 64)()
            check(lib.brgpu_set_slice_ipc_export(self.set._h, h), self.ctx._h)
            words = [n_mine.value] + np.frombuffer(bytes(h), dtype=np.int64).tolist()
        if not self.stream_ordered:
            self.ctx.synchronize()
        mine = torch.tensor(words, dtype=torch.int64, device=self.dev)
        rows = torch.empty(self.world * 9, dtype=torch.int64, device=self.dev)
        self.dist.all_gather_into_tensor(rows, mine, group=self.group)
        rows = rows.cpu().numpy().reshape(self.world, 9)
        counts = [int(c) for c in rows[:, 0]]
        total = sum(counts)
        if min(counts) < 0 or total * 8 > n_bytes // 2:
            return False  # compaction unavailable somewhere, or the set is too dense to gain from it
        dst = C.c_void_p()
        check(lib.brgpu_set_compact_alloc(self.set._h, total, C.byref(dst)), self.ctx._h)
        slices = [None if r == self.rank or not counts[r] else self._ipc_open_cached(rows[r, 1:].tobytes()).value
                  for r in range(self.world)]
        check(lib.brgpu_set_compact_pull(self.set._h, (C.c_void_p * self.world)(*slices), (C.c_uint64 * self.world)(*counts),
                                         self.world), self.ctx._h)
        if not self.stream_ordered:
            self.ctx.synchronize()
        self.dist.all_gather_into_tensor(summary, summary[begin // 512 : end // 512], group=self.group)
        self._compact_exchanged = True
        return True

    def finish(self, abundance):
        if self.counter is not None:
            self.counter.free()
        if self.kmers is not None:
            for part in self.parts:
                lib.brgpu_kmers_free(part)
            self.parts, self.kmers = [], None
            if getattr(self, "_compact_exchanged", False):  # blocks and summary are whole: only the rank directory is missing
                check(lib.brgpu_set_compact_commit(self.set._h), self.ctx._h)
            else:  # bitfield and summary are whole: build the lookup structures without re-reading the bitfield
                check(lib.brgpu_set_commit_slices(self.set._h, int(getattr(self, "_summary_gathered", False))), self.ctx._h)
        return self.set
