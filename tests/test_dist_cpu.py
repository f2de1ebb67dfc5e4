"""world_size-2 CPU test of the multi-GPU protocol (br_b200/dist.py) over gloo.

The protocol (count own shard -> exchange handles -> saturating merge of the owned slice ->
[all-reduce spectrum] -> threshold slice -> all-gather bitfield) is written against an `ops`
object; here the ops are a numpy fake whose counting uses the CPU oracle, so the host-side
logic (slicing, sharding, ordering, the first-minimum branch) runs without a GPU.  The merged
bitfield must equal what a single process counting all reads produces.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

torch = pytest.importorskip("torch")


def _load_dist_module():
    # br_b200/__init__ loads libbrgpu.so, which is fine on CPU (it only needs a device to compute)
    from br_b200 import dist

    return dist


class NumpyOps:
    def __init__(self, seq, off, rank, world):
        import torch.distributed as tdist

        self.tdist = tdist
        self.seq, self.off, self.rank, self.world = seq, off, rank, world

    def count_local(self, k):
        from oracle import br_oracle as o

        self.k = k
        c = o.Counter(k)
        c.count(self.seq, self.off)
        self.table = c.raw().copy()
        self.bits = np.zeros(self.table.size // 8, dtype=np.uint8)

    def exchange_handles(self):
        mine = torch.from_numpy(self.table.copy())
        allt = [torch.empty_like(mine) for _ in range(self.world)]
        self.tdist.all_gather(allt, mine)
        return [t.numpy() for t in allt]

    def barrier(self, why=""):
        self.tdist.barrier()

    def merge_slice(self, handles, begin, end):
        acc = np.zeros(end - begin, dtype=np.uint32)
        for t in handles:
            acc += t[begin:end]
        self.table[begin:end] = np.minimum(acc, 255).astype(np.uint8)

    def spectrum_slice(self, begin, end):
        return np.bincount(self.table[begin:end], minlength=256).astype(np.uint64)

    def all_reduce_sum(self, hist):
        t = torch.from_numpy(hist.astype(np.int64))
        self.tdist.all_reduce(t)
        return t.numpy().astype(np.uint64)

    @staticmethod
    def first_minimum(hist):
        for i in range(255):
            if hist[i + 1] > hist[i]:
                return i
        return None

    @staticmethod
    def spectrum_threshold(hist, abundance_selection, percent):
        from oracle import br_oracle as o

        return o.Counter.spectrum_threshold(hist, abundance_selection, percent)

    def threshold_slice(self, abundance, begin, end):
        self.bits[begin // 8 : end // 8] = np.packbits(self.table[begin:end] > abundance, bitorder="little")

    def all_gather_bitfield(self, begin, end, n_bits, bounds=None):
        assert bounds is not None and tuple(bounds[self.rank]) == (begin, end)
        assert bounds[0][0] == 0 and bounds[-1][1] == n_bits
        for r, (b, e) in enumerate(bounds):
            t = torch.from_numpy(self.bits[b // 8 : e // 8].copy())
            self.tdist.broadcast(t, src=r)
            self.bits[b // 8 : e // 8] = t.numpy()

    # ---- bucketed k-mer protocol (k >= 15): the fake keeps canonical indices instead of residues ----
    supports_kmers = True

    def partition_local(self, k):
        from kmer_numpy import numpy_canonical_indices

        self.k = k
        self.idx = np.sort(numpy_canonical_indices(self.seq, self.off, k))
        self.bits = np.zeros((1 << (2 * k - 1)) // 8, dtype=np.uint8)

    def exchange_kmer_handles(self):
        n = torch.tensor([self.idx.size], dtype=torch.int64)
        sizes = [torch.empty_like(n) for _ in range(self.world)]
        self.tdist.all_gather(sizes, n)
        m = max(int(x) for x in sizes)
        mine = torch.full((m,), -1, dtype=torch.int64)
        mine[: self.idx.size] = torch.from_numpy(self.idx)
        allt = [torch.empty_like(mine) for _ in range(self.world)]
        self.tdist.all_gather(allt, mine)
        return [t.numpy()[: int(sz)] for t, sz in zip(allt, sizes)]

    def open_peers(self, handles):
        self.all_idx = np.concatenate(handles)

    def count_range(self, b0, b1, abundance):
        dist = _load_dist_module()
        lo, hi = b0 << dist.BUCKET_BITS, b1 << dist.BUCKET_BITS
        sel = self.all_idx[(self.all_idx >= lo) & (self.all_idx < hi)] - lo
        counts = np.minimum(np.bincount(sel, minlength=hi - lo), 255)
        if abundance is not None:
            self.bits[lo // 8 : hi // 8] = np.packbits(counts > abundance, bitorder="little")
        return np.bincount(counts, minlength=256).astype(np.uint64)

    def finish(self, abundance):
        return abundance, self.bits


def _worker(rank, world, port, k, selection, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as tdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dist = _load_dist_module()
        from br_b200 import synth

        genome = synth.make_genome(6000, seed=42)
        seq, off, _ = synth.make_reads(genome, 20, 0.05, seed=43, mean_len=400, min_len=50)
        lo, hi = dist.shard_records(off, world, rank)
        sub = off[lo : hi + 1]
        ops = NumpyOps(seq[int(sub[0]) : int(sub[-1])], sub - sub[0], rank, world)
        if selection == "explicit":
            ab, bits = dist.build_set_sharded(ops, k, abundance=2)
        elif selection == "first-minimum":
            ab, bits = dist.build_set_sharded(ops, k, abundance_selection="first-minimum")
        else:
            ab, bits = dist.build_set_sharded(ops, k, abundance_selection=selection, percent=0.2)
        q.put((rank, lo, hi, ab, bits.tobytes()))
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("k", [9, 15])  # 9: table protocol, 15: bucketed k-mer protocol
@pytest.mark.parametrize("selection", ["explicit", "first-minimum", "percent-least"])
def test_sharded_set_equals_single_process(selection, k):
    import torch.multiprocessing as mp

    from br_b200 import synth
    from oracle import br_oracle as o

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"explicit": 0, "first-minimum": 1, "percent-least": 4}[selection] + (0 if k == 9 else 2)
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, selection, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    genome = synth.make_genome(6000, seed=42)
    seq, off, _ = synth.make_reads(genome, 20, 0.05, seed=43, mean_len=400, min_len=50)
    c = o.Counter(k)
    c.count(seq, off)
    if selection == "explicit":
        ab = 2
    elif selection == "first-minimum":
        ab = o.Counter.first_minimum(c.spectrum())
    else:
        ab = o.Counter.spectrum_threshold(c.spectrum(), selection, 0.2)
    expect = c.to_solid(ab).bits().tobytes()
    results.sort()
    # shards are contiguous, cover every record once, and every rank ends with the full bitfield
    assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == off.size - 1
    for _, _, _, got_ab, bits in results:
        assert got_ab == ab
        assert bits == expect


def test_slice_bounds_and_shards():
    dist = _load_dist_module()
    n = 1 << 33
    for world in (1, 2, 3, 4, 8):
        b = [dist.slice_bounds(n, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(x % 1024 == 0 for s in b for x in s)
    assert dist.slice_bounds(32, 4, 0) == (0, 32) and dist.slice_bounds(32, 4, 3) == (32, 32)  # tiny tables: rank 0 owns all
    for world in (1, 2, 3, 8):
        bb = [dist.bucket_bounds(1 << 18, world, r) for r in range(world)]
        assert bb[0][0] == 0 and bb[-1][1] == 1 << 18 and all(bb[i][1] == bb[i + 1][0] for i in range(world - 1))
    off = np.array([0, 10, 10, 250, 300, 1000, 1001], dtype=np.uint64)
    cuts = [dist.shard_records(off, 3, r) for r in range(3)]
    assert cuts[0][0] == 0 and cuts[-1][1] == off.size - 1
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(2))
