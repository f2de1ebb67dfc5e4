"""Device-side synthetic read generator (brgpu_reads_synth) against its numpy mirror, and the
profiling variant of the scan kernels against the product variant."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import br_b200

    ctx = br_b200.Context(0)
    yield br_b200, ctx
    ctx.close()


def test_device_generator_equals_the_numpy_mirror(gpu):
    """Same bytes from the kernel and from br_b200.synth.host_reads: both strands, templates longer
    than one 2048-position tile, a read-id offset (a rank's shard of a global read list), all three
    error classes, and the degenerate rates 0 and 1."""
    br, ctx = gpu
    from br_b200 import synth

    start, tlen, strand = synth.read_descriptors(300_000, 6, seed=11, mean_len=4000, min_len=50, max_len=20_000)
    assert tlen.max() > 2048 * 2 and strand.min() == 0 and strand.max() == 1
    for error, first in ((0.10, 0), (0.12, 1000), (0.0, 5), (1.0, 7)):
        thr = synth.error_thresholds(error)
        reads = br.Reads.synth(ctx, 42, 43, first, start, tlen, strand, thr)
        got, got_off = reads.download()
        exp, exp_off = synth.host_reads(42, 43, first, start, tlen, strand, thr)
        assert np.array_equal(got_off, exp_off), (error, first)
        assert np.array_equal(got, exp), (error, first)
        if error == 0.0:  # error-free reads are the genome (or its reverse complement)
            g = synth.host_genome(42, 300_000)
            for r in (0, 1, 2):
                t = g[int(start[r]) : int(start[r]) + int(tlen[r])]
                if strand[r]:
                    t = synth._COMP[t[::-1]]
                assert np.array_equal(got[int(got_off[r]) : int(got_off[r + 1])], t)
        reads.free()
    # a shard generated on its own equals the same range of the whole list
    lo, hi = synth.shard_descriptors(tlen, 3, 1)
    thr = synth.error_thresholds(0.1)
    part = br.Reads.synth(ctx, 42, 43, lo, start[lo:hi], tlen[lo:hi], strand[lo:hi], thr)
    pg, po = part.download()
    eg, eo = synth.host_reads(42, 43, 0, start, tlen, strand, thr)
    assert np.array_equal(pg, eg[int(eo[lo]) : int(eo[hi])])
    part.free()
    empty = br.Reads.synth(ctx, 42, 43, 0, start[:0], tlen[:0], strand[:0], thr)
    assert len(empty) == 0
    empty.free()


def test_profiling_variant_gives_the_same_bytes_and_counts_lookups(gpu):
    """The product path runs scan kernels compiled without the KmerSet::get bookkeeping; with
    profiling on, the counting variant runs.  Same output, and the per-kernel table reports lookups for
    the scans (forward and reversed separately) and for the bitmap pass."""
    br, ctx = gpu
    from br_b200 import synth

    start, tlen, strand = synth.read_descriptors(100_000, 20, seed=3, mean_len=3000)
    reads = br.Reads.synth(ctx, 1, 2, 0, start, tlen, strand, synth.error_thresholds(0.08))
    solid = br.Pcon.from_reads(ctx, reads, 15, abundance=2)
    methods = br.build_methods(["one", "two", "graph", "greedy", "gap_size"], solid, 3, 7)
    plain = br.correct_reads(methods, reads)
    a, ao = plain.download()
    ctx.profile_reset()
    ctx.profile_enable(True)
    counted = br.correct_reads(methods, reads)
    prof = ctx.profile()
    ctx.profile_enable(False)
    b, bo = counted.download()
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    n_kmers = int(np.maximum(np.diff(reads.download()[1].astype(np.int64)) - 15 + 1, 0).sum())
    assert prof["solid_bitmap"]["lookups"] >= n_kmers  # first forward launch looks every k-mer up
    for name in ("scan_one", "scan_two", "scan_graph", "scan_greedy", "scan_gap_size"):
        assert prof[name]["lookups"] > 0, name
        assert name + "_rev" in prof and "merge" + name[4:] in prof
    assert ctx.scan_lookups >= sum(p["lookups"] for p in prof.values())
    for h in (plain, counted, solid, reads):
        h.free()


def test_two_bit_transport_round_trip_and_echo_of_original_bytes(gpu, oracle):
    """brgpu_reads_upload_packed / _download_packed: reads that crossed PCIe as 2 bits per base plus an
    exception list are, on the device, byte for byte the reads uploaded as ASCII — lower case, N and
    arbitrary bytes included — and the corrected output comes back through the packed download identical
    to the ASCII download (Corrector::correct echoes original bytes, src/correct/mod.rs:91,100).  Also the
    asynchronous forms, empty reads and a capacity that is too small."""
    br, ctx = gpu
    from br_b200 import synth
    from br_b200.runtime import pack_2bit, unpack_2bit

    rng = np.random.default_rng(5)
    genome = synth.make_genome(30_000, seed=9)
    seq, off, _ = synth.make_reads(genome, 12, 0.08, seed=10, mean_len=900, min_len=1)
    seq = seq.copy()
    weird = rng.choice(seq.size, size=seq.size // 40, replace=False)
    seq[weird] = rng.choice(np.frombuffer(b"acgtNnRY-*\x00\xff", dtype=np.uint8), size=weird.size)
    off = np.concatenate([off[:5], [off[5]], off[5:]]).astype(np.uint64)  # an empty read in the middle
    packed, exc_pos, exc_byte = pack_2bit(seq)
    assert exc_pos.size >= weird.size * 0.5 and np.array_equal(unpack_2bit(packed, seq.size, exc_pos, exc_byte), seq)
    plain = br.Reads.upload(ctx, seq, off)
    via2 = br.Reads.upload_packed(ctx, packed, off, exc_pos, exc_byte)
    a, ao = plain.download()
    b, bo = via2.download()
    assert np.array_equal(ao, bo) and np.array_equal(a, b) and np.array_equal(a, seq)
    # correct both; the packed download of the result equals the ASCII one
    solid = br.Pcon.from_reads(ctx, (genome, np.array([0, genome.size], dtype=np.uint64)), 13, abundance=0)
    methods = br.build_methods(["one", "two", "graph", "greedy", "gap_size"], solid, 3, 7)
    c1 = br.correct_reads(methods, plain)
    c2 = br.correct_reads(methods, via2)
    g, go = c1.download()
    p2, o2, ep, eb, cnt = c2.download_packed(exc_pos=np.empty(exc_pos.size, np.uint64), exc_byte=np.empty(exc_pos.size, np.uint8))
    assert int(cnt[0]) == g.size and np.array_equal(o2, go) and 0 < int(cnt[1]) <= exc_pos.size
    back = unpack_2bit(p2, int(cnt[0]), ep[: int(cnt[1])], eb[: int(cnt[1])])
    assert np.array_equal(back, g)
    assert (g != a[: g.size]).any() if g.size == a.size else True  # the chain edited something
    # oracle on the same input: the transport is invisible
    osolid = oracle.Solid.from_bitfield(13, solid.bitfield())
    exp, exp_off = osolid.run_correction([0, 1, 2, 3, 4], seq, off, confirm=3, max_search=7, threads=8)
    assert np.array_equal(exp_off, go) and np.array_equal(exp, g)
    # asynchronous forms
    import torch

    hp = torch.from_numpy(packed).pin_memory()
    ho = torch.from_numpy(off.view(np.int64)).pin_memory()
    up = br.Reads.upload_packed(ctx, hp, ho, exc_pos, exc_byte, asynchronous=True)
    c3 = br.correct_reads(methods, up)
    bufs = (torch.empty(g.size // 4 + 8, dtype=torch.uint8).pin_memory(), torch.empty(off.size, dtype=torch.int64).pin_memory(),
            torch.empty(exc_pos.size, dtype=torch.int64).pin_memory(), torch.empty(exc_pos.size, dtype=torch.uint8).pin_memory(),
            torch.zeros(2, dtype=torch.int64).pin_memory())
    c3.download_packed(*bufs, asynchronous=True)
    c3.download_wait()
    n_b, n_e = int(bufs[4][0]), int(bufs[4][1])
    assert n_b == g.size and n_e == int(cnt[1])
    back3 = unpack_2bit(bufs[0].numpy(), n_b, bufs[2].numpy()[:n_e].view(np.uint64), bufs[3].numpy()[:n_e])
    assert np.array_equal(back3, g)
    # too small an exception buffer: E_OVERFLOW, the count says how many there are
    with pytest.raises(br.BrgpuError) as ei:
        c2.download_packed(exc_pos=np.empty(1, np.uint64), exc_byte=np.empty(1, np.uint8))
    assert ei.value.status == 5
    for h in (plain, via2, c1, c2, c3, up, solid):
        h.free()
