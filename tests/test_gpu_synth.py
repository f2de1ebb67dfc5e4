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
