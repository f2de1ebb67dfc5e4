"""GPU parity, part 1: counting -> spectrum -> threshold -> bitfield and KmerSet::get, through
the C ABI (br_b200 is a ctypes shim over libbrgpu.so), against the CPU oracle and the
reference's fixtures.  Bit-exact: this is integer / index work."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import br_b200

    ctx = br_b200.Context(0)
    yield br_b200, ctx
    ctx.close()


def all_forward_kmers(seq, off, k):
    code = ((seq >> 1) & 3).astype(np.uint64)
    out = []
    for r in range(off.size - 1):
        c = code[int(off[r]) : int(off[r + 1])]
        n = c.size - k + 1
        if n <= 0:
            continue
        f = np.zeros(n, dtype=np.uint64)
        for t in range(k):
            f = (f << np.uint64(2)) | c[t : t + n]
        out.append(f)
    return np.concatenate(out)


def test_set_kats_on_gpu(gpu, oracle, kats):
    br, ctx = gpu
    t = kats["set"][0]
    k, seq = t["k"], t["seq"].encode()
    s = br.Pcon.new(ctx, k)
    fwd = np.array([oracle.seq2bit(seq[i : i + k]) for i in range(len(seq) - k + 1)], dtype=np.uint64)
    cano = np.array([oracle.canonical(int(x), k) for x in fwd], dtype=np.uint64)
    s.insert(cano)  # pcon.rs:205-216
    assert s.get_batch(cano).all()
    assert s.get_batch(fwd).all()  # pcon.rs:218-230: get canonicalises
    assert not s.get(0)  # pcon.rs:232-242
    assert s.k() == 11  # pcon.rs:244-254
    # the same set through the oracle: identical bitfield
    o = oracle.Solid(k)
    for x in cano:
        o.set(int(x))
    assert np.array_equal(s.bitfield(), o.bits())


def test_alt_nucs_kat_via_get(gpu, oracle, kats):
    br, ctx = gpu
    h = kats["helpers"][0]
    s = br.Pcon.new(ctx, h["k"])
    s.insert(np.array([oracle.seq2bit(q.encode()) for q in h["insert_kmers"]], dtype=np.uint64))
    km = oracle.seq2bit(h["alt_nucs_of"].encode())
    mask = (1 << (2 * h["k"])) - 1
    cands = np.array([(((km >> 2) << 2) & mask) ^ a for a in range(4)], dtype=np.uint64)
    assert [a for a in range(4) if s.get_batch(cands)[a]] == h["expected"]


def test_get_batch_matches_oracle_on_fixture_set(gpu, oracle, fixture_reads, fixture_solid_payload):
    br, ctx = gpu
    seq, off = fixture_reads
    s = br.Pcon.from_pcon_solid(ctx, fixture_solid_payload)
    o = oracle.Solid.from_solid_payload(fixture_solid_payload)
    assert s.k() == 11
    assert s.to_solid_payload() == fixture_solid_payload
    rng = np.random.default_rng(1)
    rnd = rng.integers(0, 1 << 22, size=1_000_000, dtype=np.uint64)
    assert np.array_equal(s.get_batch(rnd), o.get_batch(rnd))
    km = all_forward_kmers(seq, off, 11)
    g = s.get_batch(km)
    assert np.array_equal(g, o.get_batch(km))
    assert 0.70 < g.mean() < 0.80  # 74.8 % of the forward k-mers are solid


def test_count_chain_regenerates_solid_fixture(gpu, oracle, fixture_reads, fixture_solid_payload):
    """count -> threshold on the reads fixture must give the reference's .solid fixture bit for bit."""
    br, ctx = gpu
    seq, off = fixture_reads
    reads = br.Reads.upload(ctx, seq, off)
    assert len(reads) == 206 and reads.bases == int(off[-1])
    c = br.Counter(ctx, 11)
    c.count(reads)
    oc = oracle.Counter(11)
    oc.count(seq, off, threads=4)
    assert np.array_equal(c.raw(), oc.raw())  # incl. the saturated counters
    hist = c.spectrum()
    assert np.array_equal(hist, oc.spectrum())
    assert list(hist[1:9]) == [442564, 95498, 19526, 4458, 1221, 460, 494, 810]
    assert br.Counter.first_minimum(hist) == 6
    s = c.to_set(2)
    assert s.to_solid_payload() == fixture_solid_payload
    assert np.array_equal(s.spectrum(), hist)
    # the one-call form (src/main.rs:72-115), explicit -a and first-minimum
    s2 = br.Pcon.from_reads(ctx, (seq, off), 11, abundance=2)
    assert s2.to_solid_payload() == fixture_solid_payload and s2.abundance == 2
    s3 = br.Pcon.from_reads(ctx, reads, 11, abundance_selection="first-minimum")
    assert s3.abundance == 6
    assert np.array_equal(s3.bitfield(), oc.to_solid(6).bits())
    # even k is decremented like Fasta::kmer_size (src/cli.rs:277-279)
    assert br.Pcon.from_reads(ctx, reads, 12, abundance=2).k() == 11
    with pytest.raises(br.BrgpuError) as e:
        br.Pcon.from_reads(ctx, reads, 11)  # neither -a nor a method (src/main.rs:109)
    assert e.value.status == 7


def test_counter_saturates_at_255(gpu, oracle):
    br, ctx = gpu
    seq = np.frombuffer(b"A" * 400 + b"ACGTTGCATGCCGTA" * 40, dtype=np.uint8)
    off = np.array([0, 400, seq.size], dtype=np.uint64)
    for k in (3, 5, 9):
        c = br.Counter(ctx, k)
        c.count(br.Reads.upload(ctx, seq, off))
        oc = oracle.Counter(k)
        oc.count(seq, off)
        assert np.array_equal(c.raw(), oc.raw())
        assert int(c.raw()[0]) == 255
        assert np.array_equal(c.spectrum(), oc.spectrum())


def test_ragged_short_and_empty_reads(gpu, oracle):
    br, ctx = gpu
    rng = np.random.default_rng(5)
    lens = [0, 1, 10, 11, 12, 31, 32, 33, 0, 64, 1000, 5, 0]
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=sum(lens))]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    reads = br.Reads.upload(ctx, seq, off)
    d, o = reads.download()
    assert np.array_equal(o, off) and np.array_equal(d, seq)
    c = br.Counter(ctx, 11)
    c.count(reads)
    oc = oracle.Counter(11)
    oc.count(seq, off)
    assert np.array_equal(c.raw(), oc.raw())
    # no reads at all
    e = br.Reads.upload(ctx, np.empty(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    d, o = e.download()
    assert d.size == 0 and list(o) == [0]


def test_any_byte_is_a_nucleotide(gpu, oracle):
    """nuc2bit is (b >> 1) & 3 for every byte: N -> G, lower case like upper case."""
    br, ctx = gpu
    seq = np.frombuffer(b"ACGTNNacgtnRYKMacgtACGTNNNNACGTACGTAGCTAGCTAGGGATCGATCGNNNN", dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    c = br.Counter(ctx, 7)
    c.count(br.Reads.upload(ctx, seq, off))
    oc = oracle.Counter(7)
    oc.count(seq, off)
    assert np.array_equal(c.raw(), oc.raw())


def test_k17_count_and_threshold_against_oracle(gpu, oracle):
    """BASELINE config-2 shape at reduced size: k = 17 (8 GiB table, 1 GiB bitfield)."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(200_000, seed=42)
    seq, off, _ = synth.make_reads(genome, 20, 0.10, seed=43, mean_len=3000)
    s = br.Pcon.from_reads(ctx, (seq, off), 17, abundance=2)
    oc = oracle.Counter(17)
    oc.count(seq, off, threads=8)
    assert np.array_equal(s.spectrum(), oc.spectrum(threads=8))
    ob = oc.to_solid(2, threads=8).bits()
    gb = s.bitfield()
    assert hashlib.sha256(gb.tobytes()).digest() == hashlib.sha256(ob.tobytes()).digest()
    km = all_forward_kmers(seq[: int(off[20])], off[:21], 17)
    assert np.array_equal(s.get_batch(km), oracle.Solid.from_bitfield(17, ob).get_batch(km))


@pytest.mark.parametrize("one_level", [False, True])
@pytest.mark.parametrize("k", [15, 17])
def test_bucketed_and_table_counting_paths_agree(gpu, oracle, k, one_level, request):
    """Pcon.from_reads takes the bucketed counting path for k = 15/17 (two-level shared-memory
    partition, or the one-level partition with L2 atomics that k = 19 uses), Counter the literal
    table path; all must give the oracle's spectrum and bitfield, including saturated counters
    (poly-A: every k-mer of a tile in one bucket), any-byte nucleotides and the data-derived
    first-minimum threshold."""
    br, ctx = gpu
    ctx.set_option("one_level_partition", int(one_level))
    request.addfinalizer(lambda: ctx.set_option("one_level_partition", 0))
    from br_b200 import synth

    genome = synth.make_genome(150_000, seed=7)
    seq, off, _ = synth.make_reads(genome, 15, 0.08, seed=8, mean_len=2500)
    extra = np.frombuffer(b"A" * 3000 + b"ACGTNNNNacgtnnACGT" * 50, dtype=np.uint8)
    seq = np.concatenate([seq, extra])
    off = np.concatenate([off, [off[-1] + 3000, off[-1] + extra.size]]).astype(np.uint64)
    reads = br.Reads.upload(ctx, seq, off)
    oc = oracle.Counter(k)
    oc.count(seq, off, threads=8)
    ohist = oc.spectrum(threads=8)
    assert int(ohist[255]) >= 1  # the poly-A k-mer saturates
    c = br.Counter(ctx, k)
    c.count(reads)
    assert np.array_equal(c.spectrum(), ohist)
    for kwargs in ({"abundance": 2}, {"abundance": 0}, {"abundance_selection": "first-minimum"}):
        s = br.Pcon.from_reads(ctx, reads, k, **kwargs)  # bucketed
        ab = s.abundance
        assert np.array_equal(s.spectrum(), ohist), kwargs
        if "abundance" in kwargs:
            assert ab == kwargs["abundance"]
        else:
            assert ab == oracle.Counter.first_minimum(ohist)
        t = c.to_set(ab)  # table path
        gb = s.bitfield()
        assert hashlib.sha256(gb.tobytes()).digest() == hashlib.sha256(t.bitfield().tobytes()).digest(), kwargs
        assert hashlib.sha256(gb.tobytes()).digest() == hashlib.sha256(oc.to_solid(ab, threads=8).bits().tobytes()).digest()
        t.free()
        s.free()


@pytest.mark.parametrize("k", [11, 15])
def test_percent_driven_abundance_methods(gpu, oracle, k):
    """`rarefaction P`, `percent-most P`, `percent-least P` (src/cli.rs:227-241, src/main.rs:100-108):
    the threshold comes from the spectrum; table path (k = 11) and bucketed path (k = 15).  The
    pickers themselves are restated from pcon as recalled (parity unpinned)."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(50_000, seed=3)
    seq, off, _ = synth.make_reads(genome, 20, 0.08, seed=4, mean_len=2000)
    oc = oracle.Counter(k)
    oc.count(seq, off, threads=8)
    ohist = oc.spectrum(threads=8)
    for method, percent in (("rarefaction", 0.3), ("percent-most", 0.6), ("percent-least", 0.25), ("percent-least", 0.6)):
        want = oracle.Counter.spectrum_threshold(ohist, method, percent)
        assert want is not None and want < 40, (method, want)
        s = br.Pcon.from_reads(ctx, (seq, off), k, abundance_selection=method, percent=percent)
        assert s.abundance == want, (method, s.abundance, want)
        assert np.array_equal(s.bitfield(), oc.to_solid(want, threads=8).bits())
        s.free()
    with pytest.raises(br.BrgpuError):  # an impossible share has no threshold: Error::ComputeAbundanceThreshold
        br.Pcon.from_reads(ctx, (seq, off), k, abundance_selection="percent-least", percent=2.0)


def test_k19_counting_and_correction_without_the_dense_oracle(gpu):
    """k = 19 is the largest supported k (16 GiB bitfield; the oracle's dense table would need 128 GiB
    of host memory, so it cannot replay this).  Counting (one-level partition: 2^22 buckets) is
    checked against numpy's own canonical-k-mer counts; correction against a known answer by
    construction: isolated substitutions in reads of a random genome, set = all genome k-mers, must
    be repaired back to the genome by `one` (38-bit k-mer arithmetic end to end)."""
    br, ctx = gpu
    from kmer_numpy import numpy_canonical_indices

    from br_b200 import synth

    k = 19
    rng = np.random.default_rng(19)
    genome = synth.make_genome(30_000, seed=19)
    seq, off, _ = synth.make_reads(genome, 12, 0.05, seed=20, mean_len=1500, min_len=200)
    idx = numpy_canonical_indices(seq, off, k)
    uniq, counts = np.unique(idx, return_counts=True)
    s = br.Pcon.from_reads(ctx, (seq, off), k, abundance=2)
    hist = s.spectrum()
    want = np.bincount(np.minimum(counts, 255), minlength=256).astype(np.uint64)
    want[0] = (1 << (2 * k - 1)) - uniq.size
    assert np.array_equal(hist, want)
    # every k-mer of the reads: solid iff its canonical form occurs more than twice; plus random k-mers (absent)
    km = all_forward_kmers(seq[: int(off[40])], off[:41], k)
    kidx = numpy_canonical_indices(seq[: int(off[40])], off[:41], k)
    solid_idx = set(uniq[counts > 2].tolist())
    assert np.array_equal(s.get_batch(km), np.array([i in solid_idx for i in kidx.tolist()], dtype=np.uint8))
    assert not s.get_batch(rng.integers(0, 1 << 38, size=5000, dtype=np.int64).astype(np.uint64)).any()
    s.free()

    g = br.Pcon.new(ctx, k)
    g.insert_all_kmers(genome.tobytes())
    truth, reads = [], []
    for _ in range(150):
        L = int(rng.integers(400, 2500))
        st = int(rng.integers(0, genome.size - L))
        t = genome[st : st + L].copy()
        r = t.copy()
        for p in range(60, L - 60, 97):  # substitutions further apart than 2k
            r[p] = np.frombuffer(b"ACGT", dtype=np.uint8)[(np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), r[p]) + 1 + p % 3) & 3]
        truth.append(t)
        reads.append(r)
    rseq = np.concatenate(reads)
    roff = np.zeros(len(reads) + 1, dtype=np.uint64)
    roff[1:] = np.cumsum([r.size for r in reads])
    got, got_off = br.correct_batch(br.build_methods(["one"], g, 5, 7), rseq, roff)
    assert np.array_equal(got_off, roff)  # substitutions never change a length
    exact = sum(np.array_equal(got[int(roff[i]) : int(roff[i + 1])], truth[i]) for i in range(len(reads)))
    assert exact >= 145, exact  # a repair can be ambiguous by chance; nearly all reads must come back exactly
    same, _ = br.correct_batch(br.build_methods(["one", "two", "graph", "greedy", "gap_size"], g, 5, 7),
                               np.concatenate(truth), roff)
    assert np.array_equal(same, np.concatenate(truth))  # error-free reads are left alone by every method
    g.free()


@pytest.mark.parametrize("k", [13, 15, 17])
def test_set_built_from_a_stream_of_chunks_equals_the_one_shot_set(gpu, oracle, k):
    """brgpu_set_from_kmers (k >= 15: one partition per chunk, counted together) and the accumulating
    counter (k < 15): spectrum, threshold and bitfield equal those of one call over all the reads, for
    an explicit abundance and for first-minimum, with uneven chunks incl. one of a single short read."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(120_000, seed=21)
    seq, off, _ = synth.make_reads(genome, 18, 0.08, seed=22, mean_len=2500)
    seq = np.concatenate([seq, np.frombuffer(b"ACGTA", dtype=np.uint8)])
    off = np.concatenate([off, [off[-1] + 5]]).astype(np.uint64)
    n = off.size - 1
    cuts = [0, n // 7, n // 7 + 1, n // 2, n - 1, n]
    chunks = [(seq, off[a : b + 1]) for a, b in zip(cuts[:-1], cuts[1:])]
    oc = oracle.Counter(k)
    oc.count(seq, off, threads=8)
    ohist = oc.spectrum(threads=8)
    for kwargs in ({"abundance": 2}, {"abundance_selection": "first-minimum"}):
        whole = br.Pcon.from_reads(ctx, (seq, off), k, **kwargs)
        streamed = br.Pcon.from_chunks(ctx, chunks, k, **kwargs)
        assert streamed.abundance == whole.abundance
        if k >= 15:
            assert np.array_equal(streamed.spectrum(), ohist)
        a = streamed.bitfield()
        assert np.array_equal(a, whole.bitfield())
        assert np.array_equal(a, oc.to_solid(whole.abundance, threads=8).bits())
        # the streamed set corrects like the one-shot one (its summary / compacted copy came from the multi-source kernel)
        sub = off[:60]
        g1, o1 = br.correct_batch(br.build_methods(["one", "two"], streamed, 4, 7), seq, sub)
        g2, o2 = br.correct_batch(br.build_methods(["one", "two"], whole, 4, 7), seq, sub)
        assert np.array_equal(o1, o2) and np.array_equal(g1, g2)
        whole.free()
        streamed.free()


@pytest.mark.parametrize("k", [15, 17])
def test_both_shapes_of_the_bucket_counting_kernel_agree(gpu, oracle, k, request):
    """The bucket-counting kernel exists in three shapes (256 / 128 / 64 threads per bucket, holding 1024 /
    768 / 768 k-mers in registers; the poly-A and the tandem-repeat reads put thousands of k-mers into a
    handful of buckets, which overflows all of them).  Same spectrum, same bitfield, equal to the oracle's —
    for an explicit threshold, threshold 0 and first-minimum."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(200_000, seed=31)
    seq, off, _ = synth.make_reads(genome, 12, 0.08, seed=32, mean_len=2500)
    extra = np.frombuffer(b"A" * 4000 + b"ACGGT" * 1500 + b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", dtype=np.uint8)
    seq = np.concatenate([seq, extra])
    off = np.concatenate([off, [off[-1] + 4000, off[-1] + 4000 + 7500, off[-1] + extra.size]]).astype(np.uint64)
    reads = br.Reads.upload(ctx, seq, off)
    oc = oracle.Counter(k)
    oc.count(seq, off, threads=8)
    ohist = oc.spectrum(threads=8)
    request.addfinalizer(lambda: ctx.set_option("count_block_only", 0))
    for kwargs in ({"abundance": 2}, {"abundance": 0}, {"abundance_selection": "first-minimum"}):
        got = {}
        for mode in (0, 3, 2):
            ctx.set_option("count_block_only", mode)
            s = br.Pcon.from_reads(ctx, reads, k, **kwargs)
            got[mode] = (s.abundance, s.spectrum(), s.bitfield())
            s.free()
        assert got[0][0] == got[3][0] == got[2][0]
        assert all(np.array_equal(got[m][1], ohist) for m in (0, 2, 3)), kwargs
        assert np.array_equal(got[0][2], got[3][2]) and np.array_equal(got[0][2], got[2][2]), kwargs
        assert np.array_equal(got[0][2], oc.to_solid(got[0][0], threads=8).bits()), kwargs
    reads.free()
