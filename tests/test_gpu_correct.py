"""GPU parity, part 2: the correction pass through the C ABI against the CPU oracle —
the reference's own KATs, every method alone and chained on the reads fixture (config 1 of
BASELINE.json), and randomised small cases that reach the rare scenario paths.  Byte-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

METHODS = ["one", "two", "graph", "greedy", "gap_size"]


@pytest.fixture(scope="module")
def gpu():
    import br_b200

    ctx = br_b200.Context(0)
    yield br_b200, ctx
    ctx.close()


def make_corrector(br, s, c):
    m = c["method"]
    if m == "One":
        return br.One(s, c["confirm"])
    if m == "Two":
        return br.Two(s, c["confirm"])
    if m == "Graph":
        return br.Graph(s)
    if m == "Greedy":
        return br.Greedy(s, c["max_search"], c["nb_validate"])
    return br.GapSize(s, c["confirm"])


def compare_batches(name, got, got_off, exp, exp_off, seq=None, off=None):
    assert got_off.size == exp_off.size, name
    bad = []
    for r in range(exp_off.size - 1):
        g = got[int(got_off[r]) : int(got_off[r + 1])].tobytes()
        e = exp[int(exp_off[r]) : int(exp_off[r + 1])].tobytes()
        if g != e:
            p = next((i for i in range(min(len(g), len(e))) if g[i] != e[i]), min(len(g), len(e)))
            bad.append((r, len(g), len(e), p))
    if bad:
        r, lg, le, p = bad[0]
        g = got[int(got_off[r]) : int(got_off[r + 1])].tobytes()
        e = exp[int(exp_off[r]) : int(exp_off[r + 1])].tobytes()
        msg = f"{name}: {len(bad)} of {exp_off.size - 1} reads differ; first read {r}: len gpu {lg} vs oracle {le}, first diff at {p}\n"
        msg += f"  gpu    ...{g[max(0, p - 30) : p + 30]!r}\n  oracle ...{e[max(0, p - 30) : p + 30]!r}"
        if seq is not None:
            i = seq[int(off[r]) : int(off[r + 1])].tobytes()
            msg += f"\n  input  (len {len(i)}) ...{i[max(0, p - 40) : p + 40]!r}"
        pytest.fail(msg)


def test_reference_kats_on_gpu(gpu, oracle, kats):
    """The 52 active corrector KATs of the reference, through Corrector::correct on the GPU."""
    br, ctx = gpu
    for t in kats["correctors"]:
        s = br.Pcon.new(ctx, t["k"])
        for q in t["insert_all_kmers_of"]:
            s.insert_all_kmers(q.encode())
        if t["insert_kmers"]:
            s.insert(np.array([oracle.seq2bit(q.encode()) for q in t["insert_kmers"]], dtype=np.uint64))
        corr = make_corrector(br, s, t["corrector"])
        assert corr.k() == t["k"]
        for a in t["asserts"]:
            got = corr.correct(a["input"].encode()).decode()
            if t["ignored_upstream"]:
                # #[ignore]d upstream (their assertion does not hold): pin to the oracle instead
                os_ = oracle.Solid.from_bitfield(t["k"], s.bitfield())
                c = t["corrector"]
                exp = os_.correct(3, a["input"].encode(), confirm=c["nb_validate"], max_search=c["max_search"]).decode()
                assert got == exp, (t["module"], t["name"])
            else:
                assert got == a["expected"], (t["module"], t["name"], a["input"], got)


@pytest.fixture(scope="module")
def fixture_sets(gpu, oracle, fixture_solid_payload):
    br, ctx = gpu
    return br.Pcon.from_pcon_solid(ctx, fixture_solid_payload), oracle.Solid.from_solid_payload(fixture_solid_payload)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("two_side", [True, False])
def test_each_method_on_reads_fixture(gpu, oracle, fixture_reads, fixture_sets, method, two_side):
    """Config 1 of BASELINE.json (tests/data reads, k = 11) for every method, forward only and
    with the reversed pass (two_side=False is br's default, src/lib.rs:48)."""
    br, ctx = gpu
    seq, off = fixture_reads
    gs, os_ = fixture_sets
    from oracle.br_oracle import METHOD_IDS

    exp, exp_off = os_.run_correction([METHOD_IDS[method]], seq, off, confirm=5, max_search=7, two_side=two_side, threads=8)
    got, got_off = br.correct_batch(br.build_methods([method], gs, 5, 7), seq, off, two_side=two_side)
    compare_batches(f"{method} two_side={two_side}", got, got_off, exp, exp_off, seq, off)


@pytest.mark.parametrize("chain", [["one", "two"], ["graph", "greedy", "gap_size"], METHODS, ["gap_size", "one", "gap_size"]])
def test_method_chains_on_reads_fixture(gpu, oracle, fixture_reads, fixture_sets, chain):
    br, ctx = gpu
    seq, off = fixture_reads
    gs, os_ = fixture_sets
    from oracle.br_oracle import METHOD_IDS

    exp, exp_off = os_.run_correction([METHOD_IDS[m] for m in chain], seq, off, threads=8)
    got, got_off = br.correct_batch(br.build_methods(chain, gs), seq, off)
    compare_batches("+".join(chain), got, got_off, exp, exp_off, seq, off)


def test_device_resident_path_equals_host_path(gpu, fixture_reads, fixture_sets):
    br, ctx = gpu
    seq, off = fixture_reads
    gs, _ = fixture_sets
    methods = br.build_methods(["one", "gap_size"], gs)
    a, ao = br.correct_batch(methods, seq, off)
    out = br.correct_reads(methods, br.Reads.upload(ctx, seq, off))
    b, bo = out.download()
    assert np.array_equal(ao, bo) and np.array_equal(a, b)


def random_case(rng, k, genome_len, n_reads, error, alphabet=b"ACGT"):
    from br_b200 import synth

    genome = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=genome_len)]
    parts = []
    for _ in range(n_reads):
        L = int(rng.integers(1, min(genome_len, 400)))
        st = int(rng.integers(0, genome_len - L + 1))
        parts.append(synth.mutate(genome[st : st + L], error, rng))
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in parts])
    return genome, np.concatenate(parts), off


@pytest.mark.parametrize("k", [5, 7, 9, 13])
@pytest.mark.parametrize("confirm,max_search", [(1, 0), (2, 3), (5, 7), (3, 12)])
def test_randomised_small_cases(gpu, oracle, k, confirm, max_search):
    """Dense little de Bruijn graphs (small k, high error) reach the branches real data rarely
    takes: ties broken by one_more, DCI's empty emission, cycles in the walks, dirty windows that
    trigger again, read ends inside every scenario's look-ahead."""
    br, ctx = gpu
    from oracle.br_oracle import METHOD_IDS

    rng = np.random.default_rng(1000 * k + 10 * confirm + max_search)
    genome, seq, off = random_case(rng, k, genome_len=int(rng.integers(200, 3000)), n_reads=600, error=0.12)
    gs = br.Pcon.new(ctx, k)
    gs.insert_all_kmers(genome.tobytes())
    os_ = oracle.Solid.from_bitfield(k, gs.bitfield())
    for method in METHODS:
        exp, exp_off = os_.run_correction([METHOD_IDS[method]], seq, off, confirm=confirm, max_search=max_search, two_side=True, threads=8)
        got, got_off = br.correct_batch(br.build_methods([method], gs, confirm, max_search), seq, off, two_side=True)
        compare_batches(f"k={k} c={confirm} M={max_search} {method}", got, got_off, exp, exp_off, seq, off)
    exp, exp_off = os_.run_correction(list(range(5)), seq, off, confirm=confirm, max_search=max_search, threads=8)
    got, got_off = br.correct_batch(br.build_methods(METHODS, gs, confirm, max_search), seq, off)
    compare_batches(f"k={k} c={confirm} M={max_search} chain", got, got_off, exp, exp_off, seq, off)


def test_non_acgt_bytes_echo_and_short_reads(gpu, oracle):
    """Untouched positions echo the original byte, corrected ones are upper-case ACTG; reads
    shorter than k pass through."""
    br, ctx = gpu
    from oracle.br_oracle import METHOD_IDS

    rng = np.random.default_rng(77)
    genome, seq, off = random_case(rng, 7, 1500, 300, 0.08, alphabet=b"ACGTacgtN")
    gs = br.Pcon.new(ctx, 7)
    gs.insert_all_kmers(genome.tobytes())
    os_ = oracle.Solid.from_bitfield(7, gs.bitfield())
    for method in METHODS:
        exp, exp_off = os_.run_correction([METHOD_IDS[method]], seq, off, confirm=2, threads=8)
        got, got_off = br.correct_batch(br.build_methods([method], gs, 2, 7), seq, off)
        compare_batches(f"non-ACGT {method}", got, got_off, exp, exp_off, seq, off)


def test_graph_path_longer_than_the_slot_triggers_reslot(gpu, oracle):
    """A deletion of 300 bases repaired by Graph grows the read far beyond len/8 + 64: the
    library must notice the overflow, re-slot and still return the oracle's bytes."""
    br, ctx = gpu
    rng = np.random.default_rng(9)
    refe = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=800)].tobytes()
    read = refe[:60] + refe[360:440]
    k = 11
    gs = br.Pcon.new(ctx, k)
    gs.insert_all_kmers(refe)
    os_ = oracle.Solid.from_bitfield(k, gs.bitfield())
    exp = os_.correct(2, read)
    assert len(exp) > len(read) + 200  # the walk really is long
    assert br.Graph(gs).correct(read) == exp
    seq = np.frombuffer(read * 3, dtype=np.uint8)
    off = np.array([0, len(read), 2 * len(read), 3 * len(read)], dtype=np.uint64)
    got, got_off = br.correct_batch([br.Graph(gs)], seq, off, two_side=True)
    assert [got[int(got_off[i]) : int(got_off[i + 1])].tobytes() for i in range(3)] == [exp] * 3


def test_asynchronous_correction_matches_the_synchronous_call(gpu, oracle, fixture_sets, fixture_reads):
    """brgpu_correct_reads_async returns once the chain is enqueued; the first consumer waits for it.  Same
    bytes as the synchronous call on the fixture, chained into a second asynchronous correction, and in the
    overflow case (a Graph path that outgrows its slot: the chain is redone with the retry loop)."""
    br, ctx = gpu
    gs, os_ = fixture_sets
    seq, off = fixture_reads
    reads = br.Reads.upload(ctx, seq, off)
    methods = br.build_methods(["one", "two", "gap_size"], gs, 4, 7)
    a, ao = br.correct_reads(methods, reads).download()
    pending = br.correct_reads(methods, reads, asynchronous=True)
    second = br.correct_reads(br.build_methods(["one"], gs, 4, 7), pending, asynchronous=True)  # consumes `pending`
    b, bo = pending.download()
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    second.wait()
    c, co = second.download()
    exp, exp_off = os_.run_correction([oracle.METHOD_IDS[m] for m in ("one", "two", "gap_size")], seq, off, confirm=4, threads=8)
    assert np.array_equal(exp_off, ao) and np.array_equal(exp, a)
    exp2, exp2_off = os_.run_correction([oracle.ONE], exp, exp_off, confirm=4, threads=8)
    assert np.array_equal(exp2_off, co) and np.array_equal(exp2, c)
    # overflow: the long Graph walk of test_graph_path_longer_than_the_slot_triggers_reslot
    rng = np.random.default_rng(9)
    refe = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=800)].tobytes()
    read = refe[:60] + refe[360:440]
    g2 = br.Pcon.new(ctx, 11)
    g2.insert_all_kmers(refe)
    want = oracle.Solid.from_bitfield(11, g2.bitfield()).correct(2, read)
    r2 = br.Reads.upload(ctx, np.frombuffer(read * 2, dtype=np.uint8), np.array([0, len(read), 2 * len(read)], dtype=np.uint64))
    out = br.correct_reads([br.Graph(g2)], r2, two_side=True, asynchronous=True)
    got, got_off = out.download()
    assert [got[int(got_off[i]) : int(got_off[i + 1])].tobytes() for i in range(2)] == [want] * 2
    for h in (reads, pending, second, r2, out, g2):
        h.free()


def test_parameter_validation(gpu, fixture_sets):
    br, ctx = gpu
    gs, _ = fixture_sets
    seq = np.frombuffer(b"ACGTACGTACGTACGT", dtype=np.uint8)
    off = np.array([0, 16], dtype=np.uint64)
    with pytest.raises(br.BrgpuError):
        br.correct_batch([br.One(gs, 0)], seq, off)  # confirm == 0
    with pytest.raises(br.BrgpuError):
        br.Pcon.new(ctx, 12)  # even k
    with pytest.raises(br.BrgpuError):
        br.Pcon.new(ctx, 21)  # dense bitfield does not fit


def test_bio_alignment_on_gpu_matches_oracle_through_greedy(gpu, oracle):
    """Greedy with a long search exercises alignments with more than 32 rows (several rows per
    lane in the systolic sweep)."""
    br, ctx = gpu
    rng = np.random.default_rng(4242)
    genome, seq, off = random_case(rng, 13, 4000, 400, 0.10)
    gs = br.Pcon.new(ctx, 13)
    gs.insert_all_kmers(genome.tobytes())
    os_ = oracle.Solid.from_bitfield(13, gs.bitfield())
    for max_search in (25, 60):
        exp, exp_off = os_.run_correction([3], seq, off, confirm=2, max_search=max_search, two_side=True, threads=8)
        got, got_off = br.correct_batch([br.Greedy(gs, max_search, 2)], seq, off, two_side=True)
        compare_batches(f"greedy M={max_search}", got, got_off, exp, exp_off, seq, off)


def test_segment_scratch_overflow_falls_back_to_the_merge_warp(gpu, oracle):
    """A Graph repair that emits ~4000 bases overflows the 3 KiB scratch region of its segment:
    the speculative piece is marked unusable, the merge warp re-runs the segment itself, the slot
    overflows, the batch is re-slotted — and the bytes still equal the oracle's.  Long reads around
    it exercise multi-segment splicing in the same batch."""
    br, ctx = gpu
    rng = np.random.default_rng(12)
    refe = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=12000)].tobytes()
    k = 15  # 4^13 is small enough that a 4000-step walk through 24 000 solid 13-mers meets a branch
    gs = br.Pcon.new(ctx, k)
    gs.insert_all_kmers(refe)
    os_ = oracle.Solid.from_bitfield(k, gs.bitfield())
    from br_b200 import synth

    reads = [refe[:150] + refe[4150:4400], refe[2000:9000], refe[:12000]]
    reads += [synth.mutate(np.frombuffer(refe[a : a + 7000], dtype=np.uint8), 0.03, rng).tobytes() for a in (0, 2500, 5000)]
    seq = np.frombuffer(b"".join(reads), dtype=np.uint8)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    for method, mid in (("graph", 2), ("gap_size", 4), ("one", 0), ("two", 1), ("greedy", 3)):
        exp, exp_off = os_.run_correction([mid], seq, off, confirm=3, max_search=7, threads=4)
        got, got_off = br.correct_batch(br.build_methods([method], gs, 3, 7), seq, off)
        compare_batches(f"segments {method}", got, got_off, exp, exp_off, seq, off)
    e0 = os_.correct(2, reads[0])
    assert len(e0) > len(reads[0]) + 3500  # the Graph walk really is longer than a scratch region


@pytest.mark.parametrize("k", [15, 17])
def test_large_k_sets_correct_like_the_oracle_on_every_lookup_path(gpu, oracle, k, request):
    """k = 15 / 17 (BASELINE configs 2-5): the set is looked up through the rank-compacted copy
    (directory + one byte per occupied block, the 64-bit block only where it holds several k-mers)
    when it is sparse, through summary + bitfield otherwise.
    Both paths, for a set built by counting and for the same set loaded as a bitfield, must give
    the oracle's bytes for all five methods chained with the reversed pass."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(60_000, seed=5)
    seq, off, _ = synth.make_reads(genome, 25, 0.08, seed=6, mean_len=3000)
    oc = oracle.Counter(k)
    oc.count(seq, off, threads=8)
    osolid = oc.to_solid(2, threads=8)
    ids = [oracle.METHOD_IDS[m] for m in METHODS]
    exp, exp_off = osolid.run_correction(ids, seq, off, confirm=4, max_search=7, threads=8)
    changed = int((np.diff(exp_off.astype(np.int64)) != np.diff(off.astype(np.int64))).sum())
    assert changed > 10  # the chain really edits reads
    request.addfinalizer(lambda: (ctx.set_option("no_compact", 0), ctx.set_option("no_pos8", 0), ctx.set_option("compact_max_pct", 50)))
    # compacted with the one-byte form of the blocks (default) / compacted, 64-bit blocks only / summary + bitfield /
    # treated as too dense to compact: fine (one bit per 16) summary + bitfield
    for no_compact, no_pos8, max_pct in (("0", 0, 50), ("0", 1, 50), ("1", 0, 50), ("0", 0, 0)):
        ctx.set_option("no_compact", int(no_compact))
        ctx.set_option("no_pos8", no_pos8)
        ctx.set_option("compact_max_pct", max_pct)
        counted = br.Pcon.from_reads(ctx, (seq, off), k, abundance=2)
        assert np.array_equal(counted.bitfield(), osolid.bits())
        loaded = br.Pcon.from_bitfield(ctx, k, osolid.bits())
        for name, s in (("counted", counted), ("loaded", loaded)):
            got, got_off = br.correct_batch(br.build_methods(METHODS, s, 4, 7), seq, off)
            compare_batches(f"k={k} {name} no_compact={no_compact} no_pos8={no_pos8} compact_max_pct={max_pct}", got, got_off, exp, exp_off, seq, off)
        counted.free()
        loaded.free()
    # a denser set (17 % of the 64-bit blocks occupied): the rank-compacted copy is built by the
    # streaming kernel instead of the sparse one
    ctx.set_option("no_compact", 0)
    ctx.set_option("no_pos8", 0)
    ctx.set_option("compact_max_pct", 50)
    rng = np.random.default_rng(k)
    dense_bits = osolid.bits().copy()
    pos = rng.integers(0, dense_bits.size * 8, size=int(dense_bits.size * 8 * 0.003), dtype=np.int64)
    np.bitwise_or.at(dense_bits, pos >> 3, (1 << (pos & 7)).astype(np.uint8))
    dsolid = oracle.Solid.from_bitfield(k, dense_bits)
    sub = off[:41]
    exp, exp_off = dsolid.run_correction(ids, seq, sub, confirm=4, max_search=7, threads=8)
    for max_pct in (50, 10):  # 10: the same set counts as too dense to compact (fine summary + bitfield)
        ctx.set_option("compact_max_pct", max_pct)
        dense = br.Pcon.from_bitfield(ctx, k, dense_bits)
        got, got_off = br.correct_batch(br.build_methods(METHODS, dense, 4, 7), seq, sub)
        compare_batches(f"k={k} dense set compact_max_pct={max_pct}", got, got_off, exp, exp_off, seq, sub)
        dense.free()
    # a saturated set (every 64-bit block occupied, one k-mer in eight solid — configs[4]'s regime): looked up in
    # the bitfield directly, the summary would reject nothing
    ctx.set_option("compact_max_pct", 50)
    sat_bits = osolid.bits() | (rng.integers(0, 256, dense_bits.size, dtype=np.uint8) & rng.integers(0, 256, dense_bits.size, dtype=np.uint8)
                                & rng.integers(0, 256, dense_bits.size, dtype=np.uint8))
    ssolid = oracle.Solid.from_bitfield(k, sat_bits)
    sub = off[:21]
    exp, exp_off = ssolid.run_correction(ids, seq, sub, confirm=4, max_search=7, threads=8)
    sat = br.Pcon.from_bitfield(ctx, k, sat_bits)
    got, got_off = br.correct_batch(br.build_methods(METHODS, sat, 4, 7), seq, sub)
    compare_batches(f"k={k} saturated set", got, got_off, exp, exp_off, seq, sub)
    sat.free()


@pytest.mark.parametrize("mode", ["warp", "groups"])
def test_both_scan_kernels_give_the_oracle_bytes(gpu, oracle, fixture_sets, fixture_reads, mode, request):
    """One and Two exist as a warp-per-segment kernel and as a four-segments-per-warp kernel
    (8-lane groups); the library picks per method.  Both must reproduce the oracle on the
    reference's reads (chained, reversed pass on) and on a dense random case."""
    br, ctx = gpu
    ctx.set_option("scan_mode", mode)
    request.addfinalizer(lambda: ctx.set_option("scan_mode", 0))
    gs, os_ = fixture_sets
    seq, off = fixture_reads
    ids = [oracle.METHOD_IDS[m] for m in ("one", "two")]
    exp, exp_off = os_.run_correction(ids, seq, off, confirm=5, max_search=7, threads=8)
    got, got_off = br.correct_batch(br.build_methods(["one", "two"], gs, 5, 7), seq, off)
    compare_batches(f"fixture one+two ({mode})", got, got_off, exp, exp_off, seq, off)
    rng = np.random.default_rng(99)
    genome, rseq, roff = random_case(rng, 9, 3000, 300, 0.12)
    rs = br.Pcon.new(ctx, 9)
    rs.insert_all_kmers(genome.tobytes())
    ro = oracle.Solid.from_bitfield(9, rs.bitfield())
    for c in (2, 7):  # 7: more than 24 scenario items, the unbatched rounds
        exp, exp_off = ro.run_correction(ids, rseq, roff, confirm=c, max_search=7, threads=8)
        got, got_off = br.correct_batch(br.build_methods(["one", "two"], rs, c, 7), rseq, roff)
        compare_batches(f"random one+two c={c} ({mode})", got, got_off, exp, exp_off, rseq, roff)


def test_async_staging_pipeline_matches_the_synchronous_calls(gpu, fixture_sets, fixture_reads):
    """brgpu_reads_upload_async / _download_async / _download_wait: chunk c+1 goes up and chunk c-1
    comes down on the copy stream while chunk c is corrected.  Three rounds over the same buffers
    so that the allocator recycles blocks between the two streams; the bytes must equal what the
    synchronous calls return."""
    import torch

    br, ctx = gpu
    gs, _ = fixture_sets
    seq, off = fixture_reads
    n = off.size - 1
    cuts = [0, 40, 41, 100, 160, n]  # uneven chunks, one of a single read
    methods = br.build_methods(["one", "two", "gap_size"], gs, 4, 7)
    chunks, expected = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        cseq = torch.from_numpy(seq[int(off[a]) : int(off[b])].copy()).pin_memory()
        coff = torch.from_numpy((off[a : b + 1] - off[a]).astype(np.int64)).pin_memory()
        chunks.append((cseq, coff))
        expected.append(br.correct_batch(methods, cseq.numpy(), coff.numpy().astype(np.uint64)))
    cap = int(max(c[0].numel() for c in chunks) * 2 + 4096)
    bufs = [(torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(n + 1, dtype=torch.int64).pin_memory())
            for _ in range(2)]

    def check(i, count, buf):
        exp, exp_off = expected[i]
        m = chunks[i][1].numel()
        got_off = buf[1][:m].numpy().astype(np.uint64)
        assert count == exp.size and np.array_equal(got_off, exp_off), f"chunk {i}: offsets differ"
        assert np.array_equal(buf[0][:count].numpy(), exp), f"chunk {i}: bytes differ"

    for _ in range(3):
        nxt = br.Reads.upload_async(ctx, *chunks[0])
        prev = None  # (index, corrected reads, byte count)
        for i in range(len(chunks)):
            cur = nxt
            if i + 1 < len(chunks):
                nxt = br.Reads.upload_async(ctx, *chunks[i + 1])
            if prev is not None:
                prev = (prev[0], prev[1], prev[1].download_async(*bufs[prev[0] % 2]))
            out = br.correct_reads(methods, cur)
            if prev is not None:
                prev[1].download_wait()
                check(prev[0], prev[2], bufs[prev[0] % 2])
                prev[1].free()
            cur.free()
            prev = (i, out, None)
        count = prev[1].download_async(*bufs[prev[0] % 2])
        prev[1].download_wait()
        check(prev[0], count, bufs[prev[0] % 2])
        prev[1].free()
    # an upload that nobody consumes is still released cleanly
    br.Reads.upload_async(ctx, *chunks[0]).free()


def test_output_capacity_contract(gpu, fixture_sets, fixture_reads):
    """Caller-owned output buffers (include/brgpu.h, ownership row of SURVEY section 8b): too small a
    buffer is BRGPU_E_OVERFLOW with the needed size in *required, nothing is written past the
    capacity, and a second call with that size succeeds."""
    import ctypes as C

    from br_b200 import _lib
    from br_b200.runtime import _ptr

    br, ctx = gpu
    gs, _ = fixture_sets
    seq, off = fixture_reads
    sub = np.ascontiguousarray(off[:21])
    ids = np.array([_lib.ONE, _lib.TWO], dtype=np.uint8)
    exp, exp_off = br.correct_batch(br.build_methods(["one", "two"], gs, 5, 7), seq, sub)
    small = np.full(1000 + 16, 0xEE, dtype=np.uint8)
    out_off = np.zeros(sub.size, dtype=np.uint64)
    req = C.c_uint64()
    st = _lib.lib.brgpu_correct_batch(ctx._h, gs._h, _ptr(ids), ids.size, 5, 7, 0, _ptr(seq), _ptr(sub), sub.size - 1,
                                      _ptr(small), 1000, _ptr(out_off), C.byref(req))
    assert st == _lib.E_OVERFLOW and req.value == exp.size
    assert (small[1000:] == 0xEE).all()
    big = np.empty(req.value, dtype=np.uint8)
    st = _lib.lib.brgpu_correct_batch(ctx._h, gs._h, _ptr(ids), ids.size, 5, 7, 0, _ptr(seq), _ptr(sub), sub.size - 1,
                                      _ptr(big), big.size, _ptr(out_off), C.byref(req))
    assert st == _lib.OK and np.array_equal(big, exp) and np.array_equal(out_off, exp_off)
    # Corrector::correct for one read through brgpu_correct_one
    read = np.ascontiguousarray(seq[int(off[3]) : int(off[4])])
    n = C.c_uint64()
    tiny = np.empty(10, dtype=np.uint8)
    st = _lib.lib.brgpu_correct_one(ctx._h, gs._h, _lib.ONE, 5, 7, _ptr(read), read.size, _ptr(tiny), tiny.size, C.byref(n))
    assert st == _lib.E_OVERFLOW and n.value > 10
