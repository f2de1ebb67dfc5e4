"""The CPU oracle against every golden vector the reference's own tests hold.

These run without a GPU.  They are what "pins" the oracle:
  * 55 corrector KATs + the alt_nucs KAT + the Pcon set KATs (tests/golden/kats.json,
    extracted from the reference's #[test] functions by tests/golden/make_kats.py);
  * the `.solid` fixture, regenerated from the reads fixture by count -> threshold.
"""
import hashlib

import numpy as np
import pytest

METHOD = {"One": 0, "Two": 1, "Graph": 2, "Greedy": 3, "GapSize": 4}


def build_set(o, t):
    s = o.Solid(t["k"])
    for q in t.get("insert_all_kmers_of", []):
        s.insert_all_kmers(q.encode())
    for q in t.get("insert_kmers", []):
        s.set(o.seq2bit(q.encode()))
    return s


def kat_params(t):
    c = t["corrector"]
    return METHOD[c["method"]], c.get("confirm", c.get("nb_validate", 5)), c.get("max_search", 7)


from kmer_numpy import numpy_canonical_indices  # noqa: E402


def test_kat_inventory(kats):
    per = {}
    for t in kats["correctors"]:
        per[t["module"]] = per.get(t["module"], 0) + 1
    assert per == {"one": 9, "two": 15, "graph": 11, "greedy": 13, "gap_size": 7}
    assert sum(t["ignored_upstream"] for t in kats["correctors"]) == 3


def test_corrector_kats(oracle, kats):
    for t in kats["correctors"]:
        if t["ignored_upstream"]:
            continue
        s = build_set(oracle, t)
        m, conf, ms = kat_params(t)
        for a in t["asserts"]:
            got = s.correct(m, a["input"].encode(), confirm=conf, max_search=ms).decode()
            assert got == a["expected"], (t["module"], t["name"])


def test_upstream_ignored_greedy_kats_fail_like_upstream(oracle, kats):
    """greedy.rs:313,331,348 are #[ignore] upstream because `correct(read) == read` does not
    hold there.  The restatement must reproduce that: Greedy does change those reads (and leaves
    REFE alone).  This is the only result-level evidence on the bio-alignment restatement."""
    for t in kats["correctors"]:
        if not t["ignored_upstream"]:
            continue
        s = build_set(oracle, t)
        m, conf, ms = kat_params(t)
        first, second = t["asserts"]
        assert s.correct(m, first["input"].encode(), confirm=conf, max_search=ms).decode() != first["expected"]
        assert s.correct(m, second["input"].encode(), confirm=conf, max_search=ms).decode() == second["expected"]


def test_alt_nucs_kat(oracle, kats):
    h = kats["helpers"][0]
    s = build_set(oracle, h)
    assert s.alt_nucs(oracle.seq2bit(h["alt_nucs_of"].encode())) == h["expected"]


def test_set_kats(oracle, kats):
    t = kats["set"][0]
    k, seq = t["k"], t["seq"].encode()
    s = oracle.Solid(k)
    fwd = [oracle.seq2bit(seq[i : i + k]) for i in range(len(seq) - k + 1)]
    for km in fwd:
        s.set(oracle.canonical(km, k))
    assert all(s.get(oracle.canonical(km, k)) for km in fwd)  # pcon.rs:205-216
    assert all(s.get(km) for km in fwd)  # pcon.rs:218-230 (get canonicalises)
    assert not s.get(0)  # pcon.rs:232-242
    assert s.k == 11  # pcon.rs:244-254


def test_hash_set_kats(oracle, kats):
    """src/set/hash.rs:185-242: canonical / forward / absence / k through Hash::from_fasta."""
    t = kats["hash_set"][0]
    k, seq = t["k"], t["seq"].encode()
    h = oracle.Hash.from_reads(k, np.frombuffer(seq, dtype=np.uint8), np.array([0, len(seq)], dtype=np.uint64))
    fwd = [oracle.seq2bit(seq[i : i + k]) for i in range(len(seq) - k + 1)]
    assert all(h.get(oracle.canonical(km, k)) for km in fwd)  # hash.rs:192-205
    assert all(h.get(km) for km in fwd)  # hash.rs:207-219
    assert not h.get(0)  # hash.rs:221-230
    assert h.k == 11  # hash.rs:232-241
    assert len(h) == len({oracle.canonical(km, k) for km in fwd})


def test_hash_set_answers_like_the_dense_set(oracle, fixture_reads):
    """set::Hash and set::Pcon are two containers for the same membership function (odd k): built from
    the same reads they answer every get alike, so every corrector does the same on either."""
    seq, off = fixture_reads
    k = 11
    sub = off[:31]
    c = oracle.Counter(k)
    c.count(seq, sub, threads=4)
    dense = c.to_solid(0, threads=4)  # presence: count > 0
    h = oracle.Hash.from_reads(k, seq, sub)
    rng = np.random.default_rng(1)
    km = rng.integers(0, 1 << (2 * k), size=20000, dtype=np.uint64)
    assert np.array_equal(dense.get_batch(km), h.get_batch(km))
    ids = [oracle.ONE, oracle.TWO, oracle.GRAPH, oracle.GREEDY, oracle.GAP_SIZE]
    a, ao = dense.run_correction(ids, seq, off[40:61], confirm=3, threads=4)
    b, bo = h.run_correction(ids, seq, off[40:61], confirm=3, threads=4)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    # short records are skipped (src/set/hash.rs:52), even k works (large-kmer takes k as given)
    tiny = oracle.Hash.from_reads(21, np.frombuffer(b"ACGT" * 5, dtype=np.uint8), np.array([0, 20], dtype=np.uint64))
    assert len(tiny) == 0
    even = oracle.Hash.from_reads(20, np.frombuffer(b"ACGTTGCAAGGCTTAACCGGTTACA", dtype=np.uint8), np.array([0, 25], dtype=np.uint64))
    assert len(even) > 0 and even.get(oracle.seq2bit(b"ACGTTGCAAGGCTTAACCGG"))


def test_kmer_primitives(oracle):
    assert [oracle.lib().bro_nuc2bit(c) for c in b"ACTGactgN"] == [0, 1, 2, 3, 0, 1, 2, 3, 3]
    assert bytes(oracle.lib().bro_bit2nuc(i) for i in range(4)) == b"ACTG"
    assert oracle.seq2bit(b"ACTG") == 0b00011011
    k = 5
    km = oracle.seq2bit(b"ACTGA")
    assert oracle.revcomp(km, k) == oracle.seq2bit(b"TCAGT")
    assert oracle.revcomp(oracle.revcomp(km, k), k) == km
    for x in range(0, 1 << 10, 7):
        c = oracle.canonical(x, k)
        assert bin(c).count("1") % 2 == 0 and c in (x, oracle.revcomp(x, k))


@pytest.mark.parametrize("threads", [1, 4])
def test_count_chain_regenerates_solid_fixture(oracle, fixture_reads, fixture_solid_payload, manifest, threads):
    seq, off = fixture_reads
    assert int(off[-1]) == manifest["br_reads.fa"]["bases"] and off.size - 1 == manifest["br_reads.fa"]["reads"]
    assert hashlib.sha256(fixture_solid_payload).hexdigest() == manifest["br_reads.k11.a2.solid"]["sha256"]
    c = oracle.Counter(11)
    c.count(seq, off, threads=threads)
    # independent numpy recount of every canonical 11-mer (2 517 532 k-mers; a few counters saturate)
    idx = numpy_canonical_indices(seq, off, 11)
    assert idx.size == int(off[-1]) - 206 * 10
    expect = np.minimum(np.bincount(idx, minlength=1 << 21), 255).astype(np.uint8)
    assert np.array_equal(c.raw(), expect)
    solid = c.to_solid(2, threads)
    assert bytes([11]) + solid.bits().tobytes() == fixture_solid_payload
    hist = c.spectrum(threads)
    assert int(hist.sum()) == 1 << 21
    assert list(hist[1:9]) == [442564, 95498, 19526, 4458, 1221, 460, 494, 810]
    assert oracle.Counter.first_minimum(hist) == 6


def test_counter_saturates(oracle):
    seq = np.frombuffer(b"A" * 400, dtype=np.uint8)
    off = np.array([0, 400], dtype=np.uint64)
    c = oracle.Counter(5)
    c.count(seq, off)
    assert int(c.raw()[0]) == 255 and int(c.raw().astype(np.uint64).sum()) == 255


def test_run_correction_order_threads_and_reverse(oracle, fixture_reads, fixture_solid_payload):
    seq, off = fixture_reads
    s = oracle.Solid.from_solid_payload(fixture_solid_payload)
    sub = off[:41]
    d1, o1 = s.run_correction([0, 4], seq, sub, threads=1)
    d4, o4 = s.run_correction([0, 4], seq, sub, threads=4)
    assert np.array_equal(o1, o4) and np.array_equal(d1, d4)
    # two_side=True skips the reversed pass (src/lib.rs:48): equals the single-read fold
    d2, o2 = s.run_correction([0, 4], seq, sub, two_side=True)
    r0 = seq[int(off[0]) : int(off[1])].tobytes()
    fold = s.correct(4, s.correct(0, r0))
    assert d2[int(o2[0]) : int(o2[1])].tobytes() == fold
    rev = s.correct(4, s.correct(0, fold[::-1]))[::-1]
    assert d1[int(o1[0]) : int(o1[1])].tobytes() == rev


def test_short_reads_pass_through(oracle, fixture_solid_payload):
    s = oracle.Solid.from_solid_payload(fixture_solid_payload)
    for m in range(5):
        assert s.correct(m, b"ACGT") == b"ACGT"
        assert s.correct(m, b"") == b""


def test_bio_global_basics(oracle):
    assert oracle.bio_global(b"ACGT", b"ACGT") == "MMMM"
    assert oracle.bio_global(b"ACGT", b"AGGT") == "MXMM"
    # Ins consumes x (first argument), Del consumes y
    assert oracle.bio_global(b"ACGGT", b"ACGT").count("I") == 1
    assert oracle.bio_global(b"ACGT", b"ACGGT").count("D") == 1
    assert oracle.bio_global(b"", b"AC") == "DD"
    assert oracle.bio_global(b"AC", b"") == "II"
