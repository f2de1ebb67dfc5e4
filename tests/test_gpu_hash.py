"""GPU parity for set::Hash (src/set/hash.rs, br's `large-kmer` sub-command): the reference's four
hash KATs, membership against the oracle's exact set at k = 20 / 21 / 31, and all five correction
methods on a hash set at k = 21 and 31 against the oracle — byte-exact."""
import numpy as np
import pytest

from test_gpu_correct import METHODS, compare_batches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import br_b200

    ctx = br_b200.Context(0)
    yield br_b200, ctx
    ctx.close()


def test_hash_set_kats_on_gpu(gpu, oracle, kats):
    br, ctx = gpu
    t = kats["hash_set"][0]
    k, seq = t["k"], t["seq"].encode()
    off = np.array([0, len(seq)], dtype=np.uint64)
    h = br.Hash.from_reads(ctx, (np.frombuffer(seq, dtype=np.uint8), off), k)
    fwd = np.array([oracle.seq2bit(seq[i : i + k]) for i in range(len(seq) - k + 1)], dtype=np.uint64)
    cano = np.array([oracle.canonical(int(x), k) for x in fwd], dtype=np.uint64)
    assert h.get_batch(cano).all()  # hash.rs:192-205
    assert h.get_batch(fwd).all()  # hash.rs:207-219
    assert not h.get(0)  # hash.rs:221-230
    assert h.k() == 11  # hash.rs:232-241
    assert len(h) == len(set(cano.tolist()))
    with pytest.raises(br.BrgpuError):
        br.Pcon.bitfield(h)  # no bitfield behind a hash set
    # an empty Hash filled through Solid::set-style insertion, growing past its first table
    g = br.Hash.new(ctx, k, expected_kmers=4)
    g.insert(fwd)
    g.insert(fwd)  # duplicates do not count twice
    assert len(g) == len(h) and g.get_batch(cano).all() and not g.get(0)
    h.free()
    g.free()


@pytest.mark.parametrize("k", [20, 21, 31])
def test_hash_membership_matches_the_oracle(gpu, oracle, k):
    """Hash::from_fasta on the GPU and in the oracle: same size, same answers for k-mers of the reads,
    their reverse complements, mutated k-mers and random ones; chunks accumulate; reads shorter than
    k are skipped.  k = 20 is even (large-kmer takes k as given: canonical() is applied literally)."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(40_000, seed=k)
    seq, off, _ = synth.make_reads(genome, 12, 0.06, seed=k + 1, mean_len=1500, min_len=10)
    seq = np.concatenate([seq, np.frombuffer(b"ACGTN" * 3, dtype=np.uint8)])
    off = np.concatenate([off, [off[-1] + 15]]).astype(np.uint64)  # a 15-base read: shorter than k
    oh = oracle.Hash.from_reads(k, seq, off)
    gh = br.Hash.from_reads(ctx, (seq, off), k)
    assert len(gh) == len(oh)
    # the same set built from two chunks
    half = (off.size - 1) // 2
    r1 = br.Reads.upload(ctx, seq, off[: half + 1])
    r2 = br.Reads.upload(ctx, seq, off[half:])
    g2 = br.Hash.new(ctx, k, expected_kmers=1000)
    g2.add_reads(r1)
    g2.add_reads(r2)
    assert len(g2) == len(oh)
    rng = np.random.default_rng(k)
    mask = (1 << (2 * k)) - 1
    km = []
    for r in rng.integers(0, off.size - 2, size=300):
        s = seq[int(off[r]) : int(off[r + 1])].tobytes()
        if len(s) > k:
            i = int(rng.integers(0, len(s) - k))
            km.append(oracle.seq2bit(s[i : i + k]))
    km = np.array(km, dtype=np.uint64)
    probes = np.concatenate([km, np.array([oracle.revcomp(int(x), k) for x in km], dtype=np.uint64),
                             km ^ np.uint64(1), (km >> np.uint64(2)) | (np.uint64(3) << np.uint64(2 * k - 2)),
                             rng.integers(0, mask, size=2000, dtype=np.uint64), np.array([0, mask], dtype=np.uint64)])
    exp = oh.get_batch(probes)
    assert exp[: km.size].all() and not exp[-2000:].all()  # the k-mers of the reads are in (even k: not always their reverse complements)
    assert np.array_equal(gh.get_batch(probes), exp)
    assert np.array_equal(g2.get_batch(probes), exp)
    for h in (gh, g2, r1, r2):
        h.free()


@pytest.mark.parametrize("k", [21, 31])
def test_all_methods_on_a_hash_set_match_the_oracle(gpu, oracle, k):
    """The correctors only see KmerSet::get, so they run unchanged on a hash set: solid k-mers = the
    k-mers of the genome (both strands come with canonical()), reads with 6 % errors, every method
    alone and the chain with the reversed pass."""
    br, ctx = gpu
    from br_b200 import synth

    genome = synth.make_genome(30_000, seed=100 + k)
    seq, off, _ = synth.make_reads(genome, 10, 0.06, seed=200 + k, mean_len=1200, min_len=5)
    goff = np.array([0, genome.size], dtype=np.uint64)
    oh = oracle.Hash.from_reads(k, genome, goff)
    gh = br.Hash.from_reads(ctx, (genome, goff), k)
    assert len(gh) == len(oh)
    for methods in [[m] for m in METHODS] + [METHODS]:
        ids = [oracle.METHOD_IDS[m] for m in methods]
        exp, exp_off = oh.run_correction(ids, seq, off, confirm=3, max_search=7, threads=8)
        got, got_off = br.correct_batch(br.build_methods(methods, gh, 3, 7), seq, off)
        compare_batches(f"hash k={k} {'+'.join(methods)}", got, got_off, exp, exp_off, seq, off)
    changed = int((np.diff(exp_off.astype(np.int64)) != np.diff(off.astype(np.int64))).sum())
    assert changed > 10
    gh.free()
