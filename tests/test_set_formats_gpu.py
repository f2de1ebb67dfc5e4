"""GPU parity for the other set inputs (br's cargo features `fastq` / `csv`): set::Pcon::from_fastq / from_csv
(src/set/pcon.rs:27-45,114-181) and set::Hash::from_fastq / from_csv (src/set/hash.rs:20-39,102-175), through the
Python mirror and through `brgpu-cli solid|large-kmer -f fastq|csv` (src/main.rs:117-163).  What pins them: a CSV
listing the k-mers of the reference's `.solid` fixture must give that fixture back byte for byte; a FASTQ of the
reference's reads must give the presence-only set of the same reads as FASTA; corrected reads equal the oracle's."""
import gzip
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, parse_fasta

pytestmark = pytest.mark.gpu

CLI = ROOT / "br_b200" / "brgpu-cli"
LETTERS = np.frombuffer(b"ACTG", dtype=np.uint8)  # bit2nuc


@pytest.fixture(scope="module")
def built():
    import importlib.util

    spec = importlib.util.spec_from_file_location("brgpu_build", ROOT / "br_b200" / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    assert CLI.exists()


@pytest.fixture(scope="module")
def gpu(built):
    import br_b200

    ctx = br_b200.Context(0)
    yield br_b200, ctx
    ctx.close()


def kmer_strings(values, k):
    """kmer2seq for an array of 2-bit packed k-mers -> (n, k) uint8 letters."""
    v = np.asarray(values, dtype=np.uint64)
    shifts = np.uint64(2) * np.arange(k - 1, -1, -1, dtype=np.uint64)
    return LETTERS[((v[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.int64)]


def write_csv(path, values, k, header=b"kmer,count\n", quoted_every=0):
    rows = kmer_strings(values, k)
    with open(path, "wb") as f:
        f.write(header)
        for i, r in enumerate(rows):
            s = r.tobytes()
            f.write((b'"' + s + b'"' if quoted_every and i % quoted_every == 0 else s) + b",3\n")


def forward_kmers(seq, off, k):
    """Tokenizer(seq, k) over every read with len >= k, as packed u64."""
    code = ((seq >> 1) & 3).astype(np.uint64)
    out = []
    for r in range(off.size - 1):
        a, b = int(off[r]), int(off[r + 1])
        if b - a < k:
            continue
        c = code[a:b]
        v = np.zeros(b - a - k + 1, dtype=np.uint64)
        for j in range(k):
            v = (v << np.uint64(2)) | c[j : j + v.size]
        out.append(v)
    return np.concatenate(out) if out else np.empty(0, dtype=np.uint64)


def write_fastq(path, names, seq, off, crlf=False):
    eol = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as f:
        for r, name in enumerate(names):
            s = seq[int(off[r]) : int(off[r + 1])].tobytes()
            f.write(b"@" + name + eol + s + eol + b"+" + eol + b"I" * len(s) + eol)


def fixture_set_kmers(payload):
    """The canonical k-mers of a `.solid` payload: index i stands for the k-mer (i << 1 | parity bit) with an even popcount."""
    bits = np.unpackbits(np.frombuffer(payload[1:], dtype=np.uint8), bitorder="little")
    idx = np.flatnonzero(bits).astype(np.uint64)
    par = np.zeros(idx.size, dtype=np.uint64)
    x = idx.copy()
    while x.any():
        par ^= x & np.uint64(1)
        x >>= np.uint64(1)
    return (idx << np.uint64(1)) | par


def records(path):
    return parse_fasta(open(path, "rb").read())


def run(args):
    return subprocess.run([str(CLI), *map(str, args)], capture_output=True, timeout=600)


def test_pcon_from_csv_rebuilds_the_reference_solid_fixture(gpu, tmp_path, fixture_solid_payload):
    """Pcon::from_csv: Solid::new(k) + set(seq2bit(record[0])) per record.  The CSV lists the 123 072 canonical 11-mers
    of tests/data/raw.k11.a2.solid (half of them as their reverse complement: set() canonicalises), some quoted."""
    br, ctx = gpu
    k = fixture_solid_payload[0]
    cano = fixture_set_kmers(fixture_solid_payload)
    assert k == 11 and cano.size == 123072
    vals = cano.copy()
    rc = np.zeros_like(vals)
    x = vals.copy()
    for _ in range(k):  # reverse complement: complement = code ^ 2
        rc = (rc << np.uint64(2)) | ((x & np.uint64(3)) ^ np.uint64(2))
        x >>= np.uint64(2)
    vals[::2] = rc[::2]
    p = tmp_path / "set.csv"
    write_csv(p, vals, k, quoted_every=7)
    s = br.Pcon.from_csv(ctx, str(p), k)
    assert s.to_solid_payload() == fixture_solid_payload
    s.free()
    gz = tmp_path / "set.csv.gz"
    gz.write_bytes(gzip.compress(p.read_bytes(), 1))
    s = br.Pcon.from_csv(ctx, str(gz), k)
    assert s.to_solid_payload() == fixture_solid_payload
    s.free()


def test_pcon_and_hash_from_fastq_equal_the_fasta_sets(gpu, oracle, tmp_path, fixture_reads):
    br, ctx = gpu
    seq, off = fixture_reads
    names, _, _ = parse_fasta(gzip.open(GOLDEN / "br_reads.fa.gz").read())
    fq = tmp_path / "reads.fq"
    write_fastq(fq, names, seq, off, crlf=True)
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    s = br.Pcon.from_fastq(ctx, str(fq), 11)
    assert np.array_equal(s.bitfield(), c.to_solid(0, 8).bits())
    s.free()
    s = br.Pcon.from_fasta(ctx, str(GOLDEN / "br_reads.fa.gz"), 11)
    assert np.array_equal(s.bitfield(), c.to_solid(0, 8).bits())
    s.free()
    oh = oracle.Hash.from_reads(21, seq, off)
    h = br.Hash.from_fastq(ctx, str(fq), 21)
    assert len(h) == len(oh)
    probe = forward_kmers(seq[: int(off[3])], off[:4], 21)
    rng = np.random.default_rng(1)
    probe = np.concatenate([probe, probe ^ np.uint64(1), rng.integers(0, 1 << 42, size=1000, dtype=np.uint64)])
    assert np.array_equal(h.get_batch(probe), oh.get_batch(probe))
    h.free()


def test_hash_from_csv_matches_the_oracle(gpu, oracle, tmp_path, fixture_reads):
    br, ctx = gpu
    seq, off = fixture_reads
    n = 12
    sub_seq, sub_off = seq[: int(off[n])], off[: n + 1]
    k = 21
    fwd = np.unique(forward_kmers(sub_seq, sub_off, k))  # forward k-mers: from_csv canonicalises (hash.rs:30-33)
    p = tmp_path / "kmers.csv"
    write_csv(p, fwd, k, header=b"kmer,abundance\n")
    oh = oracle.Hash.from_reads(k, sub_seq, sub_off)
    h = br.Hash.from_csv(ctx, str(p), k)
    assert len(h) == len(oh)
    rng = np.random.default_rng(2)
    probe = np.concatenate([fwd[:5000], fwd[:5000] ^ np.uint64(2), rng.integers(0, 1 << 42, size=1000, dtype=np.uint64)])
    assert np.array_equal(h.get_batch(probe), oh.get_batch(probe))
    h.free()


def test_cli_solid_and_large_kmer_take_fastq_and_csv(built, oracle, tmp_path, fixture_reads, fixture_solid_payload):
    """`solid -f csv|fastq -k 11` and `large-kmer -f csv|fastq -k 21` (src/main.rs:117-163): empty stderr like
    tests/br.rs demands, the set as expected, the corrected records equal to the oracle's."""
    seq, off = fixture_reads
    names, _, _ = parse_fasta(gzip.open(GOLDEN / "br_reads.fa.gz").read())
    n = 40
    sub_seq, sub_off, sub_names = seq[: int(off[n])], off[: n + 1], names[:n]
    reads_fa = tmp_path / "reads.fa"
    with open(reads_fa, "wb") as f:
        for r in range(n):
            f.write(b">" + sub_names[r] + b"\n" + sub_seq[int(sub_off[r]) : int(sub_off[r + 1])].tobytes() + b"\n")
    fq = tmp_path / "reads.fq"
    write_fastq(fq, sub_names, sub_seq, sub_off)

    def check(out, solid, methods):
        exp, exp_off = solid.run_correction([oracle.METHOD_IDS[m] for m in methods], sub_seq, sub_off, confirm=5, max_search=7, threads=8)
        n1, s1, o1 = records(out)
        assert n1 == sub_names and np.array_equal(o1, exp_off) and np.array_equal(s1, exp)

    # solid -f csv: the k-mers of the reference's .solid fixture
    csv_path = tmp_path / "solid.csv"
    write_csv(csv_path, fixture_set_kmers(fixture_solid_payload), 11)
    out, solid_out = tmp_path / "csv.fa", tmp_path / "csv.solid"
    r = run(["-i", reads_fa, "-o", out, "-c", "one", "two", "--write-solid", solid_out, "solid", "-i", csv_path, "-f", "csv", "-k", "11"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    assert gzip.open(solid_out).read() == fixture_solid_payload
    check(out, oracle.Solid.from_solid_payload(fixture_solid_payload), ["one", "two"])
    # solid -f fastq: presence-only set of the records
    out, solid_out = tmp_path / "fq.fa", tmp_path / "fq.solid"
    r = run(["-i", reads_fa, "-o", out, "-c", "one", "--write-solid", solid_out, "solid", "-i", fq, "-f", "fastq", "-k", "11"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    c = oracle.Counter(11)
    c.count(sub_seq, sub_off, threads=8)
    presence = c.to_solid(0, 8)
    assert np.array_equal(np.frombuffer(gzip.open(solid_out).read()[1:], dtype=np.uint8), presence.bits())
    check(out, presence, ["one"])
    # large-kmer -f fastq / csv at k = 21
    oh = oracle.Hash.from_reads(21, sub_seq, sub_off)
    kmers_csv = tmp_path / "k21.csv"
    write_csv(kmers_csv, np.unique(forward_kmers(sub_seq, sub_off, 21)), 21)
    for fmt, src in (("fastq", fq), ("csv", kmers_csv)):
        out = tmp_path / f"large_{fmt}.fa"
        r = run(["-i", reads_fa, "-o", out, "-c", "one", "gap-size", "large-kmer", "-i", src, "-f", fmt, "-k", "21"])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        check(out, oh, ["one", "gap_size"])
    # the reference's errors: no -k (Error::SolidRequireKmerSize), a ragged CSV
    r = run(["-i", reads_fa, "-o", tmp_path / "x.fa", "solid", "-i", csv_path, "-f", "csv"])
    assert r.returncode == 1 and b"kmer size" in r.stderr
    bad = tmp_path / "bad.csv"
    bad.write_bytes(b"kmer,count\nACGTACGTACG,3\nACGTACGTACG\n")
    r = run(["-i", reads_fa, "-o", tmp_path / "x.fa", "solid", "-i", bad, "-f", "csv", "-k", "11"])
    assert r.returncode == 1 and b"fields" in r.stderr


def hash_kat_script(kats, corrector_ks):
    """KAT blocks for brgpu-kat with the set held as br::set::Hash: the reference's four hash-set KATs (src/set/hash.rs:185-242:
    canonical and forward k-mers present, get(0) false) and the corrector KATs with k in corrector_ks replayed over a hash set
    (the correctors only see KmerSet::get, so every expectation holds unchanged)."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    lines, n = [], 0
    for h in kats["hash_set"]:
        k, seq = h["k"], h["seq"]
        lines += [f"HASHKAT set::hash {k} One 2 7", f"ALL {seq}"]
        for i in range(len(seq) - k + 1):
            fwd = seq[i : i + k]
            lines += [f"GET {fwd} 1", f"GET {fwd.encode().translate(comp)[::-1].decode()} 1"]
        lines += [f"GET {'A' * k} 0", "END"]
        n += 1
    for c in kats["correctors"]:
        if c["ignored_upstream"] or c["k"] not in corrector_ks:
            continue
        cor = c["corrector"]
        confirm = cor.get("confirm", cor.get("nb_validate", 2))
        lines.append(f"HASHKAT {c['module']}::{c['name']} {c['k']} {cor['method']} {confirm} {cor.get('max_search', 7)}")
        lines += [f"ALL {x}" for x in c["insert_all_kmers_of"]] + [f"KMER {x}" for x in c["insert_kmers"]]
        lines += [f"CASE {a['input']} {a['expected']}" for a in c["asserts"]] + ["END"]
        n += 1
    return "\n".join(lines).encode(), n


KAT = ROOT / "br_b200" / "brgpu-kat"


def test_hash_set_kats_through_the_cpp_interface(built, kats, corrector_ks=(11,)):
    """br::set::Hash through brgpu-kat: the reference's hash-set KATs, and its k = 11 corrector KATs over a hash set
    (every k on the CPU stage, tests/test_host_cli_double_cpu.py)."""
    script, n = hash_kat_script(kats, corrector_ks)
    r = subprocess.run([str(KAT)], input=script, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-3000:] + r.stderr.decode()[-2000:]
    assert n > len(kats["hash_set"]) and f"{n} KATs".encode() in r.stdout and b" 0 failed" in r.stdout

