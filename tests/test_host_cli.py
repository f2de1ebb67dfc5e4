"""The C++ host side (br_b200/host/br.hpp: br's own set / corrector / run_correction interface over
the C ABI) and its command line `brgpu-cli` (br's flags, src/cli.rs).

CPU tests: the FASTA reader/writer pair (stands in for noodles-fasta, src/lib.rs:30-31,57-60), the
argument contract and the loud failure without a GPU.  GPU tests mirror the reference's integration
tests (tests/br.rs:8-59: `fasta ... first-minimum` and `solid -f solid` must succeed with an empty
stderr) and go further: the corrected FASTA is compared record by record with the oracle, and the
reference's unit KATs are replayed through the C++ classes.
"""
import gzip
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, parse_fasta

CLI = ROOT / "br_b200" / "brgpu-cli"
KAT = ROOT / "br_b200" / "brgpu-kat"


@pytest.fixture(scope="module", autouse=True)
def built():
    import importlib.util

    spec = importlib.util.spec_from_file_location("brgpu_build", ROOT / "br_b200" / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    assert CLI.exists() and KAT.exists()


def run(args, **kw):
    return subprocess.run([str(CLI), *map(str, args)], capture_output=True, timeout=600, **kw)


def records(path):
    data = open(path, "rb").read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    return parse_fasta(data)


# ---------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------
def test_echo_round_trips_the_reference_fixture(tmp_path):
    out = tmp_path / "echo.fa"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "echo"])
    assert r.returncode == 0 and r.stderr == b"" and r.stdout == b""
    n0, s0, o0 = records(GOLDEN / "br_reads.fa.gz")
    n1, s1, o1 = records(out)
    assert n0 == n1 and np.array_equal(s0, s1) and np.array_equal(o0, o1)
    lines = [l for l in open(out, "rb").read().split(b"\n") if l and not l.startswith(b">")]
    assert max(map(len, lines)) == 80  # noodles' default line width


def test_reader_handles_wrapped_crlf_empty_and_unterminated_records(tmp_path):
    src = tmp_path / "in.fa"
    src.write_bytes(b"\n>r1 desc with spaces\r\nACGT\r\nAC\r\n>empty\n>r3\nNNNNacgt\n\nTT\n>last\nGG")
    out = tmp_path / "out.fa"
    r = run(["-i", src, "-o", out, "echo"])
    assert r.returncode == 0 and r.stderr == b""
    assert out.read_bytes() == b">r1 desc with spaces\nACGTAC\n>empty\n>r3\nNNNNacgtTT\n>last\nGG\n"


def test_reader_on_random_framings_and_buffer_boundaries(tmp_path):
    """Random line widths, LF / CRLF, blank lines, records of 0..5000 bases, and a CRLF pair that
    straddles the reader's 4 MiB buffer: the C++ reader must see the same records as the Python
    parser, and the writer must re-wrap them at 80 columns."""
    rng = np.random.default_rng(5)
    recs, blob = [], bytearray()
    for i in range(300):
        n = int(rng.integers(0, 5000))
        seq = bytes(np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=n)])
        recs.append((b"r%d some description" % i, seq))
        eol = b"\r\n" if i % 3 == 0 else b"\n"
        blob += b">" + recs[-1][0] + eol
        w = int(rng.integers(1, 200))
        for p in range(0, n, w):
            blob += seq[p : p + w] + eol
        if i % 7 == 0:
            blob += b"\n"
    # one long CRLF-terminated line whose "\r" is the last byte of the first 4 MiB and "\n" the first of the next
    pad = (1 << 22) - 1 - (len(blob) + len(b">edge\r\n")) % (1 << 22)
    big = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=pad)])
    recs.append((b"edge", big + b"GATTACA"))
    blob += b">edge\r\n" + big
    cr = len(blob)
    blob += b"\r\n" + b"GATTACA\r\n"
    assert (cr + 1) % (1 << 22) == 0 and blob[cr : cr + 2] == b"\r\n"
    src, out = tmp_path / "in.fa", tmp_path / "out.fa"
    src.write_bytes(bytes(blob))
    r = run(["-i", src, "-o", out, "echo"])
    assert r.returncode == 0 and r.stderr == b""
    expect = b"".join(b">" + d + b"\n" + b"".join(s[p : p + 80] + b"\n" for p in range(0, len(s), 80)) for d, s in recs)
    assert out.read_bytes() == expect


def test_echo_reads_stdin_and_writes_stdout():
    r = run(["echo"], input=b">a\nAC\nGT\n")
    assert r.returncode == 0 and r.stdout == b">a\nACGT\n" and r.stderr == b""


def test_inputs_and_outputs_are_zipped_pairwise(tmp_path):
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_bytes(b">a\nAAAA\n")
    b.write_bytes(b">b\nCCCC\n")
    oa, ob = tmp_path / "oa.fa", tmp_path / "ob.fa"
    r = run(["-i", a, b, "-o", oa, ob, "echo"])  # src/lib.rs:79: inputs.zip(outputs)
    assert r.returncode == 0
    assert oa.read_bytes() == b">a\nAAAA\n" and ob.read_bytes() == b">b\nCCCC\n"
    r = run(["-i", a, "-i", b, "-o", oa, "-o", ob, "echo"])
    assert r.returncode == 0 and ob.read_bytes() == b">b\nCCCC\n"


def test_help_and_version_need_no_device():
    r = run(["--help"])
    assert r.returncode == 0 and r.stderr == b"" and b"large-kmer" in r.stdout and b"--two-side" in r.stdout
    r = run(["-V"])
    assert r.returncode == 0 and r.stdout.startswith(b"brgpu-cli (brgpu")


def test_argument_contract():
    assert run([]).returncode == 2  # a sub-command is required
    assert run(["-c", "three", "echo"]).returncode == 2
    assert run(["--bogus", "echo"]).returncode == 2
    assert run(["-C", "300", "echo"]).returncode == 2  # confirm is a u8


@pytest.mark.skipif("__import__('torch').cuda.is_available()")
def test_without_a_gpu_the_cli_fails_loudly(tmp_path):
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "o.fa", "solid", "-i", GOLDEN / "br_reads.k11.a2.solid",
             "-f", "solid"])
    assert r.returncode == 1 and b"no CUDA device" in r.stderr
    assert not (tmp_path / "o.fa").exists() or (tmp_path / "o.fa").stat().st_size == 0


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
def oracle_corrected(oracle, payload, methods, seq, off, confirm=5, max_search=7, two_side=False):
    solid = oracle.Solid.from_solid_payload(payload)
    ids = [oracle.METHOD_IDS[m] for m in methods]
    return solid.run_correction(ids, seq, off, confirm=confirm, max_search=max_search, two_side=two_side, threads=8)


def assert_same_records(path, names, exp, exp_off):
    n1, s1, o1 = records(path)
    assert n1 == names
    assert np.array_equal(o1, exp_off), "corrected lengths differ from the oracle"
    assert np.array_equal(s1, exp), "corrected bases differ from the oracle"


@pytest.mark.gpu
def test_solid_subcommand_like_tests_br_rs(tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """tests/br.rs:35-59 (`solid -i raw.k11.a2.solid -f solid`), default method chain and flags."""
    out = tmp_path / "corr.fasta"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-t", "4", "solid", "-i", GOLDEN / "br_reads.k11.a2.solid",
             "-f", "solid"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    seq, off = fixture_reads
    exp, exp_off = oracle_corrected(oracle, fixture_solid_payload, ["one", "two", "graph", "greedy", "gap_size"], seq, off)
    assert_same_records(out, names, exp, exp_off)


@pytest.mark.gpu
def test_fasta_subcommand_config1(tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """BASELINE.json configs[0]: the reference's reads, k = 11, method one — `fasta -k 11 -a 2` must
    rebuild the `.solid` fixture (checked through --write-solid) and correct like the oracle."""
    out, solid_out = tmp_path / "corr.fasta", tmp_path / "set.solid"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-c", "one", "--write-solid", solid_out, "fasta", "-i",
             GOLDEN / "br_reads.fa.gz", "-k", "11", "-a", "2"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    assert gzip.open(solid_out).read() == fixture_solid_payload
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    seq, off = fixture_reads
    exp, exp_off = oracle_corrected(oracle, fixture_solid_payload, ["one"], seq, off)
    assert_same_records(out, names, exp, exp_off)


@pytest.mark.gpu
def test_fasta_first_minimum_two_side_and_even_k(tmp_path, oracle, fixture_reads):
    """tests/br.rs:8-33 (`fasta -k 11 first-minimum`), here with -k 12 (decremented to 11,
    src/cli.rs:277-279), -s (no reversed pass) and -C 3."""
    out = tmp_path / "corr.fasta"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-s", "-c", "two", "gap-size", "-C", "3", "fasta", "-i",
             GOLDEN / "br_reads.fa.gz", "-k", "12", "first-minimum"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    seq, off = fixture_reads
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    thr = oracle.Counter.first_minimum(c.spectrum(8))
    solid = c.to_solid(thr, 8)
    exp, exp_off = solid.run_correction([oracle.METHOD_IDS["two"], oracle.METHOD_IDS["gap_size"]], seq, off, confirm=3,
                                        max_search=7, two_side=True, threads=8)
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    assert_same_records(out, names, exp, exp_off)


@pytest.mark.gpu
def test_missing_abundance_is_the_reference_error(tmp_path):
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "o.fa", "fasta", "-i", GOLDEN / "br_reads.fa.gz", "-k", "11"])
    assert r.returncode == 1 and b"abundance" in r.stderr  # Error::AbundanceThresholdOrAbundanceMethod


@pytest.mark.gpu
def test_solid_from_fasta_is_presence_only(tmp_path, oracle, fixture_reads):
    """`solid -f fasta -k 11` = set::Pcon::from_fasta (src/set/pcon.rs:47-112): every canonical
    k-mer of the file is in the set."""
    solid_out = tmp_path / "presence.solid"
    small = tmp_path / "small.fa"
    small.write_bytes(b">x\nACGTTGCA\n")
    r = run(["-i", small, "-o", tmp_path / "o.fa", "-c", "one", "--write-solid", solid_out, "solid", "-i",
             GOLDEN / "br_reads.fa.gz", "-f", "fasta", "-k", "11"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    seq, off = fixture_reads
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    payload = gzip.open(solid_out).read()
    assert payload[0] == 11
    assert np.array_equal(np.frombuffer(payload[1:], dtype=np.uint8), c.to_solid(0, 8).bits())


@pytest.mark.gpu
def test_reference_unit_kats_through_the_cpp_interface(kats):
    """The reference's #[test] vectors (tests/golden/kats.json) replayed through br::set::Pcon and
    br::correct::{One,Two,Graph,Greedy,GapSize} of br.hpp."""
    lines = []
    n = 0
    for c in kats["correctors"]:
        if c["ignored_upstream"]:
            continue
        cor = c["corrector"]
        confirm = cor.get("confirm", cor.get("nb_validate", 2))
        lines.append(f"KAT {c['module']}::{c['name']} {c['k']} {cor['method']} {confirm} {cor.get('max_search', 7)}")
        lines += [f"ALL {s}" for s in c["insert_all_kmers_of"]] + [f"KMER {s}" for s in c["insert_kmers"]]
        lines += [f"CASE {a['input']} {a['expected']}" for a in c["asserts"]]
        lines.append("END")
        n += 1
    for s in kats["set"]:  # src/set/pcon.rs:198-255: forward and canonical k-mers are present, get(0) is false
        k, seq = s["k"], s["seq"]
        lines.append(f"KAT set::pcon {k} One 2 7")
        lines.append(f"ALL {seq}")
        lines += [f"GET {seq[i:i + k]} 1" for i in range(len(seq) - k + 1)]
        lines.append(f"GET {'A' * k} 0")
        lines.append("END")
    r = subprocess.run([str(KAT)], input="\n".join(lines).encode(), capture_output=True, timeout=600)
    sys.stdout.write(r.stdout.decode())
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    assert f"{n + len(kats['set'])} KATs".encode() in r.stdout


@pytest.mark.gpu
def test_fasta_percent_least_through_the_cli(tmp_path, oracle, fixture_reads):
    """`fasta -k 11 percent-least 0.3`: the abundance sub-sub-command with its percent argument."""
    out, solid_out = tmp_path / "corr.fasta", tmp_path / "set.solid"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-c", "one", "--write-solid", solid_out, "fasta", "-i",
             GOLDEN / "br_reads.fa.gz", "-k", "11", "percent-least", "0.3"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    seq, off = fixture_reads
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    thr = oracle.Counter.spectrum_threshold(c.spectrum(8), "percent-least", 0.3)
    payload = gzip.open(solid_out).read()
    assert np.array_equal(np.frombuffer(payload[1:], dtype=np.uint8), c.to_solid(thr, 8).bits())


@pytest.mark.gpu
def test_large_kmer_subcommand_like_tests_br_rs(tmp_path, oracle, fixture_reads):
    """tests/br.rs:61-87 (`large-kmer -i FILE -f fasta -k 31`): set::Hash of every canonical 31-mer of
    the file (the reference's raw.k31.fasta is not shipped, so the reads themselves serve as the k-mer
    source), default method chain; also k = 21 with a tiny --chunk-bases so that the hash set is filled
    chunk by chunk and grows.  The corrected records must equal the oracle's."""
    seq, off = fixture_reads
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    for k, extra in ((31, []), (21, ["--chunk-bases", "300000"])):
        out = tmp_path / f"corr{k}.fasta"
        r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-t", "4", *extra, "large-kmer", "-i", GOLDEN / "br_reads.fa.gz",
                 "-f", "fasta", "-k", str(k)])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        oh = oracle.Hash.from_reads(k, seq, off)
        ids = [oracle.METHOD_IDS[m] for m in ("one", "two", "graph", "greedy", "gap_size")]
        exp, exp_off = oh.run_correction(ids, seq, off, confirm=5, max_search=7, threads=8)
        assert_same_records(out, names, exp, exp_off)
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "o.fa", "large-kmer", "-i", GOLDEN / "br_reads.fa.gz", "-f",
             "fasta", "-k", "33"])
    assert r.returncode == 1 and b"3..=31" in r.stderr


@pytest.mark.gpu
def test_fasta_subcommand_streams_its_input_in_chunks(tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """count_fasta(inputs, 8192) reads the records chunk by chunk (src/main.rs:74).  The CLI streams too:
    with --chunk-bases 200000 the 2.5 Mbase fixture becomes 13 partitions (k = 15: the bucketed path;
    k = 11: the counter accumulates), two input files are chained, and the set equals the one-shot one."""
    seq, off = fixture_reads
    half = tmp_path / "half.fa"
    data = gzip.open(GOLDEN / "br_reads.fa.gz").read()
    cut = data.find(b"\n>", len(data) // 2) + 1
    half.write_bytes(data[:cut])
    rest = tmp_path / "rest.fa"
    rest.write_bytes(data[cut:])
    for k in (11, 15):
        solid_out = tmp_path / f"set{k}.solid"
        r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "o.fa", "-c", "one", "--write-solid", solid_out,
                 "--chunk-bases", "200000", "fasta", "-i", half, rest, "-k", str(k), "-a", "2"])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        payload = gzip.open(solid_out).read()
        if k == 11:
            assert payload == fixture_solid_payload
        else:
            c = oracle.Counter(k)
            c.count(seq, off, threads=8)
            assert payload[0] == k and np.array_equal(np.frombuffer(payload[1:], dtype=np.uint8), c.to_solid(2, 8).bits())
    # first-minimum over chunks: the spectrum pass runs over all partitions
    out = tmp_path / "fm.fa"
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-c", "one", "--chunk-bases", "500000", "fasta", "-i",
             GOLDEN / "br_reads.fa.gz", "-k", "15", "first-minimum"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    c = oracle.Counter(15)
    c.count(seq, off, threads=8)
    solid = c.to_solid(oracle.Counter.first_minimum(c.spectrum(8)), 8)
    exp, exp_off = solid.run_correction([oracle.METHOD_IDS["one"]], seq, off, confirm=5, threads=8)
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    assert_same_records(out, names, exp, exp_off)


@pytest.mark.gpu
def test_both_transports_echo_non_acgt_bytes(tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """The CLI moves chunks across PCIe at 2 bits per base plus an exception list (`--transport packed`,
    the default) or as ASCII; reads with lower-case runs, N and other bytes must come out the same either
    way and equal the oracle: uncorrected positions echo the original byte (src/correct/mod.rs:91,100)."""
    seq, off = fixture_reads
    rng = np.random.default_rng(3)
    seq = seq[: int(off[40])].copy()
    off = off[:41]
    for r in range(0, 40, 3):  # a lower-case stretch and a few Ns / IUPAC codes in every third read
        a, b = int(off[r]), int(off[r + 1])
        s = a + int(rng.integers(0, max(1, b - a - 200)))
        seq[s : s + 150] |= 0x20
        hits = a + rng.integers(0, b - a, size=12)
        seq[hits] = rng.choice(np.frombuffer(b"NnRYKM", dtype=np.uint8), size=12)
    inp = tmp_path / "odd.fa"
    with open(inp, "wb") as f:
        for r in range(40):
            f.write(b">r%d\n" % r + seq[int(off[r]) : int(off[r + 1])].tobytes() + b"\n")
    exp, exp_off = oracle_corrected(oracle, fixture_solid_payload, ["one", "two", "graph", "greedy", "gap_size"], seq, off)
    names = [b"r%d" % r for r in range(40)]
    for transport in ("packed", "ascii"):
        out = tmp_path / f"corr_{transport}.fa"
        r = run(["-i", inp, "-o", out, "--transport", transport, "solid", "-i", GOLDEN / "br_reads.k11.a2.solid", "-f", "solid"])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        assert_same_records(out, names, exp, exp_off)
    assert (exp[: 0] == exp[: 0]).all() and any(c in exp.tobytes() for c in (b"n", b"N", b"a"))


@pytest.mark.gpu
def test_count_subcommand_reads_a_pcon_count_file(tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """`count -i FILE -a 2` (src/main.rs:59-70): Counter::from_stream + count2solid.  No reference fixture holds
    a count file (PARITY UNPINNED for the container; restated as recalled from pcon: one raw byte k, then the
    counters as a multi-member gzip stream) — the counters come from the oracle's count of the fixture reads, so
    the set must be the `.solid` fixture and the corrected reads the oracle's; the all-gzip and the raw-counters
    spellings are accepted too, and first-minimum works on the file's spectrum."""
    seq, off = fixture_reads
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    raw = c.raw().tobytes()
    members = gzip.compress(raw[: len(raw) // 3]) + gzip.compress(raw[len(raw) // 3 :])  # two gzip members
    spellings = {"members": bytes([11]) + members, "raw": bytes([11]) + raw, "all_gzip": gzip.compress(bytes([11]) + raw)}
    names, _, _ = records(GOLDEN / "br_reads.fa.gz")
    exp, exp_off = oracle_corrected(oracle, fixture_solid_payload, ["one"], seq, off)
    for name, blob in spellings.items():
        path = tmp_path / f"{name}.pcon"
        path.write_bytes(blob)
        out, solid_out = tmp_path / f"corr_{name}.fa", tmp_path / f"{name}.solid"
        r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", out, "-c", "one", "--write-solid", solid_out, "count", "-i", path, "-a", "2"])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        assert gzip.open(solid_out).read() == fixture_solid_payload
        assert_same_records(out, names, exp, exp_off)
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "fm.fa", "-c", "one", "--write-solid", tmp_path / "fm.solid",
             "count", "-i", tmp_path / "members.pcon", "first-minimum"])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    thr = oracle.Counter.first_minimum(c.spectrum(8))
    assert np.array_equal(np.frombuffer(gzip.open(tmp_path / "fm.solid").read()[1:], dtype=np.uint8), c.to_solid(thr, 8).bits())
    r = run(["-i", GOLDEN / "br_reads.fa.gz", "-o", tmp_path / "x.fa", "count", "-i", tmp_path / "members.pcon"])
    assert r.returncode == 1 and b"abundance" in r.stderr


def test_cpp_writer_and_python_writer_emit_the_same_bytes(tmp_path):
    """Both host mirrors restate noodles' writer (definition line verbatim, 80-column lines): the C++ `echo` sub-command and
    br_b200.fasta.write_fasta must produce identical files for the same records (empty records included)."""
    from br_b200 import fasta

    rng = np.random.default_rng(12)
    src = tmp_path / "in.fa"
    with open(src, "wb") as f:
        for i in range(60):
            n = [0, 1, 79, 80, 81, 160, int(rng.integers(0, 2000))][i % 7]
            s = bytes(np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=n)])
            f.write(b">r%d some description\n" % i + s + b"\n")
    out_cpp, out_py = tmp_path / "cpp.fa", tmp_path / "py.fa"
    r = run(["-i", src, "-o", out_cpp, "echo"])
    assert r.returncode == 0 and r.stderr == b""
    defs, seq, off = fasta.read_fasta(src)
    with open(out_py, "wb") as f:
        fasta.write_fasta(f, defs, seq, off)
    assert out_cpp.read_bytes() == out_py.read_bytes()
