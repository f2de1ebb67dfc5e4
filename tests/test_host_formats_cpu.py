"""The other inputs a solid set can be built from (br's cargo features `fastq` / `csv`: src/set/pcon.rs:27-45,
114-181, src/set/hash.rs:20-39,102-175), on the CPU: br_b200/host/formats.hpp's FASTQ record reader and CSV
first-column reader against the Python mirror (br_b200/fasta.py: a line-based FASTQ parse, the standard library's
csv module), and the framing rules both restate: four lines per FASTQ record with the first malformed record
ending the input silently (`while let Some(Ok(record))`, pcon.rs:122); `,` / `"` CSV with a header record, empty
lines skipped, LF / CRLF / CR record ends, unequal field counts an error (csv::Reader defaults, pcon.rs:33-36)."""
import gzip
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hostformats") / "host_formats_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-pthread", f"-I{ROOT / 'br_b200' / 'host'}",
                    str(ROOT / "tests" / "host_formats_check.cpp"), "-lz", "-o", str(exe)], check=True)
    return exe


def run_fastq(checker, path, prefix, chunk_records):
    r = subprocess.run([str(checker), "fastq", str(path), str(prefix), str(chunk_records)], check=True, capture_output=True, timeout=120)
    seq = np.fromfile(str(prefix) + ".seq", dtype=np.uint8)
    off = np.fromfile(str(prefix) + ".off", dtype=np.uint64)
    defs = open(str(prefix) + ".defs", "rb").read().split(b"\n")[:-1]
    return defs, seq, off, b"malformed" in r.stdout


def run_csv(checker, path, k, batch=1 << 20):
    r = subprocess.run([str(checker), "csv", str(path), str(k), str(batch)], check=True, capture_output=True, timeout=120)
    lines = r.stdout.decode().split("\n")[:-1]
    err = [l for l in lines if l.startswith("ERROR")]
    kmers = np.array([int(l) for l in lines if l and l != "-" and not l.startswith("ERROR")], dtype=np.uint64)
    return kmers, (err[0] if err else None), sum(l == "-" for l in lines)


def make_fastq(rng, n, crlf=False, final_newline=True):
    alphabet = np.frombuffer(b"ACGTACGTacgtN", dtype=np.uint8)
    names, seqs, parts = [], [], []
    for i in range(n):
        L = int(rng.integers(0, 400)) if i % 11 else 0  # empty sequences too
        s = rng.choice(alphabet, size=L).tobytes()
        q = bytes(rng.integers(33, 74, size=L, dtype=np.uint8))  # qualities may start with '@' or '+'
        name = b"read%d/1 len=%d @x +y" % (i, L)
        eol = b"\r\n" if crlf and i % 2 else b"\n"
        parts += [b"@" + name + eol, s + eol, (b"+" + name if i % 3 == 0 else b"+") + eol, q + eol]
        names.append(name)
        seqs.append(s)
    buf = b"".join(parts)
    if not final_newline and buf.endswith(b"\n"):
        buf = buf[:-1]
    return buf, names, seqs


@pytest.mark.parametrize("n,chunk_records,crlf,final_newline", [(200, 8192, False, True), (200, 7, True, False), (5000, 1000, True, True), (0, 10, False, True)])
def test_fastq_reader_matches_the_python_mirror(checker, tmp_path, n, chunk_records, crlf, final_newline):
    from br_b200.fasta import read_fastq

    buf, names, seqs = make_fastq(np.random.default_rng(n + chunk_records), n, crlf, final_newline)
    plain, gz = tmp_path / "in.fq", tmp_path / "in.fq.gz"
    plain.write_bytes(buf)
    gz.write_bytes(gzip.compress(buf, 1))
    exp_seq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    exp_off = np.zeros(n + 1, dtype=np.uint64)
    exp_off[1:] = np.cumsum([len(s) for s in seqs])
    for path in (plain, gz):
        defs, seq, off, malformed = run_fastq(checker, path, tmp_path / "out", chunk_records)
        assert defs == names and np.array_equal(seq, exp_seq) and np.array_equal(off, exp_off) and not malformed
        pdefs, pseq, poff = read_fastq(str(path))
        assert pdefs == names and np.array_equal(pseq, exp_seq) and np.array_equal(poff, exp_off)


@pytest.mark.parametrize("tail,kept", [
    (b"@r2\nACGT\n+\n", 2),               # truncated: no quality line
    (b"@r2\nACGT\n", 2),                  # truncated after the sequence
    (b"@r2\nACGT\n-\nIIII\n@r3\nAC\n+\nII\n", 2),   # third line is not a '+' line: the input ends there
    (b">r2\nACGT\n+\nIIII\n", 2),         # a FASTA record in a FASTQ stream
    (b"\n@r2\nACGT\n+\nIIII\n", 2),       # a blank line is not a record start
    (b"@r2\nACGT\n+\nIIII", 3),           # no final newline: still a record
])
def test_fastq_reader_stops_silently_at_the_first_malformed_record(checker, tmp_path, tail, kept):
    from br_b200.fasta import read_fastq

    head = b"@r0 d\nACGTAC\n+r0 d\nIIIIII\n@r1\nTTGG\n+\n@@++\n"
    p = tmp_path / "bad.fq"
    p.write_bytes(head + tail)
    defs, seq, off, malformed = run_fastq(checker, p, tmp_path / "out", 8192)
    exp_defs = [b"r0 d", b"r1", b"r2"][:kept]
    exp_seq = (b"ACGTAC" + b"TTGG" + b"ACGT")[: [0, 6, 10, 14][kept]]
    assert defs == exp_defs and seq.tobytes() == exp_seq and list(off) == [0, 6, 10, 14][: kept + 1]
    assert malformed == (kept == 2)
    pdefs, pseq, poff = read_fastq(str(p))
    assert pdefs == exp_defs and pseq.tobytes() == exp_seq and list(poff) == list(off)


def kmer_of(field):
    v = 0
    for b in field:
        v = (v << 2) | ((b >> 1) & 3)
    return v


def test_csv_first_column_matches_the_csv_module(checker, tmp_path):
    from br_b200.fasta import read_csv_first_column
    from br_b200.set import csv_kmers

    rng = np.random.default_rng(11)
    k = 15
    kmers = [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=k).tobytes() for _ in range(3000)]
    rows = [b"kmer,count,note"]
    for i, km in enumerate(kmers):
        first = b'"' + km + b'"' if i % 5 == 0 else km           # a quoted field is the same field
        note = [b"", b"x", b'"a,b"', b'"say ""hi"""', b'"two\nlines"'][i % 5]
        rows.append(first + b"," + str(int(rng.integers(1, 255))).encode() + b"," + note)
    eols = [b"\n", b"\r\n", b"\r"]
    blob = b"".join(r + eols[i % 3] + (b"\n" if i % 97 == 0 else b"") for i, r in enumerate(rows))  # empty lines in between
    for name, data in (("lf_crlf_cr.csv", blob), ("no_final_newline.csv", blob.rstrip(b"\r\n")), ("gz.csv.gz", gzip.compress(blob, 1))):
        p = tmp_path / name
        p.write_bytes(data)
        assert read_csv_first_column(str(p)) == kmers
        got, err, batches = run_csv(checker, p, k, batch=1000)
        assert err is None and batches == 3
        assert np.array_equal(got, np.array([kmer_of(x) for x in kmers], dtype=np.uint64))
        assert np.array_equal(csv_kmers(str(p), k), got)


def test_csv_header_only_empty_and_errors(checker, tmp_path):
    from br_b200.fasta import read_csv_first_column
    from br_b200.set import csv_kmers

    cases = {"empty.csv": b"", "header.csv": b"kmer,count\n", "blank.csv": b"\n\n\r\n"}
    for name, data in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        got, err, _ = run_csv(checker, p, 5)
        assert got.size == 0 and err is None and read_csv_first_column(str(p)) == []
    p = tmp_path / "one_column.csv"  # the header is a record even when it looks like a k-mer (csv::Reader has_headers)
    p.write_bytes(b"ACGTA\nCCGTA\nGGGTA\n")
    got, err, _ = run_csv(checker, p, 5)
    assert err is None and list(got) == [kmer_of(b"CCGTA"), kmer_of(b"GGGTA")]
    assert list(csv_kmers(str(p), 5)) == list(got)
    p = tmp_path / "ragged.csv"      # csv::ErrorKind::UnequalLengths (flexible = false)
    p.write_bytes(b"kmer,count\nACGTA,3\nCCGTA\n")
    got, err, _ = run_csv(checker, p, 5)
    assert err is not None and "fields" in err
    with pytest.raises(ValueError, match="fields"):
        read_csv_first_column(str(p))
    p = tmp_path / "wrong_length.csv"  # a field that is not a k-mer of the set's k
    p.write_bytes(b"kmer,count\nACGTA,3\nACGTAC,3\n")
    got, err, _ = run_csv(checker, p, 5)
    assert err is not None and "5-mer" in err
    with pytest.raises(ValueError, match="5-mer"):
        csv_kmers(str(p), 5)


def test_csv_reader_differential_fuzz_against_the_csv_module(checker, tmp_path):
    """Random soups of `A C , " LF CR space`: the C++ state machine and Python's csv module (both restating the same
    convention: quote at field start opens a quoted field, `""` inside is a quote, anything after the closing quote is
    literal, LF / CR / CRLF end a record, empty lines are no records, header first, ragged rows are an error) must see
    the same first column or both fail."""
    import io

    from br_b200.fasta import read_csv_first_column

    rng = np.random.default_rng(2)
    alphabet = np.frombuffer(b'AC,"\n\r ', dtype=np.uint8)
    p = tmp_path / "fuzz.csv"
    for _ in range(500):
        data = rng.choice(alphabet, size=int(rng.integers(0, 60)), p=[.25, .25, .15, .12, .13, .05, .05]).tobytes()
        p.write_bytes(data)
        r = subprocess.run([str(checker), "csvfields", str(p), "-", "-"], check=True, capture_output=True, timeout=60)
        lines = r.stdout.decode().split("\n")[:-1]
        try:
            expect = read_csv_first_column(io.BytesIO(data))
        except ValueError:
            assert lines and lines[-1].startswith("ERROR"), data
            continue
        assert [bytes.fromhex(l) for l in lines] == expect, data


def test_fastq_reader_differential_fuzz(checker, tmp_path):
    """Random soups of `@ + A C LF CR`: the streaming C++ reader and the line-based Python parse agree on where the
    input stops being FASTQ and on every record before that."""
    import io

    from br_b200.fasta import read_fastq

    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"@+AC\n\r", dtype=np.uint8)
    p = tmp_path / "fuzz.fq"
    for i in range(400):
        n_good = int(rng.integers(0, 4))  # a few well-formed records, then noise
        good = b"".join(b"@r%d\nAC\n+\nII\n" % j for j in range(n_good))
        data = good + rng.choice(alphabet, size=int(rng.integers(0, 40)), p=[.12, .12, .25, .25, .2, .06]).tobytes()
        p.write_bytes(data)
        defs, seq, off, _ = run_fastq(checker, p, tmp_path / "fz", 1 + i % 3)
        pdefs, pseq, poff = read_fastq(io.BytesIO(data))
        assert defs == pdefs and np.array_equal(seq, pseq) and np.array_equal(off, poff), data
