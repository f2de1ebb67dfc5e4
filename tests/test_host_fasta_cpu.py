"""Host side of the path on the CPU (no GPU, no libbrgpu): br_b200/host/fasta.hpp's reader — mapped plain file,
block reads from a pipe, gzip stream — its 2-bit packer / unpacker (word-at-a-time fast path + exception list) and
its writer, against a line-based Python parse and the numpy mirror of the transport form
(br_b200.runtime.pack_2bit).  Framing cases follow what the reference's readers accept (SURVEY §8c "FASTA
framing"): CRLF, ragged line widths, '>' inside a definition, blank lines before the first record, no final
newline, lower case and N (which must come back byte for byte: src/correct/mod.rs:91,100)."""
import gzip
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hostfasta") / "host_fasta_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", f"-I{ROOT / 'br_b200' / 'host'}",
                    str(ROOT / "tests" / "host_fasta_check.cpp"), "-lz", "-o", str(exe)], check=True)
    return exe


def make_fasta(rng, n_records, mean_len, crlf=False, final_newline=True):
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtNn", dtype=np.uint8)  # mostly upper case, some exceptions
    parts, names, seqs = [b"\n\n"], [], []
    for i in range(n_records):
        L = int(rng.integers(0, 2 * mean_len)) if i % 17 else 0  # empty records too
        s = rng.choice(alphabet, size=L).tobytes()
        name = b"read%d len=%d a>b" % (i, L)  # a '>' that does not start a line
        width = int(rng.integers(1, 200))
        eol = b"\r\n" if crlf and i % 2 else b"\n"
        parts.append(b">" + name + eol)
        for p in range(0, L, width):
            parts.append(s[p : p + width] + eol)
        names.append(name)
        seqs.append(s)
    buf = b"".join(parts)
    if not final_newline and buf.endswith(b"\n"):
        buf = buf[:-1]
    return buf, names, seqs


def run_and_load(checker, path, prefix, threads, chunk_records):
    subprocess.run([str(checker), str(path), str(prefix), str(threads), str(chunk_records)], check=True, timeout=300)
    seq = np.fromfile(str(prefix) + ".seq", dtype=np.uint8)
    off = np.fromfile(str(prefix) + ".off", dtype=np.uint64)
    defs = open(str(prefix) + ".defs", "rb").read().split(b"\n")[:-1]
    return seq, off, defs


def check_against(names, seqs, seq, off, defs):
    exp = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    exp_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    exp_off[1:] = np.cumsum([len(s) for s in seqs])
    assert defs == names
    assert np.array_equal(off, exp_off)
    assert np.array_equal(seq, exp)


@pytest.mark.parametrize("threads,chunk_records,n_records,mean_len,crlf,final_newline", [
    (1, 8192, 50, 300, False, True),
    (4, 7, 50, 300, True, False),          # chunk boundaries inside the file, CRLF, no final newline
    (8, 8192, 3000, 3000, False, True),    # > 4 MiB: the multi-threaded parse / pack / unpack / format paths
    (3, 1000, 3000, 3000, True, True),
])
def test_reader_packer_writer(checker, tmp_path, threads, chunk_records, n_records, mean_len, crlf, final_newline):
    from br_b200.runtime import pack_2bit

    rng = np.random.default_rng(n_records + threads)
    buf, names, seqs = make_fasta(rng, n_records, mean_len, crlf, final_newline)
    plain = tmp_path / "in.fa"
    plain.write_bytes(buf)
    gz = tmp_path / "in.fa.gz"
    with gzip.open(gz, "wb", compresslevel=1) as f:
        f.write(buf)
    fifo = tmp_path / "in.fifo"
    os.mkfifo(fifo)

    def feed():
        with open(fifo, "wb") as f:
            f.write(buf)

    for tag, path in (("mapped", plain), ("gzip", gz), ("pipe", fifo)):
        t = None
        if tag == "pipe":
            t = threading.Thread(target=feed)
            t.start()
        seq, off, defs = run_and_load(checker, path, tmp_path / tag, threads, chunk_records)
        if t:
            t.join()
        check_against(names, seqs, seq, off, defs)
        # the transport form equals the numpy mirror's, byte for byte
        packed, exc_pos, exc_byte = pack_2bit(seq)
        assert np.array_equal(np.fromfile(str(tmp_path / tag) + ".packed", dtype=np.uint8), packed)
        got_pos = np.fromfile(str(tmp_path / tag) + ".excpos", dtype=np.uint64)
        got_byte = np.fromfile(str(tmp_path / tag) + ".excbyte", dtype=np.uint8)
        order = np.argsort(got_pos, kind="stable")
        assert np.array_equal(got_pos[order], exc_pos) and np.array_equal(got_byte[order], exc_byte)
        assert exc_pos.size > 0 or seq.size == 0
        # what the writer wrote parses back to the same records (80-column lines)
        seq2, off2, defs2 = run_and_load(checker, str(tmp_path / tag) + ".fa", tmp_path / (tag + "2"), threads, chunk_records)
        check_against(names, seqs, seq2, off2, defs2)
        lines = open(str(tmp_path / tag) + ".fa", "rb").read().split(b"\n")
        assert max(len(l) for l in lines if not l.startswith(b">")) <= 80


def test_fasta_reader_differential_fuzz(checker, tmp_path):
    """Random soups of `> A C LF CR space`: the C++ reader (mapped file and gzip stream, chunk sizes 1..3) and the
    Python mirror agree on every record — '>' only starts a record at the start of a line, a line ends in LF or CRLF
    (one CR is dropped, any other CR is a byte like any other), text before the first record is skipped."""
    import io

    from br_b200.fasta import read_fasta

    rng = np.random.default_rng(8)
    alphabet = np.frombuffer(b">AC\n\r ", dtype=np.uint8)
    plain, gz = tmp_path / "fuzz.fa", tmp_path / "fuzz.fa.gz"
    for i in range(300):
        data = rng.choice(alphabet, size=int(rng.integers(0, 80)), p=[.1, .3, .3, .2, .05, .05]).tobytes()
        plain.write_bytes(data)
        gz.write_bytes(gzip.compress(data, 1))
        pdefs, pseq, poff = read_fasta(io.BytesIO(data))
        for path in (plain, gz):
            seq, off, defs = run_and_load(checker, path, tmp_path / "fz", 2, 1 + i % 3)
            assert defs == pdefs and np.array_equal(seq, pseq) and np.array_equal(off, poff), (path.name, data)
