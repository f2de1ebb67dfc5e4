"""Minimal FASTA reader/writer for the host side of run_correction (src/lib.rs:30-31,57-60), plus the two
other inputs a solid set can be built from: FASTQ records and a CSV of k-mers (cargo features `fastq` / `csv`:
src/set/pcon.rs:27-45,114-181, src/set/hash.rs:20-39,102-175).

The reference uses noodles-fasta 0.38 (not vendored).  Reader: `>` definition line, sequence =
the following lines joined.  Writer: definition line verbatim, sequence wrapped at 80 columns
(noodles' default line base count; recalled, see SURVEY §8c — parity tests compare sequences
per record, not file bytes).  Input may be gzip (niffler sniffing in the reference).
"""
import csv
import gzip
import io

import numpy as np

LINE_BASES = 80


def _open(path_or_file):
    if hasattr(path_or_file, "read"):
        data = path_or_file.read()
    else:
        with open(path_or_file, "rb") as f:
            data = f.read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    return data


def read_fasta(path_or_file):
    """Returns (definitions: list[bytes], seq: uint8 ndarray, offsets: uint64 ndarray)."""
    data = _open(path_or_file)
    defs, parts, lens = [], [], []
    # a record starts with '>' at the start of a line only: a '>' inside a definition line is legal
    start = data.find(b">") if not data.startswith(b">") else 0
    if start > 0 and data[start - 1 : start] != b"\n":
        start = data.find(b"\n>")
        start = -1 if start < 0 else start + 1
    records = data[start + 1 :].split(b"\n>") if start >= 0 else []
    for rec in records:
        nl = rec.find(b"\n")
        if nl < 0:
            defs.append(rec[:-1] if rec.endswith(b"\r") else rec)
            lens.append(0)
            continue
        defs.append(rec[: nl - 1] if nl and rec[nl - 1 : nl] == b"\r" else rec[:nl])
        # a line ends in LF or CRLF; a CR anywhere else is a sequence byte like any other (as in the C++ reader)
        body = b"".join(l[:-1] if l.endswith(b"\r") else l for l in rec[nl + 1 :].split(b"\n"))
        parts.append(body)
        lens.append(len(body))
    off = np.zeros(len(defs) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.empty(0, dtype=np.uint8)
    return defs, seq, off


def iter_chunks(defs, seq, off, chunk_records):
    """populate_buffer (src/lib.rs:168-188): consecutive chunks of at most `chunk_records`."""
    n = len(defs)
    for a in range(0, n, chunk_records):
        b = min(n, a + chunk_records)
        yield defs[a:b], seq, off[a : b + 1]


def write_fasta(out, defs, seq, off, line_bases=LINE_BASES):
    buf = io.BytesIO()
    data = seq.tobytes() if hasattr(seq, "tobytes") else bytes(seq)
    for i, d in enumerate(defs):
        buf.write(b">" + d + b"\n")
        s = data[int(off[i]) : int(off[i + 1])]
        for p in range(0, len(s), line_bases):
            buf.write(s[p : p + line_bases] + b"\n")
    out.write(buf.getvalue())


def read_fastq(path_or_file):
    """noodles::fastq::Reader::records() as the reference's set builders consume it
    (`while let Some(Ok(record)) = records.next()`, src/set/pcon.rs:122): four lines per record —
    `@` name [description], the sequence on one line, a `+` line, the qualities; the first malformed or
    truncated record ends the input silently.  Returns (definitions, seq, offsets) like read_fasta."""
    data = _open(path_or_file)
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # the final newline does not start a line
    defs, parts, lens = [], [], []
    for i in range(0, len(lines), 4):
        rec = [l[:-1] if l.endswith(b"\r") else l for l in lines[i : i + 4]]
        if len(rec) < 4 or not rec[0].startswith(b"@") or not rec[2].startswith(b"+"):
            break
        defs.append(rec[0][1:])
        parts.append(rec[1])
        lens.append(len(rec[1]))
    off = np.zeros(len(defs) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.empty(0, dtype=np.uint8)
    return defs, seq, off


def read_csv_first_column(path_or_file):
    """csv::Reader::from_reader(input).byte_records() reduced to `record.get(0)` (src/set/pcon.rs:34-42):
    `,` delimiter, `"` quoting, the first record is the header and is skipped, empty lines are not records,
    a record with another number of fields than the header is an error.  Returns a list of bytes."""
    text = _open(path_or_file).decode("latin-1")  # byte-transparent
    out, n_fields = [], None
    for row in csv.reader(io.StringIO(text, newline="")):
        if not row:
            continue
        if n_fields is None:
            n_fields = len(row)  # header
            continue
        if len(row) != n_fields:
            raise ValueError(f"CSV error: found record with {len(row)} fields, but the previous record has {n_fields} fields")
        out.append(row[0].encode("latin-1"))
    return out
