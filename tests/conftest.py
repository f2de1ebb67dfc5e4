import gzip
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def parse_fasta(buf: bytes):
    """Minimal FASTA parse for the fixtures: returns (names, concatenated uint8 sequence, u64 offsets)."""
    names, seqs = [], []
    for rec in buf.split(b">")[1:]:
        lines = rec.split(b"\n")
        names.append(lines[0])
        seqs.append(b"".join(lines[1:]))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return names, np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(), off


@pytest.fixture(scope="session")
def kats():
    return json.loads((GOLDEN / "kats.json").read_text())


@pytest.fixture(scope="session")
def fixture_reads():
    """The reference's tests/data/raw.fasta (206 reads) as (seq, offsets)."""
    _, seq, off = parse_fasta(gzip.open(GOLDEN / "br_reads.fa.gz").read())
    return seq, off


@pytest.fixture(scope="session")
def fixture_solid_payload():
    """gunzip(tests/data/raw.k11.a2.solid): byte 0 = k, rest = bitfield."""
    return gzip.open(GOLDEN / "br_reads.k11.a2.solid").read()


@pytest.fixture(scope="session")
def manifest():
    return json.loads((GOLDEN / "fixtures.json").read_text())


@pytest.fixture(scope="session")
def oracle():
    from oracle import br_oracle

    br_oracle.build()
    return br_oracle
