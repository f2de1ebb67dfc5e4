"""A host written in plain C (tests/abi_c_client.c) drives the hot path through include/brgpu.h the way a cgo / Rust-FFI
caller would: set from host reads, a correction batch into a caller-owned buffer that is first too small
(BRGPU_E_OVERFLOW + required size), teardown.  CPU stage: against libbrgpu.so it must stop at brgpu_ctx_create with
BRGPU_E_NO_DEVICE (no CPU path); linked against the ABI test double (tests/abi_double/) its digest must equal the
oracle's — that checks the client, not the kernels.  GPU stage: the same binary against libbrgpu.so equals the oracle."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CLIENT = ROOT / "tests" / "abi_c_client.c"


def fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def sample(fixture_reads, n=30):
    seq, off = fixture_reads
    return seq[: int(off[n])], off[: n + 1]


def expected_line(oracle, seq, off, k, abundance):
    c = oracle.Counter(k)
    c.count(seq, off, threads=4)
    solid = c.to_solid(abundance, 4)
    exp, exp_off = solid.run_correction([oracle.ONE, oracle.TWO], seq, off, confirm=5, max_search=7, two_side=False, threads=4)
    total = int(exp_off[-1])
    return (f"reads {off.size - 1} bases_in {int(off[-1])} bases_out {total} calls 2 fnv1a {fnv1a(exp[:total].tobytes()):016x} k {k}")


def stdin_of(seq, off):
    return b"".join(seq[int(off[r]) : int(off[r + 1])].tobytes() + b"\n" for r in range(off.size - 1))


def build_against_the_library(tmp_path):
    exe = tmp_path / "abi_c_client"
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(CLIENT), f"-L{ROOT / 'br_b200'}",
                    "-lbrgpu", f"-Wl,-rpath,{ROOT / 'br_b200'}", "-o", str(exe)], check=True)
    return exe


@pytest.mark.skipif("__import__('torch').cuda.is_available()")
def test_c_host_without_a_device_stops_at_ctx_create(tmp_path, fixture_reads):
    import br_b200  # noqa: F401  builds nothing, but fails loudly if libbrgpu.so is missing

    exe = build_against_the_library(tmp_path)
    seq, off = sample(fixture_reads)
    r = subprocess.run([str(exe), "11", "2"], input=stdin_of(seq, off), capture_output=True, timeout=120)
    assert r.returncode == 2 and r.stdout == b"brgpu_ctx_create: status 2\n" and r.stderr == b""


def test_c_host_against_the_abi_double(tmp_path, oracle, fixture_reads):
    obj, exe = tmp_path / "client.o", tmp_path / "abi_c_client_double"
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-Wall", "-Wextra", "-Werror", "-fsanitize=address,undefined", f"-I{ROOT / 'include'}", "-c",
                    str(CLIENT), "-o", str(obj)], check=True)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-array-bounds", "-fsanitize=address,undefined", str(obj),
                    str(ROOT / "tests" / "abi_double" / "brgpu_double.cpp"), str(ROOT / "oracle" / "br_oracle.cpp"), "-lz", "-o", str(exe)],
                   check=True)
    seq, off = sample(fixture_reads)
    probe = subprocess.run([str(exe), "11", "2"], input=b"", capture_output=True, timeout=120)
    if b"Sanitizer" in probe.stderr:  # the sanitizer runtime cannot start here (ptrace / address-space restrictions): plain build
        subprocess.run(["gcc", "-std=gnu99", "-O1", f"-I{ROOT / 'include'}", "-c", str(CLIENT), "-o", str(obj)], check=True)
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-array-bounds", str(obj),
                        str(ROOT / "tests" / "abi_double" / "brgpu_double.cpp"), str(ROOT / "oracle" / "br_oracle.cpp"), "-lz", "-o", str(exe)],
                       check=True)
    for k, abundance in ((11, 2), (15, 1)):
        r = subprocess.run([str(exe), str(k), str(abundance)], input=stdin_of(seq, off), capture_output=True, timeout=300)
        assert r.returncode == 0 and r.stderr == b"", r.stderr.decode()[-2000:]
        assert r.stdout.decode().strip() == expected_line(oracle, seq, off, k, abundance)
    assert np.diff(off.astype(np.int64)).min() > 11  # every read of the sample is longer than k


@pytest.mark.gpu
def test_c_host_on_the_gpu_equals_the_oracle(tmp_path, oracle, fixture_reads):
    exe = build_against_the_library(tmp_path)
    seq, off = sample(fixture_reads)
    for k, abundance in ((11, 2), (15, 1)):
        r = subprocess.run([str(exe), str(k), str(abundance)], input=stdin_of(seq, off), capture_output=True, timeout=300)
        assert r.returncode == 0 and r.stderr == b"", r.stdout.decode() + r.stderr.decode()
        assert r.stdout.decode().strip() == expected_line(oracle, seq, off, k, abundance)
