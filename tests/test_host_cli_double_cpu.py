"""The C++ host side on the CPU: br_b200/host/{cli.cpp, kat_runner.cpp} (and with them br.hpp, fasta.hpp,
formats.hpp) are linked, unmodified, against tests/abi_double/brgpu_double.cpp — a TEST DOUBLE of include/brgpu.h
over the oracle — instead of libbrgpu.so, and the command-line tests of tests/test_host_cli.py and
tests/test_set_formats_gpu.py are replayed through that binary.

What this covers (the driver's CPU stage has no GPU): br's argument contract, the streamed set builders
(`fasta` over several files and chunks, `count`, `solid -f solid|fasta|fastq|csv`, `large-kmer -f fasta|fastq|csv`),
the three-stage run_correction pipeline (reader / device calls / writer threads, input order), both transports with
their exception lists, `--write-solid`, the reference's error messages, the KAT runner — i.e. the host's use of the
ABI contract.  What it does NOT cover: the CUDA kernels.  The double computes with the oracle, so "corrected reads
equal the oracle's" is by construction here; the same tests run against libbrgpu.so under `-m gpu`.
One run under ThreadSanitizer checks the pipeline's hand-overs between its threads."""
import subprocess

import pytest

import test_host_cli as cli_tests
import test_set_formats_gpu as format_tests
from conftest import GOLDEN, ROOT

SOURCES = [ROOT / "tests" / "abi_double" / "brgpu_double.cpp", ROOT / "oracle" / "br_oracle.cpp"]


def link_double(exe, host_source, extra=()):
    # /usr/bin/g++: the image's default g++ has no libgomp.spec (oracle/Makefile)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Wno-array-bounds", "-pthread", "-fopenmp", *extra, "-o", str(exe),
                    str(ROOT / "br_b200" / "host" / host_source), *map(str, SOURCES), "-lz"], check=True)
    return exe


def sanitizer_runtime_starts(exe):
    """False where the sanitizer's own runtime cannot run (LeakSanitizer needs ptrace, ThreadSanitizer a compatible address
    space layout): that is the sandbox, not the code under test."""
    r = subprocess.run([str(exe), "echo"], input=b">a\nAC\n", capture_output=True, timeout=120)
    return r.returncode == 0 and r.stdout == b">a\nAC\n" and r.stderr == b""


@pytest.fixture(scope="module")
def doubles(tmp_path_factory):
    d = tmp_path_factory.mktemp("abi_double")
    san = ("-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-g")  # every replayed test also runs under ASan + UBSan
    cli = link_double(d / "brgpu-cli-double", "cli.cpp", extra=san)
    if not sanitizer_runtime_starts(cli):
        san = ()
        cli = link_double(d / "brgpu-cli-double", "cli.cpp")
    return cli, link_double(d / "brgpu-kat-double", "kat_runner.cpp", extra=san)


@pytest.fixture
def through_the_double(doubles, monkeypatch):
    cli, kat = doubles
    monkeypatch.setattr(cli_tests, "CLI", cli)
    monkeypatch.setattr(cli_tests, "KAT", kat)
    monkeypatch.setattr(format_tests, "CLI", cli)
    monkeypatch.setattr(format_tests, "KAT", kat)


def test_solid_subcommand(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    cli_tests.test_solid_subcommand_like_tests_br_rs(tmp_path, oracle, fixture_reads, fixture_solid_payload)


def test_fasta_subcommand(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    cli_tests.test_fasta_subcommand_config1(tmp_path, oracle, fixture_reads, fixture_solid_payload)
    cli_tests.test_fasta_first_minimum_two_side_and_even_k(tmp_path, oracle, fixture_reads)
    cli_tests.test_missing_abundance_is_the_reference_error(tmp_path)
    cli_tests.test_fasta_percent_least_through_the_cli(tmp_path, oracle, fixture_reads)


def test_fasta_subcommand_streams_its_input(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    cli_tests.test_fasta_subcommand_streams_its_input_in_chunks(tmp_path, oracle, fixture_reads, fixture_solid_payload)


def test_solid_from_fasta_and_count_file(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    cli_tests.test_solid_from_fasta_is_presence_only(tmp_path, oracle, fixture_reads)
    cli_tests.test_count_subcommand_reads_a_pcon_count_file(tmp_path, oracle, fixture_reads, fixture_solid_payload)


def test_large_kmer_subcommand(through_the_double, tmp_path, oracle, fixture_reads):
    cli_tests.test_large_kmer_subcommand_like_tests_br_rs(tmp_path, oracle, fixture_reads)


def test_both_transports(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    cli_tests.test_both_transports_echo_non_acgt_bytes(tmp_path, oracle, fixture_reads, fixture_solid_payload)


def test_fastq_and_csv_set_inputs(through_the_double, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    format_tests.test_cli_solid_and_large_kmer_take_fastq_and_csv(None, oracle, tmp_path, fixture_reads, fixture_solid_payload)


def test_reference_unit_kats_through_the_cpp_interface(through_the_double, kats):
    cli_tests.test_reference_unit_kats_through_the_cpp_interface(kats)
    format_tests.test_hash_set_kats_through_the_cpp_interface(None, kats, corrector_ks=(5, 7, 11))  # the same over br::set::Hash


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_pipeline_under_sanitizers(tmp_path, oracle, fixture_reads, fixture_solid_payload, sanitizer):
    """run_correction's three stages (parse + pack, device calls, unpack + format + write) hand two buffers round between
    three threads: six chunks through the packed transport under TSan (and ASan + UBSan), output still in input order
    and correct."""
    COPIES, CUT = 200, 120
    exe = link_double(tmp_path / "brgpu-cli-san", "cli.cpp", extra=(f"-fsanitize={sanitizer}", "-fno-sanitize-recover=all", "-g", "-O1",
                                                                    "-DBRGPU_DOUBLE_SERIAL"))
    if not sanitizer_runtime_starts(exe):
        pytest.skip(f"-fsanitize={sanitizer}: the sanitizer runtime does not start in this environment")
    seq, off = fixture_reads
    # 206 reads x 200 copies = 41 200 records: six chunks of the 8192-record loop
    big = tmp_path / "many.fa"
    names = []
    with open(big, "wb") as f:
        for rep in range(COPIES):
            for r in range(off.size - 1):
                names.append(b"c%d_r%d" % (rep, r))
                f.write(b">" + names[-1] + b"\n" + seq[int(off[r]) : int(off[r + 1])].tobytes()[:CUT] + b"\n")
    out = tmp_path / "out.fa"
    r = subprocess.run([str(exe), "-i", str(big), "-o", str(out), "-c", "one", "solid", "-i", str(GOLDEN / "br_reads.k11.a2.solid"), "-f", "solid"],
                       capture_output=True, timeout=900)
    assert r.returncode == 0 and r.stderr == b"", r.stderr.decode()[-3000:]
    n1, s1, o1 = cli_tests.records(out)
    assert n1 == names
    import numpy as np

    lens = np.minimum(np.diff(off.astype(np.int64)), CUT)
    sub_off = np.zeros(off.size, dtype=np.uint64)
    sub_off[1:] = np.cumsum(lens)
    sub_seq = np.concatenate([seq[int(off[r]) : int(off[r]) + int(lens[r])] for r in range(off.size - 1)])
    exp, exp_off = cli_tests.oracle_corrected(oracle, fixture_solid_payload, ["one"], sub_seq, sub_off)
    exp_off = exp_off.astype(np.int64)
    assert o1.size - 1 == COPIES * (off.size - 1)
    per = int(exp_off[-1])
    for rep in (0, 39, 40, 117, COPIES - 1):  # every copy of the fixture comes back as the oracle's correction, in order
        a = int(o1[rep * (off.size - 1)])
        assert np.array_equal(s1[a : a + per], exp)


def test_device_list_goes_through_the_group_calls(doubles, tmp_path, oracle, fixture_reads, fixture_solid_payload):
    """`brgpu-cli -d 0,1,2 ... fasta`: run_group's own chunk loop over brgpu_group_* (three contexts of the double stand
    for three devices): records in input order, the reference's error when neither -a nor a method is given."""
    cli, _ = doubles
    seq, off = fixture_reads
    reads = GOLDEN / "br_reads.fa.gz"
    out = tmp_path / "corr.fa"
    r = subprocess.run([str(cli), "-d", "0,1,2", "-i", str(reads), "-o", str(out), "-c", "one", "two", "fasta", "-i", str(reads), "-k", "12",
                        "-a", "2"], capture_output=True, timeout=600)  # -k 12 is decremented to 11 (src/cli.rs:277-279)
    assert r.returncode == 0 and r.stderr == b"", r.stderr.decode()[-2000:]
    exp, exp_off = cli_tests.oracle_corrected(oracle, fixture_solid_payload, ["one", "two"], seq, off)
    names, _, _ = cli_tests.records(reads)
    cli_tests.assert_same_records(out, names, exp, exp_off)
    r = subprocess.run([str(cli), "-d", "0,1", "-i", str(reads), "-o", str(out), "-s", "-c", "gap-size", "fasta", "-i", str(reads), "-k", "11",
                        "first-minimum"], capture_output=True, timeout=600)
    assert r.returncode == 0 and r.stderr == b"", r.stderr.decode()[-2000:]
    c = oracle.Counter(11)
    c.count(seq, off, threads=8)
    solid = c.to_solid(oracle.Counter.first_minimum(c.spectrum(8)), 8)
    exp, exp_off = solid.run_correction([oracle.METHOD_IDS["gap_size"]], seq, off, confirm=5, max_search=7, two_side=True, threads=8)
    cli_tests.assert_same_records(out, names, exp, exp_off)
    r = subprocess.run([str(cli), "-d", "0,1", "-i", str(reads), "-o", str(out), "fasta", "-i", str(reads), "-k", "11"], capture_output=True,
                       timeout=600)
    assert r.returncode == 1 and b"abundance" in r.stderr
