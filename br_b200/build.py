"""Build br_b200/libbrgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libbrgpu.so"
SOURCES = ["brgpu.cu", "set_kernels.cu", "correct_kernels.cu"]
HEADERS = ["internal.h", "kmer.cuh", "../../include/brgpu.h"]
# C++ host side (br's own interface over the C ABI) and its command line
HOST = PKG / "host"
CLI = PKG / "brgpu-cli"
KAT = PKG / "brgpu-kat"  # known-answer-test runner over the C++ interface (tests/test_host_cli.py)
HOST_SOURCES = ["cli.cpp", "kat_runner.cpp"]
HOST_HEADERS = ["br.hpp", "fasta.hpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libbrgpu.so cannot be built (there is no CPU fallback)")
    return p


def needs_build():
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def cli_needs_build():
    if not CLI.exists() or not KAT.exists():
        return True
    t = min(CLI.stat().st_mtime, KAT.stat().st_mtime)
    return SO.stat().st_mtime > t or any((HOST / f).stat().st_mtime > t for f in HOST_SOURCES + HOST_HEADERS)


def build_cli(force=False):
    """g++ host/cli.cpp -> br_b200/brgpu-cli, linked against the in-tree libbrgpu.so ($ORIGIN rpath)."""
    if not force and not cli_needs_build():
        return CLI
    for src, exe in (("cli.cpp", CLI), ("kat_runner.cpp", KAT)):
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-o", str(exe), str(HOST / src),
               f"-L{PKG}", "-lbrgpu", "-lz", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return CLI


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_cli()
        return SO
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", str(SO), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    build_cli(force=True)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
