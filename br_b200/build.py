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


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", str(SO), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
