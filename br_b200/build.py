"""Build br_b200/libbrgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The translation units compile in parallel: brgpu.cu (C ABI), set_kernels.cu (part 1),
correct_kernels.cu (phase A + the scan's host side) and scan_methods.cu once per
(correction method, variant) — `fast` is the product path, `cnt` counts the scans' KmerSet::get
calls for profiling runs (bench.py's roofline)."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build_obj"
SO = PKG / "libbrgpu.so"
SOURCES = ["brgpu.cu", "set_kernels.cu", "correct_kernels.cu", "scan_methods.cu", "synth_kernels.cu", "hash_kernels.cu", "group.cu"]
HEADERS = ["internal.h", "kmer.cuh", "scan_common.cuh", "scan_device.cuh", "../../include/brgpu.h"]
# C++ host side (br's own interface over the C ABI) and its command line
HOST = PKG / "host"
CLI = PKG / "brgpu-cli"
KAT = PKG / "brgpu-kat"  # known-answer-test runner over the C++ interface (tests/test_host_cli.py)
HOST_SOURCES = ["cli.cpp", "kat_runner.cpp"]
HOST_HEADERS = ["br.hpp", "fasta.hpp", "formats.hpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libbrgpu.so cannot be built (there is no CPU fallback)")
    return p


def existing(names):
    return [f for f in names if (CSRC / f).exists()]


def needs_build():
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in existing(SOURCES) + HEADERS)


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
        tmp = exe.with_name(exe.name + ".tmp")
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-o", str(tmp), str(HOST / src),
               f"-L{PKG}", "-lbrgpu", "-lz", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, exe)
    return CLI


def compile_jobs(extra_defines=()):
    """(object path, nvcc argument list) for every translation unit."""
    jobs = []
    for src in existing(SOURCES):
        if src == "scan_methods.cu":
            for variant, count in (("fast", 0), ("cnt", 1)):
                for m in range(5):
                    obj = OBJ / f"scan_m{m}_{variant}.o"
                    jobs.append((obj, [f"-DBRGPU_METHOD={m}", f"-DBRGPU_VARIANT={variant}", f"-DBRGPU_COUNT_GETS={count}",
                                       *extra_defines, "-c", "-o", str(obj), str(CSRC / src)]))
        else:
            obj = OBJ / (Path(src).stem + ".o")
            jobs.append((obj, [*extra_defines, "-c", "-o", str(obj), str(CSRC / src)]))
    return jobs


def build(force=False, verbose=False, out=None, extra_defines=()):
    """out / extra_defines: A/B builds (e.g. extra_defines=["-DBRGPU_SCAN_MINB=12"], out=libbrgpu_mb12.so)."""
    target = Path(out) if out else SO
    if not force and not out and not needs_build():
        build_cli()
        return SO
    nvcc = nvcc_path()
    OBJ.mkdir(exist_ok=True)
    jobs = compile_jobs(tuple(extra_defines))

    def run(job):
        obj, args = job
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", *(["-Xptxas", "-v"] if verbose else []), *args]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, r

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(run, jobs))
    for obj, r in results:
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {obj.name}:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout + r.stderr)
    # link beside the target and rename: a process that has the old library mapped keeps its (now unlinked) file
    tmp = target.with_name(target.name + ".tmp")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", "/usr/bin/g++", "-o", str(tmp),
           *[str(obj) for obj, _ in results]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, target)
    if not out:
        build_cli(force=True)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
