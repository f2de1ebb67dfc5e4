"""CPU-side checks of the boundary: the shared library loads without a GPU, exports every
symbol include/brgpu.h declares, the ctypes table matches the header, and entry points that
need a device fail loudly instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    src = (ROOT / "include" / "brgpu.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(brgpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from br_b200 import _lib

    names = header_functions()
    assert len(names) >= 45
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table and the header agree
    for n in names:
        assert getattr(_lib.lib, n) is not None  # raises AttributeError if the .so lacks it
    assert _lib.lib.brgpu_version().startswith(b"brgpu")


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import br_b200

    with pytest.raises(br_b200.BrgpuError) as e:
        br_b200.Context(0)
    assert e.value.status == 2  # BRGPU_E_NO_DEVICE


def test_first_minimum_helper_needs_no_device():
    from br_b200 import Counter

    hist = np.zeros(256, dtype=np.uint64)
    hist[:9] = [1436018, 442564, 95498, 19526, 4458, 1221, 460, 494, 810]
    assert Counter.first_minimum(hist) == 6
    assert Counter.first_minimum(np.arange(256, 0, -1, dtype=np.uint64)) is None


def test_fasta_roundtrip_and_chunks(tmp_path):
    from br_b200 import fasta

    rec = [(b"r1 desc", b"ACGT" * 50), (b"r2", b""), (b"r3", b"N" * 81)]
    p = tmp_path / "x.fa"
    with open(p, "wb") as f:
        for d, s in rec:
            f.write(b">" + d + b"\n")
            for i in range(0, len(s), 60):
                f.write(s[i : i + 60] + b"\n")
    defs, seq, off = fasta.read_fasta(p)
    assert defs == [d for d, _ in rec]
    assert [seq[int(off[i]) : int(off[i + 1])].tobytes() for i in range(3)] == [s for _, s in rec]
    out = tmp_path / "y.fa"
    with open(out, "wb") as f:
        fasta.write_fasta(f, defs, seq, off)
    lines = out.read_bytes().split(b"\n")
    assert lines[0] == b">r1 desc" and len(lines[1]) == 80  # 80-column wrapping
    d2, s2, o2 = fasta.read_fasta(out)
    assert d2 == defs and np.array_equal(s2, seq) and np.array_equal(o2, off)
    chunks = list(fasta.iter_chunks(defs, seq, off, 2))
    assert [len(c[0]) for c in chunks] == [2, 1]


def test_synthetic_reads_are_deterministic():
    from br_b200 import synth

    g = synth.make_genome(5000, seed=42)
    a = synth.make_reads(g, 5, 0.1, seed=43, mean_len=500, min_len=50)
    b = synth.make_reads(g, 5, 0.1, seed=43, mean_len=500, min_len=50)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert set(np.unique(a[0])) <= set(b"ACGT")
    assert int(a[1][-1]) >= 5 * 5000


def test_bench_reference_arm_line_shape(monkeypatch, capsys):
    """--impl reference prints one JSON line with the contract's keys (tiny genome so it runs in seconds)."""
    import json
    import sys

    sys.path.insert(0, str(ROOT))
    import bench

    monkeypatch.setattr(bench, "K", 11)  # 2 MiB table instead of 8 GiB
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                                      "--genome-per-gpu", "20000"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "bases/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config"):
        assert key in line


def test_spectrum_threshold_is_host_arithmetic_and_matches_the_oracle():
    """brgpu_spectrum_threshold needs no device; library and oracle restate pcon's pickers
    independently and must agree, on the reference fixture's spectrum and on random ones."""
    import numpy as np

    from br_b200 import set as bset
    from oracle import br_oracle as o

    o.build()
    fixture = np.zeros(256, dtype=np.uint64)
    fixture[1:9] = [442564, 95498, 19526, 4458, 1221, 460, 494, 810]  # raw.fasta at k = 11
    fixture[0] = (1 << 21) - int(fixture.sum())
    assert bset.spectrum_threshold(fixture, "first-minimum") == 6
    rng = np.random.default_rng(1)
    cases = [fixture] + [(rng.integers(0, 10**6, 256) * (rng.random(256) < 0.4)).astype(np.uint64) for _ in range(200)]
    for h in cases:
        for m in ("rarefaction", "percent-most", "percent-least"):
            for p in (0.001, 0.05, 0.5, 0.99, 1.5):
                assert bset.spectrum_threshold(h, m, p) == o.Counter.spectrum_threshold(h, m, p), (m, p)


def header_declarations():
    """(return type, name, [parameter declarations]) of every function include/brgpu.h declares."""
    src = (ROOT / "include" / "brgpu.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = []
    for ret, name, params in re.findall(r"([\w\s\*]+?)\b(brgpu_\w+)\s*\(([^;{]*?)\)\s*;", src):
        params = params.strip()
        out.append((ret.strip(), name, [] if params in ("", "void") else [p.strip() for p in params.split(",")]))
    return out


def test_ctypes_table_matches_the_header_argument_by_argument():
    """The Python binding restates every prototype by hand (br_b200/_lib.py): arity, pointer-ness and the width of
    every scalar must be the header's — a drifted entry would corrupt arguments silently."""
    from br_b200 import _lib

    scalars = {"int": C.c_int, "uint64_t": C.c_uint64, "size_t": C.c_size_t, "double": C.c_double}

    def is_pointer(t):
        return t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents")

    decls = header_declarations()
    assert sorted(n for _, n, _ in decls) == sorted(_lib.SIGNATURES)
    for ret, name, params in decls:
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(params), name
        for p, a in zip(params, args):
            if "*" in p or "[" in p:
                assert is_pointer(a), (name, p, a)
            else:
                ctype = re.sub(r"\bconst\b", "", p).split()[0]
                assert a is scalars[ctype], (name, p, a)
        if ret == "void":
            assert res is None, name
        elif "*" in ret:
            assert is_pointer(res), name
        else:
            assert res is scalars[re.sub(r"\bconst\b", "", ret).split()[0]], (name, ret, res)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/brgpu.h must compile as C99 on its own (no C++, no CUDA, no torch types)."""
    import subprocess

    src = tmp_path / "use_header.c"
    src.write_text('#include "brgpu.h"\nint main(void) { brgpu_ctx *c = 0; (void)c; return BRGPU_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", "-c", str(src), "-o",
                    str(tmp_path / "use_header.o")], check=True)
