"""Multi-GPU parity as a pytest.  The torch.distributed test needs >= 2 visible GPUs (skipped on the
single-GPU box the driver uses for `-m gpu`; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`; the log of that run is kept under profiles/).  The group tests also run
with two contexts on ONE GPU (`-d 0,0`): the same partition / pull / count / compacted-exchange code path,
"peer" pointers being plain device pointers.  Spawns tests/multi_gpu_check.py under torch.distributed.run:
sharded set == single-GPU set (k = 17 / 15 k-mer protocol, k = 13 table protocol; explicit and
first-minimum thresholds) and every rank's corrected shard == the same records corrected alone."""
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_set_and_correction_equal_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "multi-GPU parity OK" in r.stdout


@pytest.mark.parametrize("second", [0, 1])
def test_group_api_single_process_two_gpus(oracle, fixture_reads, second):
    """brgpu_group_*: ONE process owning two GPUs (or two contexts on one GPU) builds the set (k = 15: k-mer protocol, k = 13: table
    protocol; explicit and first-minimum thresholds) and corrects a batch through the C ABI alone — no
    torch.distributed, no CUDA IPC.  Every replica's bitfield equals the oracle's, the corrected batch
    equals the oracle's bytes in input order."""
    import ctypes as C

    import numpy as np
    import torch

    if torch.cuda.device_count() < 1 + second:
        pytest.skip("needs at least 2 GPUs")
    import br_b200  # noqa: F401
    from br_b200._lib import lib

    seq, off = fixture_reads
    n = off.size - 1
    devs = (C.c_int * 2)(0, second)
    g = C.c_void_p()
    assert lib.brgpu_group_create(devs, 2, C.byref(g)) == 0
    try:
        assert lib.brgpu_group_size(g) == 2
        for k, abundance, selection in ((15, 2, 0), (15, -1, 1), (13, 2, 0), (13, -1, 1)):
            sets = (C.c_void_p * 2)()
            st = lib.brgpu_group_set_from_host_reads(g, k, abundance, selection, 0.0, seq.ctypes.data_as(C.c_void_p),
                                                     off.ctypes.data_as(C.c_void_p), n, sets)
            assert st == 0, lib.brgpu_group_last_error(g)
            oc = oracle.Counter(k)
            oc.count(seq, off, threads=8)
            thr = abundance if selection == 0 else oracle.Counter.first_minimum(oc.spectrum(8))
            osolid = oc.to_solid(thr, 8)
            for i in range(2):
                assert lib.brgpu_set_abundance(sets[i]) == thr
                nb = lib.brgpu_set_bitfield_bytes(sets[i])
                bits = np.empty(nb, dtype=np.uint8)
                assert lib.brgpu_set_export_bitfield(sets[i], bits.ctypes.data_as(C.c_void_p), nb) == 0
                assert np.array_equal(bits, osolid.bits()), (k, abundance, selection, i)
            methods = np.array([0, 1, 4], dtype=np.uint8)
            out = np.empty(int(off[-1]) * 2, dtype=np.uint8)
            out_off = np.empty(n + 1, dtype=np.uint64)
            req = C.c_uint64()
            st = lib.brgpu_group_correct_batch(g, sets, methods.ctypes.data_as(C.c_void_p), 3, 4, 7, 0, seq.ctypes.data_as(C.c_void_p),
                                               off.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), out.size,
                                               out_off.ctypes.data_as(C.c_void_p), C.byref(req))
            assert st == 0, lib.brgpu_group_last_error(g)
            exp, exp_off = osolid.run_correction([0, 1, 4], seq, off, confirm=4, max_search=7, threads=8)
            assert np.array_equal(out_off, exp_off) and np.array_equal(out[: req.value], exp)
            lib.brgpu_group_sets_free(g, sets)
    finally:
        lib.brgpu_group_destroy(g)


@pytest.mark.parametrize("second", [0, 1])
def test_cli_with_a_device_list_uses_the_group(tmp_path, oracle, fixture_reads, fixture_solid_payload, second):
    """`brgpu-cli -d 0,1 ... fasta -k 11 -a 2`: the single-process multi-GPU path of the command line."""
    import gzip

    import numpy as np
    import torch

    from conftest import GOLDEN, parse_fasta

    if torch.cuda.device_count() < 1 + second:
        pytest.skip("needs at least 2 GPUs")
    out = tmp_path / "corr.fa"
    r = subprocess.run([str(ROOT / "br_b200" / "brgpu-cli"), "-d", f"0,{second}", "-i", str(GOLDEN / "br_reads.fa.gz"), "-o", str(out), "-c", "one",
                        "two", "fasta", "-i", str(GOLDEN / "br_reads.fa.gz"), "-k", "11", "-a", "2"], capture_output=True, timeout=600)
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    seq, off = fixture_reads
    solid = oracle.Solid.from_solid_payload(fixture_solid_payload)
    exp, exp_off = solid.run_correction([0, 1], seq, off, confirm=5, max_search=7, threads=8)
    _, s1, o1 = parse_fasta(open(out, "rb").read())
    assert np.array_equal(o1, exp_off) and np.array_equal(s1, exp)
