"""Multi-GPU parity as a pytest: needs >= 2 visible GPUs (skipped on the single-GPU box the driver
uses for `-m gpu`; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the log
of that run is kept under profiles/).  Spawns tests/multi_gpu_check.py under torch.distributed.run:
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
