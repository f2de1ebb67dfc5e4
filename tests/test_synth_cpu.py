"""Host mirror of the counter-based read generator (br_b200/synth.py): determinism, shape of the
data set, shard arithmetic.  The device kernel is compared with it in tests/test_gpu_synth.py."""
import importlib.util
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def load():
    spec = importlib.util.spec_from_file_location("brgpu_synth", ROOT / "br_b200" / "synth.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_descriptors_cover_the_genome_at_the_asked_depth():
    sy = load()
    start, tlen, strand = sy.read_descriptors(1_000_000, 30, seed=43)
    total = int(tlen.astype(np.int64).sum())
    assert 30_000_000 <= total < 30_000_000 + 100_000
    assert int(tlen[:-1].astype(np.int64).sum()) < 30_000_000  # the last read is the one that reaches the target
    assert tlen.min() >= 500 and tlen.max() <= 100_000
    assert np.all(start.astype(np.int64) + tlen <= 1_000_000)
    assert 0.4 < strand.mean() < 0.6
    a = sy.read_descriptors(1_000_000, 30, seed=43)
    assert all(np.array_equal(x, y) for x, y in zip(a, (start, tlen, strand)))


def test_error_rates_and_shards():
    sy = load()
    start, tlen, strand = sy.read_descriptors(200_000, 10, seed=5, mean_len=2000)
    thr = sy.error_thresholds(0.10)
    seq, off = sy.host_reads(42, 43, 0, start, tlen, strand, thr)
    # expected length = template length (insertions and deletions are equally likely)
    assert abs(int(off[-1]) - int(tlen.astype(np.int64).sum())) < 0.01 * int(off[-1])
    assert set(np.unique(seq)) <= set(b"ACGT")
    clean, _ = sy.host_reads(42, 43, 0, start[:20], tlen[:20], strand[:20], sy.error_thresholds(0.0))
    noisy, noff = sy.host_reads(42, 43, 0, start[:20], tlen[:20], strand[:20], thr)
    assert clean.size == int(tlen[:20].astype(np.int64).sum()) and noisy.size != clean.size
    # shards tile the read list and generate the same bytes as the whole
    cuts = [sy.shard_descriptors(tlen, 4, r) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == tlen.size and all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
    lo, hi = cuts[2]
    part, _ = sy.host_reads(42, 43, lo, start[lo:hi], tlen[lo:hi], strand[lo:hi], thr)
    assert np.array_equal(part, seq[int(off[lo]) : int(off[hi])])
    # the genome is a pure function of the position
    g = sy.host_genome(42, 5000)
    assert np.array_equal(g[1000:1100], sy._ACGT[sy.genome_codes(42, np.arange(1000, 1100, dtype=np.uint64))])
