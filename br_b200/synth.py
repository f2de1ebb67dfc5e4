"""Synthetic genomes and ONT-like reads for the BASELINE.json configs.

Genome: i.i.d. uniform over ACGT (seed 42).  Reads (seed 43): start uniform, strand +/- with
p = 1/2 (reverse complement), length ~ Gamma(shape 2, mean `mean_len`) clipped to
[min_len, max_len], independent per-base errors at total rate `error` split sub:ins:del = 4:3:3,
until sum(len) >= coverage * genome_len.  Pure numpy; deterministic for given seeds.
"""
import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def make_genome(length, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    return ALPHABET[rng.integers(0, 4, size=int(length), dtype=np.uint8)]


def mutate(template, error, rng):
    """Apply sub/ins/del errors (4:3:3) to one template; returns the read as uint8 array."""
    n = template.size
    u = rng.random(n)
    p_sub, p_ins, p_del = 0.4 * error, 0.3 * error, 0.3 * error
    is_sub = u < p_sub
    is_ins = (u >= p_sub) & (u < p_sub + p_ins)
    is_del = (u >= p_sub + p_ins) & (u < p_sub + p_ins + p_del)
    base = template.copy()
    if is_sub.any():
        # substitute by one of the three other bases
        idx = np.searchsorted(ALPHABET_SORTED, base[is_sub])
        shift = rng.integers(1, 4, size=int(is_sub.sum()))
        base[is_sub] = ALPHABET_SORTED[(idx + shift) & 3]
    # output count per template position: deleted -> 0, insertion -> 2 (random base, then the base)
    cnt = np.ones(n, dtype=np.int64)
    cnt[is_del] = 0
    cnt[is_ins] = 2
    pos = np.cumsum(cnt) - cnt
    out = np.empty(int(cnt.sum()), dtype=np.uint8)
    keep = ~is_del
    out[(pos + cnt - 1)[keep]] = base[keep]
    if is_ins.any():
        out[pos[is_ins]] = ALPHABET[rng.integers(0, 4, size=int(is_ins.sum()))]
    return out


ALPHABET_SORTED = np.frombuffer(b"ACGT", dtype=np.uint8)  # already sorted by byte value


def make_reads(genome, coverage, error, seed=43, mean_len=10_000, min_len=500, max_len=100_000):
    """Returns (seq uint8 array, offsets uint64 array, names list[bytes])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = genome.size
    target = int(coverage * G)
    parts, lens, total = [], [], 0
    max_len = min(max_len, G)
    min_len = min(min_len, max_len)
    while total < target:
        L = int(np.clip(rng.gamma(2.0, mean_len / 2.0), min_len, max_len))
        start = int(rng.integers(0, G - L + 1))
        t = genome[start : start + L]
        if rng.random() < 0.5:
            t = _COMP[t[::-1]]
        r = mutate(t, error, rng)
        parts.append(r)
        lens.append(r.size)
        total += r.size
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    names = [b"read_%d" % i for i in range(len(parts))]
    return np.concatenate(parts), off, names
