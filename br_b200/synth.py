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


# ----------------------------------------------------------------------------------------------
# Counter-based generator: every byte is a pure function of (seed, counter), so the
# device kernel (csrc/synth_kernels.cu, brgpu_reads_synth) and this numpy mirror produce the same
# reads — a GPU can generate its shard of a 1 Gb x 30 data set without the host ever holding it,
# and the CPU reference arm can generate any sample of the same data set.  This module is plain
# numpy on purpose: bench.py's reference arm loads it by path, without the br_b200 package.
# ----------------------------------------------------------------------------------------------
_M = np.uint64
GENOME_MUL = _M(0xD1342543DE82EF95)
READ_MUL = _M(0xA24BAED4963EE407)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mix64(x):
    """splitmix64's finaliser over x + golden ratio (uint64 arrays, wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=_M) + _M(0x9E3779B97F4A7C15)
        z = (x ^ (x >> _M(30))) * _M(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _M(27))) * _M(0x94D049BB133111EB)
        return z ^ (z >> _M(31))


def genome_codes(genome_seed, pos):
    """Index into "ACGT" of the genome bases at `pos` (uint64 array); the genome is never stored."""
    with np.errstate(over="ignore"):
        return (mix64(_M(genome_seed) * GENOME_MUL + np.asarray(pos, dtype=_M)) >> _M(62)).astype(np.uint8)


def error_thresholds(error):
    """Cumulative 24-bit thresholds (substitution, insertion, deletion) for a total per-base error
    rate split sub:ins:del = 4:3:3."""
    one = 1 << 24
    t_sub = int(round(0.4 * error * one))
    t_ins = t_sub + int(round(0.3 * error * one))
    t_del = t_ins + int(round(0.3 * error * one))
    return np.array([t_sub, t_ins, t_del], dtype=np.uint32)


def read_descriptors(genome_len, coverage, seed=43, mean_len=10_000, min_len=500, max_len=100_000):
    """(start uint64, template length uint32, strand uint8) of the reads of a data set: lengths ~
    Gamma(shape 2, mean `mean_len`) clipped to [min_len, max_len], starts uniform, strands fair, as many
    reads as it takes for the template lengths to sum to coverage x genome_len (the expected read length
    equals the template length: insertions and deletions are equally likely)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = int(genome_len)
    max_len = min(max_len, G)
    min_len = min(min_len, max_len)
    target = int(coverage * G)
    lens, total = [], 0
    while total < target:
        block = np.clip(rng.gamma(2.0, mean_len / 2.0, size=max(1024, int(1.1 * (target - total) / mean_len))), min_len, max_len).astype(np.int64)
        cs = np.cumsum(block) + total
        n = int(np.searchsorted(cs, target, side="left")) + 1  # first read that reaches the target is kept
        lens.append(block[:n])
        total = int(cs[min(n, block.size) - 1])
    tlen = np.concatenate(lens)
    start = np.floor(rng.random(tlen.size) * (G - tlen + 1)).astype(np.uint64)
    strand = (rng.random(tlen.size) < 0.5).astype(np.uint8)
    return start, tlen.astype(np.uint32), strand


def shard_descriptors(tlen, world, rank):
    """Contiguous read range [lo, hi) of `rank`, balanced by template bases."""
    cs = np.concatenate([[0], np.cumsum(tlen.astype(np.int64))])
    total = int(cs[-1])
    lo = int(np.searchsorted(cs, total * rank // world, side="left"))
    hi = int(np.searchsorted(cs, total * (rank + 1) // world, side="left")) if rank < world - 1 else tlen.size
    return min(lo, tlen.size), min(hi, tlen.size)


def host_reads(genome_seed, read_seed, first_read_id, start, tlen, strand, thresholds):
    """numpy mirror of brgpu_reads_synth: returns (seq uint8, offsets uint64)."""
    t_sub, t_ins, t_del = (int(x) for x in thresholds)
    parts, lens = [], []
    with np.errstate(over="ignore"):
        for r in range(len(tlen)):
            L = int(tlen[r])
            t = np.arange(L, dtype=_M)
            gpos = (_M(int(start[r]) + L - 1) - t) if strand[r] else (_M(int(start[r])) + t)
            b = genome_codes(genome_seed, gpos)
            if strand[r]:
                b = np.uint8(3) - b
            key = mix64(_M(read_seed) * READ_MUL + _M(first_read_id + r))
            h = mix64(key + t)
            u = (h & _M(0xFFFFFF)).astype(np.int64)
            is_sub = u < t_sub
            is_ins = (u >= t_sub) & (u < t_ins)
            is_del = (u >= t_ins) & (u < t_del)
            base = b.copy()
            shift = (((h >> _M(24)) & _M(0xFFFF)) % _M(3)).astype(np.uint8)
            base[is_sub] = (b[is_sub] + np.uint8(1) + shift[is_sub]) & np.uint8(3)
            cnt = np.ones(L, dtype=np.int64)
            cnt[is_del] = 0
            cnt[is_ins] = 2
            pos = np.cumsum(cnt) - cnt
            out = np.empty(int(cnt.sum()), dtype=np.uint8)
            keep = ~is_del
            out[(pos + cnt - 1)[keep]] = _ACGT[base[keep]]
            out[pos[is_ins]] = _ACGT[((h[is_ins] >> _M(40)) & _M(3)).astype(np.uint8)]
            parts.append(out)
            lens.append(out.size)
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    return (np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)), off


def host_genome(genome_seed, length):
    """The genome as ASCII (tests; small genomes only)."""
    return _ACGT[genome_codes(genome_seed, np.arange(int(length), dtype=_M))]
