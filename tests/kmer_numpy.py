"""Pure-numpy canonical k-mer indices: a second opinion independent of the C++ oracle."""
import numpy as np


def numpy_canonical_indices(seq, off, k):
    """All canonical-k-mer table indices (canonical >> 1) of all reads, pure numpy (second opinion)."""
    code = ((seq >> 1) & 3).astype(np.uint64)
    out = []
    for r in range(off.size - 1):
        c = code[int(off[r]) : int(off[r + 1])]
        n = c.size - k + 1
        if n <= 0:
            continue
        fwd = np.zeros(n, dtype=np.uint64)
        rev = np.zeros(n, dtype=np.uint64)
        for t in range(k):
            fwd = (fwd << np.uint64(2)) | c[t : t + n]
            rev = rev | ((c[t : t + n] ^ np.uint64(2)) << np.uint64(2 * t))
        par = np.zeros(n, dtype=np.uint64)
        x = fwd.copy()
        for _ in range(2 * k):
            par ^= x & np.uint64(1)
            x >>= np.uint64(1)
        out.append(np.where(par == 0, fwd, rev) >> np.uint64(1))
    return np.concatenate(out).astype(np.int64)
