#!/usr/bin/env python3
"""Re-pack the reference's two data fixtures into tests/golden/ (build container only).

  /root/reference/tests/data/raw.fasta          -> br_reads.fa.gz   (206 Badread-style reads)
  /root/reference/tests/data/raw.k11.a2.solid   -> br_reads.k11.a2.solid
      (pcon `.solid` container: gzip( u8 k || bitfield ), k=11, abundance 2)

They are data the reference's own integration tests run on (tests/br.rs:8-59), not
source code.  The payload bytes are unchanged; only the gzip framing is re-done
(mtime=0, level 9) so the files are reproducible.  A manifest with sizes, sha256 of
the *uncompressed* payloads and the bitfield popcount is written beside them and is
what the tests pin against.
"""
import gzip, hashlib, json
from pathlib import Path

SRC = Path("/root/reference/tests/data")
DST = Path(__file__).parent

fa = (SRC / "raw.fasta").read_bytes()
solid = gzip.decompress((SRC / "raw.k11.a2.solid").read_bytes())
assert solid[0] == 11 and len(solid) == 1 + (1 << 21) // 8

with open(DST / "br_reads.fa.gz", "wb") as f:
    with gzip.GzipFile(fileobj=f, mode="wb", compresslevel=9, mtime=0, filename="") as g:
        g.write(fa)
with open(DST / "br_reads.k11.a2.solid", "wb") as f:
    with gzip.GzipFile(fileobj=f, mode="wb", compresslevel=9, mtime=0, filename="") as g:
        g.write(solid)

n_reads = fa.count(b">")
n_bases = sum(len(l) for l in fa.split(b"\n") if l and not l.startswith(b">"))
manifest = {
    "br_reads.fa": {"bytes": len(fa), "sha256": hashlib.sha256(fa).hexdigest(), "reads": n_reads, "bases": n_bases},
    "br_reads.k11.a2.solid": {
        "payload_bytes": len(solid),
        "sha256": hashlib.sha256(solid).hexdigest(),
        "k": solid[0],
        "popcount": sum(bin(b).count("1") for b in solid[1:]),
    },
}
(DST / "fixtures.json").write_text(json.dumps(manifest, indent=1) + "\n")
print(json.dumps(manifest, indent=1))
