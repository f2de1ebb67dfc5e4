#!/usr/bin/env python3
"""One GPU playing rank 0 of a multi-GPU leg: the set is built from ALL shards' reads (streamed construction,
one chunk per shard), then shard 0's reads are corrected against it.  Isolates what the size / density of the
set costs the correction kernels (no exchange involved).

    bigset_probe.py weak N [steps]                   the weak-scaled headline at N GPUs (N x 4.6 Mb genome, 30x, 10 %)
    bigset_probe.py shard GENOME_MB COV ERR N [steps]  one genome, reads sharded over N (configs[3]: 100 50 0.12 8)

BRGPU_LIBRARY / BRGPU_NO_POS8 select the variant."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import br_b200
from br_b200 import synth

mode = sys.argv[1]
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = br_b200.Context(0, stream=stream)
    if mode == "weak":
        N = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
        thr = synth.error_thresholds(0.10)
        descs = []
        for r in range(N):
            start, tlen, strand = synth.read_descriptors(4_600_000 * N, 30 / N, seed=43 + r)
            descs.append((43 + r, 0, start, tlen, strand))
        label = f"weak N={N}"
    else:
        genome, cov, err, N = int(float(sys.argv[2]) * 1e6), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
        steps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
        thr = synth.error_thresholds(err)
        start, tlen, strand = synth.read_descriptors(genome, cov, seed=43)
        descs = []
        for r in range(N):
            lo, hi = synth.shard_descriptors(tlen, N, r)
            descs.append((43, lo, start[lo:hi], tlen[lo:hi], strand[lo:hi]))
        label = f"shard {genome/1e6:.0f} Mb {cov}x {err} N={N}"
    first = br_b200.Reads.synth(ctx, 42, descs[0][0], descs[0][1], *descs[0][2:], thr)
    def chunks():
        yield first
        prev = None
        for d in descs[1:]:
            if prev is not None:
                prev.free()
            prev = br_b200.Reads.synth(ctx, 42, d[0], d[1], *d[2:], thr)
            yield prev
        if prev is not None:
            prev.free()
    solid = br_b200.Pcon.from_chunks(ctx, chunks(), 17, abundance=2)
    methods = br_b200.build_methods(["one", "two"], solid, 5, 7)
    def run():
        o = br_b200.correct_reads(methods, first); o.free()
    for _ in range(2): run()
    ctx.profile_enable(True); ctx.profile_reset(); run(); prof = ctx.profile(); ctx.profile_enable(False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps): run()
    b.record(stream); torch.cuda.synchronize()
    print(f"{label} lib={os.environ.get('BRGPU_LIBRARY','default').split('/')[-1]} no_pos8={os.environ.get('BRGPU_NO_POS8','0')}: "
          f"correction {a.elapsed_time(b)/steps:.3f} ms/step ({int(first.bases)/1e6:.0f} Mbases)  "
          + " ".join(f"{k}={v['ms']/max(1,v['launches']):.3f}" for k, v in prof.items() if k.startswith(("solid", "scan"))))
