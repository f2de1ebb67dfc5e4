#!/usr/bin/env python3
"""One GPU playing rank 0 of the weak-scaled N-GPU headline: the set is built from the reads of all N ranks
(N x 4.6 Mb genome, streamed construction over N chunks), then rank 0's reads are corrected against it.
Isolates what the larger set costs the correction kernels (no exchange involved).
usage: bigset_probe.py N [steps]   (BRGPU_LIBRARY / BRGPU_NO_POS8 select the variant)"""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import br_b200
from br_b200 import synth

N = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = br_b200.Context(0, stream=stream)
    thr = synth.error_thresholds(0.10)
    chunks = []
    for r in range(N):
        start, tlen, strand = synth.read_descriptors(4_600_000 * N, 30 / N, seed=43 + r)
        chunks.append(br_b200.Reads.synth(ctx, 42, 43 + r, 0, start, tlen, strand, thr))
    solid = br_b200.Pcon.from_chunks(ctx, chunks, 17, abundance=2)
    methods = br_b200.build_methods(["one", "two"], solid, 5, 7)
    def run():
        o = br_b200.correct_reads(methods, chunks[0]); o.free()
    for _ in range(2): run()
    ctx.profile_enable(True); ctx.profile_reset(); run(); prof = ctx.profile(); ctx.profile_enable(False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps): run()
    b.record(stream); torch.cuda.synchronize()
    print(f"N={N} lib={os.environ.get('BRGPU_LIBRARY','default').split('/')[-1]} no_pos8={os.environ.get('BRGPU_NO_POS8','0')}: "
          f"correction {a.elapsed_time(b)/steps:.3f} ms/step ({int(chunks[0].bases)/1e6:.0f} Mbases)  "
          + " ".join(f"{k}={v['ms']/max(1,v['launches']):.3f}" for k, v in prof.items() if k.startswith(("solid", "scan"))))
