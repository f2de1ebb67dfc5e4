import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import br_b200
from br_b200 import synth
stream = torch.cuda.Stream(); ctx = br_b200.Context(0, stream=stream)
genome = synth.make_genome(4_600_000, seed=42); seq, off, _ = synth.make_reads(genome, 30, 0.10, seed=43)
h_seq = torch.from_numpy(seq).pin_memory(); h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
n=off.size-1; nb=int(off[-1])
bufs=[(torch.empty(nb+nb//8+64*n+64,dtype=torch.uint8).pin_memory(), torch.empty(n+1,dtype=torch.int64).pin_memory()) for _ in range(2)]
def step(reads):
    solid = br_b200.Pcon.from_reads(ctx, reads, 17, abundance=2)
    out = br_b200.correct_reads(br_b200.build_methods(["one","two"], solid, 5, 7), reads)
    solid.free(); return out
def run(n_steps, overlap):
    nxt = br_b200.Reads.upload_async(ctx, h_seq, h_off); prev=None
    for i in range(n_steps):
        cur = nxt
        if i+1<n_steps:
            nxt = br_b200.Reads.upload_async(ctx, h_seq, h_off)
            if not overlap: torch.cuda.synchronize()
        if prev is not None:
            prev.download_async(*bufs[(i-1)%2])
            if not overlap: prev.download_wait()
        out = step(cur); cur.free()
        if prev is not None: prev.download_wait(); prev.free()
        prev = out
    prev.download_async(*bufs[(n_steps-1)%2]); prev.download_wait(); prev.free()
with torch.cuda.stream(stream):
    run(4, True)
    for overlap in (False, True):
        ctx.profile_reset(); ctx.profile_enable(True)
        run(6, overlap); torch.cuda.synchronize()
        prof = ctx.profile(); ctx.profile_enable(False)
        print("overlap", overlap, {k: round(v['ms']/v['launches'],3) for k,v in prof.items() if v['ms']/v['launches']>0.05})
