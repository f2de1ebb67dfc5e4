import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench, br_b200
from br_b200.runtime import pack_2bit
synth = bench.load_synth()
stream = torch.cuda.Stream()
ctx = br_b200.Context(0, stream=stream)
with torch.cuda.stream(stream):
    d = bench.headline_descriptors(synth, 1, 0, bench.GENOME_PER_GPU)
    reads = br_b200.Reads.synth(ctx, d["genome_seed"], d["read_seed"], d["first"], d["start"], d["tlen"], d["strand"], d["thr"])
    seq, off = reads.download()
    packed, ep, eb = pack_2bit(seq)
    hp = torch.from_numpy(packed).pin_memory(); ho = torch.from_numpy(off.view(np.int64)).pin_memory()
    outp = (torch.empty(seq.size // 3, dtype=torch.uint8).pin_memory(), torch.empty(off.size, dtype=torch.int64).pin_memory(),
            torch.empty(16, dtype=torch.int64).pin_memory(), torch.empty(16, dtype=torch.uint8).pin_memory(), torch.zeros(2, dtype=torch.int64).pin_memory())
    solid = br_b200.Pcon.from_reads(ctx, reads, 17, abundance=2)
    m = br_b200.build_methods(["one", "two"], solid, 5, 7)
    out = br_b200.correct_reads(m, reads)
    tu, td, tw = [], [], []
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); r = br_b200.Reads.upload_packed(ctx, hp, ho, ep, eb, asynchronous=True); t1 = time.perf_counter()
        out.download_packed(*outp, asynchronous=True); t2 = time.perf_counter()
        out.download_wait(); t3 = time.perf_counter()
        torch.cuda.synchronize(); t4 = time.perf_counter()
        tu.append((t1 - t0) * 1e3); td.append((t2 - t1) * 1e3); tw.append((t4 - t2) * 1e3)
        r.free()
    print("upload_packed_async host ms:", [round(x, 3) for x in tu])
    print("download_packed_async host ms:", [round(x, 3) for x in td])
    print("wait+sync ms:", [round(x, 3) for x in tw])
    # per-kernel times of the staging path (profiling on): upload -> unpack, download -> pack
    ctx.profile_reset(); ctx.profile_enable(True)
    for it in range(4):
        r = br_b200.Reads.upload_packed(ctx, hp, ho, ep, eb)
        o2 = br_b200.correct_reads(m, r)
        o2.download_packed(*outp)
        r.free(); o2.free()
    prof = ctx.profile(); ctx.profile_enable(False)
    for k, v in prof.items():
        if not (k.startswith("scan_") or k.startswith("merge_") or k.startswith("solid_")):
            print(f"  {k:20s} x{v['launches']/4:.0f} {v['ms']/max(1,v['launches']):.4f} ms")
