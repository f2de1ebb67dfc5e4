#!/usr/bin/env python3
"""Wall-clock breakdown of the sharded set construction per protocol step (run under torchrun)."""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as tdist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import br_b200
from br_b200 import dist as bdist, synth

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = br_b200.Context(local, stream=stream)
        if "weak" in sys.argv[1:]:  # the bench's weak-scaled headline: every rank holds 30x of the SAME 4.6 Mb genome
            genome = synth.make_genome(4_600_000, seed=42)
            seq, off, _ = synth.make_reads(genome, 30, 0.10, seed=43 + rank)
        else:  # a genome that grows with the box, 30x in total
            genome = synth.make_genome(4_600_000 * world, seed=42)
            seq, off, _ = synth.make_reads(genome, 30 / world, 0.10, seed=43 + rank)
        reads = br_b200.Reads.upload(ctx, seq, off)
        acc = {}
        class DistProxy:  # times the collectives (with a device synchronise on both sides) by name and payload
            def __init__(self, d):
                self._d = d
            def __getattr__(self, name):
                f = getattr(self._d, name)
                if name not in ("all_gather_into_tensor", "all_gather", "broadcast", "all_reduce"):
                    return f
                def g(*a, **k):
                    torch.cuda.synchronize(); t = time.perf_counter()
                    r = f(*a, **k)
                    torch.cuda.synchronize()
                    out = a[0]
                    n = sum(x.numel() * x.element_size() for x in out) if isinstance(out, (list, tuple)) else out.numel() * out.element_size()
                    key = f"  nccl {name} {'>=1MB' if n >= 1 << 20 else 'small'}"
                    acc[key] = acc.get(key, 0.0) + time.perf_counter() - t
                    return r
                return g
        class Timed(bdist.GpuOps):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                self.dist = DistProxy(self.dist)
        def wrap(name):
            orig = getattr(bdist.GpuOps, name)
            def f(self, *a, **k):
                torch.cuda.synchronize(); t = time.perf_counter()
                r = orig(self, *a, **k)
                torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
                return r
            setattr(Timed, name, f)
        for n in ("partition_local", "exchange_kmer_handles", "barrier", "open_peers", "count_range", "all_gather_bitfield", "finish"):
            wrap(n)
        for it in range(6):
            if it == 2:
                acc.clear()
            torch.cuda.synchronize(); tdist.barrier(); t0 = time.perf_counter()
            s = bdist.build_set_sharded(Timed(ctx, reads), 17, abundance=2)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            s.free()
            torch.cuda.synchronize(); t2 = time.perf_counter()
            if rank == 0:
                print(f"iter {it}: build {1e3*(t1-t0):.1f} ms, free {1e3*(t2-t1):.1f} ms", flush=True)
        if rank == 0:
            for k, v in acc.items():
                print(f"  {k:24s} {1e3*v/4:8.2f} ms/iter")
    tdist.barrier(); tdist.destroy_process_group()
main()
