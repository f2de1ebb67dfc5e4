#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank builds (a) the sharded set (own shard counted locally, saturating merge of its slice
over NVLink peer memory, NCCL all-gather of the bitfield) and (b) the single-GPU set from all the
reads; the two bitfields must be byte-identical on every rank, for an explicit threshold and for
first-minimum, and each rank's corrected shard must equal the same records corrected against (b).
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as tdist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import br_b200  # noqa: E402
from br_b200 import dist as bdist, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    ok = True
    with torch.cuda.stream(stream):
        ctx = br_b200.Context(local, stream=stream)
        genome = synth.make_genome(400_000, seed=42)
        seq, off, _ = synth.make_reads(genome, 25, 0.08, seed=43, mean_len=3000)
        lo, hi = bdist.shard_records(off, world, rank)
        sub = off[lo : hi + 1]
        my_seq, my_off = seq[int(sub[0]) : int(sub[-1])], sub - sub[0]
        for k, kwargs in ((17, {"abundance": 2}), (15, {"abundance_selection": "first-minimum"}), (13, {"abundance_selection": "first-minimum"})):
            mine = br_b200.Reads.upload(ctx, my_seq, my_off)
            sharded = bdist.build_set_sharded(bdist.GpuOps(ctx, mine), k, **kwargs)
            single = br_b200.Pcon.from_reads(ctx, (seq, off), k, **kwargs)
            same = np.array_equal(sharded.bitfield(), single.bitfield()) and sharded.abundance == single.abundance
            methods_a = br_b200.build_methods(["one", "gap_size"], sharded)
            methods_b = br_b200.build_methods(["one", "gap_size"], single)
            a, ao = br_b200.correct_reads(methods_a, mine).download()
            b, bo = br_b200.correct_batch(methods_b, my_seq, my_off)
            same_corr = np.array_equal(ao, bo) and np.array_equal(a, b)
            print(f"rank {rank}/{world} k={k} {kwargs}: abundance {sharded.abundance}, bitfield identical: {same}, "
                  f"corrected shard identical: {same_corr} ({hi - lo} reads)", flush=True)
            ok = ok and same and same_corr
        # a shard held in three chunks (how a shard larger than 2^32 slot bytes is processed): every chunk is
        # partitioned on its own, the owned bucket range is counted over all local and all peer partitions
        cuts = [0, (hi - lo) // 5, (hi - lo) // 2, hi - lo]
        for k, kwargs in ((17, {"abundance": 2}), (15, {"abundance_selection": "first-minimum"})):
            chunks = [br_b200.Reads.upload(ctx, my_seq, my_off[a : b + 1]) for a, b in zip(cuts[:-1], cuts[1:])]
            sharded = bdist.build_set_sharded(bdist.GpuOps(ctx, chunks), k, **kwargs)
            single = br_b200.Pcon.from_reads(ctx, (seq, off), k, **kwargs)
            same = np.array_equal(sharded.bitfield(), single.bitfield()) and sharded.abundance == single.abundance
            a = np.concatenate([br_b200.correct_reads(br_b200.build_methods(["one", "two"], sharded), c).download()[0] for c in chunks])
            b, _ = br_b200.correct_batch(br_b200.build_methods(["one", "two"], single), my_seq, my_off)
            same_corr = np.array_equal(a, b)
            print(f"rank {rank}/{world} k={k} {kwargs} 3 chunks per rank: bitfield identical: {same}, corrected shard identical: {same_corr}", flush=True)
            ok = ok and same and same_corr
    t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
    tdist.barrier()
    tdist.destroy_process_group()
    if int(t.item()) != 1:
        raise SystemExit(1)
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
