#!/usr/bin/env python3
"""Where a step's time goes, phase by phase (diagnostic, not the benchmark).

    python profiles/breakdown.py [--genome 4600000] [--methods one two]

Prints host-wall and CUDA-event times of: upload, set construction, correction (forward only and
with the reversed pass), download — each bracketed by synchronisation — and the per-kernel
CUDA-event table of the forward-only and the full chain, so that forward and reversed passes
can be told apart.
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import br_b200  # noqa: E402
from br_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--coverage", type=float, default=30)
    ap.add_argument("--error", type=float, default=0.10)
    ap.add_argument("--k", type=int, default=17)
    ap.add_argument("--methods", nargs="+", default=["one", "two"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-step-sync", action="store_true", help="no device synchronisation between e2e steps (as in bench.py)")
    ap.add_argument("--smi", action="store_true", help="poll nvidia-smi every 200 ms during the e2e loop, like bench.py")
    a = ap.parse_args()

    stream = torch.cuda.Stream()
    ctx = br_b200.Context(0, stream=stream)
    genome = synth.make_genome(a.genome, seed=42)
    seq, off, _ = synth.make_reads(genome, a.coverage, a.error, seed=43)
    n_bases = int(off[-1])
    n_reads = off.size - 1
    h_seq = torch.from_numpy(seq).pin_memory()
    h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    h_out = torch.empty(n_bases + n_bases // 8 + 64 * n_reads + 64, dtype=torch.uint8).pin_memory()
    h_out_off = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()

    def timed(fn, reps=a.reps):
        best_wall, best_ev = 1e9, 1e9
        res = None
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            res = fn()
            e1.record(stream)
            torch.cuda.synchronize()
            best_wall = min(best_wall, (time.perf_counter() - t0) * 1e3)
            best_ev = min(best_ev, e0.elapsed_time(e1))
        return res, best_wall, best_ev

    with torch.cuda.stream(stream):
        state = {}

        def up():
            if "reads" in state:
                state["reads"].free()
            state["reads"] = br_b200.Reads.upload(ctx, h_seq, h_off)

        def mkset():
            if "solid" in state:
                state["solid"].free()
            state["solid"] = br_b200.Pcon.from_reads(ctx, state["reads"], a.k, abundance=2)

        def corr(two_side):
            def f():
                if "out" in state:
                    state["out"].free()
                m = br_b200.build_methods(a.methods, state["solid"], 5, 7)
                state["out"] = br_b200.correct_reads(m, state["reads"], two_side=two_side)
            return f

        def down():
            state["out"].download(h_out, h_out_off)

        print(f"reads {n_reads}  bases {n_bases}  k {a.k}  methods {a.methods}")
        for name, fn in [("upload", up), ("set construction", mkset), ("correct forward only", corr(True)),
                         ("correct fwd+reversed", corr(False)), ("download", down)]:
            _, w, e = timed(fn)
            print(f"{name:24s} wall {w:8.3f} ms   events {e:8.3f} ms")

        # the e2e step of bench.py, phase by phase on the host clock (no extra synchronisation)
        for key in ("reads", "solid", "out"):
            state.pop(key).free()
        torch.cuda.synchronize()
        smi = None
        if a.smi:
            import subprocess
            smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks_event_reasons.active", "--format=csv,noheader",
                                    "-lms", "200", "-i", "0"], stdout=subprocess.DEVNULL)
            time.sleep(2.0)
        for it in range(a.e2e_steps):
            t = [time.perf_counter()]
            reads = br_b200.Reads.upload(ctx, h_seq, h_off)
            t.append(time.perf_counter())
            solid = br_b200.Pcon.from_reads(ctx, reads, a.k, abundance=2)
            t.append(time.perf_counter())
            out = br_b200.correct_reads(br_b200.build_methods(a.methods, solid, 5, 7), reads)
            t.append(time.perf_counter())
            out.download(h_out, h_out_off)
            t.append(time.perf_counter())
            out.free()
            solid.free()
            reads.free()
            t.append(time.perf_counter())
            if not a.no_step_sync:
                torch.cuda.synchronize()
            t.append(time.perf_counter())
            d = [(t[j + 1] - t[j]) * 1e3 for j in range(len(t) - 1)]
            if a.e2e_steps <= 8 or (t[-1] - t[0]) * 1e3 > 30.0:
                print("e2e step %d: upload %.2f  set %.2f  correct %.2f  download %.2f  free %.2f  sync %.2f  total %.2f ms"
                      % (it, *d, (t[-1] - t[0]) * 1e3))
        if smi is not None:
            smi.terminate()
        state["reads"] = br_b200.Reads.upload(ctx, h_seq, h_off)
        state["solid"] = br_b200.Pcon.from_reads(ctx, state["reads"], a.k, abundance=2)

        for label, ts in [("forward only", True), ("forward + reversed", False)]:
            ctx.profile_reset()
            ctx.profile_enable(True)
            l0 = ctx.scan_lookups
            corr(ts)()
            torch.cuda.synchronize()
            prof = ctx.profile()
            ctx.profile_enable(False)
            print(f"-- per kernel, {label}: scan lookups {ctx.scan_lookups - l0}")
            for kname, p in prof.items():
                print(f"   {kname:20s} launches {p['launches']:3d}  total {p['ms']:8.3f} ms")
    ctx.close()


if __name__ == "__main__":
    main()
