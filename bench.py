#!/usr/bin/env python3
"""bench.py — corrected bases/s of the br hot path (count -> threshold -> correct) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the whole hot path over one batch of synthetic reads:
    Counter::new + count + spectrum + threshold (src/main.rs:72-115)  then
    run_correction's chunk body with methods `one two`, confirm 5, reversed pass on
    (src/lib.rs:44-55) over every read.
Workload at N = 1: BASELINE.json configs[1] — synthetic 4.6 Mb genome (seed 42), 30x ONT-like
reads (seed 43), 10 % error, k = 17, -a 2.  At N > 1 the genome is N x 4.6 Mb and every rank owns
30 x 4.6 Mb of reads (weak scaling): per-rank count tables are merged with the saturating
reduce-scatter over NVLink peer memory, the bitfield slices are all-gathered with NCCL, and each
rank corrects its own reads.

`value`  : bases/s with the reads already resident in HBM (device-resident handles in and out).
`e2e`    : the same metric through the host-buffer calls — pinned host reads are copied to the
           device and the corrected reads are copied back inside the timed region, every step.
`roofline`: the dominant kernel's achieved algorithmic bytes/s over measured HBM bandwidth,
           timed with CUDA events on the launching stream inside the timed region.
`cpu_baseline` / `--impl reference`: the CPU restatement of br (oracle/, C++ + OpenMP on all host
           cores; the Rust reference cannot be built in this image) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K = 17
ABUNDANCE = 2
METHODS = ["one", "two"]
CONFIRM = 5
MAX_SEARCH = 7
GENOME_PER_GPU = 4_600_000
COVERAGE = 30
ERROR = 0.10
METRIC = "corrected bases/sec (count + threshold + correct, whole pipeline)"
UNIT = "bases/s"


def workload_name(n_gpus, genome_per_gpu):
    g = genome_per_gpu * n_gpus / 1e6
    return (f"synthetic {g:.1f} Mb genome (seed 42), {COVERAGE}x ONT-like reads (seed 43+rank), "
            f"{int(ERROR * 100)}% error, k={K}, -a {ABUNDANCE}, methods {'+'.join(METHODS)}, confirm {CONFIRM}, "
            f"reversed pass on")


def make_shard(n_gpus, rank, genome_per_gpu):
    from br_b200 import synth

    genome = synth.make_genome(genome_per_gpu * n_gpus, seed=42)
    # every rank draws reads from the whole genome; its share is COVERAGE x genome_per_gpu bases
    seq, off, _ = synth.make_reads(genome, COVERAGE / n_gpus, ERROR, seed=43 + rank)
    return seq, off


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        # nvidia-smi's start-up (NVML initialisation) holds driver locks for a second or two; wait
        # for its first sample so that none of that lands in a timed region
        t0 = time.time()
        while time.time() - t0 < 10.0 and os.path.getsize(self.tmp.name) == 0 and self.proc.poll() is None:
            time.sleep(0.05)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.tmp.flush()
        rows = []
        try:
            for line in open(self.tmp.name):
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    rows.append(f)
        finally:
            os.unlink(self.tmp.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[5 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of br) with OpenMP on all host cores
# ----------------------------------------------------------------------------------------------
def cpu_pipeline_sample(seq, off, total_bases_all_shards, n_shards, threads, correct_target_s=6.0):
    """Times the CPU restatement on a bounded sample of the workload and extrapolates linearly
    in the parts that are linear (counting after the table is touched; correction over reads).
    Returns (bases/s for the whole workload, description of the sample)."""
    from oracle import br_oracle as o

    n = off.size - 1
    t0 = time.perf_counter()
    c = o.Counter(K)
    c.count(seq, off, threads=threads)          # this shard, cold table: page faults included
    t_count_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    hist = c.spectrum(threads)                   # Spectrum::from_count (src/main.rs:93)
    solid = c.to_solid(ABUNDANCE, threads)       # Solid::from_count (src/main.rs:112-114)
    t_passes = time.perf_counter() - t0
    del hist
    t_count_warm = 0.0
    if n_shards > 1:                             # the other shards hit a touched table
        t0 = time.perf_counter()
        c.count(seq, off, threads=threads)
        t_count_warm = time.perf_counter() - t0
    del c
    # correction: grow the sample until it costs about correct_target_s
    ids = [o.METHOD_IDS[m] for m in METHODS]
    ns = max(1, n // 64)
    while True:
        t0 = time.perf_counter()
        solid.run_correction(ids, seq, off[: ns + 1], confirm=CONFIRM, max_search=MAX_SEARCH, two_side=False, threads=threads)
        t_corr = time.perf_counter() - t0
        if t_corr >= correct_target_s / 2 or ns == n:
            break
        ns = min(n, max(ns * 2, int(ns * correct_target_s / max(t_corr, 1e-3))))
    sample_bases = int(off[ns]) - int(off[0])
    t_full = t_count_cold + (n_shards - 1) * t_count_warm + t_passes + t_corr * (total_bases_all_shards / sample_bases)
    desc = (f"count of one {int(off[-1]) / 1e6:.0f} Mbase shard into a fresh 2^{2 * K - 1}-counter table ({t_count_cold:.2f} s"
            f"{', warm recount %.2f s x %d' % (t_count_warm, n_shards - 1) if n_shards > 1 else ''}) + spectrum and threshold "
            f"passes ({t_passes:.2f} s) measured in full; correction measured on the first {ns} reads "
            f"({sample_bases / 1e6:.1f} Mbases, {t_corr:.2f} s) and scaled linearly to {total_bases_all_shards / 1e6:.0f} Mbases")
    return total_bases_all_shards / t_full, desc


def run_reference(args, world, rank):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import br_oracle as o

    o.build()
    threads = o.max_threads()
    seq, off = make_shard(args.gpus, 0, args.genome_per_gpu)
    total = int(off[-1]) * args.gpus
    vals, desc = [], ""
    for it in range(args.warmup + args.steps):
        v, desc = cpu_pipeline_sample(seq, off, total, args.gpus, threads)
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, args.genome_per_gpu)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of br (oracle/, C++17 + OpenMP); the Rust reference cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def algo_bytes(name, prof, n_bases, n_kmers, table, scan_lookups_per_step):
    """SURVEY §8(d) algorithmic bytes per launch for each kernel of the step."""
    if name == "count_kmers":
        return 0.25 * n_bases + 64.0 * n_kmers
    if name == "zero_counts":
        return float(table)
    if name == "spectrum_threshold":
        return table * 1.125
    if name == "spectrum":
        return float(table)
    if name == "solid_bitmap":
        return 32.0 * n_kmers + 0.25 * n_bases + n_bases / 8.0
    if name.startswith("scan_"):
        # read + write of the ASCII bases and one 32 B sector per KmerSet::get the scan issued
        return 2.0 * n_bases + 32.0 * scan_lookups_per_step / max(1, sum(1 for k in prof if k.startswith("scan_")))
    return prof[name]["algo_bytes"] / max(1, prof[name]["launches"])


def run_ours(args, world, rank, local_rank):
    import torch

    import br_b200
    from br_b200 import dist as bdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; br_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from br_b200.runtime import bind_to_gpu_numa_node

    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation
    tdist = None
    if world > 1:
        import torch.distributed as tdist

        tdist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    stream = torch.cuda.Stream()
    ctx = br_b200.Context(local_rank, stream=stream)

    seq, off = make_shard(world, rank, args.genome_per_gpu)
    n_reads = off.size - 1
    n_bases = int(off[-1])
    lens = np.diff(off.astype(np.int64))
    n_kmers = int(np.maximum(lens - K + 1, 0).sum())
    table = 1 << (2 * K - 1)

    # pinned host buffers (what a host application would hand to the C ABI)
    h_seq = torch.from_numpy(seq).pin_memory()
    h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    out_cap = n_bases + n_bases // 8 + 64 * n_reads + 64
    h_out = torch.empty(out_cap, dtype=torch.uint8).pin_memory()
    h_out_off = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()

    def build_set(reads, c=None):
        c = c or ctx
        if world == 1:
            return br_b200.Pcon.from_reads(c, reads, K, abundance=ABUNDANCE)
        return bdist.build_set_sharded(bdist.GpuOps(c, reads), K, abundance=ABUNDANCE)

    def step_device(reads):
        solid = build_set(reads)
        out = br_b200.correct_reads(br_b200.build_methods(METHODS, solid, CONFIRM, MAX_SEARCH), reads)
        out.free()
        solid.free()

    # e2e: every step uploads its inputs from pinned host memory and downloads its result inside
    # the timed region, through the host-buffer API.  The steps are pipelined — by default on one
    # context whose copy stream carries step i+1's upload and step i-1's download while step i's
    # kernels run (`stream`; one host thread, so the collectives of the sharded set construction
    # keep one order per rank); at N = 1 alternatively on several contexts driven by host threads
    # (`lanes`); `serial` = one step at a time.
    e2e_mode = "serial" if args.no_e2e_pipeline else args.e2e_mode
    if world > 1 and e2e_mode == "lanes":
        e2e_mode = "stream"  # host threads would issue the collectives of two steps in no fixed order
    out_bufs = [(h_out, h_out_off)]
    if e2e_mode == "stream":
        out_bufs.append((torch.empty_like(h_out).pin_memory(), torch.empty_like(h_out_off).pin_memory()))
    lanes = [(ctx, h_out, h_out_off)]
    if e2e_mode == "lanes":
        for _ in range(max(1, args.e2e_lanes) - 1):
            cx = br_b200.Context(local_rank, stream=torch.cuda.Stream())
            lanes.append((cx, torch.empty_like(h_out).pin_memory(), torch.empty_like(h_out_off).pin_memory()))

    def step_e2e(lane=0):
        c, o, oo = lanes[lane]
        reads = br_b200.Reads.upload(c, h_seq, h_off)            # H2D every step
        solid = build_set(reads, c)
        out = br_b200.correct_reads(br_b200.build_methods(METHODS, solid, CONFIRM, MAX_SEARCH), reads)
        d, _ = out.download(o, oo)                                # D2H every step
        nbytes = int(d.numel())
        out.free()
        solid.free()
        reads.free()
        return nbytes

    def run_e2e_stream(n_steps, stamps=None):
        """n_steps e2e steps on ONE context with the asynchronous staging calls: step i+1's reads go
        up and step i-1's result comes down on the context's copy stream while step i's kernels run.
        Every step still uploads its own inputs and downloads its own result."""
        c = ctx
        res = 0
        nxt = br_b200.Reads.upload_async(c, h_seq, h_off)
        prev = None
        t0 = time.perf_counter()
        for i in range(n_steps):
            cur = nxt
            if i + 1 < n_steps:
                nxt = br_b200.Reads.upload_async(c, h_seq, h_off)           # H2D of the next step
            if prev is not None:
                res = max(res, prev.download_async(*out_bufs[(i - 1) % 2]))  # D2H of the previous step
            solid = build_set(cur, c)
            out = br_b200.correct_reads(br_b200.build_methods(METHODS, solid, CONFIRM, MAX_SEARCH), cur)
            solid.free()
            cur.free()
            if prev is not None:
                prev.download_wait()
                prev.free()
            prev = out
            if stamps is not None:
                t1 = time.perf_counter()
                stamps.append(round((t1 - t0) * 1e3, 2))
                t0 = t1
        res = max(res, prev.download_async(*out_bufs[(n_steps - 1) % 2]))
        prev.download_wait()
        prev.free()
        return res

    def run_e2e(n_steps, stamps=None):
        """n_steps e2e steps spread over the lanes; returns the D2H bytes of one step."""
        if e2e_mode == "stream":
            return run_e2e_stream(n_steps, stamps)
        res = [0] * len(lanes)

        def work(lane, n):
            for _ in range(n):
                t0 = time.perf_counter()
                res[lane] = step_e2e(lane)
                if stamps is not None:
                    stamps.append(round((time.perf_counter() - t0) * 1e3, 2))

        share = [n_steps // len(lanes) + (1 if i < n_steps % len(lanes) else 0) for i in range(len(lanes))]
        if len(lanes) == 1:
            work(0, n_steps)
        else:
            import threading

            ts = [threading.Thread(target=work, args=(i, share[i])) for i in range(len(lanes))]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return max(res)

    def barrier():
        torch.cuda.synchronize()
        if tdist is not None:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if tdist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # a full collection of the interpreter's heap (torch imports ~10^6 objects) costs 0.1-0.5 s and
    # would land inside a timed region: collect now, then keep the collector off while timing
    import gc

    gc.collect()
    gc.freeze()
    gc.disable()
    with torch.cuda.stream(stream):
        dev_reads = br_b200.Reads.upload(ctx, h_seq, h_off)
        # the sampler is started before the warm-up so that nvidia-smi's own start-up (NVML
        # initialisation takes driver locks) is over before the timed regions begin
        sampler = ClockSampler(local_rank)
        if rank == 0 and not os.environ.get("BRGPU_BENCH_NO_SAMPLER"):
            sampler.start()
        for _ in range(args.warmup):
            step_device(dev_reads)
        # ---- device-resident timed region: `value` ----
        launches0 = ctx.launch_count
        ms_dev = timed(lambda: step_device(dev_reads), args.steps)
        launches = ctx.launch_count - launches0
        # ---- the same K steps again with a CUDA-event pair around every kernel (per-kernel
        # durations for the roofline; the event records perturb the step a little, which is why
        # `value` is taken from the undisturbed region above) ----
        ctx.profile_reset()
        ctx.profile_enable(True)
        lookups0 = ctx.scan_lookups
        ms_prof = timed(lambda: step_device(dev_reads), args.steps)
        prof = ctx.profile()
        ctx.profile_enable(False)
        lookups = (ctx.scan_lookups - lookups0) / args.steps
        # ---- end-to-end timed region (host buffers in and out) ----
        # warm-up: at least W steps per lane, then until two consecutive rounds agree within 3 % (at
        # most 12 more): the first e2e steps size the allocator's cache for the upload/download buffers
        d2h = 0
        e2e_warm_ms = []
        for it in range(max(3, args.warmup) + 12):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_warm = 4 if e2e_mode == "stream" else len(lanes)
            d2h = max(d2h, run_e2e(n_warm))
            torch.cuda.synchronize()
            e2e_warm_ms.append((time.perf_counter() - t0) * 1e3 / n_warm)
            stable = it + 1 >= max(3, args.warmup) and abs(e2e_warm_ms[-1] - e2e_warm_ms[-2]) <= 0.03 * e2e_warm_ms[-1]
            if tdist is not None:  # the step contains collectives: every rank must take the same decision
                t = torch.tensor([1 if stable else 0], dtype=torch.int32, device=f"cuda:{local_rank}")
                tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
                stable = bool(t.item())
            if stable:
                break
        e2e_step_ms = []  # host clock, whole steps (every e2e step ends in a synchronising download)
        ms_e2e = timed(lambda: run_e2e(args.steps, e2e_step_ms), 1)
        clocks = sampler.stop() if rank == 0 else None

    total_bases = n_bases
    if tdist is not None:
        t = torch.tensor([n_bases], dtype=torch.int64, device=f"cuda:{local_rank}")
        tdist.all_reduce(t)
        total_bases = int(t.item())
    if rank != 0:
        if tdist is not None:
            tdist.barrier()
            tdist.destroy_process_group()
        return

    value = total_bases * args.steps / (ms_dev * 1e-3)
    e2e_value = total_bases * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = peaks()
    kernels = {}
    tot_kernel_ms = sum(p["ms"] for p in prof.values())
    for name, p in prof.items():
        per_launch_ms = p["ms"] / max(1, p["launches"])
        ab = algo_bytes(name, prof, n_bases, n_kmers, table, lookups)
        gbs = ab / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        kernels[name] = {"launches_per_step": p["launches"] / args.steps, "ms_per_launch": round(per_launch_ms, 4),
                         "share_of_kernel_time": round(p["ms"] / tot_kernel_ms, 4) if tot_kernel_ms else 0.0,
                         "algo_bytes_per_launch": ab, "achieved_gbs": round(gbs, 1), "frac_of_hbm": round(gbs / peak, 4)}
    dominant = max(prof, key=lambda k_: prof[k_]["ms"])
    dk = kernels[dominant]
    # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/summarize.py)
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        traffic_src = tj.get("source")
        for name, kk in kernels.items():
            if name in tj["kernels"]:
                kk["dram_bytes_per_launch_ncu"] = tj["kernels"][name]["dram_bytes_per_launch"]
        if dominant in tj["kernels"]:
            traffic = tj["kernels"][dominant]["dram_bytes_per_launch"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(world, args.genome_per_gpu), "reads_per_gpu": n_reads,
                   "bases_per_gpu": n_bases, "kmers_per_gpu": n_kmers,
                   "l2": "no flush: every step streams more than L2 holds (155 MB of read slots per pass x 4 passes, 0.55 GB of "
                         "partitioned k-mers, the 1 GiB bitfield written per step); the 69 MB rank-compacted copy of the "
                         "solid set is L2 resident by design and is rebuilt every step",
                   "parallelism": f"reads sharded over {world} GPU(s)", "rank0_numa_node": numa_node},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "warmup_ms_per_step": [round(x, 2) for x in e2e_warm_ms],
                "host_clock_ms_per_step": e2e_step_ms, "mode": e2e_mode,
                "pipeline": {"stream": "one context; step i+1's upload and step i-1's download run on its copy stream "
                                       "(brgpu_reads_upload_async / _download_async) while step i's kernels run",
                             "lanes": "%d contexts (streams) driven by %d host threads, steps alternate between them"
                                      % (len(lanes), len(lanes)),
                             "serial": "one step at a time"}[e2e_mode],
                "h2d_bytes_per_step": int(h_seq.numel() + 8 * h_off.numel()),
                "d2h_bytes_per_step": int(d2h + 8 * (n_reads + 1))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": dominant, "bound": "hbm", "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": dk["frac_of_hbm"], "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": dk["algo_bytes_per_launch"], "peak_source": peak_src,
                     "scan_lookups_per_step": lookups},
        "kernels": kernels,
        "ms_per_step_with_kernel_events": ms_prof / args.steps,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import br_oracle as o

        o.build()
        th = o.max_threads()
        v, desc = cpu_pipeline_sample(seq, off, n_bases, 1, th)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": th, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)
    if tdist is not None:
        tdist.barrier()
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--genome-per-gpu", type=int, default=GENOME_PER_GPU,
                    help="genome bases per GPU (default: the 4.6 Mb of BASELINE.json configs[1]); smaller values are "
                         "for smoke runs only and are not the benchmark")
    ap.add_argument("--methods", nargs="+", default=None,
                    help="method chain (default: one two = BASELINE.json configs[1]; `graph greedy gap_size` = configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e steps one at a time on a single context")
    ap.add_argument("--e2e-mode", choices=["stream", "lanes", "serial"], default="stream",
                    help="e2e pipeline at N = 1: `stream` = one context, copies of the neighbouring steps on its copy "
                         "stream (default); `lanes` = several contexts driven by host threads; `serial` = one step at a time")
    ap.add_argument("--e2e-lanes", type=int, default=2, help="contexts / host threads of the `lanes` e2e mode")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.methods:
        global METHODS
        METHODS = [m.replace("-", "_") for m in args.methods]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
