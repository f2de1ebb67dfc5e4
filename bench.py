#!/usr/bin/env python3
"""bench.py — corrected bases/s of the br hot path (count -> threshold -> correct) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the whole hot path over one batch of synthetic reads:
    Counter::new + count + spectrum + threshold (src/main.rs:72-115)  then
    run_correction's chunk body with methods `one two`, confirm 5, reversed pass on
    (src/lib.rs:44-55) over every read.
Headline workload at N = 1: BASELINE.json configs[1] — synthetic 4.6 Mb genome (seed 42), 30x
ONT-like reads (seed 43), 10 % error, k = 17, -a 2.  At N > 1 the genome is N x 4.6 Mb and every
rank owns 30 x 4.6 Mb of reads (weak scaling): every rank partitions its own k-mers, rank r counts
bucket range r over all ranks' partitions (pulled over NVLink peer memory), the bitfield slices are
all-gathered with NCCL, and each rank corrects its own reads.  Reads are generated on the device
(brgpu_reads_synth, counter-based; br_b200/synth.py is the numpy mirror the CPU arm uses).

`value`  : bases/s with the reads already resident in HBM (device-resident handles in and out).
`e2e`    : the same metric through the host-buffer calls — pinned host reads are copied to the
           device and the corrected reads are copied back inside the timed region, every step.
`roofline`: the dominant kernel against the ceiling that bounds it ("issue": warp instructions per
           second, "l2": random gathers per second measured in this run, "hbm": algorithmic bytes
           per second over the measured copy bandwidth); `kernels` holds the same for every kernel,
           forward and reversed passes separately, timed with CUDA events on the launching stream.
`parity_check`: untimed, after the timed regions, on the benched data itself.
`extra`  : the other BASELINE configs this run can hold — configs[2] (graph + greedy + gap_size) at
           N = 1, configs[3] (100 Mb genome, 50x, 12 %, reads sharded: strong scaling) at N >= 2.
`cpu_baseline` / `--impl reference`: the CPU restatement of br (oracle/, C++ + OpenMP on all host
           cores; the Rust reference cannot be built in this image) on a bounded sample.
`--time-budget-s` (780): an extra leg or the CPU baseline that would push the run past the budget is
           reported as {"skipped": ...}; the headline line is never affected.  `wall_s` = the run's wall clock.
"""
import argparse
import hashlib
import importlib.util
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
T_START = time.time()

K = 17
ABUNDANCE = 2
METHODS = ["one", "two"]
CONFIRM = 5
MAX_SEARCH = 7
GENOME_PER_GPU = 4_600_000
COVERAGE = 30
ERROR = 0.10
GENOME_SEED = 42
READ_SEED = 43
METRIC = "corrected bases/sec (count + threshold + correct, whole pipeline)"
UNIT = "bases/s"
CONFIG3 = {"genome": 100_000_000, "coverage": 50, "error": 0.12}  # BASELINE.json configs[3]
CONFIG2_METHODS = ["graph", "greedy", "gap_size"]                  # BASELINE.json configs[2]
CONFIG4 = {"genome": 1_000_000_000, "coverage": 30, "error": 0.10}  # BASELINE.json configs[4] (8 GPUs)
CHUNK_TEMPLATE_BASES = 1_300_000_000  # a rank's shard is held in chunks of at most this many bases (u32 slot cursors)


def load_synth():
    """br_b200/synth.py by path: the CPU arm must not load the package (and with it libbrgpu.so)."""
    spec = importlib.util.spec_from_file_location("brgpu_synth", ROOT / "br_b200" / "synth.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def workload_name(n_gpus, genome_per_gpu, methods=None):
    g = genome_per_gpu * n_gpus / 1e6
    return (f"synthetic {g:.1f} Mb genome (seed {GENOME_SEED}), {COVERAGE}x ONT-like reads (seed {READ_SEED}+rank), "
            f"{int(ERROR * 100)}% error, k={K}, -a {ABUNDANCE}, methods {'+'.join(methods or METHODS)}, confirm {CONFIRM}, "
            f"reversed pass on")


L2_NOTE = ("no flush: every step streams more than L2 holds (155 MB of read slots per pass x 4 passes, 0.55 GB of "
           "partitioned k-mers, the 1 GiB bitfield written per step); the 69 MB rank-compacted copy of the "
           "solid set is L2 resident by design and is rebuilt every step")


def headline_config(n_gpus, genome_per_gpu, n_reads, n_bases, n_kmers):
    """`config` of the headline line — the same dict on both arms (the reference arm runs on this arm's config);
    the per-GPU figures are rank 0's shard."""
    return {"workload": workload_name(n_gpus, genome_per_gpu), "reads_per_gpu": int(n_reads), "bases_per_gpu": int(n_bases),
            "kmers_per_gpu": int(n_kmers), "l2": L2_NOTE, "parallelism": f"reads sharded over {n_gpus} GPU(s)",
            "generator": "counter-based (brgpu_reads_synth on the device; br_b200/synth.py host mirror)"}


def headline_descriptors(synth, n_gpus, rank, genome_per_gpu):
    """Every rank draws its reads from the whole (N x 4.6 Mb) genome; its share is 30 x 4.6 Mb of bases."""
    start, tlen, strand = synth.read_descriptors(genome_per_gpu * n_gpus, COVERAGE / n_gpus, seed=READ_SEED + rank)
    return {"genome_seed": GENOME_SEED, "read_seed": READ_SEED + rank, "first": 0, "start": start, "tlen": tlen,
            "strand": strand, "thr": synth.error_thresholds(ERROR)}


def sharded_descriptors(synth, cfg, n_gpus, rank):
    """One genome, one global read list (configs[3]: 100 Mb, 50x, 12 %; configs[4]: 1 Gb, 30x, 10 %): rank r
    takes a contiguous range of the list, balanced by bases."""
    start, tlen, strand = synth.read_descriptors(cfg["genome"], cfg["coverage"], seed=READ_SEED)
    lo, hi = synth.shard_descriptors(tlen, n_gpus, rank)
    return {"genome_seed": GENOME_SEED, "read_seed": READ_SEED, "first": lo, "start": start[lo:hi], "tlen": tlen[lo:hi],
            "strand": strand[lo:hi], "thr": synth.error_thresholds(cfg["error"]), "n_reads_total": int(tlen.size),
            "template_bases_total": int(tlen.astype(np.int64).sum())}


def chunk_cuts(tlen, n_chunks):
    """Read-index cuts that split a shard into n_chunks pieces of about equal bases."""
    cs = np.concatenate([[0], np.cumsum(tlen.astype(np.int64))])
    return [int(np.searchsorted(cs, cs[-1] * j // n_chunks, side="left")) for j in range(n_chunks)] + [int(tlen.size)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


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
def cpu_pipeline_sample(seq, off, total_bases_all_shards, n_shards, threads, methods, correct_target_s=6.0, keep_solid=False):
    """Times the CPU restatement on a bounded sample of the workload and extrapolates linearly
    in the parts that are linear (counting after the table is touched; correction over reads).
    Returns (bases/s for the whole workload, description of the sample, oracle Solid or None)."""
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
    ids = [o.METHOD_IDS[m] for m in methods]
    ns = max(1, n // 64)
    while True:
        t0 = time.perf_counter()
        solid.run_correction(ids, seq, off[: ns + 1], confirm=CONFIRM, max_search=MAX_SEARCH, two_side=False, threads=threads)
        t_corr = time.perf_counter() - t0
        if t_corr >= correct_target_s / 2 or ns == n:
            break
        ns = min(n, max(ns * 2, int(ns * correct_target_s / max(t_corr, 1e-3))))
    sample_bases = int(off[ns]) - int(off[0])
    shard_bases = int(off[-1]) - int(off[0])
    # counting is linear in the bases once the table is touched: the measured shard stands for the rest
    other = max(0.0, total_bases_all_shards - shard_bases)
    t_count_rest = (t_count_warm if n_shards > 1 else 0.0) * other / max(1, shard_bases)
    t_full = t_count_cold + t_count_rest + t_passes + t_corr * (total_bases_all_shards / sample_bases)
    desc = (f"count of a {shard_bases / 1e6:.0f} Mbase shard into a fresh 2^{2 * K - 1}-counter table ({t_count_cold:.2f} s"
            f"{', warm recount %.2f s scaled to the other %.0f Mbases' % (t_count_warm, other / 1e6) if n_shards > 1 else ''}) "
            f"+ spectrum and threshold passes ({t_passes:.2f} s) measured in full; correction ({'+'.join(methods)}) measured on "
            f"the first {ns} reads ({sample_bases / 1e6:.1f} Mbases, {t_corr:.2f} s) and scaled linearly to "
            f"{total_bases_all_shards / 1e6:.0f} Mbases")
    return total_bases_all_shards / t_full, desc, (solid if keep_solid else None)


def run_reference(args, world, rank):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only).  The thread
    count comes from the process's CPU affinity, not from OMP_NUM_THREADS (torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers)."""
    if rank != 0:
        return
    from oracle import br_oracle as o

    o.build()
    synth = load_synth()
    threads = host_cores()
    d = headline_descriptors(synth, args.gpus, 0, args.genome_per_gpu)
    seq, off = synth.host_reads(d["genome_seed"], d["read_seed"], d["first"], d["start"], d["tlen"], d["strand"], d["thr"])
    total = int(off[-1]) * args.gpus
    n_kmers = int(np.maximum(np.diff(off.astype(np.int64)) - K + 1, 0).sum())
    vals, desc = [], ""
    # the CPU pipeline takes seconds per pass: one untimed-quality pass would cost as much as a timed one,
    # so it is measured once plus one repeat whatever --steps / --warmup say
    for _ in range(2):
        v, desc, _ = cpu_pipeline_sample(seq, off, total, args.gpus, threads, METHODS)
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": headline_config(args.gpus, args.genome_per_gpu, off.size - 1, int(off[-1]), n_kmers),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "repeats": [round(x) for x in vals]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of br (oracle/, C++17 + OpenMP); the Rust reference cannot be built in this image; "
                "measured twice (not warmup + steps times): each pass is seconds of CPU work",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def algo_bytes(name, p, n_bases, n_kmers, table):
    """SURVEY §8(d) algorithmic bytes per launch for each kernel of the step."""
    base = name[:-4] if name.endswith("_rev") else name
    if base == "count_kmers":
        return 0.25 * n_bases + 64.0 * n_kmers
    if base == "zero_counts":
        return float(table)
    if base == "spectrum_threshold":
        return table * 1.125
    if base == "spectrum":
        return float(table)
    lookups = p["lookups"] / max(1, p["launches"])
    if base == "solid_bitmap":  # one sector per lookup it really issued + the ASCII stream in + 1 bit out per position
        return 32.0 * lookups + 0.25 * n_bases + n_bases / 8.0
    if base.startswith("scan_") or base.startswith("merge_"):
        # read + write of the ASCII bases and one 32 B sector per KmerSet::get this launch issued
        return 2.0 * n_bases + 32.0 * lookups
    return p["algo_bytes"] / max(1, p["launches"])


def kernel_table(prof, steps, n_bases, n_kmers, table, hbm_peak, issue_peak, l2_gather, counters):
    """Per-kernel record: time, share, §8(d) byte view, and the ceiling that bounds the kernel."""
    out = {}
    tot = sum(p["ms"] for p in prof.values())
    for name, p in prof.items():
        per_ms = p["ms"] / max(1, p["launches"])
        ab = algo_bytes(name, p, n_bases, n_kmers, table)
        gbs = ab / (per_ms * 1e-3) / 1e9 if per_ms > 0 else 0.0
        lookups = p["lookups"] / max(1, p["launches"])
        rec = {"launches_per_step": p["launches"] / steps, "ms_per_launch": round(per_ms, 4),
               "share_of_kernel_time": round(p["ms"] / tot, 4) if tot else 0.0,
               "lookups_per_launch": lookups, "algo_bytes_per_launch": ab,
               "hbm_view": {"achieved_gbs": round(gbs, 1), "frac_of_hbm": round(gbs / hbm_peak, 4)}}
        c = counters.get(name)
        if c:
            rec["warp_inst_per_launch_ncu"] = c.get("warp_inst_per_launch")
            rec["dram_bytes_per_launch_ncu"] = c.get("dram_bytes_per_launch")
        base = name[:-4] if name.endswith("_rev") else name
        dram = c.get("dram_bytes_per_launch") if c else None
        inst = c.get("warp_inst_per_launch") if c else None
        if base == "solid_bitmap" and l2_gather:
            ach = lookups / (per_ms * 1e-3) if per_ms > 0 else 0.0
            rec["bound"] = {"bound": "l2", "achieved": round(ach / 1e9, 2), "peak": round(l2_gather / 1e9, 2),
                            "unit": "G lookups/s vs G random 8 B gathers/s (64 MiB table, measured in this run)",
                            "frac": round(ach / l2_gather, 4)}
        elif (base.startswith("scan_") or base.startswith("merge_") or (dram is not None and dram < 0.5 * ab)) and inst:
            ach = inst / (per_ms * 1e-3) if per_ms > 0 else 0.0
            rec["bound"] = {"bound": "issue", "achieved": round(ach / 1e12, 4), "peak": round(issue_peak / 1e12, 4),
                            "unit": "T warp instructions/s (instructions per launch from the committed ncu capture)",
                            "frac": round(ach / issue_peak, 4)}
        else:
            rec["bound"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                            "frac": round(gbs / hbm_peak, 4)}
        out[name] = rec
    return out


def run_ours(args, world, rank, local_rank):
    import torch

    import br_b200
    from br_b200 import dist as bdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; br_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from br_b200.runtime import bind_to_gpu_numa_node

    synth = load_synth()
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation
    tdist = None
    if world > 1:
        import torch.distributed as tdist

        tdist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    stream = torch.cuda.Stream()
    ctx = br_b200.Context(local_rank, stream=stream)
    dev = f"cuda:{local_rank}"
    table = 1 << (2 * K - 1)

    def barrier():
        torch.cuda.synchronize()
        if tdist is not None:
            tdist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if tdist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(x):
        if tdist is None:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        tdist.all_reduce(t)
        return int(t.item())

    def all_and(flag):
        if tdist is None:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
        return bool(t.item())

    def time_left(need_s):
        """True on every rank when the slowest rank still has `need_s` seconds of --time-budget-s: an extra leg that
        could push the run past the driver's per-run limit is skipped (and said so) instead of losing the whole line."""
        return all_and(time.time() - T_START + need_s <= args.time_budget_s)

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        return all_max(a.elapsed_time(b))

    def build_set(reads):
        """reads: one device-resident chunk, or the list of chunks a large shard is held in"""
        if world == 1:
            if isinstance(reads, list):
                return br_b200.Pcon.from_reads(ctx, reads[0], K, abundance=ABUNDANCE) if len(reads) == 1 else \
                    br_b200.Pcon.from_chunks(ctx, reads, K, abundance=ABUNDANCE)
            return br_b200.Pcon.from_reads(ctx, reads, K, abundance=ABUNDANCE)
        return bdist.build_set_sharded(bdist.GpuOps(ctx, reads), K, abundance=ABUNDANCE)

    class Workload:
        """Device-resident reads of one config + the pinned host buffers of its e2e steps."""

        def __init__(self, desc, n_chunks=1):
            self.desc = desc
            if n_chunks > 1:  # a shard too large for one chunk: device-resident only (no host copy, no e2e)
                cuts = chunk_cuts(desc["tlen"], n_chunks)
                self.chunks = [br_b200.Reads.synth(ctx, desc["genome_seed"], desc["read_seed"], desc["first"] + a,
                                                   desc["start"][a:b], desc["tlen"][a:b], desc["strand"][a:b], desc["thr"])
                               for a, b in zip(cuts[:-1], cuts[1:])]
                self.dev_reads = self.chunks
                self.n_reads = sum(len(c) for c in self.chunks)
                self.n_bases = sum(int(c.bases) for c in self.chunks)
                self.n_kmers = self.n_bases - (K - 1) * self.n_reads  # every synthetic read is longer than k
                self.h_seq = None
                return
            self.dev_reads = br_b200.Reads.synth(ctx, desc["genome_seed"], desc["read_seed"], desc["first"], desc["start"],
                                                 desc["tlen"], desc["strand"], desc["thr"])
            self.chunks = [self.dev_reads]
            self.n_reads = len(self.dev_reads)
            self.n_bases = int(self.dev_reads.bases)
            self.h_seq = torch.empty(max(1, self.n_bases), dtype=torch.uint8).pin_memory()
            self.h_off = torch.empty(self.n_reads + 1, dtype=torch.int64).pin_memory()
            self.dev_reads.download(self.h_seq, self.h_off)   # what a host application would hand to the C ABI
            off = self.h_off.numpy().view(np.uint64)
            lens = np.diff(off.astype(np.int64))
            self.n_kmers = int(np.maximum(lens - K + 1, 0).sum())
            out_cap = self.n_bases + self.n_bases // 8 + 64 * self.n_reads + 64
            self.out_bufs = [(torch.empty(out_cap, dtype=torch.uint8).pin_memory(),
                              torch.empty(self.n_reads + 1, dtype=torch.int64).pin_memory()) for _ in range(2)]
            # the same chunk in the 2-bit transport form (what br::fasta's reader produces while it parses):
            # four bases per byte + the exception list (empty here: the generator emits A, C, G, T only)
            from br_b200.runtime import pack_2bit

            packed, exc_pos, exc_byte = pack_2bit(self.h_seq.numpy()[: self.n_bases])
            self.h_packed = torch.from_numpy(packed).pin_memory()
            self.exc_pos, self.exc_byte = exc_pos, exc_byte
            ecap = max(16, int(exc_pos.size))
            self.out_packed = [(torch.empty(out_cap // 4 + 8, dtype=torch.uint8).pin_memory(),
                                torch.empty(self.n_reads + 1, dtype=torch.int64).pin_memory(),
                                torch.empty(ecap, dtype=torch.int64).pin_memory(), torch.empty(ecap, dtype=torch.uint8).pin_memory(),
                                torch.zeros(2, dtype=torch.int64).pin_memory()) for _ in range(2)]

        def seq_off(self):
            return self.h_seq.numpy()[: self.n_bases], self.h_off.numpy().view(np.uint64)

        def free(self):
            for c in self.chunks:
                c.free()
            self.h_seq = self.h_off = self.out_bufs = self.h_packed = self.out_packed = None

    def step_device(wl, methods, keep=False):
        solid = build_set(wl.dev_reads)
        m = br_b200.build_methods(methods, solid, CONFIRM, MAX_SEARCH)
        if len(wl.chunks) > 1:  # chunk after chunk against the same replicated set
            outs = []
            for c in wl.chunks:
                o = br_b200.correct_reads(m, c)
                if keep:
                    outs.append(o)
                else:
                    o.free()
            if keep:
                return solid, outs
            solid.free()
            return None
        out = br_b200.correct_reads(m, wl.dev_reads)
        if keep:
            return solid, out
        out.free()
        solid.free()

    def run_e2e_stream(wl, methods, n_steps, stamps=None, transport="packed"):
        """n_steps e2e steps on ONE context with the asynchronous staging calls: step i+1's reads go
        up and step i-1's result comes down on the context's copy stream while step i's kernels run.
        Every step still uploads its own inputs and downloads its own result."""
        res = 0
        if transport == "packed":
            def up():
                return br_b200.Reads.upload_packed(ctx, wl.h_packed, wl.h_off, wl.exc_pos, wl.exc_byte, asynchronous=True)

            def down(r, j):
                r.download_packed(*wl.out_packed[j % 2], asynchronous=True)
                return j % 2
        else:
            def up():
                return br_b200.Reads.upload_async(ctx, wl.h_seq, wl.h_off)

            def down(r, j):
                return r.download_async(*wl.out_bufs[j % 2])
        nxt = up()
        prev = None
        t0 = time.perf_counter()
        for i in range(n_steps):
            cur = nxt
            if i + 1 < n_steps:
                nxt = up()                                                            # H2D of the next step
            if prev is not None:
                res = max(res, down(prev, i - 1))                                     # D2H of the previous step
            solid = build_set(cur)
            out = br_b200.correct_reads(br_b200.build_methods(methods, solid, CONFIRM, MAX_SEARCH), cur)
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
        res = max(res, down(prev, n_steps - 1))
        prev.download_wait()
        prev.free()
        if transport == "packed":  # bytes that came down: packed bases + the surviving exceptions
            c = wl.out_packed[(n_steps - 1) % 2][4]
            return (int(c[0]) + 3) // 4 + 9 * int(c[1])
        return res

    def run_e2e_serial(wl, methods, n_steps, stamps=None, transport="packed"):
        res = 0
        for _ in range(n_steps):
            t0 = time.perf_counter()
            if transport == "packed":
                reads = br_b200.Reads.upload_packed(ctx, wl.h_packed, wl.h_off, wl.exc_pos, wl.exc_byte)  # H2D every step
            else:
                reads = br_b200.Reads.upload(ctx, wl.h_seq, wl.h_off)
            solid = build_set(reads)
            out = br_b200.correct_reads(br_b200.build_methods(methods, solid, CONFIRM, MAX_SEARCH), reads)
            if transport == "packed":
                c = out.download_packed(*wl.out_packed[0])[4]                 # D2H every step
                res = max(res, (int(c[0]) + 3) // 4 + 9 * int(c[1]))
            else:
                d, _ = out.download(*wl.out_bufs[0])
                res = max(res, int(d.numel()))
            out.free()
            solid.free()
            reads.free()
            if stamps is not None:
                stamps.append(round((time.perf_counter() - t0) * 1e3, 2))
        return res

    e2e_mode = "serial" if args.no_e2e_pipeline else args.e2e_mode
    run_e2e = run_e2e_stream if e2e_mode == "stream" else run_e2e_serial

    def measure(wl, methods, steps, warmup, want_e2e=True, e2e_warm_extra=12, transports=("packed",)):
        """value, per-kernel profile and e2e of `methods` over `wl`; returns a dict of raw numbers."""
        for _ in range(warmup):
            step_device(wl, methods)
        launches0 = ctx.launch_count
        ms_dev = timed(lambda: step_device(wl, methods), steps)       # ---- device-resident region: `value`
        launches = ctx.launch_count - launches0
        # the same steps again with a CUDA-event pair around every kernel and the scans' KmerSet::get
        # counters on (the event records perturb the step a little: `value` comes from the region above)
        ctx.profile_reset()
        ctx.profile_enable(True)
        ms_prof = timed(lambda: step_device(wl, methods), steps)
        prof = ctx.profile()
        ctx.profile_enable(False)
        res = {"ms_dev": ms_dev, "launches": launches, "ms_prof": ms_prof, "prof": prof}
        if not want_e2e:
            return res
        # end-to-end region: warm up until two consecutive rounds agree within 3 % (the first e2e steps
        # size the allocator's cache for the upload / download buffers)
        for tr in transports:
            d2h, warm_ms = 0, []
            for it in range(max(3, warmup) + e2e_warm_extra):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n_warm = 4 if e2e_mode == "stream" else 1
                d2h = max(d2h, run_e2e(wl, methods, n_warm, None, tr))
                torch.cuda.synchronize()
                warm_ms.append((time.perf_counter() - t0) * 1e3 / n_warm)
                stable = it + 1 >= max(3, warmup) and abs(warm_ms[-1] - warm_ms[-2]) <= 0.03 * warm_ms[-1]
                if all_and(stable):  # the step contains collectives: every rank must take the same decision
                    break
            stamps = []
            ms_e2e = timed(lambda: run_e2e(wl, methods, steps, stamps, tr), 1)
            h2d = ((wl.n_bases + 3) // 4 + 9 * int(wl.exc_pos.size) if tr == "packed" else wl.n_bases) + 8 * (wl.n_reads + 1)
            rec = {"ms_e2e": ms_e2e, "d2h": d2h + 8 * (wl.n_reads + 1) + (16 if tr == "packed" else 0), "h2d": h2d,
                   "e2e_warm_ms": warm_ms, "e2e_step_ms": stamps, "transport": tr}
            if "ms_e2e" not in res:
                res.update(rec)
            else:
                res.setdefault("e2e_other", []).append(rec)
        return res

    # a full collection of the interpreter's heap (torch imports ~10^6 objects) costs 0.1-0.5 s and
    # would land inside a timed region: collect now, then keep the collector off while timing
    import gc

    gc.collect()
    gc.freeze()
    gc.disable()
    hbm_peak, peak_src, sm_max_mhz = peaks()
    with torch.cuda.stream(stream):
        wl = Workload(headline_descriptors(synth, world, rank, args.genome_per_gpu))
        # the sampler is started before the warm-up so that nvidia-smi's own start-up (NVML
        # initialisation takes driver locks) is over before the timed regions begin
        sampler = ClockSampler(local_rank)
        if rank == 0 and not os.environ.get("BRGPU_BENCH_NO_SAMPLER"):
            sampler.start()
        m = measure(wl, METHODS, args.steps, args.warmup, transports=("packed", "ascii"))
        clocks = sampler.stop() if rank == 0 else None
        total_bases = all_sum(wl.n_bases)

        # ---- yardsticks measured in this run ----
        l2_gather = ctx.probe_random_gather(64 << 20)
        dram_gather = ctx.probe_random_gather(8 << 30)

        # ---- parity on the benched data (untimed) ----
        parity = None
        if not args.no_parity:
            parity = parity_check(args, ctx, wl, world, rank, tdist, dev, step_device, all_and)

        # ---- the other configs this run can hold ----
        extra = {}
        if not args.no_extra:
            if world == 1 and not time_left(120):
                extra["configs[2]"] = {"skipped": f"time budget ({args.time_budget_s:.0f} s, --time-budget-s)"}
            elif world == 1:
                steps2, warm2 = max(1, min(args.steps, 3)), max(1, min(args.warmup, 2))
                m2 = measure(wl, CONFIG2_METHODS, steps2, warm2, e2e_warm_extra=4)
                extra["configs[2]"] = leg_record(m2, wl, wl.n_bases, steps2, warm2, table, hbm_peak, sm_max_mhz, l2_gather,
                                                 workload_name(1, args.genome_per_gpu, CONFIG2_METHODS), "weak", world)
                if not args.no_parity:  # same reads, same set: only the replay of a read sample with this chain is new
                    p2 = parity_check(args, ctx, wl, world, rank, tdist, dev, step_device, all_and, single_gpu_rebuild=False,
                                      sample=100, methods=CONFIG2_METHODS)
                    p2["note"] = "greedy's alignment (bio 1.6.0 custom global) is restated in the oracle from memory: parity unpinned (DESIGN.md §3)"
                    extra["configs[2]"]["parity_check"] = p2
            else:
                wl.free()
                legs = [("configs[3]", CONFIG3)] + ([("configs[4]", CONFIG4)] if world == 8 and not args.no_config4 else [])
                for leg, cfg in legs:
                    if not time_left(180 if leg == "configs[3]" else 330):
                        extra[leg] = {"skipped": f"time budget ({args.time_budget_s:.0f} s, --time-budget-s)"}
                        continue
                    desc = sharded_descriptors(synth, cfg, world, rank)
                    shard_bases = int(desc["tlen"].astype(np.int64).sum())
                    n_chunks = -(-shard_bases // args.chunk_template_bases)
                    wlx = Workload(desc, n_chunks=n_chunks)
                    stepsx, warmx = 2, 1
                    mx = measure(wlx, METHODS, stepsx, warmx, want_e2e=n_chunks == 1, e2e_warm_extra=2)
                    totalx = all_sum(wlx.n_bases)
                    namex = (f"synthetic {cfg['genome'] / 1e6:.0f} Mb genome (seed {GENOME_SEED}), {cfg['coverage']}x ONT-like "
                             f"reads (seed {READ_SEED}), {int(cfg['error'] * 100)}% error, k={K}, -a {ABUNDANCE}, methods "
                             f"{'+'.join(METHODS)}, confirm {CONFIRM}, reversed pass on; reads sharded over {world} GPUs"
                             + (f", {n_chunks} chunks per GPU, generated on the device" if n_chunks > 1 else ""))
                    extra[leg] = leg_record(mx, wlx, totalx, stepsx, warmx, table, hbm_peak, sm_max_mhz, l2_gather, namex,
                                            "strong", world)
                    if not args.no_parity:
                        extra[leg]["parity_check"] = parity_check(args, ctx, wlx, world, rank, tdist, dev, step_device,
                                                                  all_and, single_gpu_rebuild=False)
                    wlx.free()

    if rank != 0:
        if tdist is not None:
            tdist.barrier()
            tdist.destroy_process_group()
        return

    line = leg_record(m, wl if world == 1 or args.no_extra else None, total_bases, args.steps, args.warmup, table, hbm_peak,
                      sm_max_mhz, l2_gather, workload_name(world, args.genome_per_gpu), "weak", world,
                      per_gpu=(wl.n_reads, wl.n_bases, wl.n_kmers))
    line["config"] = headline_config(world, args.genome_per_gpu, wl.n_reads, wl.n_bases, wl.n_kmers)
    line["rank0_numa_node"] = numa_node
    line["e2e"]["mode"] = e2e_mode
    line["clocks"] = clocks
    line["roofline"]["peak_source"] = peak_src
    line["yardsticks"] = {"hbm_copy_gbs": hbm_peak, "hbm_source": peak_src,
                          "issue_warp_inst_per_s": 148 * 4 * sm_max_mhz * 1e6,
                          "l2_random_gather_per_s_64MiB": l2_gather, "dram_random_gather_per_s_8GiB": dram_gather}
    if parity is not None:
        line["parity_check"] = parity
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline and time.time() - T_START + 120 > args.time_budget_s:
        line["cpu_baseline"] = {"skipped": f"time budget ({args.time_budget_s:.0f} s, --time-budget-s)"}
    elif world == 1 and not args.no_cpu_baseline:
        from oracle import br_oracle as o

        o.build()
        th = host_cores()
        seq, off = wl.seq_off()
        v, desc, osolid = cpu_pipeline_sample(seq, off, wl.n_bases, 1, th, METHODS, keep_solid=not args.no_parity)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": th, "kind": "port", "sample": desc}
        if osolid is not None and parity is not None:
            # the CPU leg counted the whole benched data set: its bitfield must be the GPU's
            parity["bitfield_equals_oracle"] = bool(hashlib.blake2b(osolid.bits().tobytes()).hexdigest() == parity["bitfield_blake2b"])
    line["wall_s"] = round(time.time() - T_START, 1)
    print(json.dumps(line), flush=True)
    if tdist is not None:
        tdist.barrier()
        tdist.destroy_process_group()


def load_counters():
    """Per-kernel warp instructions and DRAM bytes per launch from the committed ncu capture
    (profiles/kernel_counters.json, written by profiles/summarize_r2.py)."""
    p = ROOT / "profiles" / "kernel_counters.json"
    if not p.exists():
        return {}, None
    d = json.loads(p.read_text())
    return d.get("kernels", {}), d.get("source")


def leg_record(m, wl, total_bases, steps, warmup, table, hbm_peak, sm_max_mhz, l2_gather, workload, scaling, world,
               per_gpu=None):
    """The JSON record of one measured workload (headline line or an `extra` leg)."""
    counters, counters_src = load_counters()
    if world != 1 or "4.6 Mb" not in workload:
        counters, counters_src = {}, None  # the instruction / DRAM counts belong to the workload they were captured on
    n_reads, n_bases, n_kmers = per_gpu if per_gpu else (wl.n_reads, wl.n_bases, wl.n_kmers)
    issue_peak = 148 * 4 * sm_max_mhz * 1e6
    kernels = kernel_table(m["prof"], steps, n_bases, n_kmers, table, hbm_peak, issue_peak, l2_gather, counters)
    dominant = max(m["prof"], key=lambda k_: m["prof"][k_]["ms"])
    dk = kernels[dominant]
    # SURVEY §8(d) aggregates: the count path as a whole (0.25 B/base + 64 B/k-mer over the time of its kernels) and
    # the solidity lookups of the step (bitmap passes + every KmerSet::get of the scans) per second of the step
    count_names = ("coarse_hist", "coarse_scatter", "fine_partition", "bucket_count", "bucket_count_multi", "peer_pull",
                   "bucket_hist", "bucket_scatter", "count_kmers", "zero_counts", "spectrum_threshold", "spectrum", "summary_popc",
                   "compact_blocks", "block_bytes", "rank_directory", "build_summary", "exclusive_scan")
    count_ms = sum(m["prof"][k_]["ms"] for k_ in m["prof"] if k_ in count_names) / steps
    scan_excl = sum(m["prof"][k_]["ms"] for k_ in m["prof"] if k_ == "exclusive_scan") / steps
    count_bytes = 0.25 * n_bases + 64.0 * n_kmers
    lookups_step = sum(p_["lookups"] for p_ in m["prof"].values()) / steps
    aggregates = {
        "count_path": {"ms_per_step": round(count_ms, 4), "algorithmic_bytes": count_bytes,
                       "achieved_gbs": round(count_bytes / (count_ms * 1e-3) / 1e9, 1) if count_ms else None,
                       "frac_of_hbm": round(count_bytes / (count_ms * 1e-3) / 1e9 / hbm_peak, 4) if count_ms else None,
                       "note": "0.25 B/base + 64 B/k-mer (SURVEY §8d) over the summed time of the set-construction kernels "
                               "(the exclusive scans of the correction passes are in: %.3f ms)" % scan_excl},
        "solid_lookups": {"per_step": lookups_step, "per_s_over_step": round(lookups_step / (m["ms_prof"] / steps * 1e-3), 1),
                          "sector_view_gbs": round(32.0 * lookups_step / (m["ms_prof"] / steps * 1e-3) / 1e9, 1),
                          "note": "bitmap passes + every KmerSet::get of the scans, counted by the profiling variant of the "
                                  "kernels; sector_view = 32 B per lookup over the whole (profiled) step"},
    }
    rec = {
        "metric": METRIC, "value": total_bases * steps / (m["ms_dev"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": m["ms_dev"] / steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload, "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "kmers_per_gpu": n_kmers},
        "gpu_launches": int(m["launches"]),
        "roofline": dict(dk["bound"], kernel=dominant, traffic=dk.get("dram_bytes_per_launch_ncu"),
                         traffic_source=counters_src, algorithmic_bytes_per_launch=dk["algo_bytes_per_launch"],
                         hbm_view=dk["hbm_view"], lookups_per_launch=dk["lookups_per_launch"],
                         warp_inst_per_launch=dk.get("warp_inst_per_launch_ncu")),
        "aggregates": aggregates,
        "kernels": kernels,
        "ms_per_step_with_kernel_events": m["ms_prof"] / steps,
    }
    def e2e_rec(x):
        what = {"packed": "2-bit packed bases + exception list both ways (brgpu_reads_upload_packed_async / "
                          "_download_packed_async): what br::fasta's reader produces while it parses",
                "ascii": "1 byte per base both ways (brgpu_reads_upload_async / _download_async)"}[x["transport"]]
        return {"value": total_bases * steps / (x["ms_e2e"] * 1e-3), "unit": UNIT, "ms_per_step": x["ms_e2e"] / steps,
                "warmup_ms_per_step": [round(v, 2) for v in x["e2e_warm_ms"]], "host_clock_ms_per_step": x["e2e_step_ms"],
                "h2d_bytes_per_step": int(x["h2d"]), "d2h_bytes_per_step": int(x["d2h"]), "transport": x["transport"],
                "transport_note": what}

    if "ms_e2e" in m:
        rec["e2e"] = e2e_rec(m)
        for other in m.get("e2e_other", []):
            rec["e2e_" + other["transport"]] = e2e_rec(other)
    return rec


def parity_check(args, ctx, wl, world, rank, tdist, dev, step_device, all_and, single_gpu_rebuild=True, sample=200, methods=None):
    """Untimed checks on the benched data: (N > 1) every rank's replicated bitfield is the same and equals
    the one a single GPU builds from all ranks' reads through the literal count-table path; (any N) the
    bucketed path's bitfield equals the table path's; a sample of this rank's reads corrected on the GPU
    equals the oracle's bytes (the oracle is given the GPU's bitfield)."""
    import torch

    import br_b200
    from oracle import br_oracle as o

    o.build()
    res = {}
    methods = methods or METHODS
    solid, out = step_device(wl, methods, keep=True)
    bits = solid.bitfield()
    digest = hashlib.blake2b(bits.tobytes()).hexdigest()
    res["bitfield_blake2b"] = digest
    if world > 1:
        t = torch.tensor(list(bytes.fromhex(digest[:32])), dtype=torch.uint8, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        tdist.all_gather(allt, t)
        res["bitfield_equal_across_ranks"] = all(bool(torch.equal(x, allt[0])) for x in allt)
    # literal table path on one GPU over all ranks' reads (Counter::count per shard, Solid::from_count)
    if single_gpu_rebuild and rank == 0:
        synth = load_synth()
        c = br_b200.Counter(ctx, K)
        for r in range(world):
            if world == 1:
                c.count(wl.dev_reads)
            else:
                d = headline_descriptors(synth, world, r, args.genome_per_gpu)
                rr = br_b200.Reads.synth(ctx, d["genome_seed"], d["read_seed"], d["first"], d["start"], d["tlen"], d["strand"], d["thr"])
                c.count(rr)
                rr.free()
        ref = c.to_set(ABUNDANCE)
        same = hashlib.blake2b(ref.bitfield().tobytes()).hexdigest() == digest
        res["equals_single_gpu_table_path" if world > 1 else "bitfield_equals_table_path"] = bool(same)
        ref.free()
        c.free()
    # oracle replay of a read sample (a chunked shard: of its first chunk)
    if isinstance(out, list):
        for o_ in out[1:]:
            o_.free()
        out = out[0]
        seq, off = wl.chunks[0].download()
    else:
        seq, off = wl.seq_off()
    n = off.size - 1
    ids = np.sort(np.random.default_rng(7 + rank).choice(n, size=min(sample, n), replace=False))
    got, got_off = out.download()
    got_off = got_off.astype(np.int64)
    sub_off = np.zeros(ids.size + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum((off[ids + 1] - off[ids]).astype(np.uint64))
    sub_seq = np.concatenate([seq[int(off[i]) : int(off[i + 1])] for i in ids]) if ids.size else np.empty(0, np.uint8)
    osolid = o.Solid.from_bitfield(K, bits)
    exp, exp_off = osolid.run_correction([o.METHOD_IDS[x] for x in methods], sub_seq, sub_off, confirm=CONFIRM,
                                         max_search=MAX_SEARCH, two_side=False, threads=min(8, host_cores()))
    exp_off = exp_off.astype(np.int64)
    ok = True
    edited = 0
    for j, i in enumerate(ids):
        g = got[got_off[i] : got_off[i + 1]]
        e = exp[exp_off[j] : exp_off[j + 1]]
        ok = ok and g.size == e.size and bool(np.array_equal(g, e))
        edited += int(e.size != int(off[i + 1] - off[i]) or not np.array_equal(e, seq[int(off[i]) : int(off[i + 1])]))
    res["sample_reads_equal_oracle"] = all_and(ok)
    res["sample_reads_per_rank"] = int(ids.size)
    res["sample_reads_edited_rank0"] = edited
    out.free()
    solid.free()
    return res


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
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity checks on the benched data")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (configs[2] at N = 1, configs[3] at N >= 2)")
    ap.add_argument("--no-config4", action="store_true", help="at N = 8: skip the configs[4] leg (1 Gb genome, 30x)")
    ap.add_argument("--time-budget-s", type=float, default=780.0,
                    help="wall-clock budget of the whole run: an extra leg (or the CPU baseline) that does not fit is skipped "
                         "and reported as skipped; the headline line is never affected")
    ap.add_argument("--chunk-template-bases", type=int, default=CHUNK_TEMPLATE_BASES,
                    help="bases per device-resident chunk of a rank's shard in the sharded legs (testing the chunked path)")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e steps one at a time")
    ap.add_argument("--e2e-mode", choices=["stream", "serial"], default="stream",
                    help="`stream` = copies of the neighbouring steps on the context's copy stream (default); "
                         "`serial` = one step at a time")
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
