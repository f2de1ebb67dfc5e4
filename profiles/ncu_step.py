"""One warm step + one profiled step of the hot path for ncu (launch list, instruction and DRAM counters).

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters.csv \
        --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread \
        python profiles/ncu_step.py [one two | graph greedy gap_size]

Same workload as bench.py's headline (BASELINE.json configs[1]; the method chain is the argument list).
profiles/summarize_r2.py turns the CSV into profiles/kernel_counters.json, which bench.py reads.
Numbers printed by a run under ncu are not bench values.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import br_b200  # noqa: E402

methods = [m.replace("-", "_") for m in sys.argv[1:]] or bench.METHODS
synth = bench.load_synth()
stream = torch.cuda.Stream()
ctx = br_b200.Context(0, stream=stream)
d = bench.headline_descriptors(synth, 1, 0, bench.GENOME_PER_GPU)
reads = br_b200.Reads.synth(ctx, d["genome_seed"], d["read_seed"], d["first"], d["start"], d["tlen"], d["strand"], d["thr"])


def step():
    solid = br_b200.Pcon.from_reads(ctx, reads, bench.K, abundance=bench.ABUNDANCE)
    out = br_b200.correct_reads(br_b200.build_methods(methods, solid, bench.CONFIRM, bench.MAX_SEARCH), reads)
    out.free()
    solid.free()


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step of", "+".join(methods), "over", reads.bases, "bases")
