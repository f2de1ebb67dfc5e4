"""br_b200 — B200-native hot path of natir/br behind br's own set / corrector surface.

The compute lives in br_b200/libbrgpu.so (hand-written CUDA for sm_100a, C ABI in
include/brgpu.h).  Importing this package loads that library and fails if it is missing;
nothing here computes on the CPU.
"""
from . import _lib  # noqa: F401  (loads libbrgpu.so or raises)
from ._lib import BrgpuError
from .correct import (
    Corrector, GapSize, Graph, Greedy, One, Two, build_methods, correct_batch, correct_reads, run_correction,
)
from .runtime import Context, Reads
from .set import Counter, Hash, KmerSet, Pcon, seq2bit

__all__ = [
    "BrgpuError", "Context", "Reads", "Counter", "Hash", "KmerSet", "Pcon", "seq2bit", "Corrector", "One", "Two", "Graph",
    "Greedy", "GapSize", "build_methods", "correct_batch", "correct_reads", "run_correction",
]
