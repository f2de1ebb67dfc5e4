"""ctypes binding of br_b200/libbrgpu.so — the C ABI declared in include/brgpu.h.

The library is the product; this module only loads it.  If the shared object is missing
the import fails loudly: there is no Python or CPU fallback for any operation.
"""
import ctypes as C
from pathlib import Path

import os

# BRGPU_LIBRARY: another build of the same sources (A/B measurements of compile-time choices)
_SO = Path(os.environ.get("BRGPU_LIBRARY") or Path(__file__).resolve().parent / "libbrgpu.so")

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_NOMEM, E_OVERFLOW, E_NO_THRESHOLD, E_NEED_ABUNDANCE = range(8)
ONE, TWO, GRAPH, GREEDY, GAP_SIZE = range(5)
ABUNDANCE_EXPLICIT, ABUNDANCE_FIRST_MINIMUM, ABUNDANCE_RAREFACTION, ABUNDANCE_PERCENT_AT_MOST, ABUNDANCE_PERCENT_AT_LEAST = range(5)

_STATUS = {
    E_INVALID: "invalid argument",
    E_NO_DEVICE: "no CUDA device (brgpu has no CPU path)",
    E_CUDA: "CUDA error",
    E_NOMEM: "out of memory",
    E_OVERFLOW: "output buffer too small",
    E_NO_THRESHOLD: "can't compute the abundance threshold",
    E_NEED_ABUNDANCE: "need an abundance threshold or an abundance method",
}

vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
pvp, pu64 = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)

# name -> (restype, argtypes): every symbol include/brgpu.h declares
SIGNATURES = {
    "brgpu_ctx_create": (C.c_int, [C.c_int, vp, pvp]),
    "brgpu_ctx_destroy": (None, [vp]),
    "brgpu_ctx_synchronize": (C.c_int, [vp]),
    "brgpu_ctx_set_option": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "brgpu_last_error": (C.c_char_p, [vp]),
    "brgpu_version": (C.c_char_p, []),
    "brgpu_host_alloc": (C.c_int, [vp, sz, pvp]),
    "brgpu_host_free": (None, [vp, vp]),
    "brgpu_reads_upload": (C.c_int, [vp, vp, vp, u64, pvp]),
    "brgpu_reads_upload_packed": (C.c_int, [vp, vp, vp, u64, vp, vp, u64, pvp]),
    "brgpu_reads_upload_packed_async": (C.c_int, [vp, vp, vp, u64, vp, vp, u64, pvp]),
    "brgpu_reads_download_packed": (C.c_int, [vp, vp, u64, vp, vp, vp, u64, vp]),
    "brgpu_reads_download_packed_async": (C.c_int, [vp, vp, u64, vp, vp, vp, u64, vp]),
    "brgpu_reads_synth": (C.c_int, [vp, u64, u64, u64, vp, vp, vp, u64, vp, pvp]),
    "brgpu_reads_count": (u64, [vp]),
    "brgpu_reads_bases": (u64, [vp]),
    "brgpu_reads_download": (C.c_int, [vp, vp, u64, vp, pu64]),
    "brgpu_reads_free": (None, [vp]),
    "brgpu_reads_upload_async": (C.c_int, [vp, vp, vp, u64, pvp]),
    "brgpu_reads_download_async": (C.c_int, [vp, vp, u64, vp, pu64]),
    "brgpu_reads_download_wait": (C.c_int, [vp]),
    "brgpu_counts_create": (C.c_int, [vp, C.c_int, pvp]),
    "brgpu_counts_add_reads": (C.c_int, [vp, vp]),
    "brgpu_counts_spectrum": (C.c_int, [vp, vp]),
    "brgpu_spectrum_first_minimum": (C.c_int, [vp]),
    "brgpu_spectrum_threshold": (C.c_int, [vp, C.c_int, C.c_double]),
    "brgpu_counts_upload": (C.c_int, [vp, vp, u64]),
    "brgpu_counts_download": (C.c_int, [vp, vp, u64]),
    "brgpu_counts_device_ptr": (vp, [vp]),
    "brgpu_counts_len": (u64, [vp]),
    "brgpu_counts_free": (None, [vp]),
    "brgpu_set_from_counts": (C.c_int, [vp, C.c_int, pvp]),
    "brgpu_set_from_reads": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, pvp]),
    "brgpu_set_from_host_reads": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, u64, pvp]),
    "brgpu_set_from_reads_ex": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, pvp]),
    "brgpu_set_from_host_reads_ex": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp, u64, pvp]),
    "brgpu_set_from_bitfield": (C.c_int, [vp, C.c_int, vp, u64, pvp]),
    "brgpu_set_from_solid_payload": (C.c_int, [vp, vp, u64, pvp]),
    "brgpu_set_new": (C.c_int, [vp, C.c_int, pvp]),
    "brgpu_set_insert_batch": (C.c_int, [vp, vp, u64]),
    "brgpu_set_hash_new": (C.c_int, [vp, C.c_int, u64, pvp]),
    "brgpu_set_hash_add_reads": (C.c_int, [vp, vp]),
    "brgpu_set_hash_from_reads": (C.c_int, [vp, C.c_int, vp, pvp]),
    "brgpu_set_hash_from_host_reads": (C.c_int, [vp, C.c_int, vp, vp, u64, pvp]),
    "brgpu_set_is_hash": (C.c_int, [vp]),
    "brgpu_set_hash_size": (u64, [vp]),
    "brgpu_set_k": (C.c_int, [vp]),
    "brgpu_set_abundance": (C.c_int, [vp]),
    "brgpu_set_bitfield_bytes": (u64, [vp]),
    "brgpu_set_export_bitfield": (C.c_int, [vp, vp, u64]),
    "brgpu_set_get_batch": (C.c_int, [vp, vp, u64, vp]),
    "brgpu_set_spectrum": (C.c_int, [vp, vp]),
    "brgpu_set_device_ptr": (vp, [vp]),
    "brgpu_set_new_sliced": (C.c_int, [vp, C.c_int, pvp]),
    "brgpu_set_slice_compact": (C.c_int, [vp, u64, u64, pvp, pu64]),
    "brgpu_set_compact_alloc": (C.c_int, [vp, u64, pvp]),
    "brgpu_set_compact_commit": (C.c_int, [vp]),
    "brgpu_set_slice_ipc_export": (C.c_int, [vp, vp]),
    "brgpu_set_compact_pull": (C.c_int, [vp, vp, vp, C.c_int]),
    "brgpu_set_summary_ptr": (vp, [vp, pu64]),
    "brgpu_set_commit_slices": (C.c_int, [vp, C.c_int]),
    "brgpu_set_free": (None, [vp]),
    "brgpu_correct_reads": (C.c_int, [vp, vp, vp, u64, C.c_int, C.c_int, C.c_int, vp, pvp]),
    "brgpu_correct_reads_async": (C.c_int, [vp, vp, vp, u64, C.c_int, C.c_int, C.c_int, vp, pvp]),
    "brgpu_reads_wait": (C.c_int, [vp]),
    "brgpu_correct_batch": (C.c_int, [vp, vp, vp, u64, C.c_int, C.c_int, C.c_int, vp, vp, u64, vp, u64, vp, pu64]),
    "brgpu_correct_one": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, u64, vp, u64, pu64]),
    "brgpu_counts_ipc_export": (C.c_int, [vp, vp]),
    "brgpu_ipc_open": (C.c_int, [vp, vp, pvp]),
    "brgpu_ipc_close": (C.c_int, [vp, vp]),
    "brgpu_counts_merge_slice": (C.c_int, [vp, vp, C.c_int, u64, u64]),
    "brgpu_counts_spectrum_slice": (C.c_int, [vp, u64, u64, vp]),
    "brgpu_set_threshold_slice": (C.c_int, [vp, vp, C.c_int, u64, u64]),
    "brgpu_kmers_create": (C.c_int, [vp, C.c_int, vp, pvp]),
    "brgpu_kmers_buckets": (u64, [vp]),
    "brgpu_kmers_offsets_ptr": (vp, [vp]),
    "brgpu_kmers_ipc_export": (C.c_int, [vp, vp]),
    "brgpu_kmers_count_range": (C.c_int, [vp, vp, vp, C.c_int, u64, u64, C.c_int, vp, vp]),
    "brgpu_kmers_offsets_at": (C.c_int, [vp, vp, u64, vp]),
    "brgpu_kmers_count_range_staged": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, u64, u64, C.c_int, vp, vp]),
    "brgpu_kmers_count_parts": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, u64, u64, C.c_int, vp, vp]),
    "brgpu_set_from_kmers": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, pvp]),
    "brgpu_kmers_free": (None, [vp]),
    "brgpu_group_create": (C.c_int, [vp, C.c_int, pvp]),
    "brgpu_group_destroy": (None, [vp]),
    "brgpu_group_size": (C.c_int, [vp]),
    "brgpu_group_ctx": (vp, [vp, C.c_int]),
    "brgpu_group_last_error": (C.c_char_p, [vp]),
    "brgpu_group_set_from_reads": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp]),
    "brgpu_group_set_from_host_reads": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp, u64, vp]),
    "brgpu_group_sets_free": (None, [vp, vp]),
    "brgpu_group_correct_batch": (C.c_int, [vp, vp, vp, u64, C.c_int, C.c_int, C.c_int, vp, vp, u64, vp, u64, vp, pu64]),
    "brgpu_profile_enable": (C.c_int, [vp, C.c_int]),
    "brgpu_profile_reset": (C.c_int, [vp]),
    "brgpu_profile_count": (C.c_int, [vp]),
    "brgpu_profile_get": (C.c_int, [vp, C.c_int, C.c_char_p, sz, C.POINTER(C.c_double), pu64, C.POINTER(C.c_double)]),
    "brgpu_profile_get_lookups": (C.c_int, [vp, C.c_int, pu64]),
    "brgpu_probe_random_gather": (C.c_int, [vp, u64, C.POINTER(C.c_double)]),
    "brgpu_launch_count": (u64, [vp]),
    "brgpu_scan_lookups": (u64, [vp]),
}


class BrgpuError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = _STATUS.get(status, f"status {status}")
        super().__init__(f"brgpu: {msg}" + (f" ({detail})" if detail else ""))


def _load():
    if not _SO.exists():
        raise ImportError(
            f"{_SO} is missing: build it with `python br_b200/build.py` (nvcc, sm_100a). "
            "br_b200 has no CPU or pure-Python fallback."
        )
    lib = C.CDLL(str(_SO))
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    return lib


lib = _load()


def check(status, ctx_handle=None):
    if status != OK:
        detail = ""
        if ctx_handle:
            d = lib.brgpu_last_error(ctx_handle)
            detail = d.decode() if d else ""
        raise BrgpuError(status, detail)
