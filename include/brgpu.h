/*
 * brgpu.h — C ABI of the B200-native hot path of natir/br ("Brutal Rewrite").
 *
 * This is the drop-in boundary: the entry points are what a Rust `brgpu-sys` FFI crate
 * (or any other host) binds to replace, inside br,
 *   part 1  the k-mer count -> solidity threshold -> canonical bitfield pass
 *           (src/main.rs:72-115, src/set/pcon.rs, pcon::{counter,spectrum,solid}), and
 *   part 2  the per-read correction pass
 *           (src/correct/ modules, the chunk loop of src/lib.rs:72-139).
 * Every function is `extern "C"`, takes plain pointers and sizes, returns an int status
 * and never throws, aborts or prints across the boundary (tests/br.rs:30 demands an empty
 * stderr).  All buffers named `host` below are caller-owned host memory (pinned memory
 * makes the copies faster, pageable memory works); the library never frees caller memory.
 * There is NO CPU fallback: without a CUDA device every entry point that needs one fails
 * with BRGPU_E_NO_DEVICE.
 *
 * Supported k: odd, 3 <= k <= 19 for the dense bitfield of 2^(2k-1) bits (set::Pcon; br's `fasta`
 * sub-command forces k odd, src/cli.rs:277-279; the parity-canonical form needs it), and
 * 3 <= k <= 31 for the hash set (set::Hash, br's `large-kmer` sub-command).
 */
#ifndef BRGPU_H
#define BRGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (replace error::Error / anyhow at the boundary, src/error.rs:12-45) ---- */
enum {
    BRGPU_OK = 0,
    BRGPU_E_INVALID = 1,      /* bad argument (NULL, even k, confirm == 0, unknown method, ...) */
    BRGPU_E_NO_DEVICE = 2,    /* no usable CUDA device: the product has no CPU path */
    BRGPU_E_CUDA = 3,         /* a CUDA call failed; see brgpu_last_error */
    BRGPU_E_NOMEM = 4,        /* host or device allocation failed */
    BRGPU_E_OVERFLOW = 5,     /* caller's output buffer too small; *required holds the size */
    BRGPU_E_NO_THRESHOLD = 6, /* Error::ComputeAbundanceThreshold (src/main.rs:97) */
    BRGPU_E_NEED_ABUNDANCE = 7 /* Error::AbundanceThresholdOrAbundanceMethod (src/main.rs:109) */
};

/* ---- correction methods: cli::CorrectionMethod in declaration order (src/cli.rs:121-131) ---- */
enum { BRGPU_ONE = 0, BRGPU_TWO = 1, BRGPU_GRAPH = 2, BRGPU_GREEDY = 3, BRGPU_GAP_SIZE = 4 };

/* ---- abundance selection: cli::AbundanceSelection (src/cli.rs:227-241) and the pcon
 * ThresholdMethod each maps to (src/main.rs:97-108): FirstMinimum, Rarefaction{percent},
 * PercentMost{percent} -> PercentAtMost, PercentLeast{percent} -> PercentAtLeast ---- */
enum {
    BRGPU_ABUNDANCE_EXPLICIT = 0,
    BRGPU_ABUNDANCE_FIRST_MINIMUM = 1,
    BRGPU_ABUNDANCE_RAREFACTION = 2,
    BRGPU_ABUNDANCE_PERCENT_AT_MOST = 3,
    BRGPU_ABUNDANCE_PERCENT_AT_LEAST = 4
};

typedef struct brgpu_ctx brgpu_ctx;       /* one per process per GPU: device, stream, scratch */
typedef struct brgpu_set brgpu_set;       /* Box<dyn KmerSet> (src/set.rs:17-23), device resident */
typedef struct brgpu_reads brgpu_reads;   /* a chunk of records (src/lib.rs:90), device resident */
typedef struct brgpu_counts brgpu_counts; /* pcon::counter::Counter<u8>, device resident */
typedef struct brgpu_kmers brgpu_kmers;   /* a chunk's canonical k-mers, partitioned by table-index range */

/* ------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------- */
/* `cuda_stream` may be NULL (the library creates its own) or an existing cudaStream_t so
 * that the host's own events/allocator order against the library's work. */
int brgpu_ctx_create(int device, void *cuda_stream, brgpu_ctx **out);
void brgpu_ctx_destroy(brgpu_ctx *ctx);
int brgpu_ctx_synchronize(brgpu_ctx *ctx);
/* Switches for tests and A/B measurements; none changes a result.  Defaults come from the
 * environment once, at brgpu_ctx_create (BRGPU_NO_COMPACT, BRGPU_ONE_LEVEL_PARTITION,
 * BRGPU_COUNT_BLOCK_ONLY, BRGPU_NO_POS8, BRGPU_COMPACT_MAX_PCT, BRGPU_NO_FINE_SUMMARY, BRGPU_SCAN=warp|groups).  name: "no_compact" (0/1: solidity lookups through summary + bitfield
 * even for sparse sets), "one_level_partition" (0/1: the k = 19 partition path for k <= 17),
 * "no_pos8" (0/1: lookups of a rank-compacted set read its 64-bit blocks instead of the one-byte-per-block form),
 * "compact_max_pct" (0..100, default 50: a set is held rank-compacted while its occupied 64-bit blocks take at most this
 * share of the bitfield's bytes), "no_fine_summary" (0/1: denser sets without the one-bit-per-16 occupancy summary),
 * "count_block_only" (threads per bucket of the counting kernel: 0 or 1 = 256, the default; 2 = 128; 3 = 64),
 * "scan_mode" (0 per-method default, 1 warp per segment, 2 four segments per warp, for One/Two). */
int brgpu_ctx_set_option(brgpu_ctx *ctx, const char *name, int value);
const char *brgpu_last_error(const brgpu_ctx *ctx);
const char *brgpu_version(void);

/* Page-locked host memory for the buffers handed to the calls below (optional: pageable
 * memory works, pinned memory lets the copies run at PCIe speed and overlap with kernels). */
int brgpu_host_alloc(brgpu_ctx *ctx, size_t bytes, void **out);
void brgpu_host_free(brgpu_ctx *ctx, void *p);

/* ------------------------------------------------------------------------------------------
 * reads — the record chunk run_correction forms (src/lib.rs:90, populate_buffer :168-188).
 * `seq` is the concatenation of the sequences (any bytes; nuc2bit is (b>>1)&3),
 * `offsets` has n_reads+1 entries.
 * ---------------------------------------------------------------------------------------- */
int brgpu_reads_upload(brgpu_ctx *ctx, const uint8_t *seq_host, const uint64_t *offsets_host, uint64_t n_reads,
                       brgpu_reads **out);
uint64_t brgpu_reads_count(const brgpu_reads *reads);
uint64_t brgpu_reads_bases(const brgpu_reads *reads); /* sum of lengths (syncs the stream) */
/* Copy back as concatenated bytes + offsets (n_reads+1).  BRGPU_E_OVERFLOW if seq_cap is too
 * small; *required (may be NULL) receives the byte count either way. */
int brgpu_reads_download(brgpu_reads *reads, uint8_t *seq_host, uint64_t seq_cap, uint64_t *offsets_host,
                         uint64_t *required);
void brgpu_reads_free(brgpu_reads *reads);
/* Asynchronous staging for pipelined chunks (what a host loop over run_correction's chunks,
 * src/lib.rs:86-136, uses to hide PCIe behind the kernels): the context owns a copy stream;
 * brgpu_reads_upload_async returns at once and any later call that takes the reads waits for the
 * copy on the device; brgpu_reads_download_async enqueues the copy back (the call that produced
 * `reads` must have returned) and brgpu_reads_download_wait blocks until the bytes are in the
 * host buffers.  Host buffers should come from brgpu_host_alloc (pageable memory makes the copies
 * synchronous) and must not be touched between the call and the wait / the first consumer. */
int brgpu_reads_upload_async(brgpu_ctx *ctx, const uint8_t *seq_host, const uint64_t *offsets_host, uint64_t n_reads,
                             brgpu_reads **out);
int brgpu_reads_download_async(brgpu_reads *reads, uint8_t *seq_host, uint64_t seq_cap, uint64_t *offsets_host,
                               uint64_t *required);
int brgpu_reads_download_wait(brgpu_reads *reads);

/* 2-bit transport (BASELINE.json north_star: "2-bit packed reads"): the same chunk at a quarter
 * of the PCIe bytes in both directions.  `packed` holds the concatenated sequences at 2 bits per base
 * (code (b >> 1) & 3, four bases per byte, first base in the two high bits; offsets[] are in bases and start
 * at 0); exc_pos / exc_byte list every byte that is not the upper-case letter of its own code (lower case, N,
 * any other byte) by its base position in the concatenation.  On the device the reads are ASCII again, so a
 * position the correctors leave alone keeps its original byte (src/correct/mod.rs:91,100).  The download
 * returns packed bases, offsets and the exceptions that survive (a subset of the input's: corrections only
 * emit A, C, T, G — so exc_cap = the input's exception count always suffices), in no particular order;
 * counts_host[0] = bases, counts_host[1] = exceptions found.  The _async forms follow
 * brgpu_reads_upload_async / _download_async (brgpu_reads_download_wait completes either download; the
 * counts are valid after the wait).  br::fasta (br_b200/host/fasta.hpp) packs while it parses and expands
 * while it formats; br_b200.runtime has the numpy equivalents. */
int brgpu_reads_upload_packed(brgpu_ctx *ctx, const uint8_t *packed_host, const uint64_t *offsets_host, uint64_t n_reads,
                              const uint64_t *exc_pos_host, const uint8_t *exc_byte_host, uint64_t n_exc, brgpu_reads **out);
int brgpu_reads_upload_packed_async(brgpu_ctx *ctx, const uint8_t *packed_host, const uint64_t *offsets_host,
                                    uint64_t n_reads, const uint64_t *exc_pos_host, const uint8_t *exc_byte_host,
                                    uint64_t n_exc, brgpu_reads **out);
int brgpu_reads_download_packed(brgpu_reads *reads, uint8_t *packed_host, uint64_t packed_cap, uint64_t *offsets_host,
                                uint64_t *exc_pos_host, uint8_t *exc_byte_host, uint64_t exc_cap, uint64_t counts_host[2]);
int brgpu_reads_download_packed_async(brgpu_reads *reads, uint8_t *packed_host, uint64_t packed_cap,
                                      uint64_t *offsets_host, uint64_t *exc_pos_host, uint8_t *exc_byte_host,
                                      uint64_t exc_cap, uint64_t counts_host[2]);

/* Measurement support (bench.py): synthetic ONT-like reads generated on the device, so
 * that the 5 and 30 Gbase workloads of BASELINE.json's configs[3]/[4] never cross PCIe.  The genome is
 * a pure function of (genome_seed, position) and is never stored; read r of this call is read number
 * first_read_id + r of the data set: template = genome[start, start + tlen), reverse-complemented when
 * strand != 0, then per-base substitutions / insertions / deletions drawn from a counter-based hash
 * of (read_seed, read number, template position) against the three cumulative 24-bit thresholds.
 * br_b200/synth.py holds the numpy mirror that produces the same bytes. */
int brgpu_reads_synth(brgpu_ctx *ctx, uint64_t genome_seed, uint64_t read_seed, uint64_t first_read_id,
                      const uint64_t *start_host, const uint32_t *tlen_host, const uint8_t *strand_host,
                      uint64_t n_reads, const uint32_t thresholds[3], brgpu_reads **out);

/* ------------------------------------------------------------------------------------------
 * part 1 — counting and the solid set
 * ---------------------------------------------------------------------------------------- */
/* pcon Counter::<u8>::new(k) (src/main.rs:73): 2^(2k-1) zeroed saturating u8 counters in HBM */
int brgpu_counts_create(brgpu_ctx *ctx, int k, brgpu_counts **out);
/* Counter::count_fasta (src/main.rs:74): counts[canonical(kmer)>>1] = min(255, +1) for every
 * k-mer of every read with len >= k.  May be called repeatedly (chunks accumulate). */
int brgpu_counts_add_reads(brgpu_counts *counts, const brgpu_reads *reads);
/* Spectrum::from_count (src/main.rs:93): 256-bin histogram over all counters incl. zeros */
int brgpu_counts_spectrum(brgpu_counts *counts, uint64_t hist_host[256]);
/* Spectrum::get_threshold(FirstMinimum): returns the threshold or -1 for None */
int brgpu_spectrum_first_minimum(const uint64_t hist[256]);
/* Spectrum::get_threshold(method, percent) for any BRGPU_ABUNDANCE_* method but EXPLICIT
 * (src/main.rs:95-108); -1 for None.  Host arithmetic on the 256 bins the GPU produced. */
int brgpu_spectrum_threshold(const uint64_t hist[256], int selection, double percent);
/* Counter::from_stream (src/main.rs:59-70, the `count` sub-command): the raw table of a pcon count file —
 * the host reads and decompresses the file — replaces the table's content; n must be 2^(2k-1) */
int brgpu_counts_upload(brgpu_counts *counts, const uint8_t *counts_host, uint64_t n);
/* Copy the raw table to the host (Counter::raw, src/main.rs:76-78); n must be 2^(2k-1) */
int brgpu_counts_download(brgpu_counts *counts, uint8_t *out_host, uint64_t n);
/* device pointer and element count of the table, for multi-GPU merge plumbing */
void *brgpu_counts_device_ptr(brgpu_counts *counts);
uint64_t brgpu_counts_len(const brgpu_counts *counts);
void brgpu_counts_free(brgpu_counts *counts);

/* Solid::from_count (src/main.rs:112-114): bit i = counts[i] > abundance, LSB-first */
int brgpu_set_from_counts(brgpu_counts *counts, int abundance, brgpu_set **out);
/* the whole `fasta` sub-command glue (src/main.rs:72-115): count -> spectrum -> threshold ->
 * bitfield.  selection = BRGPU_ABUNDANCE_EXPLICIT uses `abundance` (>= 0); FIRST_MINIMUM
 * ignores it.  abundance < 0 with EXPLICIT -> BRGPU_E_NEED_ABUNDANCE. */
int brgpu_set_from_reads(brgpu_ctx *ctx, int k, int abundance, int selection, const brgpu_reads *reads,
                         brgpu_set **out);
/* same with the percent argument of Rarefaction / PercentAtMost / PercentAtLeast */
int brgpu_set_from_reads_ex(brgpu_ctx *ctx, int k, int abundance, int selection, double percent,
                            const brgpu_reads *reads, brgpu_set **out);
int brgpu_set_from_host_reads_ex(brgpu_ctx *ctx, int k, int abundance, int selection, double percent,
                                 const uint8_t *seq_host, const uint64_t *offsets_host, uint64_t n_reads,
                                 brgpu_set **out);
/* same, from host buffers (the call the Rust shim makes) */
int brgpu_set_from_host_reads(brgpu_ctx *ctx, int k, int abundance, int selection, const uint8_t *seq_host,
                              const uint64_t *offsets_host, uint64_t n_reads, brgpu_set **out);
/* set::Pcon::new(Solid) (src/set/pcon.rs:183) from a raw bitfield of 2^(2k-1)/8 bytes */
int brgpu_set_from_bitfield(brgpu_ctx *ctx, int k, const uint8_t *bits_host, uint64_t n_bytes, brgpu_set **out);
/* set::Pcon::from_pcon_solid (src/set/pcon.rs:18-25) after the host gunzipped the stream:
 * payload[0] = k, payload[1..] = bitfield */
int brgpu_set_from_solid_payload(brgpu_ctx *ctx, const uint8_t *payload_host, uint64_t n_bytes, brgpu_set **out);
/* empty set, Solid::new(k), and Solid::set on a batch of (forward or canonical) k-mers —
 * what the reference's unit tests build their sets with */
int brgpu_set_new(brgpu_ctx *ctx, int k, brgpu_set **out);
int brgpu_set_insert_batch(brgpu_set *set, const uint64_t *kmers_host, uint64_t n);

/* set::Hash (src/set/hash.rs:14-186, wired at src/main.rs:147-163: the `large-kmer` sub-command): the
 * solid set for k-mers too large for a dense bitfield — 3 <= k <= 31, any parity — as a device hash
 * table of canonical k-mers behind the same handle type: brgpu_set_k / _get_batch / _insert_batch /
 * _free and all the correction calls take it; the bitfield calls return BRGPU_E_INVALID for it.
 * _hash_new = an empty Hash (expected_kmers sizes the first table; it grows on demand),
 * _hash_add_reads = Hash::from_fasta's loop over one chunk of records (presence only: every canonical
 * k-mer of every record with len >= k, src/set/hash.rs:52-57); may be called once per chunk. */
int brgpu_set_hash_new(brgpu_ctx *ctx, int k, uint64_t expected_kmers, brgpu_set **out);
int brgpu_set_hash_add_reads(brgpu_set *set, const brgpu_reads *reads);
int brgpu_set_hash_from_reads(brgpu_ctx *ctx, int k, const brgpu_reads *reads, brgpu_set **out);
int brgpu_set_hash_from_host_reads(brgpu_ctx *ctx, int k, const uint8_t *seq_host, const uint64_t *offsets_host,
                                   uint64_t n_reads, brgpu_set **out);
int brgpu_set_is_hash(const brgpu_set *set);
uint64_t brgpu_set_hash_size(const brgpu_set *set);         /* distinct canonical k-mers held */

int brgpu_set_k(const brgpu_set *set);                      /* KmerSet::k */
int brgpu_set_abundance(const brgpu_set *set);              /* threshold used, -1 if not built from counts */
uint64_t brgpu_set_bitfield_bytes(const brgpu_set *set);    /* 2^(2k-1)/8 */
int brgpu_set_export_bitfield(brgpu_set *set, uint8_t *out_host, uint64_t cap);
/* KmerSet::get over a batch; forward k-mers are canonicalised inside (src/set/pcon.rs:189) */
int brgpu_set_get_batch(brgpu_set *set, const uint64_t *kmers_host, uint64_t n, uint8_t *out_host);
/* spectrum seen when the set was built from reads/counts (zeros if not) */
int brgpu_set_spectrum(const brgpu_set *set, uint64_t hist[256]);
void *brgpu_set_device_ptr(brgpu_set *set);
/* Sharded construction fills a set slice by slice: _new_sliced = Solid::new(k) without the zero-fill and with
 * the occupancy summary allocated, so that brgpu_kmers_count_* write the slice's summary words next to its
 * bitfield bits; _summary_ptr exposes the summary (one bit per 64 bitfield bits; slices are proportional to
 * the bitfield's) for the same all-gather; _commit_slices declares bitfield (and summary, if
 * summary_complete) whole and builds the lookup structures — no pass over the 1 GiB bitfield. */
int brgpu_set_new_sliced(brgpu_ctx *ctx, int k, brgpu_set **out);
/* A sparse set travels cheaper in rank-compacted form (§2 of DESIGN.md: occupied 64-bit blocks in index order):
 * _slice_compact compacts this GPU's slice [bit_begin, bit_end) (multiples of 2048) and returns its blocks and
 * their number (*blocks_dev = NULL: not available for this set, exchange the bitfield); _compact_alloc sizes the
 * replica's block array for the total over all slices and returns it for the host's exchange to fill in slice
 * order; _compact_commit (summary slices gathered too) builds the rank directory.  The dense bitfield of such a
 * set is rebuilt on demand (brgpu_set_export_bitfield, brgpu_set_insert_batch, brgpu_set_device_ptr). */
int brgpu_set_slice_compact(brgpu_set *set, uint64_t bit_begin, uint64_t bit_end, void **blocks_dev, uint64_t *n_blocks);
int brgpu_set_compact_alloc(brgpu_set *set, uint64_t n_blocks_total, void **blocks_dev);
int brgpu_set_compact_commit(brgpu_set *set);
/* The exchange itself over peer memory instead of a collective: _slice_ipc_export hands out the CUDA-IPC handle of
 * the compacted slice (processes) — inside one process the peer's pointer is used as it is; _compact_pull, after
 * _compact_alloc, copies every slice (slice_blocks[i] == NULL: this GPU's own) to its place in the block array with
 * ONE kernel that has all peers' NVLink loads in flight together.  The caller orders it after every peer's
 * _slice_compact and keeps the slices alive until every peer has pulled (dist.py: the two small collectives either
 * side of it; group.cu: the joins). */
int brgpu_set_slice_ipc_export(brgpu_set *set, uint8_t handle_out[64]);
int brgpu_set_compact_pull(brgpu_set *set, void *const *slice_blocks, const uint64_t *n_blocks, int n_slices);
void *brgpu_set_summary_ptr(brgpu_set *set, uint64_t *n_bytes);
int brgpu_set_commit_slices(brgpu_set *set, int summary_complete);
void brgpu_set_free(brgpu_set *set);

/* ------------------------------------------------------------------------------------------
 * part 2 — correction.  Semantics of one record: src/lib.rs:44-55 — fold `methods` left to
 * right with Corrector::correct (src/correct/mod.rs:53-107); unless two_side, reverse the
 * bytes, fold again, reverse back.  Records come back in input order (the serial path's order,
 * src/lib.rs:21-69).  confirm: -C (One/Two/GapSize `c`, Greedy `nb_validate`, src/lib.rs:149-159),
 * max_search: -M (Greedy).  confirm must be in 1..=255, max_search in 0..=255.
 * ---------------------------------------------------------------------------------------- */
int brgpu_correct_reads(brgpu_ctx *ctx, const brgpu_set *set, const uint8_t *methods, uint64_t n_methods, int confirm,
                        int max_search, int two_side, const brgpu_reads *in, brgpu_reads **out);

/* brgpu_correct_reads without the wait at its end: the chain is enqueued and the call returns, so that the
 * host can stage the next chunk and issue the previous chunk's download while the kernels run.  The first
 * call that takes *out (a download, brgpu_reads_bases, another correction, brgpu_reads_wait) waits for the
 * chain and, if a read outgrew its slot, redoes it synchronously — `set` and `in` must stay alive until
 * then.  brgpu_reads_wait returns the status of that step. */
int brgpu_correct_reads_async(brgpu_ctx *ctx, const brgpu_set *set, const uint8_t *methods, uint64_t n_methods,
                              int confirm, int max_search, int two_side, const brgpu_reads *in, brgpu_reads **out);
int brgpu_reads_wait(brgpu_reads *reads);

int brgpu_correct_batch(brgpu_ctx *ctx, const brgpu_set *set, const uint8_t *methods, uint64_t n_methods, int confirm,
                        int max_search, int two_side, const uint8_t *seq_host, const uint64_t *offsets_host,
                        uint64_t n_reads, uint8_t *out_host, uint64_t out_cap, uint64_t *out_offsets_host,
                        uint64_t *required);

/* Corrector::correct for a single read and a single method (KAT parity; a Rust
 * `GpuCorrector: Corrector` would call this). */
int brgpu_correct_one(brgpu_ctx *ctx, const brgpu_set *set, int method, int confirm, int max_search,
                      const uint8_t *seq_host, uint64_t len, uint8_t *out_host, uint64_t out_cap, uint64_t *out_len);

/* ------------------------------------------------------------------------------------------
 * multi-GPU (one process per GPU): merge per-rank count tables over NVLink peer memory.
 * Each rank exports its table with brgpu_counts_ipc_export, the host exchanges the 64-byte
 * handles (torch.distributed / MPI / anything), every rank opens its peers' tables and then
 * calls brgpu_counts_merge_slice: counts[i] = min(255, sum over ranks) on this rank's
 * 1/world slice of the index space, reading the peers' slices straight over NVLink.
 * ---------------------------------------------------------------------------------------- */
int brgpu_counts_ipc_export(brgpu_counts *counts, uint8_t handle_out[64]);
int brgpu_ipc_open(brgpu_ctx *ctx, const uint8_t handle[64], void **peer_dev_ptr);
int brgpu_ipc_close(brgpu_ctx *ctx, void *peer_dev_ptr);
/* saturating-add the peers' [begin,end) slices into this rank's table slice */
int brgpu_counts_merge_slice(brgpu_counts *counts, void *const *peer_tables, int n_peers, uint64_t begin, uint64_t end);
/* spectrum / threshold restricted to a slice [begin,end) of the index space; the bitfield
 * slice is written into set's bitfield at the same bit range (begin, end multiples of 1024) */
int brgpu_counts_spectrum_slice(brgpu_counts *counts, uint64_t begin, uint64_t end, uint64_t hist_host[256]);
int brgpu_set_threshold_slice(brgpu_set *set, brgpu_counts *counts, int abundance, uint64_t begin, uint64_t end);

/* Sharded set construction without count tables (the default multi-GPU path, k >= 15).
 * brgpu_kmers_create partitions a chunk's canonical k-mers into buckets of 2^15 consecutive
 * table indices (16-bit residues).  Every rank exports its partition (two CUDA-IPC handles:
 * residues, bucket offsets), owns a contiguous bucket range, and brgpu_kmers_count_range counts
 * that range over its own and all peers' partitions — the peers' residues are read straight
 * out of their HBM over NVLink inside the counting kernel.  The spectrum of the range is
 * returned (sum it over ranks); if `set` is given the range is thresholded into set's bitfield
 * (bits [bucket_begin << 15, bucket_end << 15)); the host then all-gathers the bitfield slices
 * through brgpu_set_device_ptr. */
int brgpu_kmers_create(brgpu_ctx *ctx, int k, const brgpu_reads *reads, brgpu_kmers **out);
uint64_t brgpu_kmers_buckets(const brgpu_kmers *kmers);
/* device pointer of the bucket offsets (buckets + 1 u64): a host that exchanges them with a device-side
 * collective needs no round trip through brgpu_kmers_offsets_at */
void *brgpu_kmers_offsets_ptr(brgpu_kmers *kmers);
int brgpu_kmers_ipc_export(brgpu_kmers *kmers, uint8_t handles_out[128]); /* [0..64) residues, [64..128) offsets */
int brgpu_kmers_count_range(brgpu_kmers *kmers, void *const *peer_residues, void *const *peer_offsets, int n_peers,
                            uint64_t bucket_begin, uint64_t bucket_end, int abundance, brgpu_set *set,
                            uint64_t hist_host[256]);
/* Offsets (in k-mers) at which the given buckets start inside this rank's partition: the rank
 * that owns bucket range [b0, b1) needs residues [offset(b0), offset(b1)) of every peer.  The host
 * exchanges these along with the IPC handles. */
int brgpu_kmers_offsets_at(brgpu_kmers *kmers, const uint64_t *buckets_host, uint64_t n, uint64_t *offsets_host);
/* brgpu_kmers_count_range with the peers' residues of the range — [peer_first[p], peer_last[p]) of
 * peer p's residue array, contiguous because buckets are stored in order — first pulled into local
 * HBM by bulk peer-to-peer copies over NVLink (full-bandwidth transfers instead of ~1 KB remote
 * loads per bucket and peer), then counted from local memory. */
int brgpu_kmers_count_range_staged(brgpu_kmers *kmers, void *const *peer_residues, void *const *peer_offsets,
                                   const uint64_t *peer_first, const uint64_t *peer_last, int n_peers,
                                   uint64_t bucket_begin, uint64_t bucket_end, int abundance, brgpu_set *set,
                                   uint64_t hist_host[256]);
/* The general form: bucket range over `n_local` partitions of this context (one per chunk of reads —
 * a GPU whose shard does not fit one chunk holds several) plus `n_peers` peer partitions; peer_first /
 * peer_last may be NULL (remote loads instead of staged copies).  n_local + n_peers <= 64. */
int brgpu_kmers_count_parts(brgpu_ctx *ctx, brgpu_kmers *const *local, int n_local, void *const *peer_residues,
                            void *const *peer_offsets, const uint64_t *peer_first, const uint64_t *peer_last, int n_peers,
                            uint64_t bucket_begin, uint64_t bucket_end, int abundance, brgpu_set *set,
                            uint64_t hist_host[256]);
/* The `fasta` sub-command over a stream of chunks (src/main.rs:72-78, count_fasta(inputs, 8192) reads the
 * records chunk by chunk): upload a chunk, brgpu_kmers_create, free the chunk, repeat; then build the set
 * from all partitions at once.  Peak device memory is 2 B per k-mer of the input plus one chunk, and the
 * result equals brgpu_set_from_reads over all the reads.  k >= 15 (smaller k: brgpu_counts_add_reads per
 * chunk + brgpu_counts_spectrum + brgpu_set_from_counts, whose table is at most 128 MiB).  n_parts <= 64. */
int brgpu_set_from_kmers(brgpu_ctx *ctx, brgpu_kmers *const *parts, int n_parts, int abundance, int selection,
                         double percent, brgpu_set **out);
void brgpu_kmers_free(brgpu_kmers *kmers);

/* ------------------------------------------------------------------------------------------
 * multi-GPU, one process that owns several GPUs (SURVEY §8b "Threading"): a group is one context per
 * device with peer access between every pair; no torch.distributed, no CUDA IPC.  Part 1 is the sharded
 * protocol above, driven inside the library (device i counts bucket range i over all devices' partitions,
 * peer residues pulled over NVLink, slices pushed to every peer; k < 15: count tables + saturating merge);
 * the result is one replicated set per device, bit-identical to the single-GPU set.  Part 2 shards the
 * records by bases, one host thread per device, and returns them in input order.
 * ---------------------------------------------------------------------------------------- */
typedef struct brgpu_group brgpu_group;
int brgpu_group_create(const int *devices, int n, brgpu_group **out); /* BRGPU_E_NO_DEVICE without peer access */
void brgpu_group_destroy(brgpu_group *group);
int brgpu_group_size(const brgpu_group *group);
brgpu_ctx *brgpu_group_ctx(brgpu_group *group, int i); /* for uploads into device i (brgpu_reads_upload, ...) */
const char *brgpu_group_last_error(const brgpu_group *group);
/* the `fasta` sub-command glue (src/main.rs:72-115) over the group: reads[i] lives in context i;
 * out_sets[i] receives device i's replica (free them with brgpu_group_sets_free or brgpu_set_free) */
int brgpu_group_set_from_reads(brgpu_group *group, int k, int abundance, int selection, double percent,
                               brgpu_reads *const *reads, brgpu_set **out_sets);
/* same from host buffers: the records are cut into contiguous ranges balanced by bases, one per device */
int brgpu_group_set_from_host_reads(brgpu_group *group, int k, int abundance, int selection, double percent,
                                    const uint8_t *seq_host, const uint64_t *offsets_host, uint64_t n_reads,
                                    brgpu_set **out_sets);
void brgpu_group_sets_free(brgpu_group *group, brgpu_set **sets);
/* brgpu_correct_batch over the group (same arguments, sets[i] = device i's replica) */
int brgpu_group_correct_batch(brgpu_group *group, brgpu_set *const *sets, const uint8_t *methods, uint64_t n_methods,
                              int confirm, int max_search, int two_side, const uint8_t *seq_host,
                              const uint64_t *offsets_host, uint64_t n_reads, uint8_t *out_host, uint64_t out_cap,
                              uint64_t *out_offsets_host, uint64_t *required);

/* ------------------------------------------------------------------------------------------
 * instrumentation (bench.py): per-kernel CUDA-event timings on the library's stream
 * ---------------------------------------------------------------------------------------- */
int brgpu_profile_enable(brgpu_ctx *ctx, int on);
int brgpu_profile_reset(brgpu_ctx *ctx);
/* number of distinct kernels seen; i-th entry: name, accumulated ms, launches, algorithmic bytes */
int brgpu_profile_count(brgpu_ctx *ctx);
int brgpu_profile_get(brgpu_ctx *ctx, int i, char *name_out, size_t name_cap, double *ms, uint64_t *launches,
                      double *algo_bytes);
/* KmerSet::get calls the i-th kernel issued (scan / merge kernels; 0 for the others).  Counted only
 * while profiling is enabled: the product path runs kernels compiled without the bookkeeping. */
int brgpu_profile_get_lookups(brgpu_ctx *ctx, int i, uint64_t *lookups);
/* Random 8-byte gathers per second over a table of `table_bytes` (8 loads in flight per thread, the
 * access pattern of the solidity lookups): an L2-resident table gives the L2 gather ceiling, a table
 * much larger than L2 the DRAM random-sector ceiling.  bench.py measures both in the run it reports. */
int brgpu_probe_random_gather(brgpu_ctx *ctx, uint64_t table_bytes, double *gathers_per_s);
uint64_t brgpu_launch_count(const brgpu_ctx *ctx); /* kernels launched since ctx creation */
/* KmerSet::get calls issued by the correction scans while profiling was enabled, since the last
 * brgpu_profile_reset (excludes the one lookup per base of the bitmap pass); syncs the stream */
uint64_t brgpu_scan_lookups(brgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* BRGPU_H */
