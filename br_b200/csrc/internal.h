// internal.h — host-side objects behind the opaque handles of include/brgpu.h and the
// kernel launchers the API layer calls.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/brgpu.h"

namespace brgpu {

struct ProfEntry {
    std::string name;
    double ms = 0.0;
    uint64_t launches = 0;
    unsigned next_flag_slot = 0; // rotating slots of h_pinned for the overflow flags of asynchronous corrections
    double bytes = 0.0;
    uint64_t lookups = 0; // KmerSet::get calls the kernel issued (scan kernels, counted while profiling)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

} // namespace brgpu

struct brgpu_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // second stream for the asynchronous host<->device staging calls (brgpu_reads_upload_async /
    // _download_async): copies of the neighbouring chunks run while this chunk's kernels do
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_fence = nullptr; // orders the two streams against each other
    // third stream: memsets that run beside the kernels of the compute stream (the bitfield of a set under construction)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_aux_in = nullptr, ev_aux_out = nullptr;
    std::string err;
    bool profiling = false;
    std::vector<brgpu::ProfEntry> prof;
    std::vector<cudaEvent_t> event_pool;
    uint64_t launches = 0;
    unsigned next_flag_slot = 0; // rotating slots of h_pinned for the overflow flags of asynchronous corrections
    // small device scratch reused by every call
    uint32_t *d_flags = nullptr; // [0] work-queue cursor, [1] overflow flag, [2..] spare
    uint64_t *d_hist = nullptr;  // 256 bins
    uint64_t *h_pinned = nullptr; // 512 x u64 pinned staging for small readbacks
    // per-kernel-name KmerSet::get counters (profiling runs only): slot i belongs to prof[i]; the last
    // slot collects launches that have no entry
    static constexpr int GET_SLOTS = 64;
    unsigned long long *d_getcnt = nullptr;
    // test / A-B switches: defaults from the environment at context creation (BRGPU_NO_COMPACT,
    // BRGPU_ONE_LEVEL_PARTITION, BRGPU_SCAN=warp|groups), changed with brgpu_ctx_set_option
    int opt_no_compact = 0;          // lookups through summary + bitfield even for sparse sets
    int opt_one_level_partition = 0; // the k = 19 partition path for k <= 17
    int opt_count_block_only = 0;    // the 256-thread shape of the counting kernel even for sparse buckets
    int opt_scan_mode = 0;           // 0: per method default, 1: warp per segment, 2: four segments per warp
    int opt_compact_max_pct = 50;    // the compacted form is built while its 64-bit blocks take at most this share of the bitfield
    int opt_keep_summary = 0;        // A/B: saturated sets are still looked up through the summary
    int opt_fine_in_scans = 0;       // A/B: the scans, too, look dense sets up through the fine summary
    int opt_no_fine_summary = 0;     // A/B: dense sets are looked up through the one-bit-per-64 summary
    int opt_no_pos8 = 0;             // A/B: lookups read the 64-bit blocks instead of the one-byte form
    // caching device allocator (brgpu.cu): blocks handed out (ptr -> bytes) and cached free blocks
    std::unordered_map<void *, uint64_t> pool_live;
    std::vector<std::pair<void *, uint64_t>> pool_free;
    uint64_t pool_free_bytes = 0;
    // blocks whose CUDA-IPC handle went out (peers may keep them mapped): never cudaFree'd before the context dies
    std::unordered_map<void *, bool> pool_exported;
};

namespace brgpu {

// Where every read lives in the slot layout: read r owns bytes [slot_off[r], slot_off[r+1]),
// slot_off[r] % 32 == 0, capacity = len + slack so that a corrected read can grow in place.
struct Layout {
    brgpu_ctx *ctx = nullptr;
    uint64_t n = 0;
    uint64_t total_slots = 0;        // bytes, multiple of 32
    uint64_t *d_slot_off = nullptr;  // n + 1
    uint32_t *d_word2read = nullptr; // total_slots / 32
    uint32_t *d_order = nullptr;     // read ids, longest first (work queue order)
    ~Layout();
};

} // namespace brgpu

struct brgpu_reads {
    brgpu_ctx *ctx = nullptr;
    std::shared_ptr<brgpu::Layout> layout;
    uint8_t *d_seq = nullptr;  // total_slots bytes
    uint32_t *d_len = nullptr; // n
    std::vector<uint32_t> h_len; // host copy of the lengths (uploaded reads only)
    uint64_t sum_len = 0;        // sum of lengths at upload (bookkeeping hint afterwards)
    // asynchronous staging (copy stream)
    cudaEvent_t ready = nullptr;     // upload in flight: consumers make the compute stream wait for it
    std::vector<void *> deferred;    // upload temporaries, released once a consumer has ordered itself after `ready`
    cudaEvent_t dl_done = nullptr;   // download in flight
    // asynchronous correction (brgpu_correct_reads_async): the chain is enqueued, its overflow flag has not been
    // looked at yet; the first consumer (reads_ready) synchronises and, in the rare overflow case, redoes the
    // chain synchronously from the input the caller promised to keep alive
    bool pending = false;
    int flag_slot = 0; // index into ctx->h_pinned
    const brgpu_set *redo_set = nullptr;
    const brgpu_reads *redo_in = nullptr;
    std::vector<uint8_t> redo_methods;
    int redo_confirm = 0, redo_max_search = 0, redo_two_side = 0;
    int resolve_status = 0; // BRGPU_OK, or what the redo returned
    void *dl_tight = nullptr, *dl_toff = nullptr;
    std::vector<void *> dl_more;     // further device buffers of an in-flight packed download
};

struct brgpu_counts {
    brgpu_ctx *ctx = nullptr;
    int k = 0;
    uint64_t n = 0; // 2^(2k-1)
    uint8_t *d_counts = nullptr;
};

// a chunk's k-mers partitioned by table-index range: 16-bit residues grouped by index >> 15
struct brgpu_kmers {
    brgpu_ctx *ctx = nullptr;
    int k = 0;
    uint64_t n_buckets = 0;
    uint64_t capacity = 0;       // residues allocated (u16 each)
    uint16_t *d_res = nullptr;   // cudaMalloc'ed: exported over CUDA IPC
    uint64_t *d_base = nullptr;  // n_buckets + 1, cudaMalloc'ed
    double n_kmers_hint = 0;
};

struct brgpu_set {
    brgpu_ctx *ctx = nullptr;
    int k = 0;
    int abundance = -1;
    uint64_t n_bytes = 0;
    uint8_t *d_bits = nullptr;
    uint64_t hist[256] = {0};
    // L2-resident occupancy summary of d_bits (see SolidView in kmer.cuh); rebuilt lazily
    uint32_t *d_summary = nullptr;
    uint64_t summary_bytes = 0;
    int summary_shift = 0;
    bool summary_valid = false;
    // rank-compacted copy of a sparse bitfield (see SolidView in kmer.cuh); rebuilt with the summary
    void *d_dir = nullptr;       // uint2 per summary word
    uint64_t dir_bytes = 0;
    uint64_t *d_blocks = nullptr;
    uint64_t blocks_bytes = 0;   // allocation size
    uint64_t n_occupied = 0;     // occupied 64-bit blocks
    uint8_t *d_pos8 = nullptr;   // one byte per occupied block (SolidView::pos8); nullptr: not built
    uint32_t *d_fine = nullptr;  // dense sets: occupancy summary at one bit per 16 bitfield bits (nullptr: not built)
    uint64_t fine_bytes = 0;
    uint64_t pos8_bytes = 0;
    bool compact_valid = false;  // d_dir/d_blocks describe the current bitfield
    // sharded construction may leave only this GPU's slice of the dense bitfield written and hold the whole set in
    // its rank-compacted form; the dense form is then rebuilt when somebody asks for it (export, insert, ...)
    bool bits_complete = true;
    uint64_t *d_slice_blocks = nullptr; // this GPU's slice, compacted (what it sends to its peers)
    uint64_t slice_blocks_bytes = 0;
    // set::Hash (src/set/hash.rs): open-addressing table of canonical k-mers instead of a bitfield
    bool is_hash = false;
    uint64_t *d_hash = nullptr;
    uint64_t hash_slots = 0;     // power of two
    uint64_t hash_size = 0;      // distinct k-mers held
};

namespace brgpu {

// RAII-less helper: record an error string and return a status
int fail(brgpu_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess);

// profiling hooks around a launch
void prof_begin(brgpu_ctx *ctx, const char *name, double algo_bytes, bool is_kernel = true);
void prof_end(brgpu_ctx *ctx);
// device counter for the KmerSet::get calls of the kernel whose ProfScope is open
unsigned long long *prof_counter_slot(brgpu_ctx *ctx);

// cudaFuncSetAttribute applies to the CURRENT device only: a process that drives several GPUs (brgpu_group_*)
// configures a kernel once per device, not once per process
struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    bool need(int device) const { return !((done.load(std::memory_order_acquire) >> (device & 63)) & 1ULL); }
    void mark(int device) { done.fetch_or(1ULL << (device & 63), std::memory_order_release); }
};

struct ProfScope {
    brgpu_ctx *c;
    ProfScope(brgpu_ctx *ctx, const char *name, double bytes, bool is_kernel = true) : c(ctx) {
        prof_begin(c, name, bytes, is_kernel);
    }
    ~ProfScope() { prof_end(c); }
};

// small device -> host results through mapped pinned memory (h_mapped_dst inside ctx->h_pinned); the
// caller synchronises the stream before reading.  bytes: multiple of 4
void launch_readback(brgpu_ctx *ctx, void *h_mapped_dst, const void *d_src, size_t bytes);

// ---- layout / relayout kernels (set_kernels.cu) ----
void launch_fill_word2read(brgpu_ctx *ctx, const Layout &L);
void launch_scatter_to_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_tight, const uint64_t *d_tight_off,
                             uint8_t *d_slots);
void launch_gather_from_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_slots, const uint32_t *d_len,
                              const uint64_t *d_tight_off, uint8_t *d_tight, bool reverse);
// 2-bit transport: packed tight layout (+ exception list) <-> ASCII slot layout
void launch_unpack_to_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_packed, const uint64_t *d_tight_off,
                            uint8_t *d_slots, const uint64_t *d_exc_pos, const uint8_t *d_exc_byte, uint64_t n_exc);
void launch_pack_tight(brgpu_ctx *ctx, const uint8_t *d_tight, uint64_t n_bases, uint8_t *d_packed, uint64_t *d_exc_pos,
                       uint8_t *d_exc_byte, uint64_t exc_cap, unsigned long long *d_n_exc);
void launch_reverse_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_in, const uint32_t *d_len, uint8_t *d_out);
void launch_exclusive_scan_u32(brgpu_ctx *ctx, const uint32_t *d_in, uint64_t n, uint64_t *d_out /* n+1 */,
                               uint64_t *d_tmp /* >= n/4096 + 2 */);

// ---- synthetic reads (synth_kernels.cu): seeds = {genome seed, read seed, first read id}, thr = cumulative
// 24-bit thresholds {substitution, insertion, deletion} ----
int synth_tile_positions();
void launch_synth_count(brgpu_ctx *ctx, const uint64_t seeds[3], const uint32_t thr[3], const uint64_t *d_start,
                        const uint32_t *d_tlen, const uint8_t *d_strand, const uint64_t *d_tile_first, uint64_t n_reads,
                        uint64_t n_tiles, uint32_t *d_tile_bytes);
void launch_synth_write(brgpu_ctx *ctx, const uint64_t seeds[3], const uint32_t thr[3], const uint64_t *d_start,
                        const uint32_t *d_tlen, const uint8_t *d_strand, const uint64_t *d_tile_first, uint64_t n_reads,
                        uint64_t n_tiles, const uint64_t *d_tile_off, uint8_t *d_out, uint64_t *d_read_off);

// ---- part 1 kernels (set_kernels.cu) ----
void launch_count(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                  uint8_t *d_counts, double n_bases_hint);
// histogram of counts[begin,end) into d_hist (256 x u64, accumulated) and, if d_bits != nullptr,
// bit i = counts[i] > abundance for the same range.  begin/end multiples of 1024.
void launch_spectrum_threshold(brgpu_ctx *ctx, const uint8_t *d_counts, uint64_t begin, uint64_t end, uint64_t *d_hist,
                               uint8_t *d_bits, int abundance);
// bucketed counting (see set_kernels.cu)
// d_coarse_kmers (u32 per k-mer) + d_coarse_base (u64 x 513): scratch of the two-level partition;
// nullptr (or more than 2^18 buckets) selects the one-level partition.  d_fill: max(n_buckets, 1024) words.
bool bucket_partition_two_level(uint64_t n_buckets);
void launch_bucket_partition(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                             uint64_t n_buckets, uint32_t *d_fill, uint64_t *d_base, uint64_t *d_scan_tmp,
                             uint16_t *d_residues, uint32_t *d_coarse_kmers, uint64_t *d_coarse_base, double n_kmers);
void launch_bucket_count(brgpu_ctx *ctx, const uint16_t *d_residues, const uint64_t *d_base, uint64_t n_buckets,
                         int abundance, uint8_t *d_bits, uint32_t *d_summary, int summary_shift, uint64_t *d_hist,
                         double n_kmers, bool prezeroed = false);
void launch_bucket_count_multi(brgpu_ctx *ctx, const uint16_t *const *d_res, const uint64_t *const *d_base, int n_src,
                               uint64_t b0, uint64_t b1, int abundance, uint8_t *d_bits, uint32_t *d_summary,
                               uint64_t *d_hist, double n_kmers);
constexpr int BRGPU_MAX_KMER_SOURCES = 64; // == BUCKET_MAX_SOURCES in set_kernels.cu
// one launch copying up to 128 (peer) segments into local HBM; src[q] and dst[q] congruent modulo 16, bytes[q] even
struct PullSegments {
    const uint8_t *src[2 * BRGPU_MAX_KMER_SOURCES];
    uint8_t *dst[2 * BRGPU_MAX_KMER_SOURCES];
    uint64_t bytes[2 * BRGPU_MAX_KMER_SOURCES];
    uint32_t first_chunk[2 * BRGPU_MAX_KMER_SOURCES + 1]; // filled by launch_peer_pull
    int n;
};
void launch_peer_pull(brgpu_ctx *ctx, PullSegments &segs, double bytes);
void launch_get_batch(brgpu_ctx *ctx, const uint8_t *d_bits, int k, const uint64_t *d_kmers, uint64_t n,
                      uint8_t *d_out);
void launch_insert_batch(brgpu_ctx *ctx, uint8_t *d_bits, int k, const uint64_t *d_kmers, uint64_t n);
void launch_merge_slice(brgpu_ctx *ctx, uint8_t *d_counts, void *const *peers, int n_peers, uint64_t begin,
                        uint64_t end);

void launch_probe_gather(brgpu_ctx *ctx, const uint64_t *d_tab, uint64_t n_words_pow2, uint64_t n_per_thread,
                         uint64_t *d_sink, uint64_t *n_gathers);

// occupancy summary of a bitfield: one bit per 2^shift bitfield bits
void launch_build_summary(brgpu_ctx *ctx, const uint8_t *d_bits, uint64_t n_bytes, int shift, uint32_t *d_summary);
// rank directory over a shift-6 summary: d_pop[g] = popc(summary[g]), d_rank = exclusive scan (n + 1)
void launch_summary_rank(brgpu_ctx *ctx, const uint32_t *d_summary, uint64_t n_words, uint32_t *d_pop, uint64_t *d_rank,
                         uint64_t *d_scan_tmp);
// dir[g] = {summary[g], rank[g]}; blocks[rank[g] + i] = i-th occupied 64-bit block of group g
void launch_compact_blocks(brgpu_ctx *ctx, const uint32_t *d_summary, const uint64_t *d_rank, const uint8_t *d_bits,
                           uint64_t n_words, uint64_t n_occupied, void *d_dir, uint64_t *d_blocks);

void launch_dir_only(brgpu_ctx *ctx, const uint32_t *d_summary, const uint64_t *d_rank, uint64_t n_words, void *d_dir);
void launch_fine_summary(brgpu_ctx *ctx, const uint8_t *d_bits, uint64_t n_blocks64, uint32_t *d_fine);
void launch_block_bytes(brgpu_ctx *ctx, const uint64_t *d_blocks, uint64_t n_occupied, uint8_t *d_pos8);
void launch_expand_blocks(brgpu_ctx *ctx, const void *d_dir, const uint64_t *d_blocks, uint64_t n_words, uint8_t *d_bits);

// device view of a set for the correction kernels
struct SetView {
    const uint8_t *bits;
    const uint32_t *summary;
    int shift;
    int k;
    const void *dir = nullptr;        // uint2 *
    const uint64_t *blocks = nullptr;
    const uint8_t *pos8 = nullptr;
    const uint64_t *hash = nullptr;   // set::Hash table (then everything above is unused)
    uint64_t hash_mask = 0;
};

// ---- set::Hash (hash_kernels.cu) ----
void launch_hash_insert_reads(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                              uint64_t *d_table, uint64_t n_slots, unsigned long long *d_distinct, double n_kmers);
void launch_hash_insert_keys(brgpu_ctx *ctx, const uint64_t *d_keys, uint64_t n, int k, bool canonicalise,
                             uint64_t *d_table, uint64_t n_slots, unsigned long long *d_distinct);
void launch_get_batch_view(brgpu_ctx *ctx, const SetView &set, const uint64_t *d_kmers, uint64_t n, uint8_t *d_out);

// ---- part 2 kernels (correct_kernels.cu) ----
struct CorrectParams {
    int k;
    int method;
    int confirm;
    int max_search;
    int reversed = 0; // the pass runs over the byte-reversed reads (src/lib.rs:48-55): reporting only
};
// solidity bit of the k-mer ending at every slot position of every read (0 where undefined);
// d_changed != nullptr: only for the reads it marks, the others keep the words already in d_bitmap
void launch_solid_bitmap(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len,
                         const SetView &set, uint32_t *d_bitmap, const uint8_t *d_changed, double n_bases_hint,
                         bool reversed = false);
// per-pass work areas of the segmented scan (sized once per correction call)
struct ScanWork {
    uint32_t *d_n_seg = nullptr;     // n reads
    uint64_t *d_seg_first = nullptr; // n + 1
    uint64_t *d_scan_tmp = nullptr;  // n / 4096 + 4
    uint8_t *d_seg_out = nullptr;    // scan_seg_out_bytes(L)
    void *d_seg_recs = nullptr;      // scan_seg_rec_bytes(L)
    uint8_t *d_changed = nullptr;    // n reads: 1 iff the pass edited the read (the next pass's bitmap skips the others)
};
uint64_t scan_max_segments(const Layout &L);
size_t scan_seg_out_bytes(const Layout &L);
size_t scan_seg_rec_bytes(const Layout &L);
// Corrector::correct over all reads: speculative per-segment scan + per-read merge; writes
// min(len,cap) bytes per read, the true output length to d_len_out and sets d_flags[1] on overflow.
void launch_scan(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_in, const uint32_t *d_len_in, uint8_t *d_out,
                 uint32_t *d_len_out, const uint32_t *d_bitmap, const SetView &set, const CorrectParams &p,
                 uint8_t *d_scratch, size_t scratch_per_warp, int n_warps_total, const ScanWork &w,
                 double n_bases_hint);
int scan_grid_warps(brgpu_ctx *ctx);
size_t scan_scratch_per_warp(const CorrectParams &p);

} // namespace brgpu
