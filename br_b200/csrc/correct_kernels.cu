// correct_kernels.cu — part 2 of br's hot path on sm_100a: the per-read correction pass.
//
// One pass of one method over all reads is two kernels:
//
//   solid_bitmap_kernel   ("phase A") — embarrassingly parallel and HBM-sector bound: one thread
//       per 32 read positions rolls the k-mers in registers and gathers their solidity bits
//       (32 independent random 32 B sector reads in flight per thread) into a per-read bitmap.
//   scan_kernel           ("phase B") — Corrector::correct (src/correct/mod.rs:53-107) is
//       sequential inside a read (kmer / previous / i carry across events), so one warp owns one
//       read and walks it in order; reads are handed out longest-first from an atomic work
//       queue.  While the rolling k-mer consists of input bases only, the scan just looks for
//       the next solid->weak transition in the phase-A bitmap (1024 positions per step); at a
//       transition the lanes enumerate the candidate lookups of the method in parallel
//       (alternatives, scenario scores, successor sets) and the winner is picked with
//       ballots.  After a successful correction the next k-1 k-mers contain corrected bases,
//       so they are looked up directly, 32 positions per round.
//
// Reference functions restated here (Rust; there is no reference kernel):
//   Corrector::correct, alt_nucs, next_nucs, error_len      src/correct/mod.rs:53-152
//   Exist::correct_error, Scenario::{get_score,one_more}    src/correct/exist/mod.rs:21-149
//   ScenarioOne / ScenarioTwo                                src/correct/exist/one.rs:57-71, two.rs:89-325
//   Graph::correct_error                                     src/correct/graph.rs:44-85
//   Greedy::correct_error + bio 1.6.0 global alignment       src/correct/greedy.rs:56-173
//   GapSize::correct_error, ins_sub_correction               src/correct/gap_size.rs:44-108
#include <cstdlib>

#include "internal.h"
#include "kmer.cuh"

#include "scan_common.cuh"

namespace brgpu {


// ------------------------------------------------------------------------------------------
// phase A
// ------------------------------------------------------------------------------------------
// KT: compile-time k (0 = take it from the set).  k = 17 is the size every BASELINE config uses;
// with a constant k the 64-bit shifts and masks of the k-mer arithmetic become immediates.
// ARM: 0 = the form of the set is a run-time matter, 1 = rank-compacted, 2 = summary + bitfield (the other arms of
// the lookup are compiled out, as in the scan kernels)
template <int KT, int ARM>
__global__ void __launch_bounds__(256)
    solid_bitmap_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                        const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                        uint64_t n_words, SolidView set, uint32_t *__restrict__ bitmap,
                        const uint8_t *__restrict__ changed, unsigned long long *get_counter) {
    const int k = KT ? KT : set.k;
    if (ARM == 1) {
        set.hash = nullptr;
        set.summary = nullptr;
        __builtin_assume(set.dir != nullptr);
    } else if (ARM == 2) {
        set.hash = nullptr;
        set.dir = nullptr;
    }
    const uint8_t *__restrict__ bits = set.bits;
    const uint64_t mask = kmask(k);
    uint32_t n_looked = 0; // k-mers this thread looked up (reported by profiling runs only)
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = __ldg(word2read + w);
        // the previous method left this read as it was: its bitmap words are still valid
        if (changed && !__ldg(changed + r)) continue;
        uint64_t sb = w << 5;
        uint32_t p0 = (uint32_t)(sb - __ldg(slot_off + r));
        uint32_t L = __ldg(len + r);
        uint32_t out = 0;
        if (p0 < L && L >= (uint32_t)k) {
            uint64_t prev, cur;
            load_window(seq, sb, p0, prev, cur);
            int t_lo = p0 >= (uint32_t)(k - 1) ? 0 : (k - 1 - (int)p0);
            int t_hi = (L - p0) < 32u ? (int)(L - p0) : 32;
            n_looked += (uint32_t)(t_hi > t_lo ? t_hi - t_lo : 0);
            // 4 rounds of 8 independent gathers: all loads of a round are issued before the first
            // one is consumed.  Round part 1 asks the L2-resident summary; only k-mers whose block
            // is occupied go on to the bitfield byte in HBM (part 2).
#pragma unroll
            for (int g = 0; g < 32; g += 8) {
                uint64_t idx[8];
                uint32_t byte[8];
                bool go[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    idx[j] = canonical_index(window_kmer(prev, cur, g + j, mask), k);
                    go[j] = g + j >= t_lo && g + j < t_hi;
                }
                if (set.hash) { // set::Hash: probe the table (k > 19: no dense form exists)
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (go[j] && hash_contains(set.hash, set.hash_mask, canonical_kmer(window_kmer(prev, cur, g + j, mask), k)))
                            out |= 1u << (g + j);
                    continue;
                }
                if (set.dir) {
                    // rank-compacted set: directory entry (8 B, L2), then the occupied block (8 B)
                    uint2 e[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        e[j] = make_uint2(0u, 0u);
                        if (go[j]) e[j] = set_ld(set.dir + (idx[j] >> 11));
                    }
                    uint32_t r[8], pv[8]; // then the block's byte (1 B) and, for a block with several k-mers, the block (8 B)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t b = (uint32_t)(idx[j] >> 6) & 31u;
                        r[j] = e[j].y + __popc(e[j].x & ((1u << b) - 1u));
                        pv[j] = POS8_NONE;
                        if ((e[j].x >> b) & 1u) pv[j] = set.pos8 ? set_ld(set.pos8 + r[j]) : POS8_MULTI;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        uint32_t hit = pv[j] == (uint32_t)(idx[j] & 63);
                        if (pv[j] == POS8_MULTI) hit = (uint32_t)((set_ld(set.blocks + r[j]) >> (idx[j] & 63)) & 1ULL);
                        out |= hit << (g + j);
                    }
                    continue;
                }
                if (set.summary) {
                    uint32_t sw[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        sw[j] = 0;
                        if (go[j]) sw[j] = __ldg(set.summary + (idx[j] >> (set.shift + 5)));
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) go[j] = (sw[j] >> ((idx[j] >> set.shift) & 31)) & 1u;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    byte[j] = 0;
                    if (go[j]) byte[j] = __ldg(bits + (idx[j] >> 3));
                }
#pragma unroll
                for (int j = 0; j < 8; j++) out |= ((byte[j] >> (idx[j] & 7)) & 1u) << (g + j);
            }
        }
        bitmap[w] = out;
    }
    if (get_counter) {
        n_looked = __reduce_add_sync(FULL, n_looked);
        if ((threadIdx.x & 31) == 0 && n_looked) atomicAdd(get_counter, (unsigned long long)n_looked);
    }
}

void launch_solid_bitmap(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len,
                         const SetView &set, uint32_t *d_bitmap, const uint8_t *d_changed, double n_bases_hint,
                         bool reversed) {
    uint64_t n_words = L.total_slots >> 5;
    if (!n_words) return;
    // algorithmic bytes per position: 32 B sector + 1 B ASCII in + 1/8 B bit out
    ProfScope ps(ctx, reversed ? "solid_bitmap_rev" : "solid_bitmap", n_bases_hint * 33.125);
    uint64_t need = (n_words + 255) / 256;
    uint64_t capb = (uint64_t)ctx->sm_count * 8;
    const SolidView sv = solid_view(set);
    const unsigned grid = (unsigned)(need < capb ? need : capb);
    unsigned long long *gc = ctx->profiling ? prof_counter_slot(ctx) : nullptr;
    auto go = [&](auto kernel) {
        kernel<<<grid, 256, 0, ctx->stream>>>(d_seq, d_len, L.d_slot_off, L.d_word2read, n_words, sv, d_bitmap, d_changed, gc);
    };
    if (set.k != 17) go(solid_bitmap_kernel<0, 0>);
    else if (sv.hash) go(solid_bitmap_kernel<17, 0>);
    else if (sv.dir) go(solid_bitmap_kernel<17, 1>);
    else go(solid_bitmap_kernel<17, 2>);
}

// per-warp scratch of the scan kernels (Greedy's alignment; layout in scan_device.cuh: greedy_scratch)
static inline size_t greedy_dim_host(int k, int max_search) { return (size_t)(k - 1 + max_search + 2); }
size_t scan_scratch_per_warp(const CorrectParams &p) {
    if (p.method != BRGPU_GREEDY) return 0;
    size_t d = greedy_dim_host(p.k, p.max_search);
    size_t bytes = ((size_t)p.max_search + 2) * 8 + d * d * 2 + 4 + 3 * d * 4 + 2 * d + d + d;
    return (bytes + 127) & ~(size_t)127;
}

__global__ void seg_count_kernel(const uint32_t *__restrict__ len, uint32_t n_reads, uint32_t k,
                                 uint32_t *__restrict__ n_seg) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x)
        n_seg[r] = seg_count(len[r], k);
}

// one warp per spliced piece: scratch region -> its place in the output slot
__global__ void __launch_bounds__(256)
    scan_splice_kernel(const SegCopy *__restrict__ copies, uint64_t n_seg, const uint8_t *__restrict__ seg_out,
                       uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_seg; g += n_warps) {
        const SegCopy cp = copies[g];
        if (cp.n) warp_copy(out + cp.dst, seg_out + g * SEG_CAP + cp.skip, cp.n, lane);
    }
}

void launch_scan_splice(brgpu_ctx *ctx, const SegCopy *d_copies, uint64_t n_seg, const uint8_t *d_seg_out, uint8_t *d_out) {
    uint64_t blocks = (n_seg + 7) / 8, cap = (uint64_t)ctx->sm_count * 8;
    ctx->launches += 1;
    scan_splice_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(d_copies, n_seg, d_seg_out, d_out);
}

// upper bound on resident warps over all methods (sizes the per-warp scratch)
int scan_grid_warps(brgpu_ctx *ctx) { return ctx->sm_count * 16 * SCAN_WARPS_PER_BLOCK; }

uint64_t scan_max_segments(const Layout &L) { return L.total_slots / SEG + L.n + 1; }
size_t scan_seg_out_bytes(const Layout &L) { return (size_t)scan_max_segments(L) * SEG_CAP + 64; } // + slack: warp_copy reads whole source words
size_t scan_seg_rec_bytes(const Layout &L) { return (size_t)scan_max_segments(L) * (sizeof(SegRec) + sizeof(SegCopy)); }

void launch_scan(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_in, const uint32_t *d_len_in, uint8_t *d_out,
                 uint32_t *d_len_out, const uint32_t *d_bitmap, const SetView &set, const CorrectParams &p,
                 uint8_t *d_scratch, size_t scratch_per_warp, int n_warps_total, const ScanWork &w,
                 double n_bases_hint) {
    if (!L.n) return;
    // work-queue cursors: flags[0] reads (merge), flags[4..5] segments (spec, 64 bit)
    cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(ctx->d_flags + 4, 0, 2 * sizeof(uint32_t), ctx->stream);
    // segments per read -> first segment of every read
    {
        ProfScope ps(ctx, "seg_count", (double)L.n * 8.0);
        unsigned blocks = (unsigned)((L.n + 255) / 256);
        if (blocks > (unsigned)ctx->sm_count * 8) blocks = (unsigned)ctx->sm_count * 8;
        seg_count_kernel<<<blocks, 256, 0, ctx->stream>>>(d_len_in, (uint32_t)L.n, (uint32_t)p.k, w.d_n_seg);
    }
    launch_exclusive_scan_u32(ctx, w.d_n_seg, L.n, w.d_seg_first, w.d_scan_tmp);
    const SolidView sv = solid_view(set);
    // kernel names carry the orientation: the reversed pass sees almost no events (the forward pass
    // repaired them), so averaging the two hides what a launch costs
    static const char *spec_names[2][5] = {{"scan_one", "scan_two", "scan_graph", "scan_greedy", "scan_gap_size"},
                                           {"scan_one_rev", "scan_two_rev", "scan_graph_rev", "scan_greedy_rev", "scan_gap_size_rev"}};
    static const char *merge_names[2][5] = {{"merge_one", "merge_two", "merge_graph", "merge_greedy", "merge_gap_size"},
                                            {"merge_one_rev", "merge_two_rev", "merge_graph_rev", "merge_greedy_rev", "merge_gap_size_rev"}};
    const int m = p.method >= BRGPU_ONE && p.method <= BRGPU_GAP_SIZE ? p.method : BRGPU_GAP_SIZE;
    const int rev = p.reversed ? 1 : 0;
    const ScanArgs a{ctx, &L, d_in, d_len_in, d_out, d_len_out, d_bitmap, sv, p, d_scratch, scratch_per_warp,
                     n_warps_total, &w, n_bases_hint, spec_names[rev][m], merge_names[rev][m]};
    // profiling runs use the kernels that count their KmerSet::get calls; the product path does not pay for it
    using Fn = void (*)(const ScanArgs &);
    static const Fn table[2][5] = {
        {fast::launch_scan_m0, fast::launch_scan_m1, fast::launch_scan_m2, fast::launch_scan_m3, fast::launch_scan_m4},
        {cnt::launch_scan_m0, cnt::launch_scan_m1, cnt::launch_scan_m2, cnt::launch_scan_m3, cnt::launch_scan_m4}};
    table[ctx->profiling ? 1 : 0][m](a);
}

} // namespace brgpu
