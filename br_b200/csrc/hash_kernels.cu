// hash_kernels.cu — set::Hash on the device (src/set/hash.rs:14-186): the solid set for k-mers too
// large for a dense bitfield (17 < k <= 31; any k the caller asks for, br's `large-kmer` sub-command,
// src/main.rs:147-163).  The reference keeps an FxHashSet<u64> of canonical k-mers and answers
// KmerSet::get with `contains(canonical(kmer, k))`; only membership matters, so the device form is an
// open-addressing table of 64-bit keys with linear probing:
//   slot(key) = mix64(key) & (slots - 1), empty = all ones (no k-mer with k <= 31 has its top bits set),
//   insertion by atomicCAS (duplicates resolve to the slot that already holds the key), load <= 1/2.
// Lookups go through the same SolidView seam as the dense forms (kmer.cuh), so every correction
// kernel works on either kind of set.  Construction from reads is Hash::from_fasta: presence only.
#include "internal.h"
#include "kmer.cuh"

namespace brgpu {

// every canonical k-mer of every read with len >= k (src/set/hash.rs:52-57) -> table
__global__ void __launch_bounds__(256)
    hash_insert_reads_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                             const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                             uint64_t n_words, int k, uint64_t *__restrict__ table, uint64_t slot_mask,
                             unsigned long long *__restrict__ n_distinct) {
    const uint64_t mask = kmask(k);
    uint32_t fresh = 0;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = __ldg(word2read + w);
        const uint64_t sb = w << 5;
        const uint32_t p0 = (uint32_t)(sb - __ldg(slot_off + r));
        const uint32_t L = __ldg(len + r);
        if (p0 >= L || L < (uint32_t)k) continue;
        uint64_t prev, cur;
        load_window(seq, sb, p0, prev, cur);
        const int t_lo = p0 >= (uint32_t)(k - 1) ? 0 : (k - 1 - (int)p0);
        const int t_hi = (L - p0) < 32u ? (int)(L - p0) : 32;
        for (int t = t_lo; t < t_hi; t++) {
            const uint64_t key = canonical_kmer(window_kmer(prev, cur, t, mask), k);
            uint64_t h = hash_mix64(key) & slot_mask;
            for (;;) {
                const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(table + h), HASH_EMPTY, key);
                if (old == HASH_EMPTY) fresh++;
                if (old == HASH_EMPTY || old == key) break;
                h = (h + 1) & slot_mask;
            }
        }
    }
    fresh = __reduce_add_sync(FULL, fresh);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(n_distinct, (unsigned long long)fresh);
}

// keys[] are k-mers (canonicalise = 1: Solid::set / Hash insertion of forward k-mers) or table entries of
// an older, smaller table (canonicalise = 0: rehash; empty entries are skipped)
__global__ void hash_insert_keys_kernel(const uint64_t *__restrict__ keys, uint64_t n, int k, int canonicalise,
                                        uint64_t *__restrict__ table, uint64_t slot_mask,
                                        unsigned long long *__restrict__ n_distinct) {
    const uint64_t mask = kmask(k);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        if (!canonicalise && key == HASH_EMPTY) continue;
        if (canonicalise) key = canonical_kmer(key & mask, k);
        uint64_t h = hash_mix64(key) & slot_mask;
        for (;;) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(table + h), HASH_EMPTY, key);
            if (old == HASH_EMPTY) atomicAdd(n_distinct, 1ULL);
            if (old == HASH_EMPTY || old == key) break;
            h = (h + 1) & slot_mask;
        }
    }
}

__global__ void get_batch_view_kernel(SolidView set, const uint64_t *__restrict__ kmers, uint64_t n, uint8_t *__restrict__ out) {
    const uint64_t mask = kmask(set.k);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = solid(set, __ldg(kmers + i) & mask) ? 1 : 0;
}

static inline unsigned hash_grid(brgpu_ctx *ctx, uint64_t items) {
    const uint64_t need = (items + 255) / 256, cap = (uint64_t)ctx->sm_count * 8;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}

void launch_hash_insert_reads(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                              uint64_t *d_table, uint64_t n_slots, unsigned long long *d_distinct, double n_kmers) {
    const uint64_t n_words = L.total_slots >> 5;
    if (!n_words) return;
    ProfScope ps(ctx, "hash_insert", n_kmers * 64.25); // one random sector read + write-back per k-mer
    hash_insert_reads_kernel<<<hash_grid(ctx, n_words), 256, 0, ctx->stream>>>(d_seq, d_len, L.d_slot_off, L.d_word2read,
                                                                              n_words, k, d_table, n_slots - 1, d_distinct);
}

void launch_hash_insert_keys(brgpu_ctx *ctx, const uint64_t *d_keys, uint64_t n, int k, bool canonicalise,
                             uint64_t *d_table, uint64_t n_slots, unsigned long long *d_distinct) {
    if (!n) return;
    ProfScope ps(ctx, canonicalise ? "hash_insert_batch" : "hash_rehash", (double)n * 72.0);
    hash_insert_keys_kernel<<<hash_grid(ctx, n), 256, 0, ctx->stream>>>(d_keys, n, k, canonicalise ? 1 : 0, d_table,
                                                                       n_slots - 1, d_distinct);
}

void launch_get_batch_view(brgpu_ctx *ctx, const SetView &set, const uint64_t *d_kmers, uint64_t n, uint8_t *d_out) {
    if (!n) return;
    ProfScope ps(ctx, "get_batch", (double)n * 41.0);
    const SolidView sv = solid_view(set);
    get_batch_view_kernel<<<hash_grid(ctx, n), 256, 0, ctx->stream>>>(sv, d_kmers, n, d_out);
}

} // namespace brgpu
