// kmer.cuh — 2-bit k-mer arithmetic in registers (device side).
//
// Semantics follow cocktail::kmer as br uses it (call sites
// src/correct/mod.rs:61,71,110-112 and pcon's Solid::get behind src/set/pcon.rs:189):
//   nuc2bit(b) = (b >> 1) & 3            A=0 C=1 T=2 G=3 (any byte is a nucleotide)
//   canonical  = whichever of {kmer, revcomp} has even popcount (k odd)
//   table index = canonical >> 1
//   bitfield is LSB-first: bit i lives in byte i>>3 at position i&7
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace brgpu {

constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ __forceinline__ uint64_t kmask(int k) { return (1ULL << (2 * k)) - 1ULL; }

__host__ __device__ __forceinline__ uint32_t nuc2bit(uint8_t b) { return (b >> 1) & 3u; }

// 0->A 1->C 2->T 3->G, packed in one constant: 'A'|'C'<<8|'T'<<16|'G'<<24
__host__ __device__ __forceinline__ uint8_t bit2nuc(uint32_t c) {
    return (uint8_t)((0x47544341u >> (8 * (c & 3u))) & 0xffu);
}

// add_nuc_to_end (src/correct/mod.rs:110-112)
__host__ __device__ __forceinline__ uint64_t push(uint64_t kmer, uint32_t nuc, uint64_t mask) {
    return ((kmer << 2) & mask) ^ (uint64_t)nuc;
}

// replace the last base: add_nuc_to_end(kmer >> 2, a, k)
__host__ __device__ __forceinline__ uint64_t replace_last(uint64_t kmer, uint32_t a, uint64_t mask) {
    return (((kmer >> 2) << 2) & mask) ^ (uint64_t)a;
}

__device__ __forceinline__ uint64_t revcomp(uint64_t kmer, int k) {
    if (k == 17) {
        // 34 bits: the low word of the result is the reversed-and-swapped image of bits 2..33, the
        // two top bits are the first base complemented — 32-bit operations only
        const uint32_t x = __brev((uint32_t)(kmer >> 2));
        const uint32_t lo = (((x >> 1) & 0x55555555u) | ((x << 1) & 0xAAAAAAAAu)) ^ 0xAAAAAAAAu;
        const uint32_t hi = ((uint32_t)kmer & 3u) ^ 2u;
        return ((uint64_t)hi << 32) | lo;
    }
    // reverse all 64 bits, swap the two bits inside every group back, complement (xor 10 per
    // group), then drop the 64-2k low garbage bits.
    uint64_t r = __brevll(kmer);
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    r ^= 0xAAAAAAAAAAAAAAAAULL;
    return r >> (64 - 2 * k);
}

__device__ __forceinline__ uint64_t canonical_index(uint64_t kmer, int k) {
    uint64_t c = (__popcll(kmer) & 1) ? revcomp(kmer, k) : kmer;
    return c >> 1;
}

// cocktail::kmer::canonical for any k <= 31 (set::Hash stores the canonical k-mer itself, src/set/hash.rs:179)
__device__ __forceinline__ uint64_t canonical_kmer(uint64_t kmer, int k) {
    return (__popcll(kmer) & 1) ? revcomp(kmer, k) : kmer;
}

// set::Hash on the device (hash_kernels.cu): open addressing, linear probing, empty = all ones
constexpr uint64_t HASH_EMPTY = ~0ULL;
constexpr uint32_t POS8_MULTI = 0xFFu; // SolidView::pos8: the block holds more than one solid k-mer
constexpr uint32_t POS8_NONE = 0xFEu;  // in-register marker of a lookup that needs no second load
__host__ __device__ __forceinline__ uint64_t hash_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// KmerSet::get — one random byte (32 B sector) of the HBM-resident bitfield.
__device__ __forceinline__ bool solid(const uint8_t *__restrict__ bits, uint64_t kmer, int k) {
    uint64_t idx = canonical_index(kmer, k);
    return (__ldg(bits + (idx >> 3)) >> (idx & 7)) & 1;
}

// The solid set as the kernels see it: the dense bitfield in HBM plus (for k >= 15) a coarse
// occupancy summary small enough to stay in L2 — summary bit j is set iff any of the bitfield
// bits [j << shift, (j + 1) << shift) is.  A random byte of the bitfield costs a DRAM row
// activation (measured ceiling 43 G/s, profiles/microbench_random_access_r1.txt); a random word
// of the summary is an L2 hit (285 G/s).  Most k-mers a corrector asks about are weak, and for a
// sparse set almost all of them fall into empty blocks, so they never leave L2.
struct SolidView {
    const uint8_t *bits;
    const uint32_t *summary; // nullptr: no summary (small k: the bitfield itself is cache resident)
    int shift;               // log2(bitfield bits per summary bit): 6 or more; 4 for the fine summary of a dense set
    int k;
    // Rank-compacted copy of a sparse bitfield (nullptr: not built).  dir[g] = {occupancy of the
    // 32 64-bit blocks 32g .. 32g+31, number of occupied blocks before block 32g}; blocks[] holds
    // the occupied 64-bit blocks in index order.  A lookup is one 8 B load of dir (L2) and, only
    // when the block is occupied, one 8 B load of blocks — for a 4.6 Mb genome at k = 17 both
    // arrays together are 69 MB and stay in L2, where the bitfield costs a DRAM access per hit.
    const uint2 *dir;
    const uint64_t *blocks;
    // One byte per occupied block (nullptr: not built): the position of the block's only set bit, or POS8_MULTI
    // when it has several — then (and only then) the 64-bit block itself is read.  A sparse set has one solid
    // k-mer in almost every occupied block (98 % at 4.6 M k-mers, 87 % at 37 M), so the second load of a lookup
    // lands in an array eight times smaller: 32 MiB of directory + 1 B per block stay in L2 where 8 B per
    // block (294 MB for the eight-genome set of the weak-scaled N = 8 run) did not.
    const uint8_t *pos8;
    // set::Hash (nullptr: a dense set): table of canonical k-mers, hash_mask = slots - 1
    const uint64_t *hash;
    uint64_t hash_mask;
};

// host side: the kernels' view of a brgpu::SetView (internal.h)
template <class SV> inline SolidView solid_view(const SV &s) {
    return SolidView{s.bits, s.summary, s.shift, s.k, (const uint2 *)s.dir, s.blocks, s.pos8, s.hash, s.hash_mask};
}

// Loads of the lookup structures (directory, block bytes, blocks).  BRGPU_SET_LOAD_MODE: 0 = ld.global.nc (__ldg),
// 1 = ld.global.cg (L2 only), 2 = ld.global.nc with an L2 evict_last cache hint — the set is what every lookup of a
// correction pass returns to, the reads / bitmap / scratch stream past it once; the policy is a compile-time
// constant that ptxas folds into the load's descriptor.
#ifndef BRGPU_SET_LOAD_MODE
#define BRGPU_SET_LOAD_MODE 0
#endif
#if BRGPU_SET_LOAD_MODE == 2
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint2 set_ld(const uint2 *p) {
    uint2 r;
    asm("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(l2_keep_policy()));
    return r;
}
__device__ __forceinline__ uint32_t set_ld(const uint8_t *p) {
    uint32_t r;
    asm("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(l2_keep_policy()));
    return r;
}
__device__ __forceinline__ uint64_t set_ld(const uint64_t *p) {
    uint64_t r;
    asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(l2_keep_policy()));
    return r;
}
#elif BRGPU_SET_LOAD_MODE == 1
__device__ __forceinline__ uint2 set_ld(const uint2 *p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t set_ld(const uint8_t *p) { return __ldcg(p); }
__device__ __forceinline__ uint64_t set_ld(const uint64_t *p) { return (uint64_t)__ldcg(reinterpret_cast<const unsigned long long *>(p)); }
#else
__device__ __forceinline__ uint2 set_ld(const uint2 *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t set_ld(const uint8_t *p) { return __ldg(p); }
__device__ __forceinline__ uint64_t set_ld(const uint64_t *p) { return __ldg(p); }
#endif

__device__ __forceinline__ bool hash_contains(const uint64_t *__restrict__ table, uint64_t slot_mask, uint64_t key) {
    uint64_t h = hash_mix64(key) & slot_mask;
    for (;;) {
        const uint64_t e = __ldg(table + h);
        if (e == key) return true;
        if (e == HASH_EMPTY) return false;
        h = (h + 1) & slot_mask;
    }
}

__device__ __forceinline__ bool solid(const SolidView &v, uint64_t kmer) {
    if (v.hash) return hash_contains(v.hash, v.hash_mask, canonical_kmer(kmer, v.k)); // src/set/hash.rs:179-181
    uint64_t idx = canonical_index(kmer, v.k);
    if (v.dir) {
        const uint64_t j = idx >> 6;
        const uint2 e = set_ld(v.dir + (j >> 5));
        const uint32_t b = (uint32_t)j & 31u;
        if (!((e.x >> b) & 1u)) return false;
        const uint32_t r = e.y + __popc(e.x & ((1u << b) - 1u));
        if (v.pos8) {
            const uint32_t p = set_ld(v.pos8 + r);
            if (p != POS8_MULTI) return p == (uint32_t)(idx & 63);
        }
        return (set_ld(v.blocks + r) >> (idx & 63)) & 1ULL;
    }
    if (v.summary) {
        uint64_t j = idx >> v.shift;
        if (!((__ldg(v.summary + (j >> 5)) >> (j & 31)) & 1u)) return false;
    }
    return (__ldg(v.bits + (idx >> 3)) >> (idx & 7)) & 1;
}

// Pack the 2-bit codes of 16 ASCII bases held in a uint4 (memory order) into 32 bits, first
// base in the most significant pair.
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    // w holds 4 bytes b0 (lowest address, bits 0-7) .. b3.  codes = (b>>1)&3
    uint32_t x = (w >> 1) & 0x03030303u;
    // want c0<<6 | c1<<4 | c2<<2 | c3
    return ((x & 0x3u) << 6) | (((x >> 8) & 0x3u) << 4) | (((x >> 16) & 0x3u) << 2) | ((x >> 24) & 0x3u);
}

__device__ __forceinline__ uint32_t pack16(uint4 v) {
    return (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w);
}

// ------------------------------------------------------------------------------------------
// The 32-position window every streaming kernel uses: slot word w of a read holds positions
// p0..p0+31; `cur` packs their 2-bit codes (first base in the top pair), `prev` the 32 before.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_window(const uint8_t *__restrict__ seq, uint64_t slot_byte, uint32_t p0,
                                            uint64_t &prev, uint64_t &cur) {
    const uint4 *q = reinterpret_cast<const uint4 *>(seq + slot_byte);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    cur = ((uint64_t)pack16(a) << 32) | (uint64_t)pack16(b);
    prev = 0;
    if (p0 >= 32) {
        uint4 c = __ldg(q - 2), d = __ldg(q - 1);
        prev = ((uint64_t)pack16(c) << 32) | (uint64_t)pack16(d);
    }
}

// k-mer ending at window position t (0..31)
__device__ __forceinline__ uint64_t window_kmer(uint64_t prev, uint64_t cur, int t, uint64_t mask) {
    int s = 2 * (31 - t);
    uint64_t v = cur >> s;
    if (s) v |= prev << (64 - s);
    return v & mask;
}

// 64-bit OR-reduction across the warp
__device__ __forceinline__ uint64_t warp_or64(uint64_t v) {
    uint32_t lo = __reduce_or_sync(FULL, (uint32_t)v);
    uint32_t hi = __reduce_or_sync(FULL, (uint32_t)(v >> 32));
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

} // namespace brgpu
