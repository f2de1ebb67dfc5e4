// set_kernels.cu — part 1 of br's hot path on sm_100a: k-mer counting into an HBM-resident
// table of saturating u8 counters, the 256-bin spectrum, the solidity threshold into the dense
// canonical bitfield, batched KmerSet::get — plus the slot-layout plumbing (scatter/gather/
// reverse/scan) the correction pass shares.
//
// None of this is GEMM-shaped: it is integer hashing and random 32-byte-sector access, so the
// roofline is HBM and the levers are coalesced vector loads for the streams,
// many independent random accesses in flight per thread, and grids sized from the SM count.
//
// Reference semantics replaced (Rust, not CUDA — there is no reference kernel):
//   pcon Counter::<u8>::count_fasta      call site src/main.rs:73-74
//   pcon Spectrum::from_count            call site src/main.rs:93
//   pcon Solid::from_count               call site src/main.rs:112-114
//   pcon Solid::get via set::Pcon::get   src/set/pcon.rs:188-191
#include "internal.h"
#include "kmer.cuh"

namespace brgpu {

static inline int grid_for(brgpu_ctx *ctx, uint64_t work_items, int block, int max_blocks_per_sm) {
    uint64_t need = (work_items + (uint64_t)block - 1) / (uint64_t)block;
    uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------------------------------
// layout plumbing
// ------------------------------------------------------------------------------------------
__global__ void fill_word2read_kernel(const uint64_t *__restrict__ slot_off, uint64_t n_reads, uint64_t n_words,
                                      uint32_t *__restrict__ word2read) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t pos = w << 5;
        uint64_t lo = 0, hi = n_reads; // find r with slot_off[r] <= pos < slot_off[r+1]
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (__ldg(slot_off + mid) <= pos)
                lo = mid;
            else
                hi = mid;
        }
        word2read[w] = (uint32_t)lo;
    }
}

void launch_fill_word2read(brgpu_ctx *ctx, const Layout &L) {
    uint64_t n_words = L.total_slots >> 5;
    if (!n_words) return;
    ProfScope ps(ctx, "fill_word2read", (double)n_words * 4.0);
    fill_word2read_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(L.d_slot_off, L.n, n_words,
                                                                                   L.d_word2read);
}

// tight (concatenated) host layout -> slot layout; slack bytes are zeroed
__global__ void scatter_to_slots_kernel(const uint8_t *__restrict__ tight, const uint64_t *__restrict__ tight_off,
                                        const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                                        uint64_t n_quads, uint32_t *__restrict__ slots) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_quads;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = g << 2;
        uint32_t r = __ldg(word2read + (s >> 5));
        uint64_t p = s - __ldg(slot_off + r);
        uint64_t t0 = __ldg(tight_off + r);
        uint64_t len = __ldg(tight_off + r + 1) - t0;
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (p + j < len) v |= (uint32_t)__ldg(tight + t0 + p + j) << (8 * j);
        slots[g] = v;
    }
}

void launch_scatter_to_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_tight, const uint64_t *d_tight_off,
                             uint8_t *d_slots) {
    uint64_t n_quads = L.total_slots >> 2;
    if (!n_quads) return;
    ProfScope ps(ctx, "scatter_to_slots", (double)L.total_slots * 2.0);
    scatter_to_slots_kernel<<<grid_for(ctx, n_quads, 256, 8), 256, 0, ctx->stream>>>(
        d_tight, d_tight_off, L.d_slot_off, L.d_word2read, n_quads, (uint32_t *)d_slots);
}

// slot layout -> tight layout (optionally byte-reversing each read)
__global__ void gather_from_slots_kernel(const uint32_t *__restrict__ slots, const uint32_t *__restrict__ len,
                                         const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                                         const uint64_t *__restrict__ tight_off, uint64_t n_quads,
                                         uint8_t *__restrict__ tight, int reverse) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_quads;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = g << 2;
        uint32_t r = __ldg(word2read + (s >> 5));
        uint64_t p = s - __ldg(slot_off + r);
        uint64_t L = __ldg(len + r);
        if (p >= L) continue;
        uint64_t t0 = __ldg(tight_off + r);
        uint32_t v = __ldg(slots + g);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (p + j < L) tight[t0 + (reverse ? (L - 1 - (p + j)) : (p + j))] = (uint8_t)(v >> (8 * j));
    }
}

void launch_gather_from_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_slots, const uint32_t *d_len,
                              const uint64_t *d_tight_off, uint8_t *d_tight, bool reverse) {
    uint64_t n_quads = L.total_slots >> 2;
    if (!n_quads) return;
    ProfScope ps(ctx, "gather_from_slots", (double)L.total_slots * 2.0);
    gather_from_slots_kernel<<<grid_for(ctx, n_quads, 256, 8), 256, 0, ctx->stream>>>(
        (const uint32_t *)d_slots, d_len, L.d_slot_off, L.d_word2read, d_tight_off, n_quads, d_tight, reverse ? 1 : 0);
}

// out[p] = in[len-1-p] inside every slot (the reversed pass of src/lib.rs:49,111 is a plain
// byte reversal, not a reverse complement)
__global__ void reverse_slots_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ len,
                                     const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                                     uint64_t n_quads, uint32_t *__restrict__ out) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_quads;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = g << 2;
        uint32_t r = __ldg(word2read + (s >> 5));
        uint64_t base = __ldg(slot_off + r);
        uint64_t p = s - base;
        uint64_t L = __ldg(len + r);
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (p + j < L) v |= (uint32_t)__ldg(in + base + (L - 1 - (p + j))) << (8 * j);
        out[g] = v;
    }
}

void launch_reverse_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_in, const uint32_t *d_len, uint8_t *d_out) {
    uint64_t n_quads = L.total_slots >> 2;
    if (!n_quads) return;
    ProfScope ps(ctx, "reverse_slots", (double)L.total_slots * 2.0);
    reverse_slots_kernel<<<grid_for(ctx, n_quads, 256, 8), 256, 0, ctx->stream>>>(d_in, d_len, L.d_slot_off,
                                                                                  L.d_word2read, n_quads,
                                                                                  (uint32_t *)d_out);
}

// ------------------------------------------------------------------------------------------
// 2-bit transport (BASELINE.json north_star, first bullet).  Across PCIe a chunk travels as
// 2 bits per base — tight layout, read r occupies base positions [off[r], off[r+1]), four bases per
// byte, first base in the two high bits — plus an exception list (position, byte) for every byte that
// is not the upper-case letter of its own 2-bit code (lower case, N, anything else).  On the device the
// reads live in the slot layout as ASCII: Corrector::correct echoes the *original* byte of every
// position it does not rewrite (src/correct/mod.rs:91,100), and corrections only ever emit A, C, T, G,
// so the exceptions of the output are a subset of the exceptions of the input.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t packed_code(const uint8_t *__restrict__ packed, uint64_t t) {
    return (__ldg(packed + (t >> 2)) >> (2u * (3u - (uint32_t)(t & 3)))) & 3u;
}

__global__ void unpack_to_slots_kernel(const uint8_t *__restrict__ packed, const uint64_t *__restrict__ tight_off,
                                       const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                                       uint64_t n_quads, uint32_t *__restrict__ slots) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_quads;
         g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = g << 2;
        const uint32_t r = __ldg(word2read + (s >> 5));
        const uint64_t p = s - __ldg(slot_off + r);
        const uint64_t t0 = __ldg(tight_off + r);
        const uint64_t len = __ldg(tight_off + r + 1) - t0;
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (p + j < len) v |= (uint32_t)bit2nuc(packed_code(packed, t0 + p + j)) << (8 * j);
        slots[g] = v;
    }
}

// read r with off[r] <= t < off[r+1] (off has n + 1 entries, t < off[n])
__device__ __forceinline__ uint64_t read_of_position(const uint64_t *__restrict__ off, uint64_t n, uint64_t t) {
    uint64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= t)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__global__ void apply_exceptions_kernel(const uint64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte,
                                        uint64_t n_exc, const uint64_t *__restrict__ tight_off,
                                        const uint64_t *__restrict__ slot_off, uint64_t n_reads, uint8_t *__restrict__ slots) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_exc; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = __ldg(exc_pos + i);
        if (t >= __ldg(tight_off + n_reads)) continue; // out of range: ignored
        const uint64_t r = read_of_position(tight_off, n_reads, t);
        slots[__ldg(slot_off + r) + (t - __ldg(tight_off + r))] = __ldg(exc_byte + i);
    }
}

// tight ASCII -> packed: one thread per 16 bases (one 16 B load, one 4 B store); exceptions are appended in no
// particular order.  (The first version packed straight out of the slot layout, one thread per output byte
// with its own read lookup: 0.64 ms per step on the E. coli configuration; gather + this pass: see DESIGN §2.)
__global__ void __launch_bounds__(256)
    pack_tight_kernel(const uint8_t *__restrict__ tight, uint64_t n_bases, uint32_t *__restrict__ packed32,
                      uint64_t *__restrict__ exc_pos, uint8_t *__restrict__ exc_byte, uint64_t exc_cap,
                      unsigned long long *__restrict__ n_exc) {
    const uint64_t n_words = (n_bases + 15) >> 4;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t0 = w << 4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t0 + 16 <= n_bases) {
            v = __ldcs(reinterpret_cast<const uint4 *>(tight + t0)); // the staging buffer is 16 B aligned and padded
        } else {
            uint32_t q[4] = {0u, 0u, 0u, 0u};
            for (uint64_t t = t0; t < n_bases; t++) q[(t - t0) >> 2] |= (uint32_t)tight[t] << (8 * ((t - t0) & 3));
            v = make_uint4(q[0], q[1], q[2], q[3]);
        }
        const uint32_t codes = pack16(v); // first base in the top pair
        // a byte is an exception unless it is the upper-case letter of its code: rebuild the letters and compare
        const uint32_t in[4] = {v.x, v.y, v.z, v.w};
        uint32_t bad = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t letters = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) letters |= (uint32_t)bit2nuc((codes >> (2 * (15 - (4 * q + j)))) & 3u) << (8 * j);
            const uint32_t diff = in[q] ^ letters;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((diff >> (8 * j)) & 0xffu) bad |= 1u << (4 * q + j);
        }
        if (t0 + 16 > n_bases) bad &= (1u << (n_bases - t0)) - 1u;
        while (bad) {
            const int j = __ffs(bad) - 1;
            bad &= bad - 1;
            const unsigned long long at = atomicAdd(n_exc, 1ULL);
            if (at < exc_cap) {
                exc_pos[at] = t0 + (uint64_t)j;
                exc_byte[at] = (uint8_t)(in[j >> 2] >> (8 * (j & 3)));
            }
        }
        // bytes of the packed stream in memory order: base 0..3 in byte 0 -> the big-endian image of `codes`
        packed32[w] = __byte_perm(codes, 0u, 0x0123);
    }
}

void launch_unpack_to_slots(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_packed, const uint64_t *d_tight_off,
                            uint8_t *d_slots, const uint64_t *d_exc_pos, const uint8_t *d_exc_byte, uint64_t n_exc) {
    const uint64_t n_quads = L.total_slots >> 2;
    if (!n_quads) return;
    ProfScope ps(ctx, "unpack_to_slots", (double)L.total_slots * 1.25);
    unpack_to_slots_kernel<<<grid_for(ctx, n_quads, 256, 8), 256, 0, ctx->stream>>>(d_packed, d_tight_off, L.d_slot_off,
                                                                                   L.d_word2read, n_quads, (uint32_t *)d_slots);
    if (n_exc) {
        ctx->launches++;
        apply_exceptions_kernel<<<grid_for(ctx, n_exc, 256, 8), 256, 0, ctx->stream>>>(d_exc_pos, d_exc_byte, n_exc, d_tight_off,
                                                                                    L.d_slot_off, L.n, d_slots);
    }
}

void launch_pack_tight(brgpu_ctx *ctx, const uint8_t *d_tight, uint64_t n_bases, uint8_t *d_packed, uint64_t *d_exc_pos,
                       uint8_t *d_exc_byte, uint64_t exc_cap, unsigned long long *d_n_exc) {
    const uint64_t n_words = (n_bases + 15) >> 4;
    if (!n_words) return;
    ProfScope ps(ctx, "pack_tight", (double)n_bases * 1.25);
    pack_tight_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(d_tight, n_bases, reinterpret_cast<uint32_t *>(d_packed),
                                                                              d_exc_pos, d_exc_byte, exc_cap, d_n_exc);
}

// ------------------------------------------------------------------------------------------
// Small results (spectrum, flags, totals) go back through mapped pinned memory, written by a
// one-block kernel, not through a copy engine: a 2 KB cudaMemcpyAsync queues behind whatever the
// device-to-host engine is doing, and with the asynchronous staging calls that is a 138 MB
// download of the previous chunk (measured: +2.1 ms per step of waiting in the compute stream).
// ------------------------------------------------------------------------------------------
__global__ void readback_kernel(uint32_t *__restrict__ host_mapped, const uint32_t *__restrict__ src, int n32) {
    for (int t = threadIdx.x; t < n32; t += blockDim.x) host_mapped[t] = src[t];
    __threadfence_system();
}

void launch_readback(brgpu_ctx *ctx, void *h_mapped_dst, const void *d_src, size_t bytes) {
    ctx->launches++;
    readback_kernel<<<1, 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(h_mapped_dst),
                                                reinterpret_cast<const uint32_t *>(d_src), (int)(bytes / 4));
}

// ------------------------------------------------------------------------------------------
// exclusive scan of u32 lengths into u64 offsets (n + 1 outputs).  Three small kernels:
// per-tile sums, one-block scan of the tile sums, per-tile scan with the tile's base added.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_TILE = 4096; // elements per block
constexpr int SCAN_THREADS = 256;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total, uint64_t *smem /* 32 */) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) smem[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(FULL, s, d);
            if (lane >= d) s += y;
        }
        smem[lane] = s; // inclusive over warps
    }
    __syncthreads();
    uint64_t warp_base = wid ? smem[wid - 1] : 0;
    *total = smem[(blockDim.x >> 5) - 1];
    uint64_t res = warp_base + x - v;
    __syncthreads();
    return res;
}

__global__ void scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ tile_sums) {
    __shared__ uint64_t sm[32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t acc = 0;
    for (int j = threadIdx.x; j < SCAN_TILE; j += SCAN_THREADS)
        if (base + j < n) acc += in[base + j];
    uint64_t total;
    block_exclusive_scan(acc, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_tile_bases_kernel(uint64_t *tile_sums, uint64_t n_tiles) {
    // one block; turns tile sums into exclusive bases in place, total in tile_sums[n_tiles]
    __shared__ uint64_t sm[32];
    uint64_t carry = 0;
    for (uint64_t b = 0; b < n_tiles; b += blockDim.x) {
        uint64_t idx = b + threadIdx.x;
        uint64_t v = idx < n_tiles ? tile_sums[idx] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan(v, &total, sm);
        if (idx < n_tiles) tile_sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void scan_apply_kernel(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ tile_bases,
                                  uint64_t n_tiles, uint64_t *__restrict__ out) {
    __shared__ uint64_t sm[32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t carry = tile_bases[blockIdx.x];
    constexpr int PER = SCAN_TILE / SCAN_THREADS; // 16 consecutive elements per thread
    uint64_t first = base + (uint64_t)threadIdx.x * PER;
    uint32_t v[PER];
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        v[j] = (first + j < n) ? in[first + j] : 0;
        acc += v[j];
    }
    uint64_t total;
    uint64_t ex = block_exclusive_scan(acc, &total, sm) + carry;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (first + j < n) out[first + j] = ex;
        ex += v[j];
    }
    if (blockIdx.x == n_tiles - 1 && threadIdx.x == 0) out[n] = tile_bases[n_tiles];
}

void launch_exclusive_scan_u32(brgpu_ctx *ctx, const uint32_t *d_in, uint64_t n, uint64_t *d_out, uint64_t *d_tmp) {
    if (n == 0) {
        cudaMemsetAsync(d_out, 0, sizeof(uint64_t), ctx->stream);
        return;
    }
    uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    ProfScope ps(ctx, "exclusive_scan", (double)n * 12.0);
    scan_tile_sums_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, d_tmp);
    scan_tile_bases_kernel<<<1, 1024, 0, ctx->stream>>>(d_tmp, n_tiles);
    scan_apply_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, d_tmp, n_tiles, d_out);
    ctx->launches += 2;
}

// ------------------------------------------------------------------------------------------
// Counting: one thread per 32 read positions, up to 32 independent random read-modify-writes
// in flight per thread.  CUDA has no 8-bit atomics, so the saturating increment is a CAS on
// the aligned 32-bit word that holds the counter; the first attempt guesses the word is still
// zero, which is true for the large majority of updates at k = 17 (2^33 counters).
// min(255, n) is order independent, so any exact scheme gives the table the CPU would.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sat_inc_fixup(uint32_t *w, uint32_t sh, uint32_t old) {
    const uint32_t one = 1u << sh;
    while (((old >> sh) & 0xffu) != 0xffu) {
        uint32_t assumed = old;
        old = atomicCAS(w, assumed, assumed + one);
        if (old == assumed) break;
    }
}

__global__ void __launch_bounds__(256)
    count_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                 const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read, uint64_t n_words, int k,
                 uint8_t *__restrict__ counts) {
    const uint64_t mask = kmask(k);
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = __ldg(word2read + w);
        uint64_t sb = w << 5;
        uint32_t p0 = (uint32_t)(sb - __ldg(slot_off + r));
        uint32_t L = __ldg(len + r);
        if (p0 >= L || L < (uint32_t)k) continue; // slack, or read shorter than k (src/set/pcon.rs:58)
        uint64_t prev, cur;
        load_window(seq, sb, p0, prev, cur);
        int t_lo = p0 >= (uint32_t)(k - 1) ? 0 : (k - 1 - (int)p0);
        int t_hi = (L - p0) < 32u ? (int)(L - p0) : 32;
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            uint32_t *wp[8];
            uint32_t sh[8], old[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int t = g + j;
                old[j] = 0;
                wp[j] = nullptr;
                sh[j] = 0;
                if (t >= t_lo && t < t_hi) {
                    uint64_t idx = canonical_index(window_kmer(prev, cur, t, mask), k);
                    wp[j] = reinterpret_cast<uint32_t *>(counts + (idx & ~3ULL));
                    sh[j] = (uint32_t)(idx & 3) * 8u;
                    old[j] = atomicCAS(wp[j], 0u, 1u << sh[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (old[j] != 0) sat_inc_fixup(wp[j], sh[j], old[j]);
        }
    }
}

void launch_count(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                  uint8_t *d_counts, double n_bases_hint) {
    uint64_t n_words = L.total_slots >> 5;
    if (!n_words) return;
    // algorithmic bytes: 0.25 B/base stream + one 32 B sector read + one 32 B sector
    // write-back per k-mer.  n_bases_hint ~ number of k-mers (sum len - n(k-1), clamped).
    ProfScope ps(ctx, "count_kmers", n_bases_hint * 64.25);
    count_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(d_seq, d_len, L.d_slot_off, L.d_word2read,
                                                                          n_words, k, d_counts);
}

// ------------------------------------------------------------------------------------------
// Bucketed counting (k >= 15): the same counts without a single random DRAM access.
//
// A random read-modify-write of the 2^(2k-1)-byte table costs two DRAM row activations and the
// whole chip sustains only ~20 G of them per second (profiles/microbench_random_access_r1.txt).
// So the k-mers are first partitioned by the top bits of their table index into buckets that
// each cover 2^15 consecutive counters — 16-bit residues, one streaming pass that writes 2 B
// per k-mer through L2-combined runs — and then every bucket is counted by one thread block in
// 32 KiB of shared memory (7 blocks per SM hide each other's load latency): zero, saturating u8 increments (CAS on the 32-bit shared word), then
// one sweep over the slice that feeds the spectrum and emits the slice's 8 KiB of the bitfield
// and its summary bits with plain coalesced stores.  The full table is never materialised and
// no atomics ever reach HBM.  The result (spectrum, bitfield) is identical to the table
// path's by construction: both compute min(255, occurrences) per canonical k-mer.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t msb_nibble(uint32_t cmp) {
    // cmp has 0xff/0x00 per byte; gather one bit per byte: byte j -> bit j
    uint32_t y = (cmp & 0x80808080u) >> 7;
    return (y | (y >> 7) | (y >> 14) | (y >> 21)) & 0xfu;
}

// Spectrum tallies + threshold nibble of one 32-bit word holding 4 counters that is known to be
// non-zero.  Deliberately not inlined: sparse tables make this the rare path, and inlining it
// eight times into the sweep loops turns the common all-zero case into ~50 predicated
// instructions per word (ncu: 3.5 G warp instructions for a 1 GiB sweep before this split).
__device__ __noinline__ uint32_t tally_word(uint32_t w, uint32_t thr, uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                            uint32_t &c3, unsigned int *sh_hist) {
    uint32_t z0 = __vcmpeq4(w, 0u), z1 = __vcmpeq4(w, 0x01010101u);
    uint32_t z2 = __vcmpeq4(w, 0x02020202u), z3 = __vcmpeq4(w, 0x03030303u);
    c0 += __popc(z0) >> 3;
    c1 += __popc(z1) >> 3;
    c2 += __popc(z2) >> 3;
    c3 += __popc(z3) >> 3;
    uint32_t rest = ~(z0 | z1 | z2 | z3);
    while (rest) {
        int q = (__ffs(rest) - 1) >> 3;
        atomicAdd(&sh_hist[(w >> (8 * q)) & 0xffu], 1u);
        rest &= ~(0xffu << (8 * q));
    }
    return msb_nibble(__vcmpgtu4(w, thr));
}

constexpr int BUCKET_BITS = 15;                  // counters per bucket = 2^15 (32 KiB of shared memory)
constexpr int BUCKET_COUNTERS = 1 << BUCKET_BITS;
constexpr int BUCKET_THREADS = 256;

template <class F>
__device__ __forceinline__ void for_each_kmer_index8(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                                                     const uint64_t *__restrict__ slot_off,
                                                     const uint32_t *__restrict__ word2read, uint64_t w, int k,
                                                     uint64_t mask, F f) {
    // calls f(idx[8], valid[8]) four times: 8 independent k-mers per call
    uint32_t r = __ldg(word2read + w);
    uint64_t sb = w << 5;
    uint32_t p0 = (uint32_t)(sb - __ldg(slot_off + r));
    uint32_t L = __ldg(len + r);
    if (p0 >= L || L < (uint32_t)k) return;
    uint64_t prev, cur;
    load_window(seq, sb, p0, prev, cur);
    int t_lo = p0 >= (uint32_t)(k - 1) ? 0 : (k - 1 - (int)p0);
    int t_hi = (L - p0) < 32u ? (int)(L - p0) : 32;
#pragma unroll
    for (int g = 0; g < 32; g += 8) {
        uint64_t idx[8];
        bool ok[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            idx[j] = canonical_index(window_kmer(prev, cur, g + j, mask), k);
            ok[j] = g + j >= t_lo && g + j < t_hi;
        }
        f(idx, ok);
    }
}

// pass 1: bucket sizes
__global__ void __launch_bounds__(256)
    bucket_hist_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                       const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read, uint64_t n_words,
                       int k, uint32_t *__restrict__ fill) {
    const uint64_t mask = kmask(k);
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x)
        for_each_kmer_index8(seq, len, slot_off, word2read, w, k, mask, [&](const uint64_t idx[8], const bool ok[8]) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (ok[j]) atomicAdd(fill + (idx[j] >> BUCKET_BITS), 1u); // result unused: compiles to RED
        });
}

// pass 2: scatter the 16-bit residues; the per-bucket cursors live in L2
__global__ void __launch_bounds__(256)
    bucket_scatter_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                          const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                          uint64_t n_words, int k, uint32_t *__restrict__ cursor, uint16_t *__restrict__ residues) {
    const uint64_t mask = kmask(k);
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x)
        for_each_kmer_index8(seq, len, slot_off, word2read, w, k, mask, [&](const uint64_t idx[8], const bool ok[8]) {
            // cursor[b] starts at the bucket's first slot, so one L2 atomic yields the position
            uint32_t pos[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                pos[j] = 0;
                if (ok[j]) pos[j] = atomicAdd(cursor + (idx[j] >> BUCKET_BITS), 1u);
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (ok[j]) residues[pos[j]] = (uint16_t)(idx[j] & (BUCKET_COUNTERS - 1));
        });
}

// Saturating u8 increment in a (shared-memory) word; returns true iff this call took the counter
// from 0 to 1, i.e. the caller is the k-mer's first occurrence in the bucket ("owner").
__device__ __forceinline__ bool sat_inc_owner(uint32_t *wp, uint32_t sh) {
    const uint32_t one = 1u << sh;
    uint32_t old = atomicCAS(wp, 0u, one);
    if (old == 0) return true;
    for (;;) {
        uint32_t b = (old >> sh) & 0xffu;
        if (b == 0xffu) return false;
        uint32_t assumed = old;
        old = atomicCAS(wp, assumed, assumed + one);
        if (old == assumed) return b == 0;
    }
}

__global__ void bucket_cursor_kernel(const uint64_t *__restrict__ base, uint64_t n_buckets,
                                     uint32_t *__restrict__ cursor) {
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_buckets;
         b += (uint64_t)gridDim.x * blockDim.x)
        cursor[b] = (uint32_t)base[b]; // total k-mers < 2^32 (checked by the caller)
}

// pass 3: one block per bucket, counters in shared memory.  Work per bucket is proportional to
// its k-mers, not to its 2^15 counters: the first occurrence of every distinct k-mer ("owner",
// known from the CAS that took the counter from 0 to 1) comes back after the block barrier,
// reads the final count, tallies the spectrum, sets the bit in a 4 KiB shared bitfield slice
// if count > abundance, and stores 0 to its counter byte — so the slice is clean again for
// the next bucket without ever being swept.  hist[0] is 2^15 minus the owners.
constexpr int BUCKET_REG_ROUNDS = 4; // residues kept in registers per thread between the two phases (multi-source kernel)

// THREADS x ROUNDS residues of a bucket are held in registers between the two phases; the rest (large
// buckets) is read twice; the next bucket's residues are requested while the owners of this one finish.
// The kernel is a chain of short phases per bucket (zero the slice, CAS increments, barrier, owners,
// barrier, write the slice, barrier: ~6 us per bucket at 256 threads) and what hides that chain is the
// number of buckets in flight per SM, which the 32 KiB counter array caps at 5.  Measured on the E. coli
// configuration (262 144 buckets of ~530 k-mers), profiles/count_shapes_r2.txt:
//   256 threads x 4 rounds  1.52 ms  (the default for every bucket size)
//   128 threads x 6 rounds  2.00 ms,  64 threads x 12 rounds  3.57 ms  — fewer threads per bucket lengthen
//       every phase and do not buy more buckets in flight (shared memory, not threads, is the limit);
//   a warp per bucket with a 1024-entry shared-memory hash table instead of the counter array (9.6 KB per
//       bucket, 20 buckets in flight per SM) 2.3-2.6 ms: per-lane linear probing diverges (8.7 of 32 lanes
//       active per instruction, 1.22 G warp instructions against 0.92 G).
// DIRECT: bitfield and summary were zeroed beforehand (a memset that runs beside the partition kernels) and
// the owners of solid k-mers set their bits in global memory (a few million `red.or` per step, absorbed by
// L2) — no slice in shared memory, nothing to zero or to write per bucket, two barriers instead of five,
// and one more bucket in flight per SM.
template <int THREADS, int ROUNDS, bool DIRECT>
__global__ void __launch_bounds__(THREADS)
    bucket_count_kernel(const uint16_t *__restrict__ residues, const uint64_t *__restrict__ base, uint64_t n_buckets,
                        int abundance, uint32_t *__restrict__ bitfield32, uint32_t *__restrict__ summary32,
                        int summary_shift, unsigned long long *__restrict__ g_hist) {
    extern __shared__ uint32_t cnt[];          // BUCKET_COUNTERS / 4 words of u8 counters
    __shared__ uint64_t sh_bits64[DIRECT ? 1 : BUCKET_COUNTERS / 64]; // the bucket's slice of the bitfield
    __shared__ unsigned int sh_hist[256];
    uint32_t *sh_bits = reinterpret_cast<uint32_t *>(sh_bits64);
    const bool emit = bitfield32 != nullptr;
    const bool emit_summary = summary32 != nullptr && summary_shift == 6;
    auto set_bit = [&](uint64_t b, uint32_t r) { // r: residue of a solid k-mer of bucket b
        if (DIRECT) {
            if (emit) {
                const uint64_t bit = b * BUCKET_COUNTERS + r;
                atomicOr(bitfield32 + (bit >> 5), 1u << (bit & 31));
                if (emit_summary) atomicOr(summary32 + (bit >> 11), 1u << ((bit >> 6) & 31));
            }
        } else {
            atomicOr(&sh_bits[r >> 5], 1u << (r & 31));
        }
    };
    uint8_t *cnt8 = reinterpret_cast<uint8_t *>(cnt);
    for (int t = threadIdx.x; t < 256; t += THREADS) sh_hist[t] = 0;
    for (int t = threadIdx.x; t < BUCKET_COUNTERS / 4; t += THREADS) cnt[t] = 0; // once: owners keep it clean
    uint32_t c1 = 0, c2 = 0, c3 = 0, owners = 0; // per-thread tallies of the values that hold the mass
    unsigned long long zeros = 0;
    uint64_t begin = 0, end = 0;
    uint32_t nres[ROUNDS]; // the register-held residues of the bucket the loop is about to enter
    auto fetch = [&](uint64_t from, uint64_t to) {
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) {
            const uint64_t j = from + threadIdx.x + (uint64_t)q * THREADS;
            nres[q] = j < to ? (uint32_t)__ldcs(residues + j) : 0xffffffffu;
        }
    };
    if (blockIdx.x < n_buckets) {
        begin = __ldg(base + blockIdx.x);
        end = __ldg(base + blockIdx.x + 1);
        fetch(begin, end);
    }
    __syncthreads();
    for (uint64_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        uint32_t res[ROUNDS];
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) res[q] = nres[q];
        // bounds of this block's next bucket: requested now, needed after phase 1
        uint64_t nb_begin = 0, nb_end = 0;
        const bool has_next = b + gridDim.x < n_buckets;
        if (has_next) {
            nb_begin = __ldg(base + b + gridDim.x);
            nb_end = __ldg(base + b + gridDim.x + 1);
        }
        if (!DIRECT)
            for (int t = threadIdx.x; t < BUCKET_COUNTERS / 32; t += THREADS) sh_bits[t] = 0;
        // ---- phase 1: saturating increments ----
        uint32_t own = 0; // bit q: res[q] is an owner
#pragma unroll
        for (int q = 0; q < ROUNDS; q++)
            if (res[q] != 0xffffffffu && sat_inc_owner(cnt + (res[q] >> 2), (res[q] & 3u) * 8u)) own |= 1u << q;
        const uint64_t rest_begin = begin + (uint64_t)ROUNDS * THREADS;
        for (uint64_t j = rest_begin + threadIdx.x; j < end; j += THREADS) {
            uint32_t r = __ldcs(residues + j);
            sat_inc_owner(cnt + (r >> 2), (r & 3u) * 8u);
        }
        if (has_next) fetch(nb_begin, nb_end); // the next bucket's residues travel while this one is finished
        __syncthreads();
        // ---- phase 2: owners read the final count, tally, threshold and clear ----
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) {
            if ((own >> q) & 1u) {
                const uint32_t r = res[q];
                const uint32_t c = cnt8[r];
                cnt8[r] = 0;
                owners++;
                if (c == 1) c1++;
                else if (c == 2) c2++;
                else if (c == 3) c3++;
                else atomicAdd(&sh_hist[c], 1u);
                if (c > (uint32_t)abundance) set_bit(b, r);
            }
        }
        // big buckets: the overflow occurrences claim their counter with an atomic swap-to-zero —
        // after the register-held owners have cleared theirs (block-uniform condition)
        if (end > rest_begin) __syncthreads();
        for (uint64_t j = rest_begin + threadIdx.x; j < end; j += THREADS) {
            const uint32_t r = __ldcs(residues + j);
            const uint32_t sh = (r & 3u) * 8u;
            const uint32_t old = atomicAnd(cnt + (r >> 2), ~(0xffu << sh));
            const uint32_t c = (old >> sh) & 0xffu;
            if (c) { // not yet claimed by a register-held owner or another overflow occurrence
                owners++;
                atomicAdd(&sh_hist[c], 1u);
                if (c > (uint32_t)abundance) set_bit(b, r);
            }
        }
        __syncthreads(); // every counter of this bucket is clean again (DIRECT: the next bucket may start)
        // ---- write the slice of the bitfield (+ its summary bits) ----
        if (!DIRECT && bitfield32) {
            uint64_t *bitfield64 = reinterpret_cast<uint64_t *>(bitfield32) + b * (BUCKET_COUNTERS / 64);
            uint32_t *summary = summary32 && summary_shift == 6 ? summary32 + b * (BUCKET_COUNTERS / 64 / 32) : nullptr;
#pragma unroll 4
            for (int t = threadIdx.x; t < BUCKET_COUNTERS / 64; t += THREADS) {
                const uint64_t out = sh_bits64[t];
                bitfield64[t] = out;
                if (summary) { // one summary bit per 64-bit block
                    const uint32_t m = __ballot_sync(FULL, out != 0);
                    if ((threadIdx.x & 31) == 0) summary[t >> 5] = m;
                }
            }
        }
        if (threadIdx.x == 0) zeros += BUCKET_COUNTERS;
        begin = nb_begin;
        end = nb_end;
        if (!DIRECT) __syncthreads(); // sh_bits is re-zeroed at the top of the next iteration
    }
    // hist[0] = counters nobody touched = all counters of this block's buckets - owners
    c1 = __reduce_add_sync(FULL, c1);
    c2 = __reduce_add_sync(FULL, c2);
    c3 = __reduce_add_sync(FULL, c3);
    owners = __reduce_add_sync(FULL, owners);
    __shared__ unsigned long long sh_owners;
    if (threadIdx.x == 0) sh_owners = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (c1) atomicAdd(&sh_hist[1], c1);
        if (c2) atomicAdd(&sh_hist[2], c2);
        if (c3) atomicAdd(&sh_hist[3], c3);
        if (owners) atomicAdd(&sh_owners, (unsigned long long)owners);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 256; t += THREADS)
        if (t > 0 && sh_hist[t]) atomicAdd(g_hist + t, (unsigned long long)sh_hist[t]);
    if (threadIdx.x == 0 && zeros >= sh_owners) atomicAdd(g_hist + 0, zeros - sh_owners);
}

// Multi-source variant for the multi-GPU path: the k-mers of bucket b come from this rank's
// residue array and from the residue arrays of the peer ranks, which are CUDA-IPC mappings of
// the peers' HBM — the exchange step of the sharded set construction is these loads, over
// NVLink, fused into the counting kernel (no staging copy, no collective).  Bucket offsets of
// all sources are local (the peers' slices are copied in beforehand, they are tiny).
constexpr int BUCKET_MAX_SOURCES = 64; // local partitions (one per chunk of reads) + peer partitions
struct BucketSources {
    const uint16_t *res[BUCKET_MAX_SOURCES];
    const uint64_t *base[BUCKET_MAX_SOURCES]; // indexable by absolute bucket id in [b0, b1]
    int n;
};

__global__ void __launch_bounds__(BUCKET_THREADS)
    bucket_count_multi_kernel(BucketSources src, uint64_t b0, uint64_t b1, int abundance,
                              uint32_t *__restrict__ bitfield32, uint32_t *__restrict__ summary32,
                              unsigned long long *__restrict__ g_hist) {
    extern __shared__ uint32_t cnt[];
    __shared__ uint64_t sh_bits64[BUCKET_COUNTERS / 64];
    __shared__ unsigned int sh_hist[256];
    __shared__ uint64_t sh_beg[BUCKET_MAX_SOURCES];
    __shared__ uint32_t sh_len[BUCKET_MAX_SOURCES];
    __shared__ uint32_t sh_cum[BUCKET_MAX_SOURCES + 1];
    uint32_t *sh_bits = reinterpret_cast<uint32_t *>(sh_bits64);
    __shared__ unsigned long long sh_owners;
    uint8_t *cnt8 = reinterpret_cast<uint8_t *>(cnt);
    for (int t = threadIdx.x; t < 256; t += BUCKET_THREADS) sh_hist[t] = 0;
    for (int t = threadIdx.x; t < BUCKET_COUNTERS / 4; t += BUCKET_THREADS) cnt[t] = 0;
    if (threadIdx.x == 0) sh_owners = 0;
    uint32_t owners = 0;
    unsigned long long zeros = 0;
    __syncthreads();
    for (uint64_t b = b0 + blockIdx.x; b < b1; b += gridDim.x) {
        if (threadIdx.x < src.n) { // every source's piece of this bucket: one thread per source
            const uint64_t beg = __ldg(src.base[threadIdx.x] + b);
            sh_beg[threadIdx.x] = beg;
            sh_len[threadIdx.x] = (uint32_t)(__ldg(src.base[threadIdx.x] + b + 1) - beg);
        }
        for (int t = threadIdx.x; t < BUCKET_COUNTERS / 32; t += BUCKET_THREADS) sh_bits[t] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int q = 0; q < src.n; q++) {
                sh_cum[q] = acc;
                acc += sh_len[q];
            }
            sh_cum[src.n] = acc;
        }
        __syncthreads();
        const uint32_t total = sh_cum[src.n];
        auto fetch = [&](uint32_t e) -> uint32_t { // e-th k-mer of the bucket over all sources
            int q = 0;
            while (e >= sh_cum[q + 1]) q++;
            return (uint32_t)__ldcs(src.res[q] + sh_beg[q] + (e - sh_cum[q]));
        };
        uint32_t res[BUCKET_REG_ROUNDS];
        uint32_t own = 0;
#pragma unroll
        for (int q = 0; q < BUCKET_REG_ROUNDS; q++) {
            uint32_t e = threadIdx.x + (uint32_t)q * BUCKET_THREADS;
            res[q] = e < total ? fetch(e) : 0xffffffffu;
        }
#pragma unroll
        for (int q = 0; q < BUCKET_REG_ROUNDS; q++)
            if (res[q] != 0xffffffffu && sat_inc_owner(cnt + (res[q] >> 2), (res[q] & 3u) * 8u)) own |= 1u << q;
        for (uint32_t e = BUCKET_REG_ROUNDS * BUCKET_THREADS + threadIdx.x; e < total; e += BUCKET_THREADS) {
            uint32_t r = fetch(e);
            sat_inc_owner(cnt + (r >> 2), (r & 3u) * 8u);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BUCKET_REG_ROUNDS; q++) {
            if ((own >> q) & 1u) {
                const uint32_t r = res[q];
                const uint32_t c = cnt8[r];
                cnt8[r] = 0;
                owners++;
                atomicAdd(&sh_hist[c], 1u);
                if (c > (uint32_t)abundance) atomicOr(&sh_bits[r >> 5], 1u << (r & 31));
            }
        }
        if (total > BUCKET_REG_ROUNDS * BUCKET_THREADS) __syncthreads(); // block-uniform
        for (uint32_t e = BUCKET_REG_ROUNDS * BUCKET_THREADS + threadIdx.x; e < total; e += BUCKET_THREADS) {
            const uint32_t r = fetch(e);
            const uint32_t sh = (r & 3u) * 8u;
            const uint32_t old = atomicAnd(cnt + (r >> 2), ~(0xffu << sh));
            const uint32_t c = (old >> sh) & 0xffu;
            if (c) {
                owners++;
                atomicAdd(&sh_hist[c], 1u);
                if (c > (uint32_t)abundance) atomicOr(&sh_bits[r >> 5], 1u << (r & 31));
            }
        }
        __syncthreads();
        if (bitfield32) {
            uint64_t *bitfield64 = reinterpret_cast<uint64_t *>(bitfield32);
            for (int t = threadIdx.x; t < BUCKET_COUNTERS / 64; t += BUCKET_THREADS) {
                const uint64_t out = sh_bits64[t];
                const uint64_t block = b * (BUCKET_COUNTERS / 64) + (uint64_t)t;
                bitfield64[block] = out;
                if (summary32) { // one summary bit per 64-bit block (shift 6)
                    const uint32_t m = __ballot_sync(FULL, out != 0);
                    if ((threadIdx.x & 31) == 0) summary32[block >> 5] = m;
                }
            }
        }
        if (threadIdx.x == 0) zeros += BUCKET_COUNTERS;
        __syncthreads();
    }
    owners = __reduce_add_sync(FULL, owners);
    if ((threadIdx.x & 31) == 0 && owners) atomicAdd(&sh_owners, (unsigned long long)owners);
    __syncthreads();
    for (int t = threadIdx.x; t < 256; t += BUCKET_THREADS)
        if (t > 0 && sh_hist[t]) atomicAdd(g_hist + t, (unsigned long long)sh_hist[t]);
    if (threadIdx.x == 0 && zeros >= sh_owners) atomicAdd(g_hist + 0, zeros - sh_owners);
}

void launch_bucket_count_multi(brgpu_ctx *ctx, const uint16_t *const *d_res, const uint64_t *const *d_base, int n_src,
                               uint64_t b0, uint64_t b1, int abundance, uint8_t *d_bits, uint32_t *d_summary,
                               uint64_t *d_hist, double n_kmers) {
    static PerDeviceOnce configured;
    if (configured.need(ctx->device)) {
        cudaFuncSetAttribute(bucket_count_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BUCKET_COUNTERS);
        cudaFuncSetAttribute(bucket_count_multi_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        configured.mark(ctx->device);
    }
    if (b1 <= b0) return;
    BucketSources src;
    src.n = n_src;
    for (int q = 0; q < BUCKET_MAX_SOURCES; q++) {
        src.res[q] = q < n_src ? d_res[q] : nullptr;
        src.base[q] = q < n_src ? d_base[q] : nullptr;
    }
    ProfScope ps(ctx, "bucket_count_multi", n_kmers * 2.0 + (double)(b1 - b0) * (BUCKET_COUNTERS / 8));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bucket_count_multi_kernel, BUCKET_THREADS,
                                                      BUCKET_COUNTERS) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
    uint64_t nb = b1 - b0;
    bucket_count_multi_kernel<<<(unsigned)(nb < cap ? nb : cap), BUCKET_THREADS, BUCKET_COUNTERS, ctx->stream>>>(
        src, b0, b1, abundance, reinterpret_cast<uint32_t *>(d_bits), d_summary, reinterpret_cast<unsigned long long *>(d_hist));
}

// Occupancy of a dense bitfield at 16-bit granularity: summary bit j <=> any of the bitfield bits [16 j, 16 j + 16).
// A thread turns 8 consecutive 64-bit blocks (two 32 B sectors) into one 32-bit word.
__global__ void __launch_bounds__(256) fine_summary_kernel(const uint4 *__restrict__ bits, uint64_t n_words, uint32_t *__restrict__ fine) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t out = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { // uint4 = two 64-bit blocks = eight 16-bit granules
            const uint4 v = __ldcs(bits + w * 4 + q);
            const uint32_t g = (v.x & 0xffffu ? 1u : 0u) | (v.x >> 16 ? 2u : 0u) | (v.y & 0xffffu ? 4u : 0u) | (v.y >> 16 ? 8u : 0u) |
                               (v.z & 0xffffu ? 16u : 0u) | (v.z >> 16 ? 32u : 0u) | (v.w & 0xffffu ? 64u : 0u) | (v.w >> 16 ? 128u : 0u);
            out |= g << (8 * q);
        }
        fine[w] = out;
    }
}

void launch_fine_summary(brgpu_ctx *ctx, const uint8_t *d_bits, uint64_t n_blocks64, uint32_t *d_fine) {
    const uint64_t n_words = n_blocks64 / 8;
    if (!n_words) return;
    ProfScope ps(ctx, "fine_summary", (double)n_blocks64 * 8.5);
    const uint64_t want = (n_words + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
    fine_summary_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(d_bits), n_words, d_fine);
}

// SolidView::pos8 from the compacted blocks: position of the only set bit, or POS8_MULTI
__global__ void __launch_bounds__(256) block_bytes_kernel(const uint64_t *__restrict__ blocks, uint64_t n, uint8_t *__restrict__ pos8) {
    // a thread turns 4 consecutive blocks into one 32-bit store
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t i = q * 4 + j;
            uint32_t v = POS8_MULTI;
            if (i < n) {
                const uint64_t b = __ldcs(blocks + i);
                if (__popcll(b) == 1) v = (uint32_t)__ffsll((long long)b) - 1u;
            }
            w |= v << (8 * j);
        }
        reinterpret_cast<uint32_t *>(pos8)[q] = w;
    }
}

void launch_block_bytes(brgpu_ctx *ctx, const uint64_t *d_blocks, uint64_t n_occupied, uint8_t *d_pos8) {
    if (!n_occupied) return;
    ProfScope ps(ctx, "block_bytes", (double)n_occupied * 9.0);
    const uint64_t quads = (n_occupied + 3) / 4, want = (quads + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
    block_bytes_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(d_blocks, n_occupied, d_pos8);
}

// Staged variant of the exchange: ONE launch pulls this rank's bucket range out of every peer partition —
// the residues (2 B per k-mer) and the slice of bucket offsets — with 16 B loads over NVLink, all peers in
// flight together (NVSwitch gives the reader its full inbound bandwidth whatever the number of sources; one
// cudaMemcpyAsync per peer, in stream order, ran at a third of it: 7 x 0.07 ms for 7 x 23 MB at N = 8).
// Source and destination of a segment are congruent modulo 16 (the caller pads the destination) or differ by 8
// (compacted slices land at their block offset: the body is then stored as two 8 B halves); sizes are
// multiples of 2: head and tail of a segment move as 2 B pieces, the body as 16 B loads.
constexpr int PULL_THREADS = 256;
constexpr int PULL_UNROLL = 4;
constexpr uint32_t PULL_CHUNK_VECS = PULL_THREADS * PULL_UNROLL * 4; // 64 KiB per block and turn

__global__ void __launch_bounds__(PULL_THREADS) peer_pull_kernel(PullSegments segs) {
    const uint32_t total = segs.first_chunk[segs.n];
    for (uint32_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        int lo = 0, hi = segs.n - 1; // last segment whose first chunk is <= chunk
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (segs.first_chunk[mid] <= chunk) lo = mid;
            else hi = mid - 1;
        }
        const uint8_t *src = segs.src[lo];
        uint8_t *dst = segs.dst[lo];
        const uint64_t bytes = segs.bytes[lo];
        const uint32_t local = chunk - segs.first_chunk[lo];
        uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u;
        if (head > bytes) head = bytes;
        const uint64_t n_vec = (bytes - head) >> 4;
        const uint64_t tail = bytes - head - (n_vec << 4);
        const uint4 *sv = reinterpret_cast<const uint4 *>(src + head);
        uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
        uint2 *dh = reinterpret_cast<uint2 *>(dst + head);
        const bool halves = ((reinterpret_cast<uintptr_t>(dst) - reinterpret_cast<uintptr_t>(src)) & 15u) != 0;
        const uint64_t v0 = (uint64_t)local * PULL_CHUNK_VECS;
        const uint64_t v1 = v0 + PULL_CHUNK_VECS < n_vec ? v0 + PULL_CHUNK_VECS : n_vec;
        for (uint64_t v = v0 + threadIdx.x; v < v1; v += PULL_THREADS * PULL_UNROLL) {
            uint4 r[PULL_UNROLL];
#pragma unroll
            for (int u = 0; u < PULL_UNROLL; u++)
                if (v + (uint64_t)u * PULL_THREADS < v1) r[u] = __ldcs(sv + v + (uint64_t)u * PULL_THREADS);
#pragma unroll
            for (int u = 0; u < PULL_UNROLL; u++) {
                const uint64_t at = v + (uint64_t)u * PULL_THREADS;
                if (at >= v1) continue;
                if (!halves) {
                    dv[at] = r[u];
                } else {
                    dh[2 * at] = make_uint2(r[u].x, r[u].y);
                    dh[2 * at + 1] = make_uint2(r[u].z, r[u].w);
                }
            }
        }
        if (local == 0) {
            const uint16_t *s2 = reinterpret_cast<const uint16_t *>(src);
            uint16_t *d2 = reinterpret_cast<uint16_t *>(dst);
            if (threadIdx.x < (head >> 1)) d2[threadIdx.x] = s2[threadIdx.x];
            const uint64_t t0 = (head + (n_vec << 4)) >> 1;
            if (threadIdx.x < (tail >> 1)) d2[t0 + threadIdx.x] = s2[t0 + threadIdx.x];
        }
    }
}

void launch_peer_pull(brgpu_ctx *ctx, PullSegments &segs, double bytes) {
    uint32_t at = 0;
    for (int q = 0; q < segs.n; q++) {
        segs.first_chunk[q] = at;
        uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(segs.src[q]) & 15u)) & 15u;
        if (head > segs.bytes[q]) head = segs.bytes[q];
        const uint64_t n_vec = (segs.bytes[q] - head) >> 4;
        const uint64_t chunks = (n_vec + PULL_CHUNK_VECS - 1) / PULL_CHUNK_VECS;
        at += (uint32_t)(chunks ? chunks : 1); // a segment without a body still has its head / tail turn
    }
    segs.first_chunk[segs.n] = at;
    if (!segs.n || !at) return;
    ProfScope ps(ctx, "peer_pull", bytes);
    const uint32_t cap = (uint32_t)ctx->sm_count * 8u;
    peer_pull_kernel<<<at < cap ? at : cap, PULL_THREADS, 0, ctx->stream>>>(segs);
}

// ------------------------------------------------------------------------------------------
// Two-level partition (k <= 17: at most 2^18 buckets).  The one-level scheme above pays one L2
// atomic per k-mer twice (sizes, then cursors: 138 M `red` + 138 M `atom` on the E. coli config,
// 0.8 + 1.9 ms) because with 2^18 destinations nothing can be combined inside a block.  Splitting
// the 18 bucket bits 9 + 9 makes both levels block-local problems:
//   coarse_hist     per-block shared-memory histogram over the 512 coarse buckets, one global add
//                   per block and bucket;
//   coarse_scatter  a block takes a tile of 8192 k-mers, ranks them inside the tile with shared-
//                   memory atomics, reserves each coarse bucket's run with one global atomic per
//                   tile and bucket (512 instead of 8192), stages the tile sorted by bucket in
//                   shared memory and writes it out as contiguous runs (u32: index bits below the
//                   coarse bits);
//   fine_partition  one block per coarse bucket (1 MB of k-mers): shared histogram over its 512
//                   fine buckets, scan, scatter of the 15-bit residues into bucket order — also
//                   produces the bucket offsets.
// The output (residues grouped by bucket, bucket offsets) is exactly what bucket_scatter produces,
// up to the order inside a bucket, which counting does not see.
// ------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_MAX_COARSE = 512;
constexpr int CP_TILE_KMERS = CP_THREADS * 32;

// KT: compile-time k (0 = runtime): with k = 17 (every BASELINE config) the shifts and masks of the
// k-mer arithmetic are immediates
template <int KT>
__global__ void __launch_bounds__(CP_THREADS)
    coarse_hist_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                       const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read, uint64_t n_words,
                       int k_, int coarse_shift_, uint32_t *__restrict__ g_hist) {
    const int k = KT ? KT : k_;
    const int coarse_shift = KT == 17 ? 24 : coarse_shift_;
    __shared__ uint32_t sh_cnt[CP_MAX_COARSE];
    for (int t = threadIdx.x; t < CP_MAX_COARSE; t += CP_THREADS) sh_cnt[t] = 0;
    __syncthreads();
    const uint64_t mask = kmask(k);
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x)
        for_each_kmer_index8(seq, len, slot_off, word2read, w, k, mask, [&](const uint64_t idx[8], const bool ok[8]) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (ok[j]) atomicAdd(&sh_cnt[idx[j] >> coarse_shift], 1u);
        });
    __syncthreads();
    for (int t = threadIdx.x; t < CP_MAX_COARSE; t += CP_THREADS)
        if (sh_cnt[t]) atomicAdd(g_hist + t, sh_cnt[t]);
}

template <int KT>
__global__ void __launch_bounds__(CP_THREADS)
    coarse_scatter_kernel(const uint8_t *__restrict__ seq, const uint32_t *__restrict__ len,
                          const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ word2read,
                          uint64_t n_words, int k_, int coarse_shift_, int n_coarse, uint32_t *__restrict__ cursor,
                          uint32_t *__restrict__ out) {
    const int k = KT ? KT : k_;
    const int coarse_shift = KT == 17 ? 24 : coarse_shift_; // k = 17: 18 bucket bits split 9 + 9, 9 + 15 below
    __shared__ uint32_t sh_cnt[CP_MAX_COARSE];
    __shared__ uint32_t sh_off[CP_MAX_COARSE];
    __shared__ uint32_t sh_gbase[CP_MAX_COARSE];
    __shared__ uint64_t sh_scan[32];
    __shared__ uint32_t sh_total;
    extern __shared__ uint32_t stage[];                 // CP_TILE_KMERS values, then CP_TILE_KMERS destinations
    uint32_t *stage_dst = stage + CP_TILE_KMERS;
    const uint64_t mask = kmask(k);
    const uint32_t rem_mask = (1u << coarse_shift) - 1u;
    const uint64_t n_tiles = (n_words + CP_THREADS - 1) / CP_THREADS;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t w = tile * CP_THREADS + threadIdx.x;
        for (int t = threadIdx.x; t < CP_MAX_COARSE; t += CP_THREADS) sh_cnt[t] = 0;
        // this thread's word: 32 positions of one read
        uint64_t prev = 0, cur = 0;
        int t_lo = 0, t_hi = 0;
        if (w < n_words) {
            const uint32_t r = __ldg(word2read + w);
            const uint64_t sb = w << 5;
            const uint32_t p0 = (uint32_t)(sb - __ldg(slot_off + r));
            const uint32_t L = __ldg(len + r);
            if (p0 < L && L >= (uint32_t)k) {
                load_window(seq, sb, p0, prev, cur);
                t_lo = p0 >= (uint32_t)(k - 1) ? 0 : (k - 1 - (int)p0);
                t_hi = (L - p0) < 32u ? (int)(L - p0) : 32;
            }
        }
        __syncthreads();
        // the 32 table indices of this thread, computed once: low words in registers, bit 32 in `hi`
        // (an index has 2k - 1 <= 33 bits); pass 1: how many k-mers go to every coarse bucket
        uint32_t lo[32];
        uint32_t hi = 0, okm = 0;
#pragma unroll
        for (int t = 0; t < 32; t++) {
            const uint64_t idx = canonical_index(window_kmer(prev, cur, t, mask), k);
            lo[t] = (uint32_t)idx;
            hi |= (uint32_t)(idx >> 32) << t;
            if (t >= t_lo && t < t_hi) {
                okm |= 1u << t;
                atomicAdd(&sh_cnt[idx >> coarse_shift], 1u);
            }
        }
        __syncthreads();
        // exclusive scan over the buckets (two per thread) = where each bucket's run starts in the
        // staging area; one global atomic per bucket reserves the run in the output
        {
            const uint32_t a = sh_cnt[2 * threadIdx.x], b = sh_cnt[2 * threadIdx.x + 1];
            uint64_t total;
            const uint32_t ex = (uint32_t)block_exclusive_scan((uint64_t)(a + b), &total, sh_scan);
            if (threadIdx.x == 0) sh_total = (uint32_t)total;
            sh_off[2 * threadIdx.x] = ex;
            sh_off[2 * threadIdx.x + 1] = ex + a;
            if (a) sh_gbase[2 * threadIdx.x] = atomicAdd(cursor + 2 * threadIdx.x, a);
            if (b) sh_gbase[2 * threadIdx.x + 1] = atomicAdd(cursor + 2 * threadIdx.x + 1, b);
            sh_cnt[2 * threadIdx.x] = 0;
            sh_cnt[2 * threadIdx.x + 1] = 0;
        }
        __syncthreads();
        // pass 2: rank inside the bucket's run, stage (8 independent shared atomics in flight)
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            uint32_t c[8], r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint64_t idx = ((uint64_t)((hi >> (g + j)) & 1u) << 32) | lo[g + j];
                c[j] = (uint32_t)(idx >> coarse_shift);
                r[j] = 0;
                if ((okm >> (g + j)) & 1u) r[j] = atomicAdd(&sh_cnt[c[j]], 1u);
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((okm >> (g + j)) & 1u) {
                    const uint32_t p = sh_off[c[j]] + r[j];
                    stage[p] = lo[g + j] & rem_mask;
                    stage_dst[p] = sh_gbase[c[j]] + r[j];
                }
        }
        __syncthreads();
        // write-out: staged position j goes to stage_dst[j]; neighbours in a bucket's run are
        // neighbours in the output, so the stores of a warp fall into a few 64 B runs
        for (uint32_t j = threadIdx.x; j < sh_total; j += CP_THREADS) out[stage_dst[j]] = stage[j];
        __syncthreads();
    }
}

constexpr int FP_THREADS = 512;
constexpr int FP_ILP = 8;                         // independent loads in flight per thread
constexpr int FP_TILE = FP_THREADS * FP_ILP;      // k-mers staged per round (4096)

// One block owns one coarse bucket.  Pass A: histogram over its fine buckets -> bucket offsets.
// Pass B: tiles of 4096 k-mers are ranked per fine bucket with shared-memory atomics, staged in
// bucket order and written as runs (a run = the tile's k-mers of one fine bucket, ~8 residues =
// one 16 B piece of a sector) — scattering the 2-byte residues one by one costs an L2 write
// transaction per k-mer and was 3x slower.
__global__ void __launch_bounds__(FP_THREADS)
    fine_partition_kernel(const uint32_t *__restrict__ coarse_kmers, const uint64_t *__restrict__ coarse_base,
                          int n_coarse, int fine_bits, uint64_t *__restrict__ base, uint16_t *__restrict__ residues) {
    __shared__ uint32_t sh_cnt[CP_MAX_COARSE];  // pass A: bucket sizes; pass B: ranks inside the tile
    __shared__ uint32_t sh_off[CP_MAX_COARSE];  // pass B: start of the bucket's run in the staging area
    __shared__ uint32_t sh_cur[CP_MAX_COARSE];  // k-mers of the bucket already written (relative to begin)
    __shared__ uint64_t sh_scan[32];
    __shared__ uint16_t stage[FP_TILE];
    __shared__ uint32_t stage_dst[FP_TILE];      // destination (relative to `begin`) of every staged residue
    __shared__ uint32_t sh_total;
    const int n_fine = 1 << fine_bits;
    for (int c = blockIdx.x; c < n_coarse; c += gridDim.x) {
        const uint64_t begin = __ldg(coarse_base + c), end = __ldg(coarse_base + c + 1);
        for (int t = threadIdx.x; t < CP_MAX_COARSE; t += FP_THREADS) sh_cnt[t] = 0;
        __syncthreads();
        for (uint64_t e0 = begin; e0 < end; e0 += FP_TILE) {
            uint32_t v[FP_ILP];
#pragma unroll
            for (int u = 0; u < FP_ILP; u++) {
                const uint64_t e = e0 + (uint64_t)u * FP_THREADS + threadIdx.x;
                v[u] = e < end ? __ldg(coarse_kmers + e) : 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < FP_ILP; u++)
                if (v[u] != 0xffffffffu) atomicAdd(&sh_cnt[v[u] >> BUCKET_BITS], 1u);
        }
        __syncthreads();
        {
            const uint32_t a = threadIdx.x < (unsigned)n_fine ? sh_cnt[threadIdx.x] : 0u;
            uint64_t total;
            const uint32_t ex = (uint32_t)block_exclusive_scan((uint64_t)a, &total, sh_scan);
            if (threadIdx.x < (unsigned)n_fine) {
                sh_cur[threadIdx.x] = ex;
                base[(uint64_t)c * n_fine + threadIdx.x] = begin + ex;
            }
            if (c == n_coarse - 1 && threadIdx.x == 0) base[(uint64_t)n_coarse * n_fine] = end;
        }
        for (uint64_t e0 = begin; e0 < end; e0 += FP_TILE) {
            if (threadIdx.x < CP_MAX_COARSE) sh_cnt[threadIdx.x] = 0;
            uint32_t v[FP_ILP], r[FP_ILP];
#pragma unroll
            for (int u = 0; u < FP_ILP; u++) {
                const uint64_t e = e0 + (uint64_t)u * FP_THREADS + threadIdx.x;
                v[u] = e < end ? __ldcs(coarse_kmers + e) : 0xffffffffu; // second and last read
            }
            __syncthreads(); // counters zeroed (and sh_cur of the previous tile updated)
#pragma unroll
            for (int u = 0; u < FP_ILP; u++) {
                r[u] = 0;
                if (v[u] != 0xffffffffu) r[u] = atomicAdd(&sh_cnt[v[u] >> BUCKET_BITS], 1u);
            }
            __syncthreads();
            {
                const uint32_t a = threadIdx.x < (unsigned)n_fine ? sh_cnt[threadIdx.x] : 0u;
                uint64_t total;
                const uint32_t ex = (uint32_t)block_exclusive_scan((uint64_t)a, &total, sh_scan);
                if (threadIdx.x < (unsigned)n_fine) sh_off[threadIdx.x] = ex;
                if (threadIdx.x == 0) sh_total = (uint32_t)total;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < FP_ILP; u++)
                if (v[u] != 0xffffffffu) {
                    const uint32_t f = v[u] >> BUCKET_BITS;
                    const uint32_t p = sh_off[f] + r[u];
                    stage[p] = (uint16_t)(v[u] & (BUCKET_COUNTERS - 1));
                    stage_dst[p] = sh_cur[f] + r[u];
                }
            __syncthreads();
            // staged position j goes to begin + stage_dst[j]: runs of ~8 consecutive residues
            for (uint32_t j = threadIdx.x; j < sh_total; j += FP_THREADS) residues[begin + stage_dst[j]] = stage[j];
            __syncthreads();
            if (threadIdx.x < (unsigned)n_fine) sh_cur[threadIdx.x] += sh_cnt[threadIdx.x];
        }
        __syncthreads();
    }
}

bool bucket_partition_two_level(uint64_t n_buckets) { return n_buckets >= 4 && n_buckets <= (uint64_t)CP_MAX_COARSE * CP_MAX_COARSE; }

void launch_bucket_partition(brgpu_ctx *ctx, const Layout &L, const uint8_t *d_seq, const uint32_t *d_len, int k,
                             uint64_t n_buckets, uint32_t *d_fill, uint64_t *d_base, uint64_t *d_scan_tmp,
                             uint16_t *d_residues, uint32_t *d_coarse_kmers, uint64_t *d_coarse_base, double n_kmers) {
    uint64_t n_words = L.total_slots >> 5;
    if (d_coarse_kmers && bucket_partition_two_level(n_buckets)) {
        int nb_bits = 0;
        while ((1ULL << nb_bits) < n_buckets) nb_bits++;
        const int coarse_bits = (nb_bits + 1) / 2, fine_bits = nb_bits - coarse_bits;
        const int n_coarse = 1 << coarse_bits;
        const int coarse_shift = fine_bits + BUCKET_BITS;
        uint32_t *d_hist = d_fill, *d_cursor = d_fill + CP_MAX_COARSE; // d_fill holds n_buckets >= 2 * 512 words... or fewer
        cudaMemsetAsync(d_fill, 0, 2 * CP_MAX_COARSE * sizeof(uint32_t), ctx->stream);
        {
            ProfScope ps(ctx, "coarse_hist", n_kmers * 1.0); // ASCII stream in
            if (k == 17)
                coarse_hist_kernel<17><<<grid_for(ctx, n_words, CP_THREADS, 8), CP_THREADS, 0, ctx->stream>>>(
                    d_seq, d_len, L.d_slot_off, L.d_word2read, n_words, k, coarse_shift, d_hist);
            else
                coarse_hist_kernel<0><<<grid_for(ctx, n_words, CP_THREADS, 8), CP_THREADS, 0, ctx->stream>>>(
                    d_seq, d_len, L.d_slot_off, L.d_word2read, n_words, k, coarse_shift, d_hist);
        }
        launch_exclusive_scan_u32(ctx, d_hist, (uint64_t)n_coarse, d_coarse_base, d_scan_tmp);
        {
            ProfScope ps(ctx, "coarse_scatter", n_kmers * 5.0); // ASCII in + 4 B out
            bucket_cursor_kernel<<<1, 512, 0, ctx->stream>>>(d_coarse_base, (uint64_t)n_coarse, d_cursor);
            const uint64_t n_tiles = (n_words + CP_THREADS - 1) / CP_THREADS;
            int per_sm = 0;
            const size_t cs_smem = 2 * CP_TILE_KMERS * sizeof(uint32_t);
            static PerDeviceOnce cs_configured; // 64 KiB of dynamic shared memory: opt-in, per device
            if (cs_configured.need(ctx->device)) {
                cudaFuncSetAttribute(coarse_scatter_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem);
                cudaFuncSetAttribute(coarse_scatter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem);
                cs_configured.mark(ctx->device);
            }
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coarse_scatter_kernel<17>, CP_THREADS, cs_smem) != cudaSuccess ||
                per_sm < 1) {
                cudaGetLastError();
                per_sm = 1;
            }
            const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
            const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
            if (k == 17)
                coarse_scatter_kernel<17><<<grid, CP_THREADS, cs_smem, ctx->stream>>>(d_seq, d_len, L.d_slot_off, L.d_word2read, n_words,
                                                                                k, coarse_shift, n_coarse, d_cursor, d_coarse_kmers);
            else
                coarse_scatter_kernel<0><<<grid, CP_THREADS, cs_smem, ctx->stream>>>(d_seq, d_len, L.d_slot_off, L.d_word2read, n_words,
                                                                               k, coarse_shift, n_coarse, d_cursor, d_coarse_kmers);
            ctx->launches += 1;
        }
        {
            ProfScope ps(ctx, "fine_partition", n_kmers * 6.0); // 4 B in + 2 B out
            fine_partition_kernel<<<n_coarse, FP_THREADS, 0, ctx->stream>>>(d_coarse_kmers, d_coarse_base, n_coarse, fine_bits,
                                                                            d_base, d_residues);
        }
        return;
    }
    cudaMemsetAsync(d_fill, 0, n_buckets * sizeof(uint32_t), ctx->stream);
    {
        ProfScope ps(ctx, "bucket_hist", n_kmers * 1.0); // ASCII stream in; the bucket counters stay in L2
        bucket_hist_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(d_seq, d_len, L.d_slot_off,
                                                                                    L.d_word2read, n_words, k, d_fill);
    }
    launch_exclusive_scan_u32(ctx, d_fill, n_buckets, d_base, d_scan_tmp);
    {
        ProfScope ps(ctx, "bucket_scatter", n_kmers * 3.0); // ASCII in + 2 B residue out
        bucket_cursor_kernel<<<grid_for(ctx, n_buckets, 256, 8), 256, 0, ctx->stream>>>(d_base, n_buckets, d_fill);
        bucket_scatter_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(
            d_seq, d_len, L.d_slot_off, L.d_word2read, n_words, k, d_fill, d_residues);
        ctx->launches += 1;
    }
}

template <int THREADS, int ROUNDS, bool DIRECT>
static void launch_bucket_count_shape(brgpu_ctx *ctx, const uint16_t *d_residues, const uint64_t *d_base, uint64_t n_buckets,
                                      int abundance, uint8_t *d_bits, uint32_t *d_summary, int summary_shift, uint64_t *d_hist) {
    static PerDeviceOnce configured;
    if (configured.need(ctx->device)) {
        cudaFuncSetAttribute(bucket_count_kernel<THREADS, ROUNDS, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BUCKET_COUNTERS);
        cudaFuncSetAttribute(bucket_count_kernel<THREADS, ROUNDS, DIRECT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        configured.mark(ctx->device);
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bucket_count_kernel<THREADS, ROUNDS, DIRECT>, THREADS, BUCKET_COUNTERS) !=
            cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
    bucket_count_kernel<THREADS, ROUNDS, DIRECT><<<(unsigned)(n_buckets < cap ? n_buckets : cap), THREADS, BUCKET_COUNTERS, ctx->stream>>>(
        d_residues, d_base, n_buckets, abundance, reinterpret_cast<uint32_t *>(d_bits), d_summary, summary_shift,
        reinterpret_cast<unsigned long long *>(d_hist));
}

void launch_bucket_count(brgpu_ctx *ctx, const uint16_t *d_residues, const uint64_t *d_base, uint64_t n_buckets,
                         int abundance, uint8_t *d_bits, uint32_t *d_summary, int summary_shift, uint64_t *d_hist,
                         double n_kmers, bool prezeroed) {
    // 2 B residue in per k-mer + the slice's share of the bitfield out (a sweep that reads the
    // residues twice for the rare buckets that overflow the register rounds is not counted)
    ProfScope ps(ctx, "bucket_count", n_kmers * 2.0 + (double)n_buckets * (BUCKET_COUNTERS / 8));
    const double mean = n_kmers / (double)(n_buckets ? n_buckets : 1);
    const int shape = ctx->opt_count_block_only; // 0 / 1: 256 threads, 2: 128, 3: 64 (A/B runs, tests)
    (void)mean;
    // prezeroed: the caller zeroed bitfield and summary (or wants no output): owners write their bits directly
    if (prezeroed && (d_summary == nullptr || summary_shift == 6) && shape != 1)
        launch_bucket_count_shape<BUCKET_THREADS, BUCKET_REG_ROUNDS, true>(ctx, d_residues, d_base, n_buckets, abundance, d_bits,
                                                                           d_summary, summary_shift, d_hist);
    else if (shape == 3)
        launch_bucket_count_shape<64, 12, false>(ctx, d_residues, d_base, n_buckets, abundance, d_bits, d_summary, summary_shift, d_hist);
    else if (shape == 2)
        launch_bucket_count_shape<128, 6, false>(ctx, d_residues, d_base, n_buckets, abundance, d_bits, d_summary, summary_shift, d_hist);
    else
        launch_bucket_count_shape<BUCKET_THREADS, BUCKET_REG_ROUNDS, false>(ctx, d_residues, d_base, n_buckets, abundance, d_bits,
                                                                            d_summary, summary_shift, d_hist);
}

// ------------------------------------------------------------------------------------------
// Spectrum + threshold: one streaming pass over the count table.  Each thread takes 16
// counters (one 16 B load, lanes contiguous), tallies values 0..3 with byte-SIMD compares in
// registers (they hold almost all the mass), sends the rest to a shared-memory histogram, and
// writes 16 bits of the LSB-first bitfield.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
    spectrum_threshold_kernel(const uint4 *__restrict__ counts16, uint64_t n16, unsigned long long *__restrict__ hist,
                              uint16_t *__restrict__ bits16, int abundance) {
    __shared__ unsigned int sh_hist[256];
    __shared__ unsigned long long sh_low[4];
    sh_hist[threadIdx.x] = 0;
    if (threadIdx.x < 4) sh_low[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t thr = (uint32_t)abundance * 0x01010101u;
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n16;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = __ldcs(counts16 + g);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t out = 0;
        if ((v.x | v.y | v.z | v.w) == 0) {
            c0 += 16;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (w[j] == 0)
                    c0 += 4;
                else
                    out |= tally_word(w[j], thr, c0, c1, c2, c3, sh_hist) << (4 * j);
            }
        }
        if (bits16) bits16[g] = (uint16_t)out;
    }
    c0 = __reduce_add_sync(FULL, c0);
    c1 = __reduce_add_sync(FULL, c1);
    c2 = __reduce_add_sync(FULL, c2);
    c3 = __reduce_add_sync(FULL, c3);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_low[0], (unsigned long long)c0);
        atomicAdd(&sh_low[1], (unsigned long long)c1);
        atomicAdd(&sh_low[2], (unsigned long long)c2);
        atomicAdd(&sh_low[3], (unsigned long long)c3);
    }
    __syncthreads();
    unsigned long long mine = sh_hist[threadIdx.x];
    if (threadIdx.x < 4) mine += sh_low[threadIdx.x];
    if (mine) atomicAdd(hist + threadIdx.x, mine);
}

void launch_spectrum_threshold(brgpu_ctx *ctx, const uint8_t *d_counts, uint64_t begin, uint64_t end, uint64_t *d_hist,
                               uint8_t *d_bits, int abundance) {
    if (end <= begin) return;
    uint64_t n16 = (end - begin) >> 4;
    double bytes = (double)(end - begin) * (d_bits ? 1.125 : 1.0);
    ProfScope ps(ctx, d_bits ? "spectrum_threshold" : "spectrum", bytes);
    // per-thread u32 tallies stay far below 2^32: >= sm_count*8*256 threads share <= 2^37 counters
    spectrum_threshold_kernel<<<grid_for(ctx, n16, 256, 8), 256, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(d_counts + begin), n16, reinterpret_cast<unsigned long long *>(d_hist),
        d_bits ? reinterpret_cast<uint16_t *>(d_bits + (begin >> 3)) : nullptr, abundance);
}

// tables smaller than 16 counters (k = 3: 32 counters is fine; k < 3 is rejected by the API)

// ------------------------------------------------------------------------------------------
// KmerSet::get / Solid::set over batches
// ------------------------------------------------------------------------------------------
__global__ void get_batch_kernel(const uint8_t *__restrict__ bits, int k, const uint64_t *__restrict__ kmers, uint64_t n,
                                 uint8_t *__restrict__ out) {
    const uint64_t mask = kmask(k);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = solid(bits, __ldg(kmers + i) & mask, k) ? 1 : 0;
}

void launch_get_batch(brgpu_ctx *ctx, const uint8_t *d_bits, int k, const uint64_t *d_kmers, uint64_t n,
                      uint8_t *d_out) {
    if (!n) return;
    ProfScope ps(ctx, "get_batch", (double)n * 41.0); // 8 B k-mer + 32 B sector + 1 B out
    get_batch_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d_bits, k, d_kmers, n, d_out);
}

__global__ void insert_batch_kernel(uint32_t *bits32, int k, const uint64_t *__restrict__ kmers, uint64_t n) {
    const uint64_t mask = kmask(k);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t idx = canonical_index(__ldg(kmers + i) & mask, k);
        atomicOr(bits32 + (idx >> 5), 1u << (idx & 31)); // little-endian u32 == LSB-first bytes
    }
}

void launch_insert_batch(brgpu_ctx *ctx, uint8_t *d_bits, int k, const uint64_t *d_kmers, uint64_t n) {
    if (!n) return;
    ProfScope ps(ctx, "insert_batch", (double)n * 72.0);
    insert_batch_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(d_bits), k,
                                                                           d_kmers, n);
}

// ------------------------------------------------------------------------------------------
// Occupancy summary: summary bit j = OR of the 2^shift bitfield bits of block j.  One warp
// produces one summary word per iteration (lane l reduces block 32*g + l, ballot packs them).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    build_summary_kernel(const uint32_t *__restrict__ bits32, uint64_t n_words, int words_per_block_log2,
                         uint32_t *__restrict__ summary, uint64_t n_summary_words) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t wpb = 1ULL << words_per_block_log2;
    for (uint64_t g = warp; g < n_summary_words; g += n_warps) {
        uint64_t first = ((g << 5) + (uint64_t)lane) << words_per_block_log2;
        uint32_t acc = 0;
        for (uint64_t t = 0; t < wpb; t++)
            if (first + t < n_words) acc |= __ldcs(bits32 + first + t);
        uint32_t m = __ballot_sync(FULL, acc != 0);
        if (lane == 0) summary[g] = m;
    }
}

void launch_build_summary(brgpu_ctx *ctx, const uint8_t *d_bits, uint64_t n_bytes, int shift, uint32_t *d_summary) {
    uint64_t n_words = n_bytes >> 2;
    uint64_t n_blocks = (n_bytes << 3) >> shift;
    uint64_t n_summary_words = (n_blocks + 31) >> 5;
    ProfScope ps(ctx, "build_summary", (double)n_bytes + (double)n_summary_words * 4.0);
    build_summary_kernel<<<grid_for(ctx, n_summary_words * 32, 256, 8), 256, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t *>(d_bits), n_words, shift - 5, d_summary, n_summary_words);
}

// ------------------------------------------------------------------------------------------
// Rank-compacted copy of a sparse bitfield (SolidView::dir / blocks in kmer.cuh).  Input: the
// shift-6 summary (one bit per 64-bit block).  popc per summary word -> exclusive scan = rank of
// the first block of every group; then every group copies its occupied blocks, in order, to
// blocks[rank ...] and stores {occupancy, rank} next to each other so that a lookup needs one
// 8-byte load to know whether and where.  Only occupied blocks are read from the bitfield.
// ------------------------------------------------------------------------------------------
__global__ void summary_popc_kernel(const uint32_t *__restrict__ summary, uint64_t n_words, uint32_t *__restrict__ pop) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_words;
         g += (uint64_t)gridDim.x * blockDim.x)
        pop[g] = (uint32_t)__popc(summary[g]);
}

void launch_summary_rank(brgpu_ctx *ctx, const uint32_t *d_summary, uint64_t n_words, uint32_t *d_pop, uint64_t *d_rank,
                         uint64_t *d_scan_tmp) {
    {
        ProfScope ps(ctx, "summary_popc", (double)n_words * 8.0);
        summary_popc_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(d_summary, n_words, d_pop);
    }
    launch_exclusive_scan_u32(ctx, d_pop, n_words, d_rank, d_scan_tmp);
}

__global__ void __launch_bounds__(256)
    compact_blocks_kernel(const uint32_t *__restrict__ summary, const uint64_t *__restrict__ rank,
                          const uint64_t *__restrict__ bits64, uint64_t n_words, uint2 *__restrict__ dir,
                          uint64_t *__restrict__ blocks) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_words;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t occ = summary[g];
        uint64_t r = rank[g];
        if (dir) dir[g] = make_uint2(occ, (uint32_t)r);
        while (occ) {
            const int i = __ffs(occ) - 1;
            blocks[r++] = __ldcs(bits64 + (g << 5) + (uint64_t)i);
            occ &= occ - 1;
        }
    }
}

// Streaming variant for sets that occupy more than a few percent of their blocks: a warp takes one
// group (32 blocks = one coalesced 256 B load), lanes with a non-empty block store it at
// rank + (number of occupied blocks before theirs).  Reads the whole bitfield once, at HBM speed.
__global__ void __launch_bounds__(256)
    compact_blocks_stream_kernel(const uint32_t *__restrict__ summary, const uint64_t *__restrict__ rank,
                                 const uint64_t *__restrict__ bits64, uint64_t n_words, uint2 *__restrict__ dir,
                                 uint64_t *__restrict__ blocks) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_words; g += n_warps) {
        const uint64_t blk = __ldcs(bits64 + (g << 5) + (uint64_t)lane);
        const uint32_t occ = __ldg(summary + g);
        const uint64_t r = __ldg(rank + g);
        if (lane == 0 && dir) dir[g] = make_uint2(occ, (uint32_t)r);
        if ((occ >> lane) & 1u) blocks[r + (uint64_t)__popc(occ & ((1u << lane) - 1u))] = blk;
    }
}

// dir[g] = {summary[g], rank[g]} alone: the blocks came from elsewhere (the sharded construction gathers every
// rank's compacted slice, in index order, instead of the 1 GiB bitfield)
__global__ void dir_only_kernel(const uint32_t *__restrict__ summary, const uint64_t *__restrict__ rank, uint64_t n_words,
                                uint2 *__restrict__ dir) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_words; g += (uint64_t)gridDim.x * blockDim.x)
        dir[g] = make_uint2(summary[g], (uint32_t)rank[g]);
}

void launch_dir_only(brgpu_ctx *ctx, const uint32_t *d_summary, const uint64_t *d_rank, uint64_t n_words, void *d_dir) {
    ProfScope ps(ctx, "rank_directory", (double)n_words * 20.0);
    dir_only_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(d_summary, d_rank, n_words, reinterpret_cast<uint2 *>(d_dir));
}

// the dense bitfield back from its rank-compacted form (the bitfield must be zero): a warp per group
__global__ void __launch_bounds__(256)
    expand_blocks_kernel(const uint2 *__restrict__ dir, const uint64_t *__restrict__ blocks, uint64_t n_words,
                         uint64_t *__restrict__ bits64) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_words; g += n_warps) {
        const uint2 e = __ldg(dir + g);
        if ((e.x >> lane) & 1u) bits64[(g << 5) + (uint64_t)lane] = __ldg(blocks + e.y + __popc(e.x & ((1u << lane) - 1u)));
    }
}

void launch_expand_blocks(brgpu_ctx *ctx, const void *d_dir, const uint64_t *d_blocks, uint64_t n_words, uint8_t *d_bits) {
    ProfScope ps(ctx, "expand_blocks", (double)n_words * 8.0);
    expand_blocks_kernel<<<grid_for(ctx, n_words * 32, 256, 8), 256, 0, ctx->stream>>>(reinterpret_cast<const uint2 *>(d_dir), d_blocks,
                                                                                    n_words, reinterpret_cast<uint64_t *>(d_bits));
}

void launch_compact_blocks(brgpu_ctx *ctx, const uint32_t *d_summary, const uint64_t *d_rank, const uint8_t *d_bits,
                           uint64_t n_words, uint64_t n_occupied, void *d_dir, uint64_t *d_blocks) {
    const bool sparse = n_occupied * 16 < n_words * 32; // fewer than 1 block in 16 occupied: read only those
    ProfScope ps(ctx, "compact_blocks", sparse ? (double)n_words * 20.0 + (double)n_occupied * 16.0
                                                : (double)n_words * 268.0 + (double)n_occupied * 8.0);
    if (sparse)
        compact_blocks_kernel<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(
            d_summary, d_rank, reinterpret_cast<const uint64_t *>(d_bits), n_words, reinterpret_cast<uint2 *>(d_dir), d_blocks);
    else
        compact_blocks_stream_kernel<<<grid_for(ctx, n_words * 32, 256, 8), 256, 0, ctx->stream>>>(
            d_summary, d_rank, reinterpret_cast<const uint64_t *>(d_bits), n_words, reinterpret_cast<uint2 *>(d_dir), d_blocks);
}

// ------------------------------------------------------------------------------------------
// Probe behind the "l2" yardstick bench.py reports for the solidity lookups: random 8-byte gathers
// over a table of the given size, 8 independent loads in flight per thread (the access pattern of
// solid_bitmap's directory / block loads).  A table that fits in L2 gives the L2 random-gather
// ceiling, one much larger than L2 the DRAM random-sector ceiling.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    probe_gather_kernel(const uint64_t *__restrict__ tab, uint64_t mask, uint64_t n_per_thread, uint64_t *__restrict__ sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += 8) {
        uint64_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t x = tid * n_per_thread + i + (uint64_t)j + 0x9E3779B97F4A7C15ULL;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
            v[j] = __ldg(tab + ((x ^ (x >> 31)) & mask));
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += v[j];
    }
    if (acc == 0xdeadbeefULL) sink[0] = acc;
}

void launch_probe_gather(brgpu_ctx *ctx, const uint64_t *d_tab, uint64_t n_words_pow2, uint64_t n_per_thread,
                         uint64_t *d_sink, uint64_t *n_gathers) {
    const unsigned blocks = (unsigned)ctx->sm_count * 8;
    ctx->launches++;
    probe_gather_kernel<<<blocks, 256, 0, ctx->stream>>>(d_tab, n_words_pow2 - 1, n_per_thread, d_sink);
    *n_gathers = (uint64_t)blocks * 256 * n_per_thread;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU merge: this rank's slice of the table += the same slice of every peer's table,
// read straight out of the peers' HBM over NVLink (the pointers are CUDA-IPC mappings), with a
// per-byte unsigned saturating add (min(255, sum) is associative and commutative, so the merged
// slice equals what one GPU counting all reads would hold).
// ------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 15;
struct PeerPtrs {
    const uint4 *p[MAX_PEERS];
};

__global__ void __launch_bounds__(256)
    merge_slice_kernel(uint4 *__restrict__ mine, PeerPtrs peers, int n_peers, uint64_t n16) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n16;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = mine[g];
        for (int q = 0; q < n_peers; q++) {
            uint4 b = __ldcs(peers.p[q] + g); // streaming: peer data is read once
            a.x = __vaddus4(a.x, b.x);
            a.y = __vaddus4(a.y, b.y);
            a.z = __vaddus4(a.z, b.z);
            a.w = __vaddus4(a.w, b.w);
        }
        mine[g] = a;
    }
}

void launch_merge_slice(brgpu_ctx *ctx, uint8_t *d_counts, void *const *peers, int n_peers, uint64_t begin,
                        uint64_t end) {
    if (end <= begin || n_peers <= 0) return;
    PeerPtrs pp;
    for (int q = 0; q < MAX_PEERS; q++)
        pp.p[q] = q < n_peers ? reinterpret_cast<const uint4 *>((const uint8_t *)peers[q] + begin) : nullptr;
    uint64_t n16 = (end - begin) >> 4;
    ProfScope ps(ctx, "merge_slice", (double)(end - begin) * (double)(n_peers + 2));
    merge_slice_kernel<<<grid_for(ctx, n16, 256, 8), 256, 0, ctx->stream>>>(
        reinterpret_cast<uint4 *>(d_counts + begin), pp, n_peers, n16);
}

} // namespace brgpu
