// synth_kernels.cu — synthetic ONT-like reads generated on the device (BASELINE.json's
// configs are synthetic genomes of 4.6 Mb .. 1 Gb with 30-50x reads; 30 Gbases of ASCII do not fit the
// host, so every GPU generates its own shard).  Measurement support, not part of br's path.
//
// Everything is a pure function of (seed, counter), so the host mirror (br_b200/synth.py, numpy) yields
// the same bytes — tests compare the two:
//   mix64(x)          splitmix64's finaliser over x + golden ratio
//   genome code g(i)  mix64(genome_seed * GENOME_MUL + i) >> 62, index into "ACGT"; never stored
//   read r            template = genome[start, start + tlen), reverse-complemented when strand = 1;
//                     key_r = mix64(read_seed * READ_MUL + read_id)
//   position t        h = mix64(key_r + t); u = h & 0xffffff against three cumulative thresholds:
//                     substitution (base + 1 + ((h >> 24) & 0xffff) % 3), insertion (a random base
//                     (h >> 40) & 3 before the base), deletion, else the base
// The per-read descriptors (start, template length, strand) come from the host (16 B per read).
#include "internal.h"
#include "kmer.cuh"

namespace brgpu {

constexpr uint64_t GENOME_MUL = 0xD1342543DE82EF95ULL;
constexpr uint64_t READ_MUL = 0xA24BAED4963EE407ULL;
constexpr int SY_THREADS = 256;
constexpr int SY_PER = 8;                       // template positions per thread
constexpr int SY_TILE = SY_THREADS * SY_PER;    // template positions per block

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct SynthParams {
    uint64_t genome_seed, read_seed, first_read_id;
    uint32_t t_sub, t_ins, t_del; // cumulative 24-bit thresholds
};

// the (up to two) bytes template position t of read r contributes; returns their number
__device__ __forceinline__ int synth_emit(const SynthParams &sp, uint64_t key, uint64_t start, uint32_t tlen, uint32_t strand,
                                          uint32_t t, uint8_t out[2]) {
    const uint64_t gpos = strand ? start + (uint64_t)(tlen - 1u - t) : start + (uint64_t)t;
    uint32_t b = (uint32_t)(mix64(sp.genome_seed * GENOME_MUL + gpos) >> 62);
    if (strand) b = 3u - b;
    const uint64_t h = mix64(key + (uint64_t)t);
    const uint32_t u = (uint32_t)h & 0xffffffu;
    const char *acgt = "ACGT";
    if (u < sp.t_sub) {
        out[0] = (uint8_t)acgt[(b + 1u + ((uint32_t)(h >> 24) & 0xffffu) % 3u) & 3u];
        return 1;
    }
    if (u < sp.t_ins) {
        out[0] = (uint8_t)acgt[(uint32_t)(h >> 40) & 3u];
        out[1] = (uint8_t)acgt[b];
        return 2;
    }
    if (u < sp.t_del) return 0;
    out[0] = (uint8_t)acgt[b];
    return 1;
}

// tile -> (read, first template position); tile_first[r] = number of tiles before read r
__device__ __forceinline__ uint32_t tile_read(const uint64_t *__restrict__ tile_first, uint32_t n_reads, uint64_t tile) {
    uint32_t lo = 0, hi = n_reads;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(tile_first + mid) <= tile)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// WRITE = false: bytes produced by every tile -> tile_bytes; WRITE = true: the bytes, at tile_off[tile]
template <bool WRITE>
__global__ void __launch_bounds__(SY_THREADS)
    synth_tiles_kernel(SynthParams sp, const uint64_t *__restrict__ start, const uint32_t *__restrict__ tlen,
                       const uint8_t *__restrict__ strand, const uint64_t *__restrict__ tile_first, uint32_t n_reads,
                       uint64_t n_tiles, uint32_t *__restrict__ tile_bytes, const uint64_t *__restrict__ tile_off,
                       uint8_t *__restrict__ out) {
    __shared__ uint32_t sh_warp[SY_THREADS / 32];
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r = tile_read(tile_first, n_reads, tile);
        const uint32_t L = __ldg(tlen + r);
        const uint64_t s0 = __ldg(start + r);
        const uint32_t st = __ldg(strand + r);
        const uint64_t key = mix64(sp.read_seed * READ_MUL + sp.first_read_id + (uint64_t)r);
        const uint32_t t0 = (uint32_t)(tile - __ldg(tile_first + r)) * SY_TILE + threadIdx.x * SY_PER;
        uint8_t bytes[2 * SY_PER];
        uint32_t n = 0;
#pragma unroll
        for (int j = 0; j < SY_PER; j++)
            if (t0 + j < L) n += (uint32_t)synth_emit(sp, key, s0, L, st, t0 + j, bytes + n);
        // block-exclusive scan of n
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t x = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) sh_warp[wid] = x;
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int q = 0; q < SY_THREADS / 32; q++) {
            if (q < wid) base += sh_warp[q];
            total += sh_warp[q];
        }
        if (WRITE) {
            uint8_t *dst = out + __ldg(tile_off + tile) + base + (x - n);
            for (uint32_t j = 0; j < n; j++) dst[j] = bytes[j];
        } else if (threadIdx.x == 0) {
            tile_bytes[tile] = total;
        }
        __syncthreads();
    }
}

__global__ void synth_read_offsets_kernel(const uint64_t *__restrict__ tile_first, const uint64_t *__restrict__ tile_off,
                                          uint64_t n_reads, uint64_t *__restrict__ read_off) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r <= n_reads; r += (uint64_t)gridDim.x * blockDim.x)
        read_off[r] = tile_off[tile_first[r]];
}

void launch_synth_count(brgpu_ctx *ctx, const uint64_t seeds[3], const uint32_t thr[3], const uint64_t *d_start,
                        const uint32_t *d_tlen, const uint8_t *d_strand, const uint64_t *d_tile_first, uint64_t n_reads,
                        uint64_t n_tiles, uint32_t *d_tile_bytes) {
    if (!n_tiles) return;
    const SynthParams sp{seeds[0], seeds[1], seeds[2], thr[0], thr[1], thr[2]};
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    ctx->launches++;
    synth_tiles_kernel<false><<<(unsigned)(n_tiles < cap ? n_tiles : cap), SY_THREADS, 0, ctx->stream>>>(
        sp, d_start, d_tlen, d_strand, d_tile_first, (uint32_t)n_reads, n_tiles, d_tile_bytes, nullptr, nullptr);
}

void launch_synth_write(brgpu_ctx *ctx, const uint64_t seeds[3], const uint32_t thr[3], const uint64_t *d_start,
                        const uint32_t *d_tlen, const uint8_t *d_strand, const uint64_t *d_tile_first, uint64_t n_reads,
                        uint64_t n_tiles, const uint64_t *d_tile_off, uint8_t *d_out, uint64_t *d_read_off) {
    const SynthParams sp{seeds[0], seeds[1], seeds[2], thr[0], thr[1], thr[2]};
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (n_tiles) {
        ctx->launches++;
        synth_tiles_kernel<true><<<(unsigned)(n_tiles < cap ? n_tiles : cap), SY_THREADS, 0, ctx->stream>>>(
            sp, d_start, d_tlen, d_strand, d_tile_first, (uint32_t)n_reads, n_tiles, nullptr, d_tile_off, d_out);
    }
    ctx->launches++;
    const uint64_t blocks = (n_reads + 256) / 256;
    synth_read_offsets_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(d_tile_first, d_tile_off,
                                                                                              n_reads, d_read_off);
}

int synth_tile_positions() { return SY_TILE; }

} // namespace brgpu
