// scan_common.cuh — what the segmented scan shares between its host side (correct_kernels.cu) and the
// per-method kernel translation units (scan_methods.cu via scan_device.cuh).
#pragma once
#include "internal.h"
#include "kmer.cuh"

namespace brgpu {

constexpr uint32_t SEG = 2048;      // input positions per segment
constexpr uint32_t SEG_CAP = 3072;  // scratch bytes per segment (overrun + growth)
constexpr uint32_t NO_HORIZON = 0xffffffffu;

struct SegRec {
    uint32_t out_len;  // bytes written to the scratch region (may exceed SEG_CAP: then `bad`)
    uint32_t q_exit;   // first clean visit at or after the nominal end (>= len: read finished)
    uint32_t horizon;  // position of the first successful correction, NO_HORIZON if none
    uint32_t bad;      // output did not fit into the scratch region: the piece is unusable
};

// What the merge warp decided for a spliced piece: the copy itself is done afterwards by
// scan_splice_kernel, all pieces in parallel (a merge warp that copied its pieces one after the
// other made the longest read the tail of the kernel).
struct SegCopy {
    uint64_t dst;   // byte offset in the output slot buffer
    uint32_t skip;  // bytes of the piece's scratch region to skip
    uint32_t n;     // bytes to copy (0: nothing — the segment was re-run, or never reached)
};


// Warp-cooperative byte copy with arbitrary alignment on both sides: the destination is
// written as aligned 32-bit words, each assembled from two aligned source words with a funnel
// shift (128 B per warp step instead of 32).  May read up to 3 bytes beyond src + n inside the
// last aligned source word; all callers copy out of 32-byte-granular slot / scratch regions.
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    if (n < 64) {
        for (uint32_t t = lane; t < n; t += 32) dst[t] = src[t];
        return;
    }
    const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u); // bytes until dst is word aligned
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint8_t *s0 = src + head;
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t n_words = (n - head) >> 2;
    const uint32_t a = (uint32_t)((uintptr_t)s0 & 3u);
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s0 - a);
    // four independent load pairs in flight per lane: a 2 KiB piece is 4 round trips, not 16 (the
    // merge of the longest read is a chain of such copies and sets the kernel's tail)
    uint32_t w = lane;
    for (; w + 96 < n_words; w += 128) {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            lo[u] = s4[w + 32 * u];
            hi[u] = a ? s4[w + 32 * u + 1] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) d4[w + 32 * u] = a ? __funnelshift_r(lo[u], hi[u], 8 * a) : lo[u];
    }
    for (; w < n_words; w += 32) {
        uint32_t lo = s4[w];
        uint32_t v = lo;
        if (a) v = __funnelshift_r(lo, s4[w + 1], 8 * a);
        d4[w] = v;
    }
    const uint32_t done = head + (n_words << 2);
    if (done + (uint32_t)lane < n) dst[done + lane] = src[done + lane]; // at most 3 tail bytes
}

constexpr int SCAN_WARPS_PER_BLOCK = 4;

// number of segments of a read
__host__ __device__ __forceinline__ uint32_t seg_count(uint32_t len, uint32_t k) {
    if (len <= k) return 1;
    return (len - k + SEG - 1) / SEG;
}



#ifndef BRGPU_SCAN8_MINB
#define BRGPU_SCAN8_MINB 8 // the same for the four-segments-per-warp kernels
#endif
#ifndef BRGPU_SCAN_MINB
#define BRGPU_SCAN_MINB 8 // resident blocks per SM the One/Two warp-per-segment kernels are compiled for (64 registers)
#endif

// everything one pass of one method needs (launch_scan in correct_kernels.cu fills it)
struct ScanArgs {
    brgpu_ctx *ctx;
    const Layout *L;
    const uint8_t *d_in;
    const uint32_t *d_len_in;
    uint8_t *d_out;
    uint32_t *d_len_out;
    const uint32_t *d_bitmap;
    SolidView sv;
    CorrectParams p;
    uint8_t *d_scratch;
    size_t scratch_per_warp;
    int n_warps_total;
    const ScanWork *w;
    double n_bases_hint;
    const char *spec_name, *merge_name;
};

// defined in correct_kernels.cu
void launch_scan_splice(brgpu_ctx *ctx, const SegCopy *d_copies, uint64_t n_seg, const uint8_t *d_seg_out, uint8_t *d_out);

// the per-(method, variant) entry points of scan_methods.cu
#define BRGPU_DECLARE_SCAN_VARIANT(NS)                                                                                 \
    namespace NS {                                                                                                     \
    void launch_scan_m0(const ScanArgs &);                                                                             \
    void launch_scan_m1(const ScanArgs &);                                                                             \
    void launch_scan_m2(const ScanArgs &);                                                                             \
    void launch_scan_m3(const ScanArgs &);                                                                             \
    void launch_scan_m4(const ScanArgs &);                                                                             \
    }
BRGPU_DECLARE_SCAN_VARIANT(cnt)
BRGPU_DECLARE_SCAN_VARIANT(fast)
#undef BRGPU_DECLARE_SCAN_VARIANT

} // namespace brgpu
