// scan_device.cuh — device side of the correction scan (phase B), included by scan_methods.cu once per
// (method, variant).  BRGPU_VARIANT names the namespace the code lands in: `cnt` counts every
// KmerSet::get the scan issues (profiling runs: the roofline needs lookups per launch), `fast` compiles
// the bookkeeping out (2 % of scan_two's instructions, profiles/hotspots_r1final).
//
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

#pragma once
#include "scan_common.cuh"

#ifndef BRGPU_VARIANT
#error "define BRGPU_VARIANT (cnt | fast) and BRGPU_COUNT_GETS (1 | 0)"
#endif

namespace brgpu {
namespace BRGPU_VARIANT {

// ------------------------------------------------------------------------------------------
// phase B — warp-level machinery.  All `Rd` fields and all scalar state in the scan are
// warp-uniform unless a comment says "per lane".
// ------------------------------------------------------------------------------------------
struct Rd {
    const uint8_t *in;   // read bytes (slot)
    uint32_t len;        // input length
    uint8_t *out;        // output slot
    uint32_t cap;        // output capacity
    const uint32_t *bm;  // phase-A bitmap of this read (bit p = solid(k-mer ending at p))
    SolidView set;       // the solid set (bitfield + L2 summary)
    int k;
    uint64_t mask;
    int lane;
    uint32_t o;          // bytes produced so far (keeps counting past cap)
    uint32_t copy_from;  // input bytes [copy_from, i) are still to be copied to out
    uint8_t *scratch;    // per-warp scratch (Greedy)
#if BRGPU_COUNT_GETS
    mutable uint32_t n_get; // per lane: KmerSet::get calls issued by this lane (bookkeeping)
#endif
    // cached 1024-position window of the phase-A bitmap: lane l holds word bm_base + l (per lane)
    uint32_t bm_w;
    uint32_t bm_base;    // 0xffffffff: nothing cached
    // 64 input bases around the current event, 2-bit packed, first base in the top pair of w0;
    // every lane holds the same copy, so a lane can cut out any sub-sequence with shifts alone
    uint64_t w0, w1;
    uint32_t w_origin;   // input position of the first base of the window; 0xffffffff: none
};

// (Re)load the 64-base window starting at input position `origin`.
__device__ __forceinline__ void load_win(Rd &rd, uint32_t origin) {
    uint32_t a = origin + (uint32_t)rd.lane, b = a + 32u;
    uint64_t c0 = a < rd.len ? (uint64_t)nuc2bit(rd.in[a]) << (2 * (31 - rd.lane)) : 0ULL;
    uint64_t c1 = b < rd.len ? (uint64_t)nuc2bit(rd.in[b]) << (2 * (31 - rd.lane)) : 0ULL;
    rd.w0 = warp_or64(c0);
    rd.w1 = warp_or64(c1);
    rd.w_origin = origin;
}

// does the window hold input positions [pos, pos + n)?
template <class R> __device__ __forceinline__ bool win_covers(const R &rd, uint32_t pos, uint32_t n) {
    return rd.w_origin != 0xffffffffu && pos >= rd.w_origin && pos + n <= rd.w_origin + 64u;
}

// the n (<= 32) 2-bit codes of input positions [pos, pos + n), first base most significant
template <class R> __device__ __forceinline__ uint64_t win_extract(const R &rd, uint32_t pos, uint32_t n) {
    if (n == 0) return 0ULL;
    const uint32_t off = pos - rd.w_origin, e = off + n;
    uint64_t v;
    if (e <= 32u)
        v = rd.w0 >> (2 * (32u - e));
    else if (off >= 32u)
        v = rd.w1 >> (2 * (64u - e));
    else
        v = (rd.w0 << (2 * (e - 32u))) | (rd.w1 >> (2 * (64u - e)));
    return n >= 32u ? v : (v & ((1ULL << (2 * n)) - 1ULL));
}

// `kmer` after pushing the input bases [pos, pos + n) taken from the window
template <class R> __device__ __forceinline__ uint64_t win_push(const R &rd, uint64_t kmer, uint32_t pos, uint32_t n) {
    uint64_t hi = n < 32u ? (kmer << (2 * n)) : 0ULL;
    return (hi | win_extract(rd, pos, n)) & rd.mask;
}

// KmerSet::get from the scan: one random sector of the bitfield, counted for the roofline report
__device__ __forceinline__ bool lookup(const Rd &rd, uint64_t kmer) {
#if BRGPU_COUNT_GETS
    rd.n_get++;
#endif
    return solid(rd.set, kmer);
}

// result of correct_error
struct Corr {
    bool some;
    uint32_t n_emit;   // bases emitted
    uint32_t codes;    // up to 3 emitted 2-bit codes, first base in the highest used pair (Exist)
    bool in_place;     // bases already written at out[o..] (walk methods); new_kmer is valid
    uint64_t new_kmer; // rolling k-mer after the emitted bases (walk methods)
    uint32_t offset;   // read bases consumed
};


__device__ __forceinline__ void copy_range(Rd &rd, uint32_t from, uint32_t to) {
    if (to <= from) return;
    uint32_t n = to - from;
    if (rd.o + n <= rd.cap) {
        warp_copy(rd.out + rd.o, rd.in + from, n, rd.lane);
    } else { // the slot overflows: keep counting, write what fits
        for (uint32_t t = rd.lane; t < n; t += 32) {
            uint32_t dst = rd.o + t;
            if (dst < rd.cap) rd.out[dst] = rd.in[from + t];
        }
    }
    rd.o += n;
}

__device__ __forceinline__ void flush_copy(Rd &rd, uint32_t upto) {
    if (upto > rd.len) upto = rd.len;
    if (upto > rd.copy_from) {
        copy_range(rd, rd.copy_from, upto);
        rd.copy_from = upto;
    }
}

__device__ __forceinline__ void emit_byte(Rd &rd, uint8_t b) {
    if (rd.lane == 0 && rd.o < rd.cap) rd.out[rd.o] = b;
    rd.o += 1;
}

// k-mer made of the input bases in[p-k+1 ..= p]
__device__ __forceinline__ uint64_t load_kmer_at(const Rd &rd, uint32_t p) {
    uint64_t v = 0;
    if (rd.lane < rd.k) v = (uint64_t)nuc2bit(rd.in[p - (uint32_t)rd.k + 1u + (uint32_t)rd.lane]) << (2 * (rd.k - 1 - rd.lane));
    return warp_or64(v);
}

// per lane l: `kmer` after pushing ptr[0..=l]; only lanes l < n (n <= 32) get a defined value
__device__ __forceinline__ uint64_t push_window(const Rd &rd, uint64_t kmer, const uint8_t *ptr, uint32_t n) {
    uint64_t c = 0;
    if ((uint32_t)rd.lane < n) c = (uint64_t)nuc2bit(ptr[rd.lane]) << (2 * (31 - rd.lane));
    uint64_t W = warp_or64(c);
    int sh = 2 * (rd.lane + 1);
    uint64_t hi = sh < 64 ? (kmer << sh) : 0ULL;
    return (hi | (W >> (2 * (31 - rd.lane)))) & rd.mask;
}

// `kmer` after pushing ptr[0..n) — scalar loop, any lane may call it on its own data
__device__ __forceinline__ uint64_t push_seq(uint64_t kmer, const uint8_t *ptr, uint32_t n, uint64_t mask) {
    for (uint32_t u = 0; u < n; u++) kmer = push(kmer, nuc2bit(ptr[u]), mask);
    return kmer;
}

// 4-bit mask of the successors of x that are solid: next_nucs(x) (src/correct/mod.rs:118-128).
// alt_nucs(y) == succ_mask(y >> 2) with the same push, because push masks the top base away.
__device__ __forceinline__ uint32_t succ_mask(const Rd &rd, uint64_t x) {
    bool s = false;
    if (rd.lane < 4) s = lookup(rd, push(x, (uint32_t)rd.lane, rd.mask));
    return __ballot_sync(FULL, s) & 0xfu;
}
__device__ __forceinline__ uint32_t alt_mask(const Rd &rd, uint64_t weak) {
    bool s = false;
    if (rd.lane < 4) s = lookup(rd, replace_last(weak, (uint32_t)rd.lane, rd.mask));
    return __ballot_sync(FULL, s) & 0xfu;
}
__device__ __forceinline__ bool uniq(uint32_t m4, uint32_t &a) {
    a = (uint32_t)(__ffs(m4) - 1);
    return __popc(m4) == 1;
}

// First position j in [i, end) with !S[j] && P[j], where S is the phase-A bitmap, P[j] = S[j-1]
// for j > i and P[i] = previous.  Returns end when there is none (end <= len).
__device__ __forceinline__ uint32_t find_transition(Rd &rd, uint32_t i, bool previous, uint32_t end) {
    if (i >= end) return end;
    uint32_t wbase = i >> 5;
    const uint32_t n_words = (end + 31) >> 5;
    uint32_t carry_in = 0;
    // events are ~60 bases apart: the 1024-position window loaded for the previous one usually
    // still covers this search, so keep it in registers instead of going back to L2
    bool cached = rd.bm_base != 0xffffffffu && wbase >= rd.bm_base && wbase < rd.bm_base + 32u;
    if (cached) wbase = rd.bm_base;
    for (;;) {
        uint32_t wi = wbase + (uint32_t)rd.lane;
        uint32_t W;
        if (cached) {
            W = rd.bm_w;
            cached = false;
        } else {
            W = wi < ((rd.len + 31) >> 5) ? __ldg(rd.bm + wi) : 0u; // cache real words, mask below
            rd.bm_w = W;
            rd.bm_base = wbase;
        }
        uint32_t up = __shfl_up_sync(FULL, W, 1);
        uint32_t carry = rd.lane ? (up >> 31) : carry_in;
        uint32_t T = ~W & ((W << 1) | carry);
        uint32_t posbase = wi << 5;
        if (posbase + 31 < i) {
            T = 0;
        } else if (posbase <= i) {
            uint32_t sh = i - posbase;
            T &= (0xffffffffu << sh);
            T &= ~(1u << sh);
            if (previous && !((W >> sh) & 1u)) T |= 1u << sh;
        }
        if (posbase >= end)
            T = 0;
        else if (posbase + 32 > end)
            T &= (1u << (end - posbase)) - 1u;
        uint32_t any = __ballot_sync(FULL, T != 0);
        if (any) {
            int fl = __ffs(any) - 1;
            uint32_t Tf = __shfl_sync(FULL, T, fl);
            return ((wbase + (uint32_t)fl) << 5) + (uint32_t)(__ffs(Tf) - 1);
        }
        wbase += 32;
        if (wbase >= n_words) return end;
        // continue in the next window: its first position has P = S[pos-1] = top bit of lane 31
        carry_in = __shfl_sync(FULL, W, 31) >> 31;
        i = wbase << 5;
        previous = carry_in != 0;
    }
}

// First position p in [from, len) whose phase-A bit is set; len when none.
__device__ __forceinline__ uint32_t find_solid(const Rd &rd, uint32_t from) {
    if (from >= rd.len) return rd.len;
    uint32_t wbase = from >> 5;
    const uint32_t n_words = (rd.len + 31) >> 5;
    for (;;) {
        uint32_t wi = wbase + (uint32_t)rd.lane;
        uint32_t W = wi < n_words ? __ldg(rd.bm + wi) : 0u;
        uint32_t posbase = wi << 5;
        if (posbase + 31 < from)
            W = 0;
        else if (posbase <= from)
            W &= 0xffffffffu << (from - posbase);
        uint32_t any = __ballot_sync(FULL, W != 0);
        if (any) {
            int fl = __ffs(any) - 1;
            uint32_t Wf = __shfl_sync(FULL, W, fl);
            uint32_t p = ((wbase + (uint32_t)fl) << 5) + (uint32_t)(__ffs(Wf) - 1);
            return p < rd.len ? p : rd.len; // bits at positions >= len are never set, but stay safe
        }
        wbase += 32;
        if (wbase >= n_words) return rd.len;
    }
}

// error_len (src/correct/mod.rs:130-152) for the weak k-mer `kmer` whose last base is in[i].
// `dr` = how many of the following pushes still produce k-mers that contain corrected bases
// (0 in the clean state): those need real lookups, everything after is in the phase-A bitmap.
// Returns elen; hit_end = the weak run reaches the end of the read (then fck is not solid).
__device__ __forceinline__ uint32_t error_len(Rd &rd, uint64_t kmer, uint32_t i, uint32_t dr, bool &hit_end,
                                              uint64_t &fck) {
    const uint32_t sublen = rd.len - i;
    hit_end = false;
    // j = 1..dr: dirty k-mers (dr <= k-2 < 32, one round)
    uint32_t nd = dr < sublen - 1 ? dr : sublen - 1; // pushes available: sub[1..sublen-1]
    if (nd > 0) {
        uint64_t km = push_window(rd, kmer, rd.in + i + 1, nd);
        bool s = (uint32_t)rd.lane < nd && lookup(rd, km);
        uint32_t m = __ballot_sync(FULL, s);
        if (m) {
            int l = __ffs(m) - 1;
            fck = shfl64(km, l);
            return (uint32_t)l + 1u;
        }
    }
    if (nd == sublen - 1) { // ran out of read inside the dirty part
        hit_end = true;
        fck = 0;
        return sublen;
    }
    uint32_t p = find_solid(rd, i + dr + 1);
    if (p >= rd.len) {
        hit_end = true;
        fck = 0;
        return sublen;
    }
    fck = load_kmer_at(rd, p); // k-mers beyond the dirty part are pure input
    return p - i;
}

// ------------------------------------------------------------------------------------------
// Exist<S>::correct_error (src/correct/exist/mod.rs:120-149) for One (3 scenarios) and Two (13).
//
// Rounds of independent lookups instead of the reference's nested calls:
//   round 1  alt_nucs(kmer)                                   4 lookups
//   round 2  (Two only) the successor sets every apply() needs 16 lookups:
//            N0 = next(K0), N1 = next(push(K0,s1)), N2 = next(push(K0,s2)), N0' = next(push(K0,s0))
//            — every alt_nucs(...) inside two.rs:98-254 reduces to one of these because
//            add_nuc_to_end masks the oldest base away
//   round 3  per scenario: get(K), the c confirmations of get_score, and the one_more k-mer
//            NS * (c + 2) lookups, flattened over the lanes
// A result needs exactly one surviving scenario, so evaluation order cannot change it.
// ------------------------------------------------------------------------------------------
struct Scen {        // per lane: lane s describes scenario s
    bool valid;      // apply() returned Some (length preconditions + uniqueness)
    uint64_t K;      // k-mer after apply
    uint32_t offa;   // apply offset (scoring)
    uint32_t offc;   // correct offset (advance)
    uint32_t n_emit; // bases correct() emits
    uint32_t codes;  // their 2-bit codes, first base highest
};

__device__ __forceinline__ Scen scen_one(int s, uint64_t K0) {
    Scen sc;
    sc.valid = s < 3;
    sc.K = K0;
    sc.offa = sc.offc = (uint32_t)(2 - s); // I:2 S:1 D:0 (one.rs:57-71)
    sc.n_emit = 1;
    sc.codes = (uint32_t)(K0 & 3);
    return sc;
}

__device__ __forceinline__ Scen scen_two(int s, uint64_t K0, uint32_t sublen, const uint32_t sb[4], uint32_t N0,
                                         uint32_t N1, uint32_t N2, uint32_t N0p, uint64_t mask) {
    Scen sc;
    sc.valid = false;
    sc.K = K0;
    sc.offa = sc.offc = 0;
    sc.n_emit = 0;
    sc.codes = 0;
    const uint32_t last = (uint32_t)(K0 & 3);
    uint32_t u0, u1, u2, u0p;
    const bool q0 = uniq(N0, u0), q1 = uniq(N1, u1), q2 = uniq(N2, u2), q0p = uniq(N0p, u0p);
    const uint64_t K1 = push(K0, sb[1], mask);
    switch (s) {
    case 0: // II  two.rs:96, :260
        sc.valid = true; sc.offa = 3; sc.offc = 2; sc.n_emit = 1; sc.codes = last; break;
    case 1: // IS  two.rs:97, :261
        sc.valid = true; sc.offa = 2; sc.offc = 2; sc.n_emit = 1; sc.codes = last; break;
    case 2: // SS  two.rs:98-114
        sc.valid = sublen >= 2 && !((N0 >> sb[1]) & 1) && q0;
        sc.K = push(K0, u0, mask); sc.offa = 2; sc.offc = 2; sc.n_emit = 2; sc.codes = (last << 2) | u0; break;
    case 3: // SD  two.rs:115-126
        sc.valid = sublen >= 1 && q0;
        sc.K = push(K0, u0, mask); sc.offa = 1; sc.offc = 1; sc.n_emit = 2; sc.codes = (last << 2) | u0; break;
    case 4: // DD  two.rs:127-134
        sc.valid = q0;
        sc.K = push(K0, u0, mask); sc.offa = 0; sc.offc = 0; sc.n_emit = 2; sc.codes = (last << 2) | u0; break;
    case 5: // ICI two.rs:135-148, :275
        sc.valid = sublen >= 4 && ((N0 >> sb[3]) & 1);
        sc.K = push(K0, sb[3], mask); sc.offa = 4; sc.offc = 3; sc.n_emit = 1; sc.codes = last; break;
    case 6: // ICS two.rs:149-166, :289-301 (correct offset = apply + 1)
        sc.valid = sublen >= 4 && !((N0 >> sb[1]) & 1) && q0;
        sc.K = push(K0, u0, mask); sc.offa = 3; sc.offc = 4; sc.n_emit = 2; sc.codes = (last << 2) | u0; break;
    case 7: // ICD two.rs:167-181, :276-288 (correct offset = apply - 1; K0's alt base is not emitted)
        sc.valid = sublen >= 4 && q2;
        sc.K = push(push(K0, sb[2], mask), u2, mask); sc.offa = 3; sc.offc = 2; sc.n_emit = 2;
        sc.codes = (sb[2] << 2) | u2; break;
    case 8: // SCI two.rs:182-191
        sc.valid = sublen >= 4;
        sc.K = push(K1, sb[3], mask); sc.offa = 4; sc.offc = 4; sc.n_emit = 3;
        sc.codes = (last << 4) | (sb[1] << 2) | sb[3]; break;
    case 9: // SCS two.rs:192-215
        sc.valid = sublen >= 3 && ((N0 >> sb[1]) & 1) && !((N1 >> sb[2]) & 1) && q1;
        sc.K = push(K1, u1, mask); sc.offa = 3; sc.offc = 3; sc.n_emit = 3;
        sc.codes = (last << 4) | (sb[1] << 2) | u1; break;
    case 10: // SCD two.rs:216-230
        sc.valid = sublen >= 2 && q1;
        sc.K = push(K1, u1, mask); sc.offa = 2; sc.offc = 2; sc.n_emit = 3;
        sc.codes = (last << 4) | (sb[1] << 2) | u1; break;
    case 11: // DCI two.rs:231-240, :323 (`_ => (vec![], 1)`)
        sc.valid = sublen >= 4;
        sc.K = push(K1, sb[3], mask); sc.offa = 4; sc.offc = 1; sc.n_emit = 0; sc.codes = 0; break;
    case 12: // DCD two.rs:241-254
        sc.valid = sublen >= 2 && q0p;
        sc.K = push(push(K0, sb[0], mask), u0p, mask); sc.offa = 1; sc.offc = 1; sc.n_emit = 3;
        sc.codes = (last << 4) | (sb[0] << 2) | u0p; break;
    default: break;
    }
    return sc;
}

template <int NS>
__device__ __forceinline__ Corr exist_correct_error(Rd &rd, uint64_t kmer, uint32_t i, uint32_t c) {
    Corr res;
    res.some = false;
    res.in_place = false;
    res.n_emit = 0;
    res.codes = 0;
    res.offset = 0;
    res.new_kmer = 0;

    uint32_t alt;
    if (!uniq(alt_mask(rd, kmer), alt)) return res; // exist/mod.rs:121-126
    const uint64_t K0 = replace_last(kmer, alt, rd.mask);
    const uint8_t *sub = rd.in + i;
    const uint32_t sublen = rd.len - i;
    // every base a scenario looks at lies in sub[0 .. c + 6): take them from the register window
    // when it reaches that far (it does for the usual confirm values), else from memory
    const bool use_win = c <= 30u && win_covers(rd, i, c + 6u);
    auto sub_push = [&](uint64_t km, uint32_t from, uint32_t n) -> uint64_t {
        return use_win ? win_push(rd, km, i + from, n) : push_seq(km, sub + from, n, rd.mask);
    };

    Scen sc;
    if (NS == 3) {
        sc = scen_one(rd.lane, K0);
    } else {
        uint32_t sb[4];
        if (use_win) { // one cut for the four bases; codes of positions >= len are zero in the window
            const uint32_t four = (uint32_t)win_extract(rd, i, 4);
#pragma unroll
            for (int t = 0; t < 4; t++) sb[t] = (four >> (2 * (3 - t))) & 3u;
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) sb[t] = (uint32_t)t < sublen ? nuc2bit(sub[t]) : 0u;
        }
        // round 2: the four successor sets
        bool s = false;
        if (rd.lane < 16) {
            int g = rd.lane >> 2;
            uint64_t base = K0;
            bool need = true;
            if (g == 1) { base = push(K0, sb[1], rd.mask); need = sublen >= 2; }
            if (g == 2) { base = push(K0, sb[2], rd.mask); need = sublen >= 3; }
            if (g == 3) { base = push(K0, sb[0], rd.mask); }
            if (need) s = lookup(rd, push(base, (uint32_t)(rd.lane & 3), rd.mask));
        }
        uint32_t m = __ballot_sync(FULL, s);
        sc = scen_two(rd.lane, K0, sublen, sb, m & 0xf, (m >> 4) & 0xf, (m >> 8) & 0xf, (m >> 12) & 0xf, rd.mask);
    }

    uint32_t bad = 0, more = 0; // per lane partial masks over scenarios
    const uint32_t valid_mask = __ballot_sync(FULL, sc.valid && rd.lane < NS);
    // get_score's length test: `if offset + c > seq.len() return 0` (exist/mod.rs:29-31)
    const uint32_t short_mask = __ballot_sync(FULL, sc.valid && rd.lane < NS && sc.offa + c > sublen);
    uint32_t cand;
    if (NS == 3) {
        // round 3, One: items (s, u), u = 0 .. c+1, all at once (3 * (c + 2) lookups, one round for
        // the usual confirm values)
        const uint32_t per = c + 2;
        const uint32_t Q = (uint32_t)NS * per;
        for (uint32_t q0 = 0; q0 < Q; q0 += 32) {
            uint32_t q = q0 + (uint32_t)rd.lane;
            int s = (int)(q / per);
            uint32_t u = q - (uint32_t)s * per;
            Scen t = scen_one(s, K0); // ScenarioOne is a function of s alone: no need to ask lane s
            if (q >= Q || ((short_mask >> s) & 1u)) continue;
            if (u <= c) {
                // u = 0: get(K) (exist/mod.rs:23); u = 1..c: the c confirmations (:35-43)
                if (!lookup(rd, sub_push(t.K, t.offa, u))) bad |= 1u << s;
            } else if (sublen > c + t.offc + 1) {
                // one_more (exist/mod.rs:49-70): uses correct()'s bases and offset, tests one k-mer
                uint64_t km = push(K0 >> 2, t.codes & 3u, rd.mask);
                if (lookup(rd, sub_push(km, t.offc, c + 1))) more |= 1u << s;
            }
        }
        bad = __reduce_or_sync(FULL, bad);
        more = __reduce_or_sync(FULL, more);
        cand = valid_mask & ~short_mask & ~bad;
    } else {
        // Two: most of the 13 scenarios die on get(K), so the rounds are taken one after the other
        // and only survivors go on — 13 + 5 * survivors lookups instead of 13 * 7.
        // round 3a: get(K) of every valid scenario (exist/mod.rs:23), lane s asks for scenario s
        bool alive = sc.valid && rd.lane < NS && !((short_mask >> rd.lane) & 1u);
        if (alive) alive = lookup(rd, sc.K);
        cand = __ballot_sync(FULL, alive);
        // round 3b: the c confirmations (:35-43) of the survivors, item (r-th survivor, u = 1..c)
        const uint32_t n_alive = (uint32_t)__popc(cand);
        const uint32_t Q = n_alive * c;
        // survivor r is scenario ids[r] (4 bits each): every surviving lane contributes its id at its
        // rank among the survivors (cheaper than one __fns per item, which is a software loop)
        uint64_t ids = 0;
        if (alive) ids = (uint64_t)rd.lane << (4 * __popc(cand & ((1u << rd.lane) - 1u)));
        ids = warp_or64(ids);
        const uint32_t inv_c = c > 1 ? 0xffffffffu / c + 1u : 0u; // q / c by multiplication (q < 2^16)
        for (uint32_t q0 = 0; q0 < Q; q0 += 32) {
            const uint32_t q = q0 + (uint32_t)rd.lane;
            const uint32_t r = c > 1 ? __umulhi(q, inv_c) : q, u = q - r * c + 1u;
            const int s = q < Q ? (int)((ids >> (4 * r)) & 0xfu) : 0;
            const uint64_t K = shfl64(sc.K, s); // all lanes take part in the shuffles
            const uint32_t offa = __shfl_sync(FULL, sc.offa, s);
            if (q < Q && !lookup(rd, sub_push(K, offa, u))) bad |= 1u << s;
        }
        cand &= ~__reduce_or_sync(FULL, bad);
        // round 3c, only on a tie: one_more (exist/mod.rs:49-70) of the tied scenarios
        if (__popc(cand) > 1) {
            bool m = false;
            if ((cand >> rd.lane) & 1u) {
                if (sublen > c + sc.offc + 1) {
                    uint64_t km = K0 >> 2;
                    for (int e = (int)sc.n_emit - 1; e >= 0; e--) km = push(km, (sc.codes >> (2 * e)) & 3u, rd.mask);
                    m = lookup(rd, sub_push(km, sc.offc, c + 1));
                }
            }
            more = __ballot_sync(FULL, m);
        }
    }

    if (cand == 0) return res;                       // exist/mod.rs:132-134
    if (__popc(cand) > 1) {                          // :138-148
        cand &= more;
        if (__popc(cand) != 1) return res;
    }
    int win = __ffs(cand) - 1;
    res.some = true;
    res.n_emit = __shfl_sync(FULL, sc.n_emit, win);
    res.codes = __shfl_sync(FULL, sc.codes, win);
    res.offset = __shfl_sync(FULL, sc.offc, win);
    return res;
}

// ------------------------------------------------------------------------------------------
// Graph::correct_error (src/correct/graph.rs:44-85).
//
// The walk x0 -> x1 -> ... follows the unique solid successor, i.e. it is a deterministic
// sequence.  The reference keeps an FxHashSet of visited k-mers and fails on the first revisit;
// for a deterministic sequence "first_correct_kmer is reached before any revisit" is the same
// as "first_correct_kmer is reached at all" (a value inside the cycle would have been met
// before the cycle closed), with the one exception x0 == first_correct_kmer, which the
// reference can only meet as a revisit.  So the visited set is replaced by Brent's cycle
// detection, which only has to guarantee termination.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Corr graph_correct_error(Rd &rd, uint64_t kmer, uint32_t i, uint32_t dr) {
    Corr res;
    res.some = false;
    res.in_place = true;
    res.n_emit = 0;
    res.codes = 0;
    res.offset = 0;
    res.new_kmer = 0;

    bool hit_end;
    uint64_t fck;
    uint32_t elen = error_len(rd, kmer, i, dr, hit_end, fck);

    uint32_t alt;
    if (!uniq(alt_mask(rd, kmer), alt)) return res;
    // weak run reaches the end of the read: first_correct_kmer is not solid, every walk k-mer is,
    // so the reference can only end in None
    if (hit_end) return res;
    uint64_t x = replace_last(kmer, alt, rd.mask);
    if (x == fck) return res;

    flush_copy(rd, i);
    const uint32_t o0 = rd.o;
    uint32_t n = 0;
    if (rd.lane == 0 && o0 + n < rd.cap) rd.out[o0 + n] = bit2nuc(alt);
    n++;

    uint64_t tortoise = x;
    uint32_t power = 1, lam = 0;
    for (;;) {
        uint32_t a;
        if (!uniq(succ_mask(rd, x), a)) return res; // dead end or branch
        x = push(x, a, rd.mask);
        lam++;
        if (x == tortoise) return res; // cycle that never meets fck
        if (lam == power) {
            tortoise = x;
            power <<= 1;
            lam = 0;
        }
        if (rd.lane == 0 && o0 + n < rd.cap) rd.out[o0 + n] = bit2nuc(a);
        n++;
        if (x == fck) break;
        if (n == 0xfffffff0u) return res; // cannot happen: 2^33 solid k-mers at most
    }
    res.some = true;
    res.n_emit = n;
    res.new_kmer = x;
    res.offset = elen + 1;
    return res;
}

// ------------------------------------------------------------------------------------------
// GapSize::ins_sub_correction (src/correct/gap_size.rs:44-89): exactly `gap` unique-successor
// steps, failing on a revisit.  For the deterministic walk "x0..x_gap are pairwise distinct" is
// equivalent to "x_gap does not occur among x0..x_{gap-1}" (once a value repeats, every later
// value is a repeat as well), so the hash set is replaced by one check of the last k-mer
// against the k-mers of the emitted path, done by all lanes in parallel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Corr ins_sub_correction(Rd &rd, uint64_t kmer, uint32_t i, uint32_t gap) {
    Corr res;
    res.some = false;
    res.in_place = true;
    res.n_emit = 0;
    res.codes = 0;
    res.offset = 0;
    res.new_kmer = 0;

    uint32_t alt;
    if (!uniq(alt_mask(rd, kmer), alt)) return res;
    const uint64_t x0 = replace_last(kmer, alt, rd.mask);
    uint64_t x = x0;

    flush_copy(rd, i);
    const uint32_t o0 = rd.o;
    uint32_t n = 0;
    if (rd.lane == 0 && o0 + n < rd.cap) rd.out[o0 + n] = bit2nuc(alt);
    n++;
    for (uint32_t t = 0; t < gap; t++) {
        uint32_t a;
        if (!uniq(succ_mask(rd, x), a)) return res;
        x = push(x, a, rd.mask);
        if (rd.lane == 0 && o0 + n < rd.cap) rd.out[o0 + n] = bit2nuc(a);
        n++;
    }
    // revisit check: is x (= x_gap) one of x_0 .. x_{gap-1}?  x_t ends with path base t, and
    // the path bases are out[o0 .. o0+gap].  If the path did not fit in the slot the batch is
    // re-run with more capacity anyway, so the answer does not matter then.
    if (gap > 0 && o0 + n <= rd.cap) {
        __syncwarp();
        const uint32_t chunk = (gap + 31) / 32; // x_0 .. x_{gap-1}: gap candidates
        uint32_t t0 = (uint32_t)rd.lane * chunk;
        uint32_t t1 = t0 + chunk < gap ? t0 + chunk : gap;
        bool hit = false;
        if (t0 < t1) {
            // roll from x0 to x_{t0}: push path bases 1..t0
            uint64_t y = x0;
            // jump: only the last k bases matter, start at most k steps before t0
            uint32_t start = t0 > (uint32_t)rd.k ? t0 - (uint32_t)rd.k : 0;
            if (start > 0) y = 0;
            for (uint32_t t = start + 1; t <= t0; t++) y = push(y, nuc2bit(rd.out[o0 + t]), rd.mask);
            if (start > 0) {
                // y now holds path bases start+1..t0 (k of them) == x_{t0} because t0 - start == k
            }
            for (uint32_t t = t0;;) {
                if (y == x) hit = true;
                if (++t >= t1) break;
                y = push(y, nuc2bit(rd.out[o0 + t]), rd.mask);
            }
        }
        if (__any_sync(FULL, hit)) return res;
    }
    res.some = true;
    res.n_emit = n;
    res.new_kmer = x;
    res.offset = n; // offset = local_corr.len() (gap_size.rs:87)
    return res;
}

// GapSize::correct_error (src/correct/gap_size.rs:97-108)
__device__ __forceinline__ Corr gap_size_correct_error(Rd &rd, uint64_t kmer, uint32_t i, uint32_t dr, uint32_t c) {
    bool hit_end;
    uint64_t fck;
    uint32_t elen = error_len(rd, kmer, i, dr, hit_end, fck);
    if (elen < (uint32_t)rd.k) return graph_correct_error(rd, kmer, i, dr);
    if (elen == (uint32_t)rd.k) return exist_correct_error<3>(rd, kmer, i, c);
    return ins_sub_correction(rd, kmer, i, elen - (uint32_t)rd.k);
}

// ------------------------------------------------------------------------------------------
// Greedy (src/correct/greedy.rs:56-173).
//
// match_alignement runs bio 1.6.0's affine global aligner (gap open -1, extend -1, match +1,
// mismatch -1) on x = before || read[..i] and y = before || path.  In `global` mode every clip
// penalty is MIN_SCORE, so no clip state can ever win and the recurrence reduces to the classic
// three layers with bio's tie rules: extension beats opening only if strictly greater; S takes
// the diagonal first, then I, then D, each only if strictly greater; an I (or D) cell opened
// from S stores the S pointer of the cell it came from.  One lane owns up to RMAX consecutive
// rows and the lanes sweep the columns as a systolic wavefront; the 12-bit traceback cells go
// to per-warp scratch and the traceback itself is replayed by all lanes uniformly.
// ------------------------------------------------------------------------------------------
enum : uint32_t { TB_START = 0, TB_INS = 1, TB_DEL = 2, TB_SUBST = 3, TB_MATCH = 4 };
enum : uint8_t { OP_MATCH = 0, OP_SUBST = 1, OP_DEL = 2, OP_INS = 3 };
constexpr int NEG = -1000000; // stands in for MIN_SCORE: always loses against a real score
constexpr int GREEDY_RMAX = 9; // rows per lane: covers k-1 + 255 + 1 rows

struct GreedyScratch {
    uint16_t *tb;     // (m+1) x (n+1) traceback cells, row-major with stride n_max+1
    uint8_t *ops;     // reversed operations
    uint8_t *x;       // before || read part
    uint8_t *y;       // before || path
    uint64_t *viewed; // visited k-mers
    int32_t *edge;    // 3 x (n_max+1): S, I, sbits of the last row of the lane above (systolic hand-off)
};

__host__ __device__ inline size_t greedy_dim(int k, int max_search) { return (size_t)(k - 1 + max_search + 2); }

__device__ __forceinline__ GreedyScratch greedy_scratch(uint8_t *base, int k, int max_search) {
    size_t d = greedy_dim(k, max_search);
    GreedyScratch g;
    size_t off = 0;
    g.viewed = reinterpret_cast<uint64_t *>(base + off);
    off += ((size_t)max_search + 2) * 8;
    g.tb = reinterpret_cast<uint16_t *>(base + off);
    off += d * d * 2;
    off = (off + 3) & ~(size_t)3;
    g.edge = reinterpret_cast<int32_t *>(base + off);
    off += 3 * d * 4;
    g.ops = base + off;
    off += 2 * d;
    g.x = base + off;
    off += d;
    g.y = base + off;
    return g;
}

// Global alignment of x[0..m) vs y[0..n); fills g.ops (reversed) and returns the op count.
__device__ __forceinline__ uint32_t bio_global(const Rd &rd, const GreedyScratch &g, uint32_t m, uint32_t n,
                                               uint32_t stride) {
    const int lane = rd.lane;
    const uint32_t rows = m + 1; // rows 0..m; row 0 is the boundary
    const uint32_t R = (rows + 31) / 32; // rows per lane (<= GREEDY_RMAX)
    const uint32_t r0 = (uint32_t)lane * R; // first row of this lane
    // per-lane state for its rows at the previous column: S and D; at the current column: S, I
    int Sp[GREEDY_RMAX], Dp[GREEDY_RMAX];
    uint32_t sbp[GREEDY_RMAX]; // S-pointer of (row, j-1)
    // column 0
#pragma unroll
    for (int q = 0; q < GREEDY_RMAX; q++) {
        uint32_t i = r0 + (uint32_t)q;
        Sp[q] = NEG;
        Dp[q] = NEG;
        sbp[q] = TB_START;
        if ((uint32_t)q < R && i < rows) {
            if (i == 0) {
                Sp[q] = 0;
                g.tb[0] = (uint16_t)((TB_START << 8) | (TB_START << 4) | TB_START);
            } else {
                Sp[q] = -1 - (int)i; // gap_open + gap_extend * i
                sbp[q] = TB_INS;
                uint32_t ib = i == 1 ? TB_START : TB_INS;
                g.tb[(size_t)i * stride] = (uint16_t)((TB_INS << 8) | (TB_START << 4) | ib);
            }
        }
    }
    // systolic sweep: at step t lane l handles column j = t - l + 1 (1..n)
    // hand-off registers from the lane above: S, I, sbits of its last row at column j and j-1
    int upS_prev = NEG; // S(r0-1, j-1)
    // for lane 0, "row above" does not exist; its first row is row 0 (boundary), handled inline
    if (lane > 0) {
        uint32_t ia = r0 - 1; // last row of the lane above, column 0
        upS_prev = ia < rows ? (ia == 0 ? 0 : -1 - (int)ia) : NEG;
    }
    int lastS = NEG, lastI = NEG; // this lane's last row at the column it just finished
    uint32_t lastSb = TB_START;
    const uint32_t steps = n + 31;
    for (uint32_t t = 0; t < steps; t++) {
        // values of the lane above for the column this lane is about to do (it did it last step)
        int upS = __shfl_up_sync(FULL, lastS, 1);
        int upI = __shfl_up_sync(FULL, lastI, 1);
        uint32_t upSb = __shfl_up_sync(FULL, lastSb, 1);
        int j = (int)t - lane + 1;
        if (j >= 1 && j <= (int)n && r0 < rows) {
            const uint8_t yc = g.y[j - 1];
            int aboveS = upS, aboveI = upI; // (i-1, j)
            uint32_t aboveSb = upSb;
            int diagS = upS_prev;           // (i-1, j-1)
#pragma unroll
            for (int q = 0; q < GREEDY_RMAX; q++) {
                uint32_t i = r0 + (uint32_t)q;
                if ((uint32_t)q < R && i < rows) {
                    int S, I, D;
                    uint32_t sb, ib, db;
                    if (i == 0) { // boundary row: only deletions
                        D = -1 - j;
                        S = D;
                        I = NEG;
                        sb = TB_DEL;
                        db = j == 1 ? TB_START : TB_DEL;
                        ib = TB_START;
                    } else {
                        int m_score = diagS + (g.x[i - 1] == yc ? 1 : -1);
                        int i_ext = aboveI - 1, i_open = aboveS - 2;
                        if (i_ext > i_open) { I = i_ext; ib = TB_INS; } else { I = i_open; ib = aboveSb; }
                        int d_ext = Dp[q] - 1, d_open = Sp[q] - 2;
                        if (d_ext > d_open) { D = d_ext; db = TB_DEL; } else { D = d_open; db = sbp[q]; }
                        S = m_score;
                        sb = g.x[i - 1] == yc ? TB_MATCH : TB_SUBST;
                        if (I > S) { S = I; sb = TB_INS; }
                        if (D > S) { S = D; sb = TB_DEL; }
                    }
                    g.tb[(size_t)i * stride + (uint32_t)j] = (uint16_t)((sb << 8) | (db << 4) | ib);
                    // next row in this lane sees this cell as "above", and the old Sp as diagonal
                    diagS = Sp[q];
                    aboveS = S;
                    aboveI = I;
                    aboveSb = sb;
                    Sp[q] = S;
                    Dp[q] = D;
                    sbp[q] = sb;
                    lastS = S;
                    lastI = I;
                    lastSb = sb;
                }
            }
            upS_prev = upS;
        }
    }
    __syncwarp();
    // traceback (uniform)
    uint32_t i = m, j = n, nops = 0;
    uint32_t layer = (g.tb[(size_t)i * stride + j] >> 8) & 0xf;
    while (layer != TB_START) {
        uint32_t cell = g.tb[(size_t)i * stride + j];
        uint32_t next;
        uint8_t op;
        if (layer == TB_INS) {
            op = OP_INS;
            next = cell & 0xf;
            i -= 1;
        } else if (layer == TB_DEL) {
            op = OP_DEL;
            next = (cell >> 4) & 0xf;
            j -= 1;
        } else {
            op = layer == TB_MATCH ? OP_MATCH : OP_SUBST;
            i -= 1;
            j -= 1;
            next = (g.tb[(size_t)i * stride + j] >> 8) & 0xf;
        }
        if (lane == 0) g.ops[nops] = op;
        nops++;
        layer = next;
    }
    __syncwarp();
    return nops;
}

// Greedy::match_alignement (greedy.rs:56-89) on the reversed op list produced above
__device__ __forceinline__ bool match_alignement(const GreedyScratch &g, uint32_t nops, uint32_t nbefore, int &off) {
    int offset = 0;
    // forward index f <-> reversed index nops-1-f
    for (uint32_t w = nbefore; w + 1 < nops; w++) {
        uint8_t op0 = g.ops[nops - 1 - w], op1 = g.ops[nops - 2 - w];
        if (op0 == OP_DEL)
            offset -= 1;
        else if (op0 == OP_INS)
            offset += 1;
        if (op0 == OP_MATCH && op1 == OP_MATCH) {
            int offset_corr = 0;
            for (uint32_t e = 0; e < nops; e++) { // operations.iter().rev() == reversed list from index 0
                uint8_t op = g.ops[e];
                if (op == OP_DEL)
                    offset_corr -= 1;
                else if (op == OP_INS)
                    offset_corr += 1;
                else
                    break;
            }
            off = offset - offset_corr;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ Corr greedy_correct_error(Rd &rd, uint64_t kmer, uint32_t i, uint32_t max_search,
                                                     uint32_t nb_validate) {
    Corr res;
    res.some = false;
    res.in_place = true;
    res.n_emit = 0;
    res.codes = 0;
    res.offset = 0;
    res.new_kmer = 0;

    uint32_t alt;
    if (!uniq(alt_mask(rd, kmer), alt)) return res; // greedy.rs:130-134
    const uint8_t *sub = rd.in + i;
    const uint32_t sublen = rd.len - i;
    GreedyScratch g = greedy_scratch(rd.scratch, rd.k, (int)max_search);
    const uint32_t stride = (uint32_t)greedy_dim(rd.k, (int)max_search);
    const uint32_t nb = (uint32_t)rd.k - 1;

    // before_seq = kmer2seq(kmer >> 2, k-1) (greedy.rs:139-141): upper-case bases of the k-1 prefix
    if ((uint32_t)rd.lane < nb) {
        uint8_t b = bit2nuc((uint32_t)((kmer >> (2 * (nb - (uint32_t)rd.lane))) & 3));
        g.x[rd.lane] = b;
        g.y[rd.lane] = b;
    }
    uint64_t x = replace_last(kmer, alt, rd.mask);
    uint32_t npath = 0;
    if (rd.lane == 0) {
        g.y[nb + npath] = bit2nuc(alt);
        g.viewed[0] = x;
    }
    npath++;
    uint32_t nviewed = 1;
    __syncwarp();

    for (uint32_t s = 0; s < max_search; s++) {
        uint32_t a;
        if (uniq(succ_mask(rd, x), a)) { // follow_graph (greedy.rs:91-102)
            x = push(x, a, rd.mask);
            if (rd.lane == 0) g.y[nb + npath] = bit2nuc(a);
            npath++;
        }
        // viewed_kmer.contains(&kmer) (greedy.rs:154-157)
        bool seen = false;
        for (uint32_t e = rd.lane; e < nviewed; e += 32) seen |= g.viewed[e] == x;
        if (__any_sync(FULL, seen)) return res;
        if (rd.lane == 0) g.viewed[nviewed] = x;
        nviewed++;
        if (sublen < s) return res; // greedy.rs:160-162
        // The reference evaluates match_alignement first and check_next_kmers second
        // (greedy.rs:163-165); both are pure and a result needs both, so the cheap one (nb_validate
        // lookups, greedy.rs:104-117) goes first and the alignment runs only when it can matter.
        bool ok = sublen - s >= nb_validate;
        if (ok) {
            bool bad = false;
            for (uint32_t v0 = 0; v0 < nb_validate; v0 += 32) {
                uint32_t v = v0 + (uint32_t)rd.lane;
                if (v < nb_validate) {
                    uint64_t km = push_seq(x, sub + s, v + 1, rd.mask);
                    if (!lookup(rd, km)) bad = true;
                }
            }
            ok = !__any_sync(FULL, bad);
        }
        if (!ok) continue;
        // x side: before || seq[..s]
        if ((uint32_t)rd.lane < s) g.x[nb + rd.lane] = sub[rd.lane];
        for (uint32_t e = 32 + rd.lane; e < s; e += 32) g.x[nb + e] = sub[e];
        __syncwarp();
        uint32_t nops = bio_global(rd, g, nb + s, nb + npath, stride);
        int off;
        if (match_alignement(g, nops, nb, off)) {
            long long o = (long long)npath + (long long)off;
            if (o < 0) o = 0; // unreachable in the reference (would wrap); see SURVEY appendix B.10
            flush_copy(rd, i);
            for (uint32_t e = rd.lane; e < npath; e += 32)
                if (rd.o + e < rd.cap) rd.out[rd.o + e] = g.y[nb + e];
            res.some = true;
            res.n_emit = npath;
            res.new_kmer = x;
            res.offset = (uint32_t)o;
            return res;
        }
    }
    return res;
}

// ------------------------------------------------------------------------------------------
// Corrector::correct (src/correct/mod.rs:53-107) for one read.  The scan is instantiated per
// method so that One/Two do not pay for the registers of Greedy's alignment.
// ------------------------------------------------------------------------------------------
template <int METHOD>
__device__ __forceinline__ Corr correct_error(Rd &rd, const CorrectParams &p, uint64_t kmer, uint32_t i, uint32_t dr) {
    if (METHOD == BRGPU_ONE) return exist_correct_error<3>(rd, kmer, i, (uint32_t)p.confirm);
    if (METHOD == BRGPU_TWO) return exist_correct_error<13>(rd, kmer, i, (uint32_t)p.confirm);
    if (METHOD == BRGPU_GRAPH) return graph_correct_error(rd, kmer, i, dr);
    if (METHOD == BRGPU_GREEDY)
        return greedy_correct_error(rd, kmer, i, (uint32_t)p.max_search, (uint32_t)p.confirm);
    return gap_size_correct_error(rd, kmer, i, dr, (uint32_t)p.confirm);
}

// ------------------------------------------------------------------------------------------
// Segmented, speculative scan.
//
// Corrector::correct is sequential inside a read, and one warp walking a 60 kb read from end
// to end takes as long as the whole rest of the batch (ncu r1g: SMs idle for half of the scan
// kernel while the longest reads finish).  The state the loop carries is small, though: when
// the rolling k-mer consists of input bases only (d == 0) and `previous` equals the bitmap bit
// S[i-1], everything that follows depends only on the position i — call that a *clean visit*.
// A run started at position q in that state reproduces the sequential run's suffix exactly.
//
//   scan_spec_kernel:  every read is cut into segments of SEG positions.  One warp per segment
//       starts at the segment boundary assuming a clean visit there, corrects up to the first
//       clean visit at or after the next boundary (q_exit), writes its output to a scratch
//       region, and records `horizon`: the position of its first successful correction.  Up to
//       and including `horizon` the run has only copied input bytes (failed corrections echo the
//       original base), so it is in a clean visit at every position of [start, horizon].
//   scan_merge_kernel: one warp per read chains the pieces in order.  The sequential run reaches
//       a clean visit at q (initially q = k).  If q <= horizon of the segment containing q, the
//       speculative run of that segment is, from q on, exactly what the sequential run would
//       do: append its output from scratch offset q - start and continue at its q_exit.
//       Otherwise (the previous piece ran ~k positions past the boundary and this segment
//       corrected something right at its start: a few percent of the boundaries) the warp
//       re-runs that one segment from q itself.
//
// The result is byte-identical to the sequential scan by construction; the critical path drops
// from "events of the longest read" to "events of one segment + a copy per segment".
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bm_bit(const Rd &rd, uint32_t p) { return (__ldg(rd.bm + (p >> 5)) >> (p & 31)) & 1u; }

// Run Corrector::correct's loop from a clean visit at `start` until the first clean visit at or
// after `limit` (or the end of the read).  Output goes to rd.out/rd.o; rd.copy_from must be set
// by the caller (start, or 0 for the piece that also carries the first k bases).
template <int METHOD>
__device__ __forceinline__ void correct_segment(Rd &rd, const CorrectParams &p, uint32_t start, uint32_t limit,
                                                uint32_t &q_exit, uint32_t &horizon) {
    const uint32_t k = (uint32_t)rd.k;
    rd.bm_base = 0xffffffffu;
    rd.w_origin = 0xffffffffu;
    horizon = NO_HORIZON;
    uint32_t i = start;
    bool previous = bm_bit(rd, start - 1); // mod.rs:67 for start == k; the clean-visit invariant otherwise
    bool canon = true;  // d == 0 and previous == S[i-1]
    uint32_t d = 0;     // pushes still to come whose k-mer contains corrected bases
    uint64_t kmer = 0;  // rolling k-mer; only maintained while d > 0 or at an event
    if (limit > rd.len) limit = rd.len;

    while (i < rd.len) {
        if (d == 0) {
            if (!canon) {
                // first position after a dirty window: `previous` is the last dirty lookup, which
                // need not equal S[i-1].  Look at this one position with the real `previous`.
                const bool Si = bm_bit(rd, i);
                if (Si || !previous) { // no event here (mod.rs:99-102): from i+1 on the state is canonical
                    previous = Si;
                    i += 1;
                    canon = true;
                    continue;
                }
            } else {
                if (i >= limit) break; // clean visit at or after the nominal end
                uint32_t j = find_transition(rd, i, previous, limit);
                if (j >= limit) {     // nothing fires before the boundary: clean visit at `limit`
                    i = limit;
                    continue;
                }
                i = j;
            }
            // one cooperative load brings in the k-mer, the bases every scenario looks at and the
            // bases of the dirty window that follows a correction
            load_win(rd, i - k + 1);
            kmer = win_extract(rd, i - k + 1, k);
        } else {
            uint32_t n = rd.len - i;
            if (n > d) n = d;
            if (n > 32) n = 32;
            uint64_t km;
            if (win_covers(rd, i, n))
                km = win_push(rd, kmer, i, (uint32_t)rd.lane + 1u <= n ? (uint32_t)rd.lane + 1u : 0u);
            else
                km = push_window(rd, kmer, rd.in + i, n);
            bool s = (uint32_t)rd.lane < n && lookup(rd, km);
            uint32_t gm = __ballot_sync(FULL, s);
            uint32_t vm = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
            uint32_t trig = ~gm & ((gm << 1) | (previous ? 1u : 0u)) & vm; // mod.rs:73
            if (trig == 0) {
                kmer = shfl64(km, (int)n - 1);
                previous = (gm >> (n - 1)) & 1u; // mod.rs:99
                i += n;
                d -= n;
                if (d == 0 && i < rd.len) canon = previous == bm_bit(rd, i - 1);
                continue;
            }
            uint32_t l = (uint32_t)(__ffs(trig) - 1);
            kmer = shfl64(km, (int)l);
            i += l;
            d -= l + 1;
            // keep the bases the scenarios will look at in the register window
            if (!win_covers(rd, i, 32u)) load_win(rd, i - k + 1);
        }
        // solid -> weak transition at input position i; kmer ends with in[i]
        Corr c = correct_error<METHOD>(rd, p, kmer, i, d);
        if (!c.some) { // mod.rs:90-96
            previous = false;
            i += 1;
            // d == 0: S[i-1] is 0 as well when the k-mer was pure input; after a dirty event it is
            // whatever the bitmap says
            if (d == 0 && i < rd.len) canon = !bm_bit(rd, i - 1);
        } else { // mod.rs:74-89
            if (horizon == NO_HORIZON) horizon = i;
            flush_copy(rd, i);
            if (c.in_place) {
                rd.o += c.n_emit;
                kmer = c.new_kmer;
            } else {
                kmer >>= 2;
                for (int e = (int)c.n_emit - 1; e >= 0; e--) {
                    uint32_t code = (c.codes >> (2 * e)) & 3u;
                    kmer = push(kmer, code, rd.mask);
                    emit_byte(rd, bit2nuc(code));
                }
            }
            previous = true;
            i += c.offset;
            rd.copy_from = i;
            d = k - 1;
            canon = false;
            if (d == 0 && i < rd.len) canon = previous == bm_bit(rd, i - 1); // k == 1 cannot happen (k >= 3)
        }
    }
    q_exit = i;
    flush_copy(rd, i < rd.len ? i : rd.len);
}

// ARM: which form of the set the kernel is compiled for.  ARM_COMPACT: the rank-compacted form (SolidView::dir),
// the usual case for a sparse set; ARM_DENSE: summary + bitfield (sets too dense to compact: configs[3] / [4]);
// ARM_ANY: decided at run time (hash sets, k != 17, the other methods).  In the specialised kernels the other
// lookup arms are compiled out — a third of the code of a kernel whose warps stall on instruction fetch as often
// as on memory (profiles/ncu_r2_scan_full.txt); growing the compacted arm by the block bytes cost the dense
// sets 10 % in the shared kernel, which is why they have their own.
constexpr int ARM_ANY = 0, ARM_COMPACT = 1, ARM_DENSE = 2;
__device__ __forceinline__ void specialise_view(SolidView &set, int arm) {
    if (arm == ARM_COMPACT) {
        set.hash = nullptr;
        set.summary = nullptr;
        __builtin_assume(set.dir != nullptr);
    } else if (arm == ARM_DENSE) {
        set.hash = nullptr;
        set.dir = nullptr;
        set.blocks = nullptr;
        set.pos8 = nullptr; // the summary stays a run-time choice: a saturated set (configs[4]) is looked up without it
    }
}
template <int METHOD, int KT, int ARM>
__global__ void __launch_bounds__(SCAN_WARPS_PER_BLOCK * 32, (METHOD == BRGPU_ONE || METHOD == BRGPU_TWO) ? BRGPU_SCAN_MINB : 1)
    scan_spec_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ len_in,
                     const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ bitmap,
                     const uint64_t *__restrict__ seg_first, uint32_t n_reads, uint8_t *__restrict__ seg_out,
                     SegRec *__restrict__ recs, uint32_t *__restrict__ flags, SolidView set, CorrectParams p,
                     uint8_t *scratch, size_t scratch_per_warp, unsigned long long *get_counter) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Rd rd;
    if (KT) p.k = KT; // compile-time k: the k-mer arithmetic below folds to immediates
    specialise_view(set, ARM);
    rd.set = set;
    rd.set.k = p.k;
    rd.k = p.k;
    rd.mask = kmask(p.k);
    rd.lane = lane;
    rd.scratch = scratch ? scratch + (size_t)warp * scratch_per_warp : nullptr;
#if BRGPU_COUNT_GETS
    rd.n_get = 0;
#endif
    const uint64_t n_seg_total = __ldg(seg_first + n_reads);
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(reinterpret_cast<unsigned long long *>(flags + 4), 1ULL);
        g = shfl64(g, 0);
        if (g >= n_seg_total) break;
        // read r with seg_first[r] <= g < seg_first[r+1]
        uint32_t lo = 0, hi = n_reads;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(seg_first + mid) <= g)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t r = lo;
        const uint32_t sidx = (uint32_t)(g - __ldg(seg_first + r));
        const uint64_t base = __ldg(slot_off + r);
        rd.in = in + base;
        rd.len = __ldg(len_in + r);
        rd.bm = bitmap + (base >> 5);
        rd.out = seg_out + g * SEG_CAP;
        rd.cap = SEG_CAP;
        rd.o = 0;
        SegRec rec;
        rec.out_len = 0;
        rec.q_exit = rd.len;
        rec.horizon = NO_HORIZON;
        rec.bad = 1;
        if (rd.len >= (uint32_t)p.k && (rd.len > (uint32_t)p.k || sidx == 0)) {
            const uint32_t start = (uint32_t)p.k + sidx * SEG;
            rd.copy_from = sidx == 0 ? 0u : start; // piece 0 carries the first k bases (mod.rs:62-65)
            if (start < rd.len) {
                correct_segment<METHOD>(rd, p, start, start + SEG, rec.q_exit, rec.horizon);
            } else { // len == k: nothing to scan, the piece is the read itself
                rec.q_exit = rd.len;
                flush_copy(rd, rd.len);
            }
            rec.out_len = rd.o;
            rec.bad = rd.o > SEG_CAP ? 1u : 0u;
        }
        __syncwarp();
        if (lane == 0) recs[g] = rec;
    }
#if BRGPU_COUNT_GETS
    uint32_t gets = __reduce_add_sync(FULL, rd.n_get);
    if (lane == 0 && gets) atomicAdd(get_counter, (unsigned long long)gets);
#endif
}

template <int METHOD, int KT, int ARM>
__global__ void __launch_bounds__(SCAN_WARPS_PER_BLOCK * 32)
    scan_merge_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ len_in, uint8_t *__restrict__ out,
                      uint32_t *__restrict__ len_out, const uint64_t *__restrict__ slot_off,
                      const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ order,
                      const uint64_t *__restrict__ seg_first, uint32_t n_reads, const uint8_t *__restrict__ seg_out,
                      const SegRec *__restrict__ recs, uint32_t *__restrict__ flags, SolidView set, CorrectParams p,
                      uint8_t *scratch, size_t scratch_per_warp, uint8_t *__restrict__ changed,
                      SegCopy *__restrict__ copies, unsigned long long *get_counter) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (KT) p.k = KT; // compile-time k: the k-mer arithmetic below folds to immediates
    const uint32_t k = (uint32_t)p.k;
    specialise_view(set, ARM);
    Rd rd;
    rd.set = set;
    rd.set.k = p.k;
    rd.k = p.k;
    rd.mask = kmask(p.k);
    rd.lane = lane;
    rd.scratch = scratch ? scratch + (size_t)warp * scratch_per_warp : nullptr;
#if BRGPU_COUNT_GETS
    rd.n_get = 0;
#endif
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(flags + 0, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_reads) break;
        const uint32_t r = __ldg(order + qi);
        const uint64_t base = __ldg(slot_off + r);
        rd.in = in + base;
        rd.out = out + base;
        rd.cap = (uint32_t)(__ldg(slot_off + r + 1) - base);
        rd.len = __ldg(len_in + r);
        rd.bm = bitmap + (base >> 5);
        rd.o = 0;
        rd.copy_from = 0;
        bool edited = false; // a correction succeeded somewhere in this read
        if (rd.len < k) { // mod.rs:56-58
            flush_copy(rd, rd.len);
        } else {
            const uint64_t g0 = __ldg(seg_first + r);
            uint32_t q = k; // the sequential run's first clean visit (mod.rs:60-67)
            bool first = true;
            for (;;) {
                const uint32_t sidx = q < rd.len ? (q - k) / SEG : 0u;
                if (!first && q >= rd.len) break;
                const SegRec rec = recs[g0 + sidx];
                const uint32_t seg_start = k + sidx * SEG;
                if (!rec.bad && q <= rec.horizon) {
                    // the speculative run of this segment is in a clean visit at q: splice it in
                    const uint32_t in0 = sidx == 0 ? 0u : seg_start;     // input position of scratch byte 0
                    const uint32_t skip = first ? 0u : q - in0;          // 1:1 copy region before q
                    const uint32_t n = rec.out_len - skip;
                    if (lane == 0) { // the bytes are moved by scan_splice_kernel; what fits if the slot overflows
                        SegCopy cp;
                        cp.dst = base + rd.o;
                        cp.skip = skip;
                        cp.n = rd.o >= rd.cap ? 0u : (rd.o + n <= rd.cap ? n : rd.cap - rd.o);
                        copies[g0 + sidx] = cp;
                    }
                    rd.o += n;
                    q = rec.q_exit;
                    edited |= rec.horizon != NO_HORIZON; // q <= horizon: the correction is inside the spliced part
                } else {
                    // re-run this segment from the true state
                    rd.copy_from = first ? 0u : q;
                    uint32_t q_exit, horizon;
                    if (q < rd.len) {
                        correct_segment<METHOD>(rd, p, q, seg_start + SEG, q_exit, horizon);
                        edited |= horizon != NO_HORIZON;
                    } else {
                        q_exit = rd.len;
                        flush_copy(rd, rd.len);
                    }
                    q = q_exit;
                }
                first = false;
                if (q >= rd.len) break;
            }
        }
        __syncwarp();
        if (lane == 0) {
            len_out[r] = rd.o;
            changed[r] = edited ? 1 : 0;
            if (rd.o > rd.cap) atomicOr(flags + 1, 1u);
        }
    }
#if BRGPU_COUNT_GETS
    uint32_t gets = __reduce_add_sync(FULL, rd.n_get);
    if (lane == 0 && gets) atomicAdd(get_counter, (unsigned long long)gets);
#endif
}


// ------------------------------------------------------------------------------------------
// Four segments per warp (One / Two).
//
// A warp that owns one segment runs ~700 warp instructions per event with most lanes idle:
// the alternatives take 4 lanes, One's scenario items 21, the dirty window 16, and everything
// between the lookup rounds is scalar bookkeeping; once the lookups had moved into L2 the forward
// scans were bound by instruction issue (profiles/ncu_r1q.txt).  Here a warp is four groups of 8
// lanes; a group owns a segment and walks it exactly like correct_segment does (same state, same
// order, so the pieces and records are interchangeable), and the four groups step through their
// events together: find the next event, load the window, alternatives, scenario rounds, apply.
// Collectives use the group's lane mask, so a group may leave a phase early (no unique
// alternative, end of segment) and rejoins the others at the next structured merge point.
// ------------------------------------------------------------------------------------------
struct G8 {
    // lane geometry
    int gl;           // lane inside the group (0..7)
    uint32_t gbase;   // first lane of the group inside the warp
    uint32_t gmask;   // lanes of the group
    // segment (uniform inside the group)
    const uint8_t *in;
    uint32_t len;
    uint8_t *out;
    const uint32_t *bm;
    uint32_t o, copy_from;
    uint64_t mask;
    int k;
    uint64_t w0, w1;
    uint32_t w_origin;
#if BRGPU_COUNT_GETS
    uint32_t n_get; // per lane
#endif
};

__device__ __forceinline__ uint32_t g_ballot(const G8 &g, bool p) { return (__ballot_sync(g.gmask, p) >> g.gbase) & 0xffu; }
__device__ __forceinline__ uint32_t g_shfl(const G8 &g, uint32_t v, int src) { return __shfl_sync(g.gmask, v, src, 8); }
__device__ __forceinline__ uint64_t g_shfl64(const G8 &g, uint64_t v, int src) {
    uint32_t lo = __shfl_sync(g.gmask, (uint32_t)v, src, 8);
    uint32_t hi = __shfl_sync(g.gmask, (uint32_t)(v >> 32), src, 8);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t g_or(const G8 &g, uint32_t v) {
    v |= __shfl_xor_sync(g.gmask, v, 1, 8);
    v |= __shfl_xor_sync(g.gmask, v, 2, 8);
    v |= __shfl_xor_sync(g.gmask, v, 4, 8);
    return v;
}
__device__ __forceinline__ bool g_lookup(G8 &g, const SolidView &set, uint64_t kmer) {
#if BRGPU_COUNT_GETS
    g.n_get++;
#endif
    return solid(set, kmer);
}
// N lookups per lane with all their loads in flight together: first the N directory words (or
// summary words), then the N blocks (or bitfield bytes) of those that are occupied.  A lane's
// rounds are latency chains of L2 round trips, so batching is what keeps 8-lane groups from
// paying one round trip per 8 items.
template <int N>
__device__ __forceinline__ void g_lookup_n(G8 &g, const SolidView &set, const uint64_t (&km)[N], const bool (&want)[N],
                                           bool (&out)[N]) {
    if (set.hash) { // set::Hash: probe the table, one item after the other
#pragma unroll
        for (int t = 0; t < N; t++) {
#if BRGPU_COUNT_GETS
            if (want[t]) g.n_get++;
#endif
            out[t] = want[t] && hash_contains(set.hash, set.hash_mask, canonical_kmer(km[t], set.k));
        }
        return;
    }
    uint64_t idx[N];
#pragma unroll
    for (int t = 0; t < N; t++) {
        idx[t] = canonical_index(km[t], set.k);
#if BRGPU_COUNT_GETS
        if (want[t]) g.n_get++;
#endif
    }
    if (set.dir) {
        uint2 e[N];
#pragma unroll
        for (int t = 0; t < N; t++) {
            e[t] = make_uint2(0u, 0u);
            if (want[t]) e[t] = set_ld(set.dir + (idx[t] >> 11));
        }
        uint32_t r[N], pv[N]; // rank of the block among the occupied ones; its byte (kmer.cuh: SolidView::pos8)
#pragma unroll
        for (int t = 0; t < N; t++) {
            const uint32_t b = (uint32_t)(idx[t] >> 6) & 31u;
            r[t] = e[t].y + __popc(e[t].x & ((1u << b) - 1u));
            pv[t] = POS8_NONE;
            if ((e[t].x >> b) & 1u) pv[t] = set.pos8 ? set_ld(set.pos8 + r[t]) : POS8_MULTI;
        }
#pragma unroll
        for (int t = 0; t < N; t++) {
            out[t] = pv[t] == (uint32_t)(idx[t] & 63);
            if (pv[t] == POS8_MULTI) out[t] = (set_ld(set.blocks + r[t]) >> (idx[t] & 63)) & 1ULL;
        }
        return;
    }
    bool go[N];
#pragma unroll
    for (int t = 0; t < N; t++) go[t] = want[t];
    if (set.summary) {
        uint32_t sw[N];
#pragma unroll
        for (int t = 0; t < N; t++) {
            sw[t] = 0;
            if (go[t]) sw[t] = __ldg(set.summary + (idx[t] >> (set.shift + 5)));
        }
#pragma unroll
        for (int t = 0; t < N; t++) go[t] = (sw[t] >> ((idx[t] >> set.shift) & 31)) & 1u;
    }
    uint32_t byte[N];
#pragma unroll
    for (int t = 0; t < N; t++) {
        byte[t] = 0;
        if (go[t]) byte[t] = __ldg(set.bits + (idx[t] >> 3));
    }
#pragma unroll
    for (int t = 0; t < N; t++) out[t] = (byte[t] >> (idx[t] & 7)) & 1u;
}

__device__ __forceinline__ bool g_bm_bit(const G8 &g, uint32_t p) { return (__ldg(g.bm + (p >> 5)) >> (p & 31)) & 1u; }

// 64 input bases from `origin` on, packed into (w0, w1) on every lane of the group: lane l reads
// bytes [8l, 8l + 8) with three aligned word loads and a funnel shift, packs them to 16 bits, and
// a 3-step butterfly concatenates the eight pieces.  Codes of positions >= len are zero.
__device__ __forceinline__ void g_load_win(G8 &g, uint32_t origin) {
    const uint8_t *p = g.in + origin + 8u * (uint32_t)g.gl;
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p - a);
    const uint32_t first = origin + 8u * (uint32_t)g.gl;
    const uint32_t nv = first >= g.len ? 0u : (g.len - first < 8u ? g.len - first : 8u);
    // a lane whose 8 bytes start inside the read touches at most 11 bytes past its first one, all
    // inside the slot's slack (>= 64 bytes); lanes past the end of the read do not load at all
    uint32_t x0 = 0, x1 = 0, x2 = 0;
    if (nv) {
        x0 = __ldg(q);
        x1 = __ldg(q + 1);
        if (a) x2 = __ldg(q + 2);
    }
    const uint32_t b0 = __funnelshift_r(x0, x1, 8 * a), b1 = __funnelshift_r(x1, x2, 8 * a);
    uint32_t v16 = (pack4(b0) << 8) | pack4(b1);
    v16 &= (0xffffu << (16u - 2u * nv)) & 0xffffu;
    const uint32_t o16 = __shfl_xor_sync(g.gmask, v16, 1, 8);
    const uint32_t t32 = (g.gl & 1) ? ((o16 << 16) | v16) : ((v16 << 16) | o16);
    const uint32_t o32 = __shfl_xor_sync(g.gmask, t32, 2, 8);
    const uint64_t t64 = (g.gl & 2) ? (((uint64_t)o32 << 32) | t32) : (((uint64_t)t32 << 32) | o32);
    const uint32_t olo = __shfl_xor_sync(g.gmask, (uint32_t)t64, 4, 8);
    const uint32_t ohi = __shfl_xor_sync(g.gmask, (uint32_t)(t64 >> 32), 4, 8);
    const uint64_t o64 = ((uint64_t)ohi << 32) | olo;
    g.w0 = (g.gl & 4) ? o64 : t64;
    g.w1 = (g.gl & 4) ? t64 : o64;
    g.w_origin = origin;
}

// group copy, arbitrary alignment on both sides (see warp_copy); 8 lanes, two word pairs in flight
__device__ __forceinline__ void g_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int gl) {
    if (n < 16) {
        for (uint32_t t = gl; t < n; t += 8) dst[t] = src[t];
        return;
    }
    const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if ((uint32_t)gl < head) dst[gl] = src[gl];
    const uint8_t *s0 = src + head;
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t n_words = (n - head) >> 2;
    const uint32_t a = (uint32_t)((uintptr_t)s0 & 3u);
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s0 - a);
    uint32_t w = gl;
    for (; w + 8 < n_words; w += 16) {
        const uint32_t l0 = s4[w], l1 = s4[w + 8];
        const uint32_t h0 = a ? s4[w + 1] : 0u, h1 = a ? s4[w + 9] : 0u;
        d4[w] = a ? __funnelshift_r(l0, h0, 8 * a) : l0;
        d4[w + 8] = a ? __funnelshift_r(l1, h1, 8 * a) : l1;
    }
    for (; w < n_words; w += 8) {
        const uint32_t lo = s4[w];
        d4[w] = a ? __funnelshift_r(lo, s4[w + 1], 8 * a) : lo;
    }
    const uint32_t done = head + (n_words << 2);
    if (done + (uint32_t)gl < n) dst[done + gl] = src[done + gl];
}

__device__ __forceinline__ void g_flush_copy(G8 &g, uint32_t upto) {
    if (upto > g.len) upto = g.len;
    if (upto <= g.copy_from) return;
    const uint32_t n = upto - g.copy_from;
    if (g.o + n <= SEG_CAP) {
        g_copy(g.out + g.o, g.in + g.copy_from, n, g.gl);
    } else { // the scratch region overflows: keep counting, write what fits (the piece is marked bad)
        for (uint32_t t = g.gl; t < n; t += 8)
            if (g.o + t < SEG_CAP) g.out[g.o + t] = g.in[g.copy_from + t];
    }
    g.o += n;
    g.copy_from = upto;
}

// find_transition for a group: 256 bitmap positions per step
__device__ __forceinline__ uint32_t g_find_transition(const G8 &g, uint32_t i, bool previous, uint32_t end) {
    if (i >= end) return end;
    uint32_t wbase = i >> 5;
    const uint32_t n_words = (end + 31) >> 5;
    const uint32_t n_words_read = (g.len + 31) >> 5;
    uint32_t carry_in = 0;
    for (;;) {
        const uint32_t wi = wbase + (uint32_t)g.gl;
        const uint32_t W = wi < n_words_read ? __ldg(g.bm + wi) : 0u;
        const uint32_t up = __shfl_up_sync(g.gmask, W, 1, 8);
        const uint32_t carry = g.gl ? (up >> 31) : carry_in;
        uint32_t T = ~W & ((W << 1) | carry);
        const uint32_t posbase = wi << 5;
        if (posbase + 31 < i) {
            T = 0;
        } else if (posbase <= i) {
            const uint32_t sh = i - posbase;
            T &= (0xffffffffu << sh);
            T &= ~(1u << sh);
            if (previous && !((W >> sh) & 1u)) T |= 1u << sh;
        }
        if (posbase >= end)
            T = 0;
        else if (posbase + 32 > end)
            T &= (1u << (end - posbase)) - 1u;
        const uint32_t any = g_ballot(g, T != 0);
        if (any) {
            const int fl = __ffs(any) - 1;
            const uint32_t Tf = g_shfl(g, T, fl);
            return ((wbase + (uint32_t)fl) << 5) + (uint32_t)(__ffs(Tf) - 1);
        }
        wbase += 8;
        if (wbase >= n_words) return end;
        carry_in = g_shfl(g, W, 7) >> 31;
        i = wbase << 5;
        previous = carry_in != 0;
    }
}

// Exist<S>::correct_error for a group (see exist_correct_error for the 32-lane version and the
// references into exist/mod.rs, one.rs, two.rs): same rounds, 8 items per round.
template <int NS>
__device__ __forceinline__ Corr g_exist_correct_error(G8 &g, const SolidView &set, uint64_t kmer, uint32_t i, uint32_t c) {
    Corr res;
    res.some = false;
    res.in_place = false;
    res.n_emit = 0;
    res.codes = 0;
    res.offset = 0;
    res.new_kmer = 0;
    const uint64_t mask = g.mask;

    bool sa = false;
    if (g.gl < 4) sa = g_lookup(g, set, replace_last(kmer, (uint32_t)g.gl, mask));
    uint32_t alt;
    if (!uniq(g_ballot(g, sa) & 0xfu, alt)) return res; // exist/mod.rs:121-126
    const uint64_t K0 = replace_last(kmer, alt, mask);
    const uint8_t *sub = g.in + i;
    const uint32_t sublen = g.len - i;
    const bool use_win = c <= 30u && win_covers(g, i, c + 6u);
    auto sub_push = [&](uint64_t km, uint32_t from, uint32_t n) -> uint64_t {
        return use_win ? win_push(g, km, i + from, n) : push_seq(km, sub + from, n, mask);
    };

    uint32_t bad = 0, more = 0, cand;
    Scen win;
    if (NS == 3) {
        uint32_t short_mask = 0;
#pragma unroll
        for (int s = 0; s < 3; s++)
            if (scen_one(s, K0).offa + c > sublen) short_mask |= 1u << s;
        const uint32_t per = c + 2, Q = 3u * per;
        if (Q <= 24u) {
            // the usual confirm values: three items per lane, their lookups in flight together
            uint64_t km[3];
            bool want[3], hit[3], is_more[3];
            int sc_of[3];
            const uint32_t inv_per = 0xffffffffu / per + 1u; // q / per by multiplication (per >= 3, q < 24)
#pragma unroll
            for (int h = 0; h < 3; h++) {
                const uint32_t q = (uint32_t)g.gl + 8u * (uint32_t)h;
                const int s = (int)__umulhi(q, inv_per);
                const uint32_t u = q - (uint32_t)s * per;
                const Scen t = scen_one(s, K0);
                sc_of[h] = s;
                want[h] = q < Q && !((short_mask >> s) & 1u);
                is_more[h] = u > c;
                km[h] = 0;
                if (want[h]) {
                    if (u <= c) {
                        km[h] = sub_push(t.K, t.offa, u);
                    } else if (sublen > c + t.offc + 1) {
                        km[h] = sub_push(push(K0 >> 2, t.codes & 3u, mask), t.offc, c + 1);
                    } else {
                        want[h] = false; // one_more is false without a lookup (exist/mod.rs:52)
                    }
                }
            }
            g_lookup_n<3>(g, set, km, want, hit);
#pragma unroll
            for (int h = 0; h < 3; h++) {
                if (!want[h]) continue;
                if (is_more[h]) {
                    if (hit[h]) more |= 1u << sc_of[h];
                } else if (!hit[h]) {
                    bad |= 1u << sc_of[h];
                }
            }
        } else {
            for (uint32_t q0 = 0; q0 < Q; q0 += 8) {
                const uint32_t q = q0 + (uint32_t)g.gl;
                const int s = (int)(q / per);
                const uint32_t u = q - (uint32_t)s * per;
                const Scen t = scen_one(s, K0);
                if (q >= Q || ((short_mask >> s) & 1u)) continue;
                if (u <= c) {
                    if (!g_lookup(g, set, sub_push(t.K, t.offa, u))) bad |= 1u << s;
                } else if (sublen > c + t.offc + 1) {
                    const uint64_t km = push(K0 >> 2, t.codes & 3u, mask);
                    if (g_lookup(g, set, sub_push(km, t.offc, c + 1))) more |= 1u << s;
                }
            }
        }
        bad = g_or(g, bad);
        more = g_or(g, more);
        cand = 7u & ~short_mask & ~bad;
        if (cand == 0) return res;
        if (__popc(cand) > 1) {
            cand &= more;
            if (__popc(cand) != 1) return res;
        }
        win = scen_one(__ffs(cand) - 1, K0);
    } else {
        uint32_t sb[4];
        if (use_win) {
            const uint32_t four = (uint32_t)win_extract(g, i, 4);
#pragma unroll
            for (int t = 0; t < 4; t++) sb[t] = (four >> (2 * (3 - t))) & 3u;
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) sb[t] = (uint32_t)t < sublen ? nuc2bit(sub[t]) : 0u;
        }
        // round 2: the four successor sets, 16 lookups = two per lane, in flight together
        uint32_t m16 = 0;
        {
            uint64_t km[2];
            bool want[2], hit[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int it = g.gl + 8 * h, grp = it >> 2;
                uint64_t base = K0;
                want[h] = true;
                if (grp == 1) { base = push(K0, sb[1], mask); want[h] = sublen >= 2; }
                if (grp == 2) { base = push(K0, sb[2], mask); want[h] = sublen >= 3; }
                if (grp == 3) { base = push(K0, sb[0], mask); }
                km[h] = push(base, (uint32_t)(it & 3), mask);
            }
            g_lookup_n<2>(g, set, km, want, hit);
#pragma unroll
            for (int h = 0; h < 2; h++) m16 |= g_ballot(g, want[h] && hit[h]) << (8 * h);
        }
        const uint32_t N0 = m16 & 0xf, N1 = (m16 >> 4) & 0xf, N2 = (m16 >> 8) & 0xf, N0p = (m16 >> 12) & 0xf;
        // The 13 scenarios are evaluated once: lane gl owns scenarios gl and gl + 8 and keeps their k-mer and
        // their four small fields (offa | offc << 8 | n_emit << 16 | codes << 24); whoever needs scenario s
        // later takes it from lane s & 7 by shuffle instead of running scen_two again.
        uint64_t myK[2];
        uint32_t myMeta[2];
        bool myValid[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int s = g.gl + 8 * h;
            myK[h] = 0;
            myMeta[h] = 0;
            myValid[h] = false;
            if (s < NS) {
                const Scen t = scen_two(s, K0, sublen, sb, N0, N1, N2, N0p, mask);
                myK[h] = t.K;
                myMeta[h] = t.offa | (t.offc << 8) | (t.n_emit << 16) | (t.codes << 24);
                myValid[h] = t.valid && !(t.offa + c > sublen); // get_score's length test (exist/mod.rs:29-31)
            }
        }
        auto scen_K = [&](int s) -> uint64_t { // all lanes of the group take part
            const uint64_t a = g_shfl64(g, myK[0], s & 7), b = g_shfl64(g, myK[1], s & 7);
            return (s >> 3) ? b : a;
        };
        auto scen_meta = [&](int s) -> uint32_t {
            const uint32_t a = g_shfl(g, myMeta[0], s & 7), b = g_shfl(g, myMeta[1], s & 7);
            return (s >> 3) ? b : a;
        };
        // round 3a: get(K) of every valid scenario that is long enough, two scenarios per lane
        cand = 0;
        {
            bool hit[2];
            g_lookup_n<2>(g, set, myK, myValid, hit);
#pragma unroll
            for (int h = 0; h < 2; h++) cand |= g_ballot(g, myValid[h] && hit[h]) << (8 * h);
        }
        // round 3b: the c confirmations of the survivors.  Survivor r is scenario (ids >> 4r) & 15: every
        // surviving scenario's owner contributes its id at its rank among the survivors.
        const uint32_t n_alive = (uint32_t)__popc(cand);
        const uint32_t Q = n_alive * c;
        uint32_t ids_lo = 0, ids_hi = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t s = (uint32_t)g.gl + 8u * (uint32_t)h;
            if ((cand >> s) & 1u) {
                const uint32_t r = (uint32_t)__popc(cand & ((1u << s) - 1u));
                if (r < 8u) ids_lo |= s << (4u * r);
                else ids_hi |= s << (4u * (r - 8u));
            }
        }
        ids_lo = g_or(g, ids_lo);
        ids_hi = g_or(g, ids_hi);
        const uint32_t inv_c = c > 1 ? 0xffffffffu / c + 1u : 0u; // q / c by multiplication (q < 2^16)
        for (uint32_t q0 = 0; q0 < Q; q0 += 24) {
            uint64_t km[3];
            bool want[3], hit[3];
            int sc_of[3];
#pragma unroll
            for (int h = 0; h < 3; h++) {
                const uint32_t q = q0 + (uint32_t)g.gl + 8u * (uint32_t)h;
                want[h] = q < Q;
                const uint32_t r = want[h] ? (c > 1 ? __umulhi(q, inv_c) : q) : 0u, u = q - r * c + 1u;
                const int s = (int)(((r < 8u ? ids_lo >> (4u * r) : ids_hi >> (4u * (r - 8u)))) & 0xfu);
                const uint64_t K = scen_K(s);
                const uint32_t offa = scen_meta(s) & 0xffu;
                sc_of[h] = s;
                km[h] = want[h] ? sub_push(K, offa, u) : 0ULL;
            }
            g_lookup_n<3>(g, set, km, want, hit);
#pragma unroll
            for (int h = 0; h < 3; h++)
                if (want[h] && !hit[h]) bad |= 1u << sc_of[h];
        }
        cand &= ~g_or(g, bad);
        if (cand == 0) return res;
        // round 3c, only on a tie: one_more of the tied scenarios, asked by their owners
        if (__popc(cand) > 1) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int s = g.gl + 8 * h;
                bool m = false;
                if (s < NS && ((cand >> s) & 1u)) {
                    const uint32_t offc = (myMeta[h] >> 8) & 0xffu, n_emit = (myMeta[h] >> 16) & 0xffu, codes = myMeta[h] >> 24;
                    if (sublen > c + offc + 1) {
                        uint64_t km = K0 >> 2;
                        for (int e = (int)n_emit - 1; e >= 0; e--) km = push(km, (codes >> (2 * e)) & 3u, mask);
                        m = g_lookup(g, set, sub_push(km, offc, c + 1));
                    }
                }
                more |= g_ballot(g, m) << (8 * h);
            }
            cand &= more;
            if (__popc(cand) != 1) return res;
        }
        {
            const uint32_t meta = scen_meta(__ffs(cand) - 1);
            win.valid = true;
            win.K = 0;
            win.offa = meta & 0xffu;
            win.offc = (meta >> 8) & 0xffu;
            win.n_emit = (meta >> 16) & 0xffu;
            win.codes = meta >> 24;
        }
    }
    res.some = true;
    res.n_emit = win.n_emit;
    res.codes = win.codes;
    res.offset = win.offc;
    return res;
}

template <int METHOD, int KT, int ARM>
__global__ void __launch_bounds__(SCAN_WARPS_PER_BLOCK * 32, BRGPU_SCAN8_MINB)
    scan_spec8_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ len_in,
                      const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ bitmap,
                      const uint64_t *__restrict__ seg_first, uint32_t n_reads, uint8_t *__restrict__ seg_out,
                      SegRec *__restrict__ recs, uint32_t *__restrict__ flags, SolidView set, CorrectParams p,
                      unsigned long long *get_counter) {
    constexpr int NS = METHOD == BRGPU_ONE ? 3 : 13;
    if (KT) p.k = KT;
    set.k = p.k;
    specialise_view(set, ARM);
    const uint32_t k = (uint32_t)p.k, c = (uint32_t)p.confirm;
    G8 g;
    g.gl = threadIdx.x & 7;
    g.gbase = threadIdx.x & 24;
    g.gmask = 0xffu << g.gbase;
    g.mask = kmask(p.k);
    g.k = p.k;
#if BRGPU_COUNT_GETS
    g.n_get = 0;
#endif
    g.in = in;
    g.out = seg_out;
    g.bm = bitmap;
    g.len = 0;
    g.o = g.copy_from = 0;
    g.w0 = g.w1 = 0;
    g.w_origin = 0xffffffffu;
    const uint64_t n_seg_total = __ldg(seg_first + n_reads);

    // state of the group's piece (correct_segment's locals), uniform inside the group
    bool active = false, exhausted = false;
    unsigned long long sg = 0;
    uint32_t i = 0, limit = 0, d = 0, horizon = NO_HORIZON;
    bool previous = false, canon = true;
    uint64_t kmer = 0;

    // One trip of this loop = every group of the warp handles one event of its piece: the four
    // groups advance to their next events in lockstep rounds, then process them side by side.
    // The warp-wide votes at the loop heads re-converge the groups.
    for (;;) {
        // ---- an idle group takes a segment
        if (!active && !exhausted) {
            if (g.gl == 0) sg = atomicAdd(reinterpret_cast<unsigned long long *>(flags + 4), 1ULL);
            sg = g_shfl64(g, sg, 0);
            if (sg >= n_seg_total) {
                exhausted = true;
            } else {
                uint32_t lo = 0, hi = n_reads;
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(seg_first + mid) <= sg)
                        lo = mid;
                    else
                        hi = mid;
                }
                const uint32_t r = lo;
                const uint32_t sidx = (uint32_t)(sg - __ldg(seg_first + r));
                const uint64_t base = __ldg(slot_off + r);
                g.in = in + base;
                g.len = __ldg(len_in + r);
                g.bm = bitmap + (base >> 5);
                g.out = seg_out + sg * SEG_CAP;
                g.o = 0;
                g.w_origin = 0xffffffffu;
                const uint32_t start = k + sidx * SEG;
                if (g.len >= k && (g.len > k || sidx == 0) && start < g.len) {
                    g.copy_from = sidx == 0 ? 0u : start; // piece 0 carries the first k bases (mod.rs:62-65)
                    i = start;
                    limit = start + SEG < g.len ? start + SEG : g.len;
                    previous = g_bm_bit(g, start - 1);
                    canon = true;
                    d = 0;
                    kmer = 0;
                    horizon = NO_HORIZON;
                    active = true;
                } else {
                    // nothing to scan: a piece of a read shorter than k (unusable, the merge copies
                    // the read), or the read of exactly k bases (the piece is the read itself)
                    SegRec rec;
                    rec.out_len = 0;
                    rec.q_exit = g.len;
                    rec.horizon = NO_HORIZON;
                    rec.bad = 1;
                    if (g.len >= k && (g.len > k || sidx == 0)) {
                        g.copy_from = 0;
                        g_flush_copy(g, g.len);
                        rec.out_len = g.o;
                        rec.bad = g.o > SEG_CAP ? 1u : 0u;
                    }
                    __syncwarp(g.gmask);
                    if (g.gl == 0) recs[sg] = rec;
                }
            }
        }
        if (__all_sync(FULL, !active && exhausted)) break;

        // ---- advance every active group to its next event (or to the end of its piece)
        bool at_event = false, clean_event = false, done = false;
        bool adv = active;
        while (__any_sync(FULL, adv)) {
            if (adv && i >= g.len) {
                done = true;
                adv = false;
            }
            if (adv && d > 0) { // k-mers that still contain corrected bases: real lookups, 16 per round
                uint32_t n = g.len - i;
                if (n > d) n = d;
                if (n > 16) n = 16;
                if (!win_covers(g, i, n)) g_load_win(g, i >= k - 1 ? i - k + 1 : 0u);
                const bool in0 = (uint32_t)g.gl < n, in1 = (uint32_t)g.gl + 8u < n;
                const uint64_t km0 = win_push(g, kmer, i, in0 ? (uint32_t)g.gl + 1u : 0u);
                const uint64_t km1 = win_push(g, kmer, i, in1 ? (uint32_t)g.gl + 9u : 0u);
                const uint64_t kmd[2] = {km0, km1};
                const bool wantd[2] = {in0, in1};
                bool hitd[2];
                g_lookup_n<2>(g, set, kmd, wantd, hitd);
                const uint32_t gm = g_ballot(g, in0 && hitd[0]) | (g_ballot(g, in1 && hitd[1]) << 8);
                const uint32_t vm = (1u << n) - 1u;
                const uint32_t trig = ~gm & ((gm << 1) | (previous ? 1u : 0u)) & vm; // mod.rs:73
                if (trig == 0) {
                    const uint32_t last = n - 1;
                    kmer = g_shfl64(g, last < 8 ? km0 : km1, (int)(last & 7));
                    previous = (gm >> last) & 1u; // mod.rs:99
                    i += n;
                    d -= n;
                    if (d == 0 && i < g.len) canon = previous == g_bm_bit(g, i - 1);
                } else {
                    const uint32_t l = (uint32_t)(__ffs(trig) - 1);
                    kmer = g_shfl64(g, l < 8 ? km0 : km1, (int)(l & 7));
                    i += l;
                    d -= l + 1;
                    if (!win_covers(g, i, 32u)) g_load_win(g, i - k + 1);
                    at_event = true;
                    adv = false;
                }
            }
            if (adv && i >= g.len) {
                done = true;
                adv = false;
            }
            if (adv && d == 0 && !canon) {
                // first position after a dirty window: `previous` is the last dirty lookup
                const bool Si = g_bm_bit(g, i);
                if (Si || !previous) { // no event here (mod.rs:99-102)
                    previous = Si;
                    i += 1;
                    canon = true;
                } else {
                    at_event = clean_event = true;
                    adv = false;
                }
            }
            if (adv && i >= g.len) {
                done = true;
                adv = false;
            }
            if (adv && d == 0 && canon) {
                uint32_t j = limit;
                if (i < limit) j = g_find_transition(g, i, previous, limit);
                if (j >= limit) { // clean visit at or after the nominal end
                    if (i < limit) i = limit;
                    done = true;
                } else {
                    i = j;
                    at_event = clean_event = true;
                }
                adv = false;
            }
        }

        // ---- pieces that ended
        if (done) {
            SegRec rec;
            rec.q_exit = i;
            g_flush_copy(g, i < g.len ? i : g.len);
            rec.horizon = horizon;
            rec.out_len = g.o;
            rec.bad = g.o > SEG_CAP ? 1u : 0u;
            __syncwarp(g.gmask);
            if (g.gl == 0) recs[sg] = rec;
            active = false;
        }

        // ---- events: solid -> weak transition at input position i; kmer ends with in[i]
        if (at_event) {
            if (clean_event) {
                g_load_win(g, i - k + 1);
                kmer = win_extract(g, i - k + 1, k);
            }
            const Corr cr = g_exist_correct_error<NS>(g, set, kmer, i, c);
            if (!cr.some) { // mod.rs:90-96
                previous = false;
                i += 1;
                if (d == 0 && i < g.len) canon = !g_bm_bit(g, i - 1);
            } else { // mod.rs:74-89
                if (horizon == NO_HORIZON) horizon = i;
                g_flush_copy(g, i);
                kmer >>= 2;
                for (int e = (int)cr.n_emit - 1; e >= 0; e--) {
                    const uint32_t code = (cr.codes >> (2 * e)) & 3u;
                    kmer = push(kmer, code, g.mask);
                    if (g.gl == 0 && g.o < SEG_CAP) g.out[g.o] = bit2nuc(code);
                    g.o += 1;
                }
                previous = true;
                i += cr.offset;
                g.copy_from = i;
                d = k - 1;
                canon = false;
            }
        }
    }
#if BRGPU_COUNT_GETS
    uint32_t gets = __reduce_add_sync(FULL, g.n_get);
    if ((threadIdx.x & 31) == 0 && gets) atomicAdd(get_counter, (unsigned long long)gets);
#endif
}

template <class K> static int occupancy_warps(brgpu_ctx *ctx, K kernel) {
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, SCAN_WARPS_PER_BLOCK * 32, 0) !=
            cudaSuccess ||
        blocks_per_sm < 1) {
        cudaGetLastError();
        blocks_per_sm = 4;
    }
    return ctx->sm_count * blocks_per_sm * SCAN_WARPS_PER_BLOCK;
}

// One pass of method M over all reads: speculative per-segment scan, per-read merge, parallel splice.
// the specialised kernels are only instantiated where they are launched (k = 17, the k of every BASELINE config)
template <int M, int KT> constexpr int special_arm(int arm) {
#ifdef BRGPU_SPECIAL_ONE_TWO_ONLY // A/B builds: Graph / Greedy / GapSize through the run-time dispatching kernel
    return (KT == 17 && (M == BRGPU_ONE || M == BRGPU_TWO)) ? arm : ARM_ANY;
#else
    return KT == 17 ? arm : ARM_ANY;
#endif
}

template <int M, int KT> static void launch_scan_method(const ScanArgs &a) {
    brgpu_ctx *ctx = a.ctx;
    const Layout &L = *a.L;
    const ScanWork &w = *a.w;
    const unsigned threads = SCAN_WARPS_PER_BLOCK * 32;
    auto grid_for_warps = [&](uint64_t resident, uint64_t items) {
        if (resident > (uint64_t)a.n_warps_total) resident = (uint64_t)a.n_warps_total;
        uint64_t warps = items < resident ? items : resident;
        if (warps < 1) warps = 1;
        return (unsigned)((warps + SCAN_WARPS_PER_BLOCK - 1) / SCAN_WARPS_PER_BLOCK);
    };
    {
        ProfScope ps(ctx, a.spec_name, a.n_bases_hint * 2.0);
        unsigned long long *gc = prof_counter_slot(ctx);
        // One runs four segments per warp (1.67 -> 1.44 ms per launch on the E. coli config); for Two
        // the same scheme measured slower than a warp per segment (its rounds are already lane-filling:
        // 2.37 vs 1.72 ms), so it is off unless asked for (ctx option "scan_mode": A/B runs, tests).
        const bool force_warp = ctx->opt_scan_mode == 1, force_groups = ctx->opt_scan_mode == 2;
        // the specialised kernels exist for the k of the BASELINE configs; other k share the general one
        const bool special = a.sv.hash == nullptr && special_arm<M, KT>(ARM_COMPACT) == ARM_COMPACT;
#ifdef BRGPU_NO_ARM_DENSE // A/B builds: dense sets through the run-time dispatching kernel
        const int arm = special && a.sv.dir != nullptr ? ARM_COMPACT : ARM_ANY;
#else
        const int arm = !special ? ARM_ANY : a.sv.dir != nullptr ? ARM_COMPACT : ARM_DENSE;
#endif
        if ((M == BRGPU_ONE && !force_warp) || (M == BRGPU_TWO && force_groups)) {
            // four segments per warp: a quarter of the warps for the same number of segments in flight
            constexpr int MG = (M == BRGPU_ONE || M == BRGPU_TWO) ? M : BRGPU_ONE;
            const uint64_t groups = scan_max_segments(L);
            const unsigned grid = grid_for_warps((uint64_t)occupancy_warps(ctx, scan_spec8_kernel<MG, KT, ARM_ANY>), (groups + 3) / 4);
            auto go = [&](auto kernel) {
                kernel<<<grid, threads, 0, ctx->stream>>>(a.d_in, a.d_len_in, L.d_slot_off, a.d_bitmap, w.d_seg_first, (uint32_t)L.n,
                                                          w.d_seg_out, (SegRec *)w.d_seg_recs, ctx->d_flags, a.sv, a.p, gc);
            };
            if (arm == ARM_COMPACT) go(scan_spec8_kernel<MG, KT, special_arm<M, KT>(ARM_COMPACT)>);
            else if (arm == ARM_DENSE) go(scan_spec8_kernel<MG, KT, special_arm<M, KT>(ARM_DENSE)>);
            else go(scan_spec8_kernel<MG, KT, ARM_ANY>);
        } else {
            const unsigned grid = grid_for_warps((uint64_t)occupancy_warps(ctx, scan_spec_kernel<M, KT, ARM_ANY>), scan_max_segments(L));
            auto go = [&](auto kernel) {
                kernel<<<grid, threads, 0, ctx->stream>>>(a.d_in, a.d_len_in, L.d_slot_off, a.d_bitmap, w.d_seg_first, (uint32_t)L.n,
                                                          w.d_seg_out, (SegRec *)w.d_seg_recs, ctx->d_flags, a.sv, a.p, a.d_scratch,
                                                          a.scratch_per_warp, gc);
            };
            if (arm == ARM_COMPACT) go(scan_spec_kernel<M, KT, special_arm<M, KT>(ARM_COMPACT)>);
            else if (arm == ARM_DENSE) go(scan_spec_kernel<M, KT, special_arm<M, KT>(ARM_DENSE)>);
            else go(scan_spec_kernel<M, KT, ARM_ANY>);
        }
    }
    SegCopy *d_copies = reinterpret_cast<SegCopy *>((uint8_t *)w.d_seg_recs + scan_max_segments(L) * sizeof(SegRec));
    cudaMemsetAsync(d_copies, 0, scan_max_segments(L) * sizeof(SegCopy), ctx->stream);
    {
        ProfScope ps(ctx, a.merge_name, a.n_bases_hint * 2.0);
        unsigned long long *gc = prof_counter_slot(ctx);
        const bool special = a.sv.hash == nullptr && special_arm<M, KT>(ARM_COMPACT) == ARM_COMPACT;
        const int arm = !special ? ARM_ANY : a.sv.dir != nullptr ? ARM_COMPACT : ARM_DENSE;
        const unsigned grid = grid_for_warps((uint64_t)occupancy_warps(ctx, scan_merge_kernel<M, KT, ARM_ANY>), L.n);
        auto go = [&](auto kernel) {
            kernel<<<grid, threads, 0, ctx->stream>>>(a.d_in, a.d_len_in, a.d_out, a.d_len_out, L.d_slot_off, a.d_bitmap, L.d_order,
                                                      w.d_seg_first, (uint32_t)L.n, w.d_seg_out, (const SegRec *)w.d_seg_recs,
                                                      ctx->d_flags, a.sv, a.p, a.d_scratch, a.scratch_per_warp, w.d_changed, d_copies,
                                                      gc);
        };
        if (arm == ARM_COMPACT) go(scan_merge_kernel<M, KT, special_arm<M, KT>(ARM_COMPACT)>);
        else if (arm == ARM_DENSE) go(scan_merge_kernel<M, KT, special_arm<M, KT>(ARM_DENSE)>);
        else go(scan_merge_kernel<M, KT, ARM_ANY>);
        // all spliced pieces at once (the merge warps only decided where they go)
        launch_scan_splice(ctx, d_copies, scan_max_segments(L), w.d_seg_out, a.d_out);
    }
}

} // namespace BRGPU_VARIANT
} // namespace brgpu
