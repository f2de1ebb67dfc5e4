/*
 * br_oracle.cpp — CPU restatement of natir/br's hot path.  TEST INFRASTRUCTURE ONLY
 * (see br_oracle.h for who may use it and for the parity status of each part).
 *
 * The code deliberately keeps the reference's shape — Vec → std::vector, Option →
 * std::optional, FxHashSet → std::unordered_set, the same loops in the same order — so
 * that each function can be read side by side with the Rust it cites.  It shares no
 * code with the product under br_b200/.
 */
#include "br_oracle.h"

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <unordered_set>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::vector<uint8_t> Bytes;

/* ------------------------------------------------------------------------------------------
 * cocktail::kmer  (git f63f0ba9; semantics verified against the .solid fixture)
 * ---------------------------------------------------------------------------------------- */

/* src/correct/mod.rs:26-42 — MASK_LOOKUP[k] = (1 << 2k) - 1 for 1 <= k < 32, [0] = 0 */
static inline uint64_t mask(int k) { return k == 0 ? 0 : ((1ULL << (2 * k)) - 1); }

static inline uint64_t nuc2bit(uint8_t b) { return (uint64_t)((b >> 1) & 3); }

static inline uint8_t bit2nuc(uint64_t x) {
    static const uint8_t t[4] = {'A', 'C', 'T', 'G'};
    return t[x & 3];
}

static inline uint64_t seq2bit(const uint8_t *s, size_t n) {
    uint64_t kmer = 0;
    for (size_t i = 0; i < n; i++) kmer = (kmer << 2) | nuc2bit(s[i]);
    return kmer;
}

static inline Bytes kmer2seq(uint64_t kmer, int k) {
    Bytes out((size_t)k);
    for (int i = k - 1; i >= 0; i--) {
        out[(size_t)i] = bit2nuc(kmer & 3);
        kmer >>= 2;
    }
    return out;
}

static inline uint64_t revcomp(uint64_t kmer, int k) {
    /* complement: A(00)<->T(10), C(01)<->G(11) == xor 0b10 on every group; then reverse groups */
    uint64_t c = kmer ^ 0xAAAAAAAAAAAAAAAAULL;
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | (c & 3);
        c >>= 2;
    }
    return r;
}

static inline bool parity_even(uint64_t x) { return (__builtin_popcountll(x) & 1) == 0; }

static inline uint64_t canonical(uint64_t kmer, int k) {
    return parity_even(kmer) ? kmer : revcomp(kmer, k);
}

/* src/correct/mod.rs:110-112 */
static inline uint64_t add_nuc_to_end(uint64_t kmer, uint64_t nuc, int k) {
    return ((kmer << 2) & mask(k)) ^ nuc;
}

/* ------------------------------------------------------------------------------------------
 * pcon::solid::Solid (git 0184ae77) — bitvec<u8, Lsb0> of 2^(2k-1) bits
 * ---------------------------------------------------------------------------------------- */
struct bro_set {
    int k;
    Bytes bits;
    /* set::Hash (src/set/hash.rs:14-17, :178-186): FxHashSet<u64> of canonical k-mers behind the same
     * KmerSet::get; only membership matters, so any exact set is equivalent.  Null for set::Pcon. */
    std::unordered_set<uint64_t> *hash = nullptr;
    ~bro_set() { delete hash; }
    inline bool get(uint64_t kmer) const {
        if (hash) return hash->count(canonical(kmer, k)) != 0; /* src/set/hash.rs:179-181 */
        uint64_t idx = canonical(kmer, k) >> 1;
        return (bits[idx >> 3] >> (idx & 7)) & 1;
    }
    inline void set(uint64_t kmer, bool v) {
        if (hash) {
            if (v)
                hash->insert(canonical(kmer, k));
            else
                hash->erase(canonical(kmer, k));
            return;
        }
        uint64_t idx = canonical(kmer, k) >> 1;
        if (v)
            bits[idx >> 3] |= (uint8_t)(1u << (idx & 7));
        else
            bits[idx >> 3] &= (uint8_t)~(1u << (idx & 7));
    }
};

static size_t nbits_for_k(int k) { return (size_t)1 << (2 * k - 1); }

/* ------------------------------------------------------------------------------------------
 * pcon::counter::Counter<u8> (feature count_u8, Cargo.toml:54,60)
 * ---------------------------------------------------------------------------------------- */
struct bro_counter {
    int k;
    size_t n;
    uint8_t *counts; /* calloc'ed: untouched pages stay unmapped for k=17 (8 GiB virtual) */
};

/* ------------------------------------------------------------------------------------------
 * src/correct/mod.rs:114-152 helpers
 * ---------------------------------------------------------------------------------------- */
static std::vector<uint64_t> next_nucs(const bro_set &s, uint64_t kmer) {
    std::vector<uint64_t> correct_nuc;
    for (uint64_t alt_nuc = 0; alt_nuc < 4; alt_nuc++)
        if (s.get(add_nuc_to_end(kmer, alt_nuc, s.k))) correct_nuc.push_back(alt_nuc);
    return correct_nuc;
}

static std::vector<uint64_t> alt_nucs(const bro_set &s, uint64_t ori) { return next_nucs(s, ori >> 2); }

static std::pair<size_t, uint64_t> error_len(const uint8_t *subseq, size_t len, uint64_t kmer, const bro_set &s) {
    size_t j = 0;
    for (;;) {
        j += 1;
        if (j >= len) break;
        kmer = add_nuc_to_end(kmer, nuc2bit(subseq[j]), s.k);
        if (s.get(kmer)) break;
    }
    return {j, kmer};
}

typedef std::optional<std::pair<Bytes, size_t>> Correction; /* Option<(Vec<u8>, usize)> */

/* ------------------------------------------------------------------------------------------
 * src/correct/exist/mod.rs — Scenario trait + Exist<S>
 * ---------------------------------------------------------------------------------------- */
struct Scenario {
    size_t c;
    int k;
    virtual ~Scenario() {}
    virtual std::optional<std::pair<uint64_t, size_t>> apply(const bro_set &s, uint64_t kmer, const uint8_t *seq,
                                                             size_t len) const = 0;
    virtual std::pair<Bytes, size_t> correct(const bro_set &s, uint64_t kmer, const uint8_t *seq,
                                             size_t len) const = 0;

    /* exist/mod.rs:21-47 */
    size_t get_score(const bro_set &s, uint64_t ori, const uint8_t *seq, size_t len) const {
        auto a = apply(s, ori, seq, len);
        if (!a) return 0;
        uint64_t kmer = a->first;
        size_t offset = a->second;
        if (!s.get(kmer)) return 0;
        if (offset + c > len) return 0;
        size_t score = 0;
        for (size_t p = offset; p < offset + c; p++) {
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[p]), s.k);
            if (s.get(kmer))
                score += 1;
            else
                break;
        }
        return score;
    }

    /* exist/mod.rs:49-70 */
    bool one_more(const bro_set &s, uint64_t kmer, const uint8_t *seq, size_t len) const {
        auto co = correct(s, kmer, seq, len);
        const Bytes &corr = co.first;
        size_t offset = co.second;
        if (len > c + offset + 1) {
            kmer >>= 2;
            for (uint8_t nuc : corr) kmer = add_nuc_to_end(kmer, nuc2bit(nuc), s.k);
            for (size_t p = offset; p < offset + c + 1; p++) kmer = add_nuc_to_end(kmer, nuc2bit(seq[p]), s.k);
            return s.get(kmer);
        }
        return false;
    }
};

/* exist/one.rs:33-74 */
struct ScenarioOne : Scenario {
    enum Kind { I, S, D } kind;
    ScenarioOne(Kind kd, size_t c_, int k_) : kind(kd) {
        c = c_;
        k = k_;
    }
    std::optional<std::pair<uint64_t, size_t>> apply(const bro_set &, uint64_t kmer, const uint8_t *,
                                                     size_t) const override {
        switch (kind) {
        case I: return std::make_pair(kmer, (size_t)2);
        case S: return std::make_pair(kmer, (size_t)1);
        default: return std::make_pair(kmer, (size_t)0);
        }
    }
    std::pair<Bytes, size_t> correct(const bro_set &, uint64_t kmer, const uint8_t *, size_t) const override {
        Bytes b{bit2nuc(kmer & 3)};
        switch (kind) {
        case I: return {b, 2};
        case S: return {b, 1};
        default: return {b, 0};
        }
    }
};

/* exist/two.rs:34-328 */
struct ScenarioTwo : Scenario {
    enum Kind { II, IS, SS, SD, DD, ICI, ICS, ICD, SCI, SCS, SCD, DCI, DCD, N_KIND } kind;
    ScenarioTwo(Kind kd, size_t c_, int k_) : kind(kd) {
        c = c_;
        k = k_;
    }
    typedef std::optional<std::pair<uint64_t, size_t>> Applied;

    Applied apply(const bro_set &s, uint64_t kmer, const uint8_t *seq, size_t len) const override {
        switch (kind) {
        case II: return std::make_pair(kmer, (size_t)3); /* two.rs:96 */
        case IS: return std::make_pair(kmer, (size_t)2); /* two.rs:97 */
        case SS: {                                       /* two.rs:98-114 */
            if (len < 2) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[1]), k);
            if (s.get(kmer)) return std::nullopt;
            auto alts = alt_nucs(s, kmer);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer >> 2, alts[0], k), (size_t)2);
        }
        case SD: { /* two.rs:115-126 */
            if (len == 0) return std::nullopt;
            auto alts = alt_nucs(s, kmer << 2);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer, alts[0], k), (size_t)1);
        }
        case DD: { /* two.rs:127-134 */
            auto alts = alt_nucs(s, kmer << 2);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer, alts[0], k), (size_t)0);
        }
        case ICI: { /* two.rs:135-148 */
            if (len < 4) return std::nullopt;
            uint64_t corr = add_nuc_to_end(kmer, nuc2bit(seq[3]), k);
            if (s.get(corr)) return std::make_pair(corr, (size_t)4);
            return std::nullopt;
        }
        case ICS: { /* two.rs:149-166 */
            if (len < 4) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[1]), k);
            if (s.get(kmer)) return std::nullopt;
            auto alts = alt_nucs(s, kmer);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer >> 2, alts[0], k), (size_t)3);
        }
        case ICD: { /* two.rs:167-181 */
            if (len < 4) return std::nullopt;
            uint64_t second = add_nuc_to_end(kmer, nuc2bit(seq[2]), k);
            auto alts = alt_nucs(s, second << 2);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(second, alts[0], k), (size_t)3);
        }
        case SCI: /* two.rs:182-191 */
        case DCI: /* two.rs:231-240 (same body) */ {
            if (len < 4) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[1]), k);
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[3]), k);
            return std::make_pair(kmer, (size_t)4);
        }
        case SCS: { /* two.rs:192-215 */
            if (len < 3) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[1]), k);
            if (s.get(kmer)) {
                kmer = add_nuc_to_end(kmer, nuc2bit(seq[2]), k);
                if (!s.get(kmer)) {
                    auto alts = alt_nucs(s, kmer);
                    if (alts.size() == 1) return std::make_pair(add_nuc_to_end(kmer >> 2, alts[0], k), (size_t)3);
                    return std::nullopt;
                }
                return std::nullopt;
            }
            return std::nullopt;
        }
        case SCD: { /* two.rs:216-230 */
            if (len < 2) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[1]), k);
            auto alts = alt_nucs(s, kmer << 2);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer, alts[0], k), (size_t)2);
        }
        case DCD: { /* two.rs:241-254 */
            if (len < 2) return std::nullopt;
            kmer = add_nuc_to_end(kmer, nuc2bit(seq[0]), k);
            auto alts = alt_nucs(s, kmer << 2);
            if (alts.size() != 1) return std::nullopt;
            return std::make_pair(add_nuc_to_end(kmer, alts[0], k), (size_t)1);
        }
        default: break;
        }
        return std::nullopt;
    }

    /* two.rs:258-325 */
    std::pair<Bytes, size_t> correct(const bro_set &s, uint64_t kmer, const uint8_t *seq, size_t len) const override {
        switch (kind) {
        case II: return {Bytes{bit2nuc(kmer & 3)}, 2};
        case IS: return {Bytes{bit2nuc(kmer & 3)}, 2};
        case SS:
        case SD:
        case DD: {
            auto a = apply(s, kmer, seq, len);
            if (!a) abort(); /* .expect("we can't failled her") */
            uint64_t corr = a->first;
            return {Bytes{bit2nuc((corr & 0xC) >> 2), bit2nuc(corr & 3)}, a->second};
        }
        case ICI: return {Bytes{bit2nuc(kmer & 3)}, 3};
        case ICD: {
            auto a = apply(s, kmer, seq, len);
            if (!a) abort();
            uint64_t corr = a->first;
            return {Bytes{bit2nuc((corr & 0xC) >> 2), bit2nuc(corr & 3)}, a->second - 1};
        }
        case ICS: {
            auto a = apply(s, kmer, seq, len);
            if (!a) abort();
            uint64_t corr = a->first;
            return {Bytes{bit2nuc((corr & 0xC) >> 2), bit2nuc(corr & 3)}, a->second + 1};
        }
        case SCI:
        case SCS:
        case SCD:
        case DCD: {
            auto a = apply(s, kmer, seq, len);
            if (!a) abort();
            uint64_t corr = a->first;
            return {Bytes{bit2nuc((corr & 0x30) >> 4), bit2nuc((corr & 0xC) >> 2), bit2nuc(corr & 3)}, a->second};
        }
        default: /* DCI: `_ => (vec![], 1)` two.rs:323 */
            return {Bytes{}, 1};
        }
    }
};

/* exist/mod.rs:97-149 */
template <class S, int NKIND>
static Correction exist_correct_error(const bro_set &set, size_t c, uint64_t kmer, const uint8_t *seq, size_t len) {
    auto alts = alt_nucs(set, kmer);
    if (alts.size() != 1) return std::nullopt;

    uint64_t corr = add_nuc_to_end(kmer >> 2, alts[0], set.k);

    std::vector<S> scenarii; /* get_scenarii: declaration order */
    for (int kd = 0; kd < NKIND; kd++) {
        S sc((typename S::Kind)kd, c, set.k);
        if (sc.get_score(set, corr, seq, len) == c) scenarii.push_back(sc);
    }

    if (scenarii.empty()) return std::nullopt;
    if (scenarii.size() == 1) return scenarii[0].correct(set, corr, seq, len);

    std::vector<S> kept;
    for (auto &sc : scenarii)
        if (sc.one_more(set, corr, seq, len)) kept.push_back(sc);
    if (kept.size() == 1) return kept[0].correct(set, corr, seq, len);
    return std::nullopt;
}

static Correction one_correct_error(const bro_set &set, size_t c, uint64_t kmer, const uint8_t *seq, size_t len) {
    return exist_correct_error<ScenarioOne, 3>(set, c, kmer, seq, len);
}
static Correction two_correct_error(const bro_set &set, size_t c, uint64_t kmer, const uint8_t *seq, size_t len) {
    return exist_correct_error<ScenarioTwo, (int)ScenarioTwo::N_KIND>(set, c, kmer, seq, len);
}

/* ------------------------------------------------------------------------------------------
 * src/correct/graph.rs:44-85
 * ---------------------------------------------------------------------------------------- */
static Correction graph_correct_error(const bro_set &set, uint64_t kmer, const uint8_t *seq, size_t len) {
    auto el = error_len(seq, len, kmer, set);
    size_t elen = el.first;
    uint64_t first_correct_kmer = el.second;

    std::unordered_set<uint64_t> viewed_kmer;
    Bytes local_corr;

    auto alts = alt_nucs(set, kmer);
    if (alts.size() != 1) return std::nullopt;

    kmer = add_nuc_to_end(kmer >> 2, alts[0], set.k);
    local_corr.push_back(bit2nuc(alts[0]));
    viewed_kmer.insert(kmer);

    while (set.get(kmer)) {
        auto nx = next_nucs(set, kmer);
        if (nx.size() != 1) return std::nullopt;

        kmer = add_nuc_to_end(kmer, nx[0], set.k);

        if (viewed_kmer.count(kmer)) return std::nullopt;
        viewed_kmer.insert(kmer);

        local_corr.push_back(bit2nuc(nx[0]));

        if (kmer == first_correct_kmer) break;
    }

    return std::make_pair(local_corr, elen + 1);
}

/* ------------------------------------------------------------------------------------------
 * bio 1.6.0  alignment::pairwise::Aligner::{custom, global}      *** PARITY UNPINNED ***
 *
 * rust-bio is not in /root/reference (Cargo.lock:137-139).  This restates the published
 * algorithm of bio::alignment::pairwise (three-layer affine DP S/I/D, column-major over y,
 * per-cell 3x4-bit traceback, clip penalties forced to MIN_SCORE in `global`).  Tie rules:
 * gap extension wins only if strictly greater than gap open; in S the order tried is
 * diagonal, I, D, each replacing the incumbent only if strictly greater.
 * ---------------------------------------------------------------------------------------- */
namespace bio {
enum Op : uint8_t { Match = 0, Subst = 1, Del = 2, Ins = 3, Xclip = 4, Yclip = 5 };

static const int32_t MIN_SCORE = -858993459; /* (i32::MIN as f32 * 0.4) as i32 */
enum : uint16_t {
    TB_START = 0,
    TB_INS = 1,
    TB_DEL = 2,
    TB_SUBST = 3,
    TB_MATCH = 4,
    TB_XCLIP_PREFIX = 5,
    TB_XCLIP_SUFFIX = 6,
    TB_YCLIP_PREFIX = 7,
    TB_YCLIP_SUFFIX = 8
};

struct Cell {
    uint16_t v = 0; /* I bits [0,4), D bits [4,8), S bits [8,12) */
    void set_i(uint16_t x) { v = (uint16_t)((v & ~0x000F) | x); }
    void set_d(uint16_t x) { v = (uint16_t)((v & ~0x00F0) | (x << 4)); }
    void set_s(uint16_t x) { v = (uint16_t)((v & ~0x0F00) | (x << 8)); }
    void set_all(uint16_t x) {
        set_i(x);
        set_d(x);
        set_s(x);
    }
    uint16_t i() const { return v & 0xF; }
    uint16_t d() const { return (v >> 4) & 0xF; }
    uint16_t s() const { return (v >> 8) & 0xF; }
};

struct Scoring {
    int32_t gap_open, gap_extend;
    int32_t xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix;
    int32_t score(uint8_t a, uint8_t b) const { return a == b ? 1 : -1; } /* greedy.rs:31-39 */
};

static std::vector<Op> custom(const Scoring &sc, const uint8_t *x, size_t m, const uint8_t *y, size_t n) {
    std::vector<Cell> tb((m + 1) * (n + 1));
    auto T = [&](size_t i, size_t j) -> Cell & { return tb[i * (n + 1) + j]; };

    std::vector<int32_t> I[2], D[2], S[2], Sn;
    std::vector<size_t> Lx, Ly;

    for (int k = 0; k < 2; k++) {
        D[k].assign(m + 1, MIN_SCORE);
        I[k].assign(m + 1, MIN_SCORE);
        S[k].assign(m + 1, MIN_SCORE);
        S[k][0] = 0;
        if (k == 0) {
            Cell c;
            c.set_all(TB_START);
            T(0, 0) = c;
            Lx.assign(n + 1, 0);
            Ly.assign(m + 1, 0);
            Sn.assign(m + 1, MIN_SCORE);
            Sn[0] = sc.yclip_suffix;
            Ly[0] = n;
        } else {
            Lx[0] = m;
        }

        for (size_t i = 1; i <= m; i++) {
            Cell c;
            c.set_all(TB_START);
            if (i == 1) {
                I[k][i] = sc.gap_open + sc.gap_extend;
                c.set_i(TB_START);
            } else {
                int32_t i_score = sc.gap_open + sc.gap_extend * (int32_t)i;
                int32_t c_score = sc.xclip_prefix + sc.gap_open + sc.gap_extend;
                if (i_score > c_score) {
                    I[k][i] = i_score;
                    c.set_i(TB_INS);
                } else {
                    I[k][i] = c_score;
                    c.set_i(TB_XCLIP_PREFIX);
                }
            }
            if (i == m)
                c.set_s(TB_XCLIP_SUFFIX);
            else
                S[k][i] = MIN_SCORE;

            if (I[k][i] > S[k][i]) {
                S[k][i] = I[k][i];
                c.set_s(TB_INS);
            }
            if (sc.xclip_prefix > S[k][i]) {
                S[k][i] = sc.xclip_prefix;
                c.set_s(TB_XCLIP_PREFIX);
            }
            if (i != m && S[k][i] + sc.xclip_suffix > S[k][m]) {
                S[k][m] = S[k][i] + sc.xclip_suffix;
                Lx[0] = m - i;
            }
            if (k == 0) T(i, 0) = c;
            if (S[k][i] + sc.yclip_suffix > Sn[i]) {
                Sn[i] = S[k][i] + sc.yclip_suffix;
                Ly[i] = n;
            }
        }
    }

    for (size_t j = 1; j <= n; j++) {
        int curr = (int)(j % 2), prev = 1 - curr;
        {
            Cell c;
            I[curr][0] = MIN_SCORE;
            if (j == 1) {
                D[curr][0] = sc.gap_open + sc.gap_extend;
                c.set_d(TB_START);
            } else {
                int32_t d_score = sc.gap_open + sc.gap_extend * (int32_t)j;
                int32_t c_score = sc.yclip_prefix + sc.gap_open + sc.gap_extend;
                if (d_score > c_score) {
                    D[curr][0] = d_score;
                    c.set_d(TB_DEL);
                } else {
                    D[curr][0] = c_score;
                    c.set_d(TB_YCLIP_PREFIX);
                }
            }
            if (D[curr][0] > sc.yclip_prefix) {
                S[curr][0] = D[curr][0];
                c.set_s(TB_DEL);
            } else {
                S[curr][0] = sc.yclip_prefix;
                c.set_s(TB_YCLIP_PREFIX);
            }
            if (j == n && Sn[0] > S[curr][0]) {
                S[curr][0] = Sn[0];
                c.set_s(TB_YCLIP_SUFFIX);
            } else if (S[curr][0] + sc.yclip_suffix > Sn[0]) {
                Sn[0] = S[curr][0] + sc.yclip_suffix;
                Ly[0] = n - j;
            }
            T(0, j) = c;
        }

        for (size_t i = 1; i <= m; i++) S[curr][i] = MIN_SCORE;

        uint8_t q = y[j - 1];
        int32_t xclip_score =
            sc.xclip_prefix + std::max(sc.yclip_prefix, sc.gap_open + sc.gap_extend * (int32_t)j);
        for (size_t i = 1; i <= m; i++) {
            uint8_t p = x[i - 1];
            Cell c;

            int32_t m_score = S[prev][i - 1] + sc.score(p, q);

            int32_t i_score = I[curr][i - 1] + sc.gap_extend;
            int32_t s_score = S[curr][i - 1] + sc.gap_open + sc.gap_extend;
            int32_t best_i_score;
            if (i_score > s_score) {
                best_i_score = i_score;
                c.set_i(TB_INS);
            } else {
                best_i_score = s_score;
                c.set_i(T(i - 1, j).s());
            }

            int32_t d_score = D[prev][i] + sc.gap_extend;
            s_score = S[prev][i] + sc.gap_open + sc.gap_extend;
            int32_t best_d_score;
            if (d_score > s_score) {
                best_d_score = d_score;
                c.set_d(TB_DEL);
            } else {
                best_d_score = s_score;
                c.set_d(T(i, j - 1).s());
            }

            c.set_s(TB_XCLIP_SUFFIX);
            int32_t best_s_score = S[curr][i];

            if (m_score > best_s_score) {
                best_s_score = m_score;
                c.set_s(p == q ? TB_MATCH : TB_SUBST);
            }
            if (best_i_score > best_s_score) {
                best_s_score = best_i_score;
                c.set_s(TB_INS);
            }
            if (best_d_score > best_s_score) {
                best_s_score = best_d_score;
                c.set_s(TB_DEL);
            }
            if (xclip_score > best_s_score) {
                best_s_score = xclip_score;
                c.set_s(TB_XCLIP_PREFIX);
            }
            int32_t yclip_score = sc.yclip_prefix + sc.gap_open + sc.gap_extend * (int32_t)i;
            if (yclip_score > best_s_score) {
                best_s_score = yclip_score;
                c.set_s(TB_YCLIP_PREFIX);
            }

            S[curr][i] = best_s_score;
            I[curr][i] = best_i_score;
            D[curr][i] = best_d_score;

            if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
                S[curr][m] = S[curr][i] + sc.xclip_suffix;
                Lx[j] = m - i;
            }
            if (S[curr][i] + sc.yclip_suffix > Sn[i]) {
                Sn[i] = S[curr][i] + sc.yclip_suffix;
                Ly[i] = n - j;
            }
            T(i, j) = c;
        }
    }

    /* suffix clipping in the j = n column */
    for (size_t i = 0; i <= m; i++) {
        size_t j = n;
        int curr = (int)(j % 2);
        if (Sn[i] > S[curr][i]) {
            S[curr][i] = Sn[i];
            T(i, j).set_s(TB_YCLIP_SUFFIX);
        }
        if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
            S[curr][m] = S[curr][i] + sc.xclip_suffix;
            Lx[j] = m - i;
            T(m, j).set_s(TB_XCLIP_SUFFIX);
        }
    }
    /* the last column of I may change because S changed */
    for (size_t i = 1; i <= m; i++) {
        size_t j = n;
        int curr = (int)(j % 2);
        int32_t s_score = S[curr][i - 1] + sc.gap_open + sc.gap_extend;
        if (s_score > I[curr][i]) {
            I[curr][i] = s_score;
            T(i, j).set_i(T(i - 1, j).s());
        }
        if (s_score > S[curr][i]) {
            S[curr][i] = s_score;
            T(i, j).set_s(TB_INS);
            if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
                S[curr][m] = S[curr][i] + sc.xclip_suffix;
                Lx[j] = m - i;
                T(m, j).set_s(TB_XCLIP_SUFFIX);
            }
        }
    }

    size_t i = m, j = n;
    std::vector<Op> ops;
    uint16_t last_layer = T(i, j).s();
    for (;;) {
        uint16_t next_layer;
        switch (last_layer) {
        case TB_START: goto done;
        case TB_INS:
            ops.push_back(Ins);
            next_layer = T(i, j).i();
            i -= 1;
            break;
        case TB_DEL:
            ops.push_back(Del);
            next_layer = T(i, j).d();
            j -= 1;
            break;
        case TB_MATCH:
            ops.push_back(Match);
            next_layer = T(i - 1, j - 1).s();
            i -= 1;
            j -= 1;
            break;
        case TB_SUBST:
            ops.push_back(Subst);
            next_layer = T(i - 1, j - 1).s();
            i -= 1;
            j -= 1;
            break;
        case TB_XCLIP_PREFIX:
            ops.push_back(Xclip);
            i = 0;
            next_layer = T(0, j).s();
            break;
        case TB_XCLIP_SUFFIX:
            ops.push_back(Xclip);
            i -= Lx[j];
            next_layer = T(i, j).s();
            break;
        case TB_YCLIP_PREFIX:
            ops.push_back(Yclip);
            j = 0;
            next_layer = T(i, 0).s();
            break;
        case TB_YCLIP_SUFFIX:
            ops.push_back(Yclip);
            j -= Ly[i];
            next_layer = T(i, j).s();
            break;
        default: fprintf(stderr, "br_oracle: corrupt traceback\n"); abort();
        }
        last_layer = next_layer;
    }
done:
    std::reverse(ops.begin(), ops.end());
    return ops;
}

/* Aligner::global: clip penalties = MIN_SCORE, then filter_clip_operations */
static std::vector<Op> global(const uint8_t *x, size_t m, const uint8_t *y, size_t n) {
    Scoring sc{-1, -1, MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE}; /* greedy.rs:63-64 */
    std::vector<Op> ops = custom(sc, x, m, y, n);
    std::vector<Op> out;
    for (Op o : ops) {
        /* a clip can never be optimal in global mode; if the restatement ever produced one the
         * unpinned part of the oracle would be wrong, so fail loudly. */
        if (o == Xclip || o == Yclip) {
            fprintf(stderr, "br_oracle: clip operation in a global alignment\n");
            abort();
        }
        out.push_back(o);
    }
    return out;
}
} // namespace bio

/* ------------------------------------------------------------------------------------------
 * src/correct/greedy.rs
 * ---------------------------------------------------------------------------------------- */
/* greedy.rs:56-89 */
static std::optional<int64_t> match_alignement(const Bytes &before_seq, const uint8_t *read, size_t nread,
                                               const Bytes &corr) {
    Bytes r = before_seq;
    r.insert(r.end(), read, read + nread);
    Bytes c = before_seq;
    c.insert(c.end(), corr.begin(), corr.end());

    std::vector<bio::Op> operations = bio::global(r.data(), r.size(), c.data(), c.size());

    int64_t offset = 0;
    if (operations.size() < before_seq.size()) abort(); /* the Rust slice would panic */
    for (size_t w = before_seq.size(); w + 1 < operations.size(); w++) { /* .windows(2) */
        bio::Op op0 = operations[w], op1 = operations[w + 1];
        if (op0 == bio::Del)
            offset -= 1;
        else if (op0 == bio::Ins)
            offset += 1;

        if (op0 == bio::Match && op0 == op1) {
            int64_t offset_corr = 0;
            for (size_t e = operations.size(); e-- > 0;) {
                if (operations[e] == bio::Del)
                    offset_corr -= 1;
                else if (operations[e] == bio::Ins)
                    offset_corr += 1;
                else
                    break;
            }
            return offset - offset_corr;
        }
    }
    return std::nullopt;
}

/* greedy.rs:91-102 */
static std::optional<std::pair<uint8_t, uint64_t>> follow_graph(const bro_set &set, uint64_t kmer) {
    auto alts = next_nucs(set, kmer);
    if (alts.size() != 1) return std::nullopt;
    kmer = add_nuc_to_end(kmer, alts[0], set.k);
    return std::make_pair(bit2nuc(alts[0]), kmer);
}

/* greedy.rs:104-117 */
static bool check_next_kmers(const bro_set &set, size_t nb_validate, uint64_t kmer, const uint8_t *seq, size_t len) {
    if (len < nb_validate) return false;
    for (size_t p = 0; p < nb_validate; p++) {
        kmer = add_nuc_to_end(kmer, nuc2bit(seq[p]), set.k);
        if (!set.get(kmer)) return false;
    }
    return true;
}

/* greedy.rs:129-173 */
static Correction greedy_correct_error(const bro_set &set, size_t max_search, size_t nb_validate, uint64_t kmer,
                                       const uint8_t *seq, size_t len) {
    auto alts = alt_nucs(set, kmer);
    if (alts.size() != 1) return std::nullopt;

    std::unordered_set<uint64_t> viewed_kmer;
    Bytes local_corr;
    Bytes before_seq = kmer2seq(kmer >> 2, set.k - 1);

    kmer = add_nuc_to_end(kmer >> 2, alts[0], set.k);

    local_corr.push_back(bit2nuc(alts[0]));
    viewed_kmer.insert(kmer);

    for (size_t i = 0; i < max_search; i++) {
        if (auto f = follow_graph(set, kmer)) {
            local_corr.push_back(f->first);
            kmer = f->second;
        }

        if (viewed_kmer.count(kmer)) return std::nullopt;
        viewed_kmer.insert(kmer);

        if (len < i) return std::nullopt;

        if (auto off = match_alignement(before_seq, seq, i, local_corr)) {
            if (check_next_kmers(set, nb_validate, kmer, seq + i, len - i)) {
                int64_t o = (int64_t)local_corr.size() + *off;
                /* a negative sum would wrap in release Rust; treat as unreachable */
                if (o < 0) {
                    fprintf(stderr, "br_oracle: negative greedy offset\n");
                    abort();
                }
                return std::make_pair(local_corr, (size_t)o);
            }
        }
    }
    return std::nullopt;
}

/* ------------------------------------------------------------------------------------------
 * src/correct/gap_size.rs
 * ---------------------------------------------------------------------------------------- */
/* gap_size.rs:44-89 */
static Correction ins_sub_correction(const bro_set &set, uint64_t kmer, size_t gap_size) {
    auto alts = alt_nucs(set, kmer);
    if (alts.size() != 1) return std::nullopt;

    uint64_t corr = add_nuc_to_end(kmer >> 2, alts[0], set.k);
    Bytes local_corr{bit2nuc(alts[0])};
    std::unordered_set<uint64_t> viewed_kmer;
    viewed_kmer.insert(corr);

    for (size_t i = 0; i < gap_size; i++) {
        alts = next_nucs(set, corr);
        if (alts.size() != 1) return std::nullopt;
        corr = add_nuc_to_end(corr, alts[0], set.k);
        if (viewed_kmer.count(corr)) return std::nullopt;
        viewed_kmer.insert(corr);
        local_corr.push_back(bit2nuc(alts[0]));
    }
    size_t offset = local_corr.size();
    return std::make_pair(local_corr, offset);
}

/* gap_size.rs:97-108 */
static Correction gap_size_correct_error(const bro_set &set, size_t c, uint64_t kmer, const uint8_t *seq, size_t len) {
    size_t elen = error_len(seq, len, kmer, set).first;
    size_t k = (size_t)set.k;
    if (elen < k) return graph_correct_error(set, kmer, seq, len);
    if (elen == k) return one_correct_error(set, c, kmer, seq, len);
    return ins_sub_correction(set, kmer, elen - k);
}

/* ------------------------------------------------------------------------------------------
 * Corrector trait: correct_error dispatch (src/lib.rs:141-164 build_methods argument mapping)
 * and the scan loop Corrector::correct (src/correct/mod.rs:53-107)
 * ---------------------------------------------------------------------------------------- */
struct Params {
    int method;
    size_t confirm;    /* One/Two/GapSize: c; Greedy: nb_validate (src/lib.rs:155) */
    size_t max_search; /* Greedy */
};

static Correction correct_error(const bro_set &set, const Params &p, uint64_t kmer, const uint8_t *seq, size_t len) {
    switch (p.method) {
    case BRO_ONE: return one_correct_error(set, p.confirm, kmer, seq, len);
    case BRO_TWO: return two_correct_error(set, p.confirm, kmer, seq, len);
    case BRO_GRAPH: return graph_correct_error(set, kmer, seq, len);
    case BRO_GREEDY: return greedy_correct_error(set, p.max_search, p.confirm, kmer, seq, len);
    case BRO_GAP_SIZE: return gap_size_correct_error(set, p.confirm, kmer, seq, len);
    }
    fprintf(stderr, "br_oracle: unknown method %d\n", p.method);
    abort();
}

static Bytes correct(const bro_set &set, const Params &p, const uint8_t *seq, size_t len) {
    Bytes correct;
    correct.reserve(len);
    size_t k = (size_t)set.k;

    if (len < k) return Bytes(seq, seq + len);

    size_t i = k;
    uint64_t kmer = seq2bit(seq, i);
    for (size_t n = 0; n < i; n++) correct.push_back(seq[n]);

    bool previous = set.get(kmer);
    while (i < len) {
        uint8_t nuc = seq[i];
        kmer = add_nuc_to_end(kmer, nuc2bit(nuc), set.k);

        if (!set.get(kmer) && previous) {
            if (auto r = correct_error(set, p, kmer, seq + i, len - i)) {
                kmer >>= 2;
                for (uint8_t b : r->first) {
                    kmer = add_nuc_to_end(kmer, nuc2bit(b), set.k);
                    correct.push_back(b);
                }
                previous = true;
                i += r->second;
            } else {
                correct.push_back(nuc);
                i += 1;
                previous = false;
            }
        } else {
            previous = set.get(kmer);
            correct.push_back(nuc);
            i += 1;
        }
    }
    return correct;
}

/* src/lib.rs:44-55 — per-record body of run_correction */
static Bytes correct_record(const bro_set &set, const std::vector<Params> &methods, bool two_side, const uint8_t *seq,
                            size_t len) {
    Bytes cur(seq, seq + len);
    for (auto &m : methods) cur = correct(set, m, cur.data(), cur.size());
    if (!two_side) {
        std::reverse(cur.begin(), cur.end());
        for (auto &m : methods) cur = correct(set, m, cur.data(), cur.size());
        std::reverse(cur.begin(), cur.end());
    }
    return cur;
}

struct bro_result {
    Bytes data;
    std::vector<uint64_t> offsets;
};

/* ------------------------------------------------------------------------------------------
 * C API
 * ---------------------------------------------------------------------------------------- */
extern "C" {

uint64_t bro_nuc2bit(uint8_t b) { return nuc2bit(b); }
uint8_t bro_bit2nuc(uint64_t x) { return bit2nuc(x); }
uint64_t bro_seq2bit(const uint8_t *seq, size_t len) { return seq2bit(seq, len); }
uint64_t bro_revcomp(uint64_t kmer, int k) { return revcomp(kmer, k); }
uint64_t bro_canonical(uint64_t kmer, int k) { return canonical(kmer, k); }

bro_set *bro_set_new(int k) {
    bro_set *s = new bro_set;
    s->k = k;
    s->bits.assign((nbits_for_k(k) + 7) / 8, 0);
    return s;
}
bro_set *bro_set_from_bitfield(int k, const uint8_t *bits, size_t n) {
    bro_set *s = bro_set_new(k);
    if (n != s->bits.size()) {
        delete s;
        return nullptr;
    }
    memcpy(s->bits.data(), bits, n);
    return s;
}
/* set::Hash — empty, then Hash::from_fasta's loop over records (src/set/hash.rs:41-60): every
 * canonical k-mer of every record with len >= k */
bro_set *bro_hash_new(int k) {
    bro_set *s = new bro_set;
    s->k = k;
    s->hash = new std::unordered_set<uint64_t>();
    return s;
}
void bro_hash_add_reads(bro_set *s, const uint8_t *seq, const uint64_t *offsets, size_t n_reads) {
    const size_t k = (size_t)s->k;
    for (size_t r = 0; r < n_reads; r++) {
        const uint8_t *p = seq + offsets[r];
        const size_t len = (size_t)(offsets[r + 1] - offsets[r]);
        if (len < k) continue; /* src/set/hash.rs:52 */
        uint64_t kmer = seq2bit(p, k - 1);
        for (size_t i = k - 1; i < len; i++) { /* cocktail::tokenizer::Canonical */
            kmer = add_nuc_to_end(kmer, nuc2bit(p[i]), s->k);
            s->hash->insert(canonical(kmer, s->k));
        }
    }
}
size_t bro_hash_size(const bro_set *s) { return s->hash ? s->hash->size() : 0; }
void bro_set_free(bro_set *s) { delete s; }
int bro_set_k(const bro_set *s) { return s->k; }
void bro_set_set(bro_set *s, uint64_t kmer, int value) { s->set(kmer, value != 0); }
int bro_set_get(const bro_set *s, uint64_t kmer) { return s->get(kmer); }
const uint8_t *bro_set_bits(const bro_set *s, size_t *nbytes) {
    *nbytes = s->bits.size();
    return s->bits.data();
}
void bro_set_get_batch(const bro_set *s, const uint64_t *kmers, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = s->get(kmers[i]);
}
void bro_set_insert_all_kmers(bro_set *s, const uint8_t *seq, size_t len) {
    /* cocktail::tokenizer::Tokenizer — every forward k-mer, rolling */
    size_t k = (size_t)s->k;
    if (len < k) return;
    uint64_t kmer = seq2bit(seq, k - 1);
    for (size_t i = k - 1; i < len; i++) {
        kmer = add_nuc_to_end(kmer, nuc2bit(seq[i]), s->k);
        s->set(kmer, true);
    }
}

bro_counter *bro_counter_new(int k) {
    bro_counter *c = new bro_counter;
    c->k = k;
    c->n = nbits_for_k(k);
    c->counts = (uint8_t *)calloc(c->n, 1);
    if (!c->counts) {
        delete c;
        return nullptr;
    }
    return c;
}
void bro_counter_free(bro_counter *c) {
    if (c) free(c->counts);
    delete c;
}

static inline void inc_sat(uint8_t *p, bool atomic) {
    if (!atomic) {
        if (*p != 255) *p += 1; /* saturating_add */
        return;
    }
    uint8_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old != 255 && !__atomic_compare_exchange_n(p, &old, (uint8_t)(old + 1), true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
}

void bro_counter_count(bro_counter *c, const uint8_t *seq, const uint64_t *offsets, size_t n_reads, int threads) {
    const int k = c->k;
    const bool atomic = threads > 1;
    (void)atomic;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 1 ? threads : 1)
    for (long r = 0; r < (long)n_reads; r++) {
        const uint8_t *s = seq + offsets[r];
        size_t len = (size_t)(offsets[r + 1] - offsets[r]);
        if (len < (size_t)k) continue; /* src/set/pcon.rs:58 and pcon count_fasta */
        /* cocktail::tokenizer::Canonical: rolling forward + reverse complement */
        uint64_t fwd = seq2bit(s, (size_t)k);
        uint64_t rev = revcomp(fwd, k);
        const int shift = 2 * (k - 1);
        for (size_t i = (size_t)k - 1;;) {
            uint64_t cano = parity_even(fwd) ? fwd : rev;
            inc_sat(&c->counts[cano >> 1], atomic);
            if (++i >= len) break;
            uint64_t n = nuc2bit(s[i]);
            fwd = add_nuc_to_end(fwd, n, k);
            rev = (rev >> 2) | ((n ^ 2) << shift);
        }
    }
}

const uint8_t *bro_counter_raw(const bro_counter *c, size_t *n) {
    *n = c->n;
    return c->counts;
}

/* pcon::spectrum::Spectrum::from_count — histogram over ALL counters incl. zeros (PARITY UNPINNED) */
void bro_spectrum(const bro_counter *c, uint64_t hist[256], int threads) {
    for (int i = 0; i < 256; i++) hist[i] = 0;
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
    {
        uint64_t local[256] = {0};
#pragma omp for schedule(static)
        for (long i = 0; i < (long)c->n; i++) local[c->counts[i]]++;
#pragma omp critical
        for (int i = 0; i < 256; i++) hist[i] += local[i];
    }
}

/* pcon Spectrum::get_threshold(FirstMinimum): first i with hist[i+1] > hist[i] (PARITY UNPINNED) */
int bro_first_minimum(const uint64_t hist[256]) {
    for (int i = 0; i + 1 < 256; i++)
        if (hist[i + 1] > hist[i]) return i;
    return -1;
}

/* pcon Spectrum::get_threshold for the three percent-driven methods br can ask for
 * (src/main.rs:100-108: Rarefaction, PercentAtLeast, PercentAtMost).  pcon @0184ae77 is not
 * vendored and no reference test calls these: restated from the published source as recalled —
 * PARITY UNPINNED.
 *   rarefaction(limit):     walk the bins with the running sum of index * count; first bin whose
 *                           count / running sum drops below limit
 *   percent_at_least(p):    first bin at which the running share of index * count exceeds p
 *   percent_at_most(p):     the bin before that one
 * method: 2 rarefaction, 3 percent_at_most, 4 percent_at_least; returns -1 for None. */
int bro_spectrum_threshold(const uint64_t hist[256], int method, double percent) {
    if (method == 2) {
        uint64_t cumulative = 0;
        for (int i = 0; i < 256; i++) {
            cumulative += (uint64_t)i * hist[i];
            if ((double)hist[i] / (double)cumulative < percent) return i;
        }
        return -1;
    }
    if (method == 3 || method == 4) {
        uint64_t total = 0;
        for (int i = 0; i < 256; i++) total += (uint64_t)i * hist[i];
        uint64_t cumulative = 0;
        for (int i = 0; i < 256; i++) {
            cumulative += (uint64_t)i * hist[i];
            if ((double)cumulative / (double)total > percent) return method == 4 ? i : (i > 0 ? i - 1 : -1);
        }
        return -1;
    }
    return -1;
}

/* pcon Solid::from_count: bit[i] = counts[i] > abundance (src/main.rs:112-114; fixture `a2` == count>=3) */
bro_set *bro_solid_from_count(const bro_counter *c, int abundance, int threads) {
    bro_set *s = bro_set_new(c->k);
    size_t nbytes = s->bits.size();
    size_t n = c->n;
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
    for (long b = 0; b < (long)nbytes; b++) {
        uint8_t v = 0;
        for (int t = 0; t < 8; t++) {
            size_t i = (size_t)b * 8 + (size_t)t;
            if (i < n && c->counts[i] > abundance) v |= (uint8_t)(1u << t);
        }
        s->bits[(size_t)b] = v;
    }
    return s;
}

int bro_alt_nucs(const bro_set *s, uint64_t kmer, uint64_t out[4]) {
    auto v = alt_nucs(*s, kmer);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}
int bro_next_nucs(const bro_set *s, uint64_t kmer, uint64_t out[4]) {
    auto v = next_nucs(*s, kmer);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

long bro_correct_error(const bro_set *s, int method, int confirm, int max_search, uint64_t kmer, const uint8_t *seq,
                       size_t len, uint8_t *out, size_t cap, size_t *offset) {
    Params p{method, (size_t)confirm, (size_t)max_search};
    auto r = correct_error(*s, p, kmer, seq, len);
    if (!r) return -1;
    if (!r->first.empty()) memcpy(out, r->first.data(), std::min(cap, r->first.size())); /* DCI emits no base */
    *offset = r->second;
    return (long)r->first.size();
}

size_t bro_correct(const bro_set *s, int method, int confirm, int max_search, const uint8_t *seq, size_t len,
                   uint8_t *out, size_t cap) {
    Params p{method, (size_t)confirm, (size_t)max_search};
    Bytes r = correct(*s, p, seq, len);
    if (!r.empty()) memcpy(out, r.data(), std::min(cap, r.size()));
    return r.size();
}

bro_result *bro_run_correction(const bro_set *s, const uint8_t *methods, size_t n_methods, int confirm, int max_search,
                               int two_side, const uint8_t *seq, const uint64_t *offsets, size_t n_reads, int threads) {
    std::vector<Params> ms;
    for (size_t i = 0; i < n_methods; i++) ms.push_back(Params{methods[i], (size_t)confirm, (size_t)max_search});
    std::vector<Bytes> outs(n_reads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
    for (long r = 0; r < (long)n_reads; r++)
        outs[(size_t)r] = correct_record(*s, ms, two_side != 0, seq + offsets[r], (size_t)(offsets[r + 1] - offsets[r]));
    bro_result *res = new bro_result;
    res->offsets.resize(n_reads + 1);
    uint64_t tot = 0;
    for (size_t r = 0; r < n_reads; r++) {
        res->offsets[r] = tot;
        tot += outs[r].size();
    }
    res->offsets[n_reads] = tot;
    res->data.resize(tot);
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
    for (long r = 0; r < (long)n_reads; r++)
        if (!outs[(size_t)r].empty())
            memcpy(res->data.data() + res->offsets[(size_t)r], outs[(size_t)r].data(), outs[(size_t)r].size());
    return res;
}
const uint8_t *bro_result_data(const bro_result *r) { return r->data.data(); }
const uint64_t *bro_result_offsets(const bro_result *r) { return r->offsets.data(); }
void bro_result_free(bro_result *r) { delete r; }

size_t bro_bio_global(const uint8_t *x, size_t m, const uint8_t *y, size_t n, uint8_t *ops, size_t cap) {
    auto v = bio::global(x, m, y, n);
    for (size_t i = 0; i < v.size() && i < cap; i++) ops[i] = (uint8_t)v[i];
    return v.size();
}

int bro_match_alignement(const uint8_t *before, size_t nb, const uint8_t *read, size_t nr, const uint8_t *corr,
                         size_t nc, long *off) {
    Bytes b(before, before + nb), c(corr, corr + nc);
    auto r = match_alignement(b, read, nr, c);
    if (!r) return 0;
    *off = (long)*r;
    return 1;
}

int bro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} /* extern "C" */
