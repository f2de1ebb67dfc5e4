// brgpu_double.cpp — TEST DOUBLE of include/brgpu.h.  TEST INFRASTRUCTURE ONLY.
//
// What it is for: the C++ host side (br_b200/host/{br.hpp, fasta.hpp, formats.hpp, cli.cpp, kat_runner.cpp}) is the
// caller of the C ABI — argument handling, the streamed chunk loops of the set builders, the three-stage
// run_correction pipeline, the 2-bit transport, the FASTA / FASTQ / CSV framing.  The driver's CPU stage has no GPU,
// so tests/test_host_cli_double_cpu.py links the UNMODIFIED host sources against this file instead of libbrgpu.so and
// runs the command-line tests there: what is under test is the host logic and its use of the ABI contract (buffer
// sizes, E_OVERFLOW + required, offsets, exception lists, status codes), under sanitizers too.
//
// What it is not: a CPU path of the product.  It lives under tests/, is never built by br_b200/build.py or
// __graft_entry__.build(), is never named libbrgpu.so, and nothing in br_b200/, include/ or bench.py refers to it.
// Every entry point the host calls is implemented over oracle/br_oracle.h (the CPU restatement, itself test
// infrastructure); results therefore say nothing about the CUDA kernels — that is what `pytest -m gpu` is for.
// Entry points the C++ host never calls are not defined (the link would fail if it started to).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/brgpu.h"
#include "../../oracle/br_oracle.h"

struct brgpu_ctx {
    std::string err;
};
struct brgpu_reads {
    brgpu_ctx *ctx;
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off; // n + 1
};
struct brgpu_counts {
    brgpu_ctx *ctx;
    int k;
    bro_counter *c;
};
struct brgpu_kmers { // a chunk's k-mers: the double keeps the chunk itself
    brgpu_ctx *ctx;
    int k;
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off;
};
struct brgpu_set {
    brgpu_ctx *ctx;
    bro_set *s;
    int k;
    int abundance;
    bool is_hash;
    uint64_t hist[256];
};
struct brgpu_group {
    std::vector<brgpu_ctx *> ctx;
    std::string err;
};

namespace {

int fail(brgpu_ctx *ctx, int st, const char *what) {
    if (ctx) ctx->err = what;
    return st;
}
bool dense_k(int k) { return k >= 3 && k <= 19 && (k & 1); }
int host_threads() {
#ifdef BRGPU_DOUBLE_SERIAL // sanitizer builds: libgomp's worker threads are not instrumented (false race reports)
    return 1;
#else
    const int t = bro_max_threads();
    return t > 8 ? 8 : (t < 1 ? 1 : t);
#endif
}
brgpu_set *wrap(brgpu_ctx *ctx, bro_set *s, int k, int abundance, bool is_hash, const uint64_t *hist) {
    brgpu_set *out = new brgpu_set{ctx, s, k, abundance, is_hash, {0}};
    if (hist) std::memcpy(out->hist, hist, sizeof(out->hist));
    return out;
}
// count2solid (src/main.rs:87-115) on a finished counter
int solid_from_counter(brgpu_ctx *ctx, bro_counter *c, int k, int abundance, int selection, double percent, brgpu_set **out) {
    uint64_t hist[256];
    bro_spectrum(c, hist, host_threads());
    if (abundance < 0) {
        if (selection == BRGPU_ABUNDANCE_EXPLICIT) return fail(ctx, BRGPU_E_NEED_ABUNDANCE, "need an abundance threshold or an abundance method");
        abundance = brgpu_spectrum_threshold(hist, selection, percent);
        if (abundance < 0) return fail(ctx, BRGPU_E_NO_THRESHOLD, "no abundance threshold");
    }
    *out = wrap(ctx, bro_solid_from_count(c, abundance, host_threads()), k, abundance, false, hist);
    return BRGPU_OK;
}

} // namespace

extern "C" {

int brgpu_ctx_create(int, void *, brgpu_ctx **out) {
    if (!out) return BRGPU_E_INVALID;
    *out = new brgpu_ctx();
    return BRGPU_OK;
}
void brgpu_ctx_destroy(brgpu_ctx *ctx) { delete ctx; }
const char *brgpu_last_error(const brgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
const char *brgpu_version(void) { return "brgpu ABI test double (tests/abi_double): not the product"; }
int brgpu_host_alloc(brgpu_ctx *, size_t bytes, void **out) {
    *out = std::malloc(bytes ? bytes : 1);
    return *out ? BRGPU_OK : BRGPU_E_NOMEM;
}
void brgpu_host_free(brgpu_ctx *, void *p) { std::free(p); }

// ---- reads ----
int brgpu_reads_upload(brgpu_ctx *ctx, const uint8_t *seq, const uint64_t *off, uint64_t n, brgpu_reads **out) {
    if (!ctx || !out || !off || (n && off[n] && !seq)) return BRGPU_E_INVALID;
    brgpu_reads *r = new brgpu_reads{ctx, {}, {}};
    r->off.assign(off, off + n + 1);
    r->seq.assign(seq + off[0], seq + off[n]);
    for (auto &o : r->off) o -= off[0];
    *out = r;
    return BRGPU_OK;
}
void brgpu_reads_free(brgpu_reads *r) { delete r; }

int brgpu_reads_upload_packed(brgpu_ctx *ctx, const uint8_t *packed, const uint64_t *off, uint64_t n, const uint64_t *exc_pos,
                              const uint8_t *exc_byte, uint64_t n_exc, brgpu_reads **out) {
    if (!ctx || !out || !off) return BRGPU_E_INVALID;
    if (n && off[0] != 0) return fail(ctx, BRGPU_E_INVALID, "packed reads: offsets must start at 0");
    static const char LETTER[4] = {'A', 'C', 'T', 'G'};
    brgpu_reads *r = new brgpu_reads{ctx, {}, {}};
    r->off.assign(off, off + n + 1);
    const uint64_t total = off[n];
    r->seq.resize(total);
    for (uint64_t t = 0; t < total; t++) r->seq[t] = (uint8_t)LETTER[(packed[t >> 2] >> (2 * (3 - (t & 3)))) & 3];
    for (uint64_t i = 0; i < n_exc; i++) {
        if (exc_pos[i] >= total) {
            delete r;
            return fail(ctx, BRGPU_E_INVALID, "packed reads: exception beyond the last base");
        }
        r->seq[exc_pos[i]] = exc_byte[i];
    }
    *out = r;
    return BRGPU_OK;
}

int brgpu_reads_download_packed(brgpu_reads *r, uint8_t *packed, uint64_t packed_cap, uint64_t *off, uint64_t *exc_pos,
                                uint8_t *exc_byte, uint64_t exc_cap, uint64_t counts[2]) {
    if (!r || !counts) return BRGPU_E_INVALID;
    static const uint8_t LETTER[4] = {'A', 'C', 'T', 'G'};
    const uint64_t total = r->seq.size();
    uint64_t n_exc = 0;
    for (uint64_t t = 0; t < total; t++) n_exc += r->seq[t] != LETTER[(r->seq[t] >> 1) & 3];
    counts[0] = total;
    counts[1] = n_exc;
    if ((total + 3) / 4 > packed_cap) return fail(r->ctx, BRGPU_E_OVERFLOW, "packed buffer too small");
    if (n_exc > exc_cap) return fail(r->ctx, BRGPU_E_OVERFLOW, "exception buffer too small");
    std::memset(packed, 0, (size_t)((total + 3) / 4));
    uint64_t e = 0;
    for (uint64_t t = 0; t < total; t++) {
        const uint8_t b = r->seq[t], code = (b >> 1) & 3;
        packed[t >> 2] |= (uint8_t)(code << (2 * (3 - (t & 3))));
        if (b != LETTER[code]) {
            exc_pos[e] = t;
            exc_byte[e++] = b;
        }
    }
    std::memcpy(off, r->off.data(), r->off.size() * sizeof(uint64_t));
    return BRGPU_OK;
}

// ---- part 1 ----
int brgpu_counts_create(brgpu_ctx *ctx, int k, brgpu_counts **out) {
    if (!ctx || !out) return BRGPU_E_INVALID;
    if (!dense_k(k)) return fail(ctx, BRGPU_E_INVALID, "k must be odd and in 3..=19");
    *out = new brgpu_counts{ctx, k, bro_counter_new(k)};
    return BRGPU_OK;
}
int brgpu_counts_add_reads(brgpu_counts *c, const brgpu_reads *r) {
    if (!c || !r) return BRGPU_E_INVALID;
    bro_counter_count(c->c, r->seq.data(), r->off.data(), r->off.size() - 1, host_threads());
    return BRGPU_OK;
}
int brgpu_counts_spectrum(brgpu_counts *c, uint64_t hist[256]) {
    if (!c || !hist) return BRGPU_E_INVALID;
    bro_spectrum(c->c, hist, host_threads());
    return BRGPU_OK;
}
int brgpu_counts_upload(brgpu_counts *c, const uint8_t *counts, uint64_t n) {
    if (!c || !counts) return BRGPU_E_INVALID;
    size_t have = 0;
    uint8_t *raw = const_cast<uint8_t *>(bro_counter_raw(c->c, &have));
    if (n != have) return fail(c->ctx, BRGPU_E_INVALID, "n must be 2^(2k-1)");
    std::memcpy(raw, counts, (size_t)n);
    return BRGPU_OK;
}
void brgpu_counts_free(brgpu_counts *c) {
    if (!c) return;
    bro_counter_free(c->c);
    delete c;
}
int brgpu_spectrum_threshold(const uint64_t hist[256], int selection, double percent) {
    if (selection == BRGPU_ABUNDANCE_FIRST_MINIMUM) return bro_first_minimum(hist);
    if (selection >= BRGPU_ABUNDANCE_RAREFACTION && selection <= BRGPU_ABUNDANCE_PERCENT_AT_LEAST)
        return bro_spectrum_threshold(hist, selection, percent);
    return -1;
}
int brgpu_set_from_counts(brgpu_counts *c, int abundance, brgpu_set **out) {
    if (!c || !out || abundance < 0) return BRGPU_E_INVALID;
    return solid_from_counter(c->ctx, c->c, c->k, abundance, BRGPU_ABUNDANCE_EXPLICIT, 0.0, out);
}
int brgpu_set_from_host_reads_ex(brgpu_ctx *ctx, int k, int abundance, int selection, double percent, const uint8_t *seq,
                                 const uint64_t *off, uint64_t n, brgpu_set **out) {
    if (!ctx || !out || !off) return BRGPU_E_INVALID;
    if (!dense_k(k)) return fail(ctx, BRGPU_E_INVALID, "k must be odd and in 3..=19");
    bro_counter *c = bro_counter_new(k);
    bro_counter_count(c, seq, off, n, host_threads());
    const int st = solid_from_counter(ctx, c, k, abundance, selection, percent, out);
    bro_counter_free(c);
    return st;
}
int brgpu_set_from_host_reads(brgpu_ctx *ctx, int k, int abundance, int selection, const uint8_t *seq, const uint64_t *off, uint64_t n,
                              brgpu_set **out) {
    return brgpu_set_from_host_reads_ex(ctx, k, abundance, selection, 0.0, seq, off, n, out);
}
int brgpu_kmers_create(brgpu_ctx *ctx, int k, const brgpu_reads *r, brgpu_kmers **out) {
    if (!ctx || !r || !out) return BRGPU_E_INVALID;
    if (!dense_k(k) || k < 15) return fail(ctx, BRGPU_E_INVALID, "bucketed k-mers need odd k in 15..=19");
    *out = new brgpu_kmers{ctx, k, r->seq, r->off};
    return BRGPU_OK;
}
void brgpu_kmers_free(brgpu_kmers *p) { delete p; }
int brgpu_set_from_kmers(brgpu_ctx *ctx, brgpu_kmers *const *parts, int n_parts, int abundance, int selection, double percent,
                         brgpu_set **out) {
    if (!ctx || !parts || n_parts < 1 || n_parts > 64 || !out) return BRGPU_E_INVALID;
    const int k = parts[0]->k;
    bro_counter *c = bro_counter_new(k);
    for (int i = 0; i < n_parts; i++) bro_counter_count(c, parts[i]->seq.data(), parts[i]->off.data(), parts[i]->off.size() - 1, host_threads());
    const int st = solid_from_counter(ctx, c, k, abundance, selection, percent, out);
    bro_counter_free(c);
    return st;
}
int brgpu_set_from_solid_payload(brgpu_ctx *ctx, const uint8_t *payload, uint64_t n, brgpu_set **out) {
    if (!ctx || !payload || !out || n < 1) return BRGPU_E_INVALID;
    const int k = payload[0];
    if (!dense_k(k)) return fail(ctx, BRGPU_E_INVALID, "k must be odd and in 3..=19");
    if (n - 1 != (1ULL << (2 * k - 1)) / 8) return fail(ctx, BRGPU_E_INVALID, "payload size does not match k");
    *out = wrap(ctx, bro_set_from_bitfield(k, payload + 1, (size_t)(n - 1)), k, -1, false, nullptr);
    return BRGPU_OK;
}
int brgpu_set_new(brgpu_ctx *ctx, int k, brgpu_set **out) {
    if (!ctx || !out) return BRGPU_E_INVALID;
    if (!dense_k(k)) return fail(ctx, BRGPU_E_INVALID, "k must be odd and in 3..=19");
    *out = wrap(ctx, bro_set_new(k), k, -1, false, nullptr);
    return BRGPU_OK;
}
int brgpu_set_insert_batch(brgpu_set *s, const uint64_t *kmers, uint64_t n) {
    if (!s || (!kmers && n)) return BRGPU_E_INVALID;
    for (uint64_t i = 0; i < n; i++) bro_set_set(s->s, kmers[i], 1);
    return BRGPU_OK;
}
int brgpu_set_hash_new(brgpu_ctx *ctx, int k, uint64_t, brgpu_set **out) {
    if (!ctx || !out) return BRGPU_E_INVALID;
    if (k < 3 || k > 31) return fail(ctx, BRGPU_E_INVALID, "hash sets hold k-mers with 3 <= k <= 31");
    *out = wrap(ctx, bro_hash_new(k), k, -1, true, nullptr);
    return BRGPU_OK;
}
int brgpu_set_hash_add_reads(brgpu_set *s, const brgpu_reads *r) {
    if (!s || !r) return BRGPU_E_INVALID;
    if (!s->is_hash) return fail(s->ctx, BRGPU_E_INVALID, "not a hash set");
    bro_hash_add_reads(s->s, r->seq.data(), r->off.data(), r->off.size() - 1);
    return BRGPU_OK;
}
uint64_t brgpu_set_hash_size(const brgpu_set *s) { return s && s->is_hash ? bro_hash_size(s->s) : 0; }
int brgpu_set_k(const brgpu_set *s) { return s ? s->k : 0; }
int brgpu_set_abundance(const brgpu_set *s) { return s ? s->abundance : -1; }
uint64_t brgpu_set_bitfield_bytes(const brgpu_set *s) { return s && !s->is_hash ? (1ULL << (2 * s->k - 1)) / 8 : 0; }
int brgpu_set_export_bitfield(brgpu_set *s, uint8_t *out, uint64_t cap) {
    if (!s || !out) return BRGPU_E_INVALID;
    if (s->is_hash) return fail(s->ctx, BRGPU_E_INVALID, "a hash set has no bitfield");
    size_t n = 0;
    const uint8_t *bits = bro_set_bits(s->s, &n);
    if (cap < n) return fail(s->ctx, BRGPU_E_OVERFLOW, "bitfield buffer too small");
    std::memcpy(out, bits, n);
    return BRGPU_OK;
}
int brgpu_set_get_batch(brgpu_set *s, const uint64_t *kmers, uint64_t n, uint8_t *out) {
    if (!s || (n && (!kmers || !out))) return BRGPU_E_INVALID;
    bro_set_get_batch(s->s, kmers, (size_t)n, out);
    return BRGPU_OK;
}
int brgpu_set_spectrum(const brgpu_set *s, uint64_t hist[256]) {
    if (!s || !hist) return BRGPU_E_INVALID;
    std::memcpy(hist, s->hist, sizeof(s->hist));
    return BRGPU_OK;
}
void brgpu_set_free(brgpu_set *s) {
    if (!s) return;
    bro_set_free(s->s);
    delete s;
}

// ---- part 2 ----
static int check_chain(brgpu_ctx *ctx, const uint8_t *methods, uint64_t n_methods, int confirm, int max_search) {
    if (confirm < 1 || confirm > 255) return fail(ctx, BRGPU_E_INVALID, "confirm must be in 1..=255");
    if (max_search < 0 || max_search > 255) return fail(ctx, BRGPU_E_INVALID, "max_search must be in 0..=255");
    for (uint64_t i = 0; i < n_methods; i++)
        if (methods[i] > BRGPU_GAP_SIZE) return fail(ctx, BRGPU_E_INVALID, "unknown method");
    return BRGPU_OK;
}
int brgpu_correct_reads(brgpu_ctx *ctx, const brgpu_set *set, const uint8_t *methods, uint64_t n_methods, int confirm, int max_search,
                        int two_side, const brgpu_reads *in, brgpu_reads **out) {
    if (!ctx || !set || !in || !out || (!methods && n_methods)) return BRGPU_E_INVALID;
    const int st = check_chain(ctx, methods, n_methods, confirm, max_search);
    if (st != BRGPU_OK) return st;
    const size_t n = in->off.size() - 1;
    bro_result *res = bro_run_correction(set->s, methods, (size_t)n_methods, confirm, max_search, two_side, in->seq.data(), in->off.data(),
                                         n, host_threads());
    brgpu_reads *r = new brgpu_reads{ctx, {}, {}};
    const uint64_t *ro = bro_result_offsets(res);
    r->off.assign(ro, ro + n + 1);
    r->seq.assign(bro_result_data(res), bro_result_data(res) + ro[n]);
    bro_result_free(res);
    *out = r;
    return BRGPU_OK;
}
int brgpu_correct_batch(brgpu_ctx *ctx, const brgpu_set *set, const uint8_t *methods, uint64_t n_methods, int confirm, int max_search,
                        int two_side, const uint8_t *seq, const uint64_t *off, uint64_t n, uint8_t *out, uint64_t out_cap,
                        uint64_t *out_off, uint64_t *required) {
    brgpu_reads *in = nullptr, *res = nullptr;
    int st = brgpu_reads_upload(ctx, seq, off, n, &in);
    if (st != BRGPU_OK) return st;
    st = brgpu_correct_reads(ctx, set, methods, n_methods, confirm, max_search, two_side, in, &res);
    brgpu_reads_free(in);
    if (st != BRGPU_OK) return st;
    if (required) *required = res->seq.size();
    if (res->seq.size() > out_cap) {
        brgpu_reads_free(res);
        return fail(ctx, BRGPU_E_OVERFLOW, "output buffer too small");
    }
    std::memcpy(out, res->seq.data(), res->seq.size());
    std::memcpy(out_off, res->off.data(), res->off.size() * sizeof(uint64_t));
    brgpu_reads_free(res);
    return BRGPU_OK;
}
int brgpu_correct_one(brgpu_ctx *ctx, const brgpu_set *set, int method, int confirm, int max_search, const uint8_t *seq, uint64_t len,
                      uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    if (!ctx || !set || !out_len) return BRGPU_E_INVALID;
    const uint8_t m = (uint8_t)method;
    const int st = check_chain(ctx, &m, 1, confirm, max_search);
    if (st != BRGPU_OK) return st;
    const size_t n = bro_correct(set->s, method, confirm, max_search, seq, (size_t)len, out, (size_t)out_cap);
    *out_len = n;
    return n > out_cap ? fail(ctx, BRGPU_E_OVERFLOW, "output buffer too small") : BRGPU_OK;
}

// ---- several GPUs in one process (brgpu_group_*): N contexts of the double stand for N devices, so that the host
// side of `brgpu-cli -d 0,1,...` (its own chunk loop, the replicas, the error path) runs on the CPU stage too ----
int brgpu_group_create(const int *devices, int n, brgpu_group **out) {
    if (!devices || n < 1 || !out) return BRGPU_E_INVALID;
    brgpu_group *g = new brgpu_group();
    for (int i = 0; i < n; i++) g->ctx.push_back(new brgpu_ctx());
    *out = g;
    return BRGPU_OK;
}
void brgpu_group_destroy(brgpu_group *g) {
    if (!g) return;
    for (auto c : g->ctx) delete c;
    delete g;
}
int brgpu_group_size(const brgpu_group *g) { return g ? (int)g->ctx.size() : 0; }
const char *brgpu_group_last_error(const brgpu_group *g) { return g ? g->err.c_str() : ""; }
int brgpu_group_set_from_host_reads(brgpu_group *g, int k, int abundance, int selection, double percent, const uint8_t *seq,
                                    const uint64_t *off, uint64_t n, brgpu_set **out_sets) {
    if (!g || !out_sets) return BRGPU_E_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++) { // one replica per device
        const int st = brgpu_set_from_host_reads_ex(g->ctx[i], k, abundance, selection, percent, seq, off, n, &out_sets[i]);
        if (st != BRGPU_OK) {
            g->err = g->ctx[i]->err;
            for (size_t j = 0; j < i; j++) {
                brgpu_set_free(out_sets[j]);
                out_sets[j] = nullptr;
            }
            return st;
        }
    }
    return BRGPU_OK;
}
void brgpu_group_sets_free(brgpu_group *g, brgpu_set **sets) {
    if (!g || !sets) return;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        brgpu_set_free(sets[i]);
        sets[i] = nullptr;
    }
}
int brgpu_group_correct_batch(brgpu_group *g, brgpu_set *const *sets, const uint8_t *methods, uint64_t n_methods, int confirm, int max_search,
                              int two_side, const uint8_t *seq, const uint64_t *off, uint64_t n, uint8_t *out, uint64_t out_cap,
                              uint64_t *out_off, uint64_t *required) {
    if (!g || !sets || !off) return BRGPU_E_INVALID;
    const size_t devs = g->ctx.size();
    std::vector<brgpu_reads *> parts(devs, nullptr);
    uint64_t total = 0, first = 0;
    int st = BRGPU_OK;
    for (size_t d = 0; d < devs && st == BRGPU_OK; d++) { // contiguous record ranges balanced by bases, one per device
        uint64_t last = first;
        const uint64_t want = off[0] + (off[n] - off[0]) * (d + 1) / devs;
        while (last < n && (d + 1 == devs || off[last + 1] <= want)) last++;
        brgpu_reads *in = nullptr;
        st = brgpu_reads_upload(g->ctx[d], seq, off + first, last - first, &in);
        if (st == BRGPU_OK) st = brgpu_correct_reads(g->ctx[d], sets[d], methods, n_methods, confirm, max_search, two_side, in, &parts[d]);
        brgpu_reads_free(in);
        if (st != BRGPU_OK) g->err = g->ctx[d]->err;
        else total += parts[d]->seq.size();
        first = last;
    }
    if (st == BRGPU_OK) {
        if (required) *required = total;
        if (total > out_cap) {
            g->err = "output buffer too small";
            st = BRGPU_E_OVERFLOW;
        }
    }
    if (st == BRGPU_OK) {
        uint64_t at = 0, r = 0;
        out_off[0] = 0;
        for (size_t d = 0; d < devs; d++) {
            std::memcpy(out + at, parts[d]->seq.data(), parts[d]->seq.size());
            for (size_t i = 1; i < parts[d]->off.size(); i++) out_off[++r] = at + parts[d]->off[i];
            at += parts[d]->seq.size();
        }
    }
    for (auto p : parts) brgpu_reads_free(p);
    return st;
}

} // extern "C"
