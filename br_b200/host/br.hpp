// br.hpp — C++ host side of the hot path, mirroring br's own interface over the C ABI.
//
// br is Rust and this image has no Rust toolchain, so the host layer above include/brgpu.h is
// written in C++ with the reference's names, argument meaning and error behaviour:
//
//   br::set::KmerSet { get(kmer) -> bool, k() -> u8 }                    src/set.rs:17-23
//   br::set::Pcon    { from_pcon_solid, from_fasta, from_fastq, from_csv, from_count, new_ }
//                                                                        src/set/pcon.rs:13-196, count2solid src/main.rs:87-115
//   br::set::Hash    { from_fasta, from_fastq, from_csv }                src/set/hash.rs:14-186
//   br::correct::Corrector { valid_kmer, k, correct }                    src/correct/mod.rs:44-108
//   br::correct::{One, Two, Graph, Greedy, GapSize}                      src/correct/*.rs
//   br::build_methods(params, solid, confirm, max_search)                src/lib.rs:141-164
//   br::run_correction(inputs, outputs, methods, two_side, buffer_len)   src/lib.rs:72-139
//
// Nothing in this file computes a correction or a count on the CPU: every call lands in
// libbrgpu.so.  Without a CUDA device Context's constructor throws (BRGPU_E_NO_DEVICE).
// Corrector::correct_error is not exposed: on the GPU it only exists inside the scan kernels.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <future>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/brgpu.h"
#include "fasta.hpp"
#include "formats.hpp"

namespace br {

// error::Error (src/error.rs:12-45) carried as the ABI status + message
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};

inline const char *status_text(int st) {
    switch (st) {
    case BRGPU_E_INVALID: return "invalid argument";
    case BRGPU_E_NO_DEVICE: return "no CUDA device (brgpu has no CPU path)";
    case BRGPU_E_CUDA: return "CUDA error";
    case BRGPU_E_NOMEM: return "out of memory";
    case BRGPU_E_OVERFLOW: return "output buffer too small";
    case BRGPU_E_NO_THRESHOLD: return "Can't compute the abundance threshold";                  // src/error.rs:32-33
    case BRGPU_E_NEED_ABUNDANCE: return "You must provide an abundance threshold or a method";  // src/error.rs:36-37
    default: return "unknown status";
    }
}

class Context {
  public:
    explicit Context(int device = 0) {
        int st = brgpu_ctx_create(device, nullptr, &h_);
        if (st != BRGPU_OK) throw Error(st, status_text(st));
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ~Context() { brgpu_ctx_destroy(h_); }
    brgpu_ctx *handle() const { return h_; }
    void check(int st) const {
        if (st == BRGPU_OK) return;
        std::string msg = status_text(st);
        const char *d = brgpu_last_error(h_);
        if (d && *d) msg += std::string(" (") + d + ")";
        throw Error(st, msg);
    }

  private:
    brgpu_ctx *h_ = nullptr;
};

// page-locked host buffer (brgpu_host_alloc): what the chunk loop stages records in
class PinnedBuffer {
  public:
    explicit PinnedBuffer(const Context &ctx) : ctx_(&ctx) {}
    PinnedBuffer(const PinnedBuffer &) = delete;
    PinnedBuffer &operator=(const PinnedBuffer &) = delete;
    ~PinnedBuffer() { brgpu_host_free(ctx_->handle(), p_); }
    uint8_t *reserve(size_t bytes) {
        if (bytes > cap_) {
            brgpu_host_free(ctx_->handle(), p_);
            p_ = nullptr;
            cap_ = 0;
            size_t want = bytes + bytes / 4 + 4096;
            void *q = nullptr;
            ctx_->check(brgpu_host_alloc(ctx_->handle(), want, &q));
            p_ = q;
            cap_ = want;
        }
        return static_cast<uint8_t *>(p_);
    }
    uint8_t *data() const { return static_cast<uint8_t *>(p_); }
    size_t capacity() const { return cap_; }

  private:
    const Context *ctx_;
    void *p_ = nullptr;
    size_t cap_ = 0;
};

namespace kmer {
// cocktail::kmer as br uses it — host helpers for building k-mers to hand to
// KmerSet::get / Pcon::set, and for Pcon::get's index into the host mirror of the bitfield.
inline uint64_t nuc2bit(uint8_t b) { return (b >> 1) & 3u; }
inline uint64_t seq2bit(const uint8_t *s, size_t n) {
    uint64_t v = 0;
    for (size_t i = 0; i < n; i++) v = (v << 2) | nuc2bit(s[i]);
    return v;
}
inline uint64_t revcomp(uint64_t kmer, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((kmer & 3u) ^ 2u);
        kmer >>= 2;
    }
    return r;
}
inline uint64_t canonical(uint64_t kmer, int k) { // parity-canonical: the form with an even popcount
    return (__builtin_popcountll(kmer) & 1) ? revcomp(kmer, k) : kmer;
}
} // namespace kmer

namespace set {

// src/set.rs:17-21
class KmerSet {
  public:
    virtual ~KmerSet() = default;
    virtual bool get(uint64_t kmer) const = 0;
    virtual uint8_t k() const = 0;
};

// cli::AbundanceSelection (src/cli.rs:227-241); None = no method given
enum class AbundanceSelection { None, FirstMinimum, Rarefaction, PercentMost, PercentLeast };

// what the two device-resident sets share: the handle every corrector and batch call takes
class DeviceSet : public KmerSet {
  public:
    DeviceSet(const Context &ctx, brgpu_set *h) : ctx_(&ctx), h_(h) {}
    DeviceSet(const DeviceSet &) = delete;
    DeviceSet &operator=(const DeviceSet &) = delete;
    ~DeviceSet() override { brgpu_set_free(h_); }
    uint8_t k() const override { return (uint8_t)brgpu_set_k(h_); }
    std::vector<uint8_t> get_batch(const std::vector<uint64_t> &kmers) const {
        std::vector<uint8_t> out(kmers.size());
        ctx_->check(brgpu_set_get_batch(h_, kmers.data(), kmers.size(), out.data()));
        return out;
    }
    brgpu_set *handle() const { return h_; }
    const Context &context() const { return *ctx_; }

  protected:
    const Context *ctx_;
    brgpu_set *h_;
};

// upload chunk after chunk: `each(chunk)` is called once per chunk of at most `chunk_bases` bases
// (RecordReader: fasta::Reader or fastq::Reader — both fill the same Chunk)
template <class RecordReader = fasta::Reader, class F>
inline void for_each_chunk(const std::vector<std::string> &paths, size_t chunk_bases, F each) {
    fasta::Chunk c;
    for (auto &p : paths) {
        RecordReader r(p);
        bool more = true;
        while (more) {
            more = r.read_chunk(c, 8192); // appends
            if (c.seq.size() >= chunk_bases) {
                each(c);
                c.clear();
            }
        }
    }
    if (c.size()) each(c);
}

// from_csv's loop (src/set/pcon.rs:34-42, src/set/hash.rs:27-36): seq2bit of the first field of every data record,
// handed over in batches (a library call per k-mer would be a GPU round trip per line).  The reference feeds whatever
// the field holds to seq2bit; a field that is not k letters long cannot name a k-mer of the set (in the dense set it
// would index outside the bitfield), so it is refused here.
template <class F> inline void for_each_csv_batch(const std::string &path, int k, F each, size_t batch = 1u << 20) {
    csv::FirstColumn rows(path);
    std::vector<uint64_t> kmers;
    std::string field;
    size_t line = 1;
    while (rows.next(field)) {
        line++;
        if ((int)field.size() != k)
            throw std::runtime_error("csv record " + std::to_string(line) + ": the first column must hold a " + std::to_string(k) + "-mer");
        kmers.push_back(kmer::seq2bit((const uint8_t *)field.data(), field.size()));
        if (kmers.size() >= batch) {
            each(kmers);
            kmers.clear();
        }
    }
    if (!kmers.empty()) each(kmers);
}

// set::Pcon (src/set/pcon.rs:13-196): the dense canonical bitfield, resident in HBM
class Pcon : public DeviceSet {
  public:
    Pcon(const Context &ctx, brgpu_set *h) : DeviceSet(ctx, h) {}

    // Pcon::new(pcon::solid::Solid::new(k)) — empty set (src/set/pcon.rs:183)
    static std::unique_ptr<Pcon> new_(const Context &ctx, int k) {
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_new(ctx.handle(), k, &h));
        return std::make_unique<Pcon>(ctx, h);
    }

    // Pcon::from_pcon_solid (src/set/pcon.rs:18-25): (gzip) stream, byte 0 = k, rest = bitfield.
    // Decompression is host I/O (niffler in the reference); the payload goes to the GPU.
    static std::unique_ptr<Pcon> from_pcon_solid(const Context &ctx, const std::string &path) {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) throw std::runtime_error("can't open " + path);
        std::vector<uint8_t> payload;
        std::vector<uint8_t> buf(1u << 20);
        int n;
        while ((n = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) payload.insert(payload.end(), buf.begin(), buf.begin() + n);
        gzclose(gz);
        if (n < 0) throw std::runtime_error("read error in " + path);
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_from_solid_payload(ctx.handle(), payload.data(), payload.size(), &h));
        return std::make_unique<Pcon>(ctx, h);
    }

    // the `fasta` sub-command (src/main.rs:72-115): Counter::new(k) + count_fasta + count2solid.
    // abundance < 0 means "not given" (Option::None); an explicit abundance wins (src/main.rs:96).
    static std::unique_ptr<Pcon> from_count(const Context &ctx, const fasta::Chunk &reads, int k, int abundance,
                                            AbundanceSelection selection, double percent = 0.0) {
        k = k - (!(k & 1) & 1); // Fasta::kmer_size forces k odd (src/cli.rs:277-279)
        int sel = BRGPU_ABUNDANCE_EXPLICIT;
        if (abundance < 0) { // count2solid's match (src/main.rs:95-110): an explicit abundance wins
            switch (selection) {
            case AbundanceSelection::FirstMinimum: sel = BRGPU_ABUNDANCE_FIRST_MINIMUM; break;
            case AbundanceSelection::Rarefaction: sel = BRGPU_ABUNDANCE_RAREFACTION; break;
            case AbundanceSelection::PercentMost: sel = BRGPU_ABUNDANCE_PERCENT_AT_MOST; break;
            case AbundanceSelection::PercentLeast: sel = BRGPU_ABUNDANCE_PERCENT_AT_LEAST; break;
            case AbundanceSelection::None: break;
            }
        }
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_from_host_reads_ex(ctx.handle(), k, abundance, sel, percent, reads.seq.data(),
                                               reads.offsets.data(), reads.size(), &h));
        return std::make_unique<Pcon>(ctx, h);
    }

    // Pcon::from_fasta (src/set/pcon.rs:47-112): presence-only set of every canonical k-mer of the
    // records with len >= k — the counting pass with threshold `count > 0`.
    static std::unique_ptr<Pcon> from_fasta(const Context &ctx, const fasta::Chunk &reads, int k) {
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_from_host_reads(ctx.handle(), k, 0, BRGPU_ABUNDANCE_EXPLICIT, reads.seq.data(),
                                            reads.offsets.data(), reads.size(), &h));
        return std::make_unique<Pcon>(ctx, h);
    }

    // Pcon::from_fastq (src/set/pcon.rs:114-181, cargo feature `fastq`): from_fasta over FASTQ records
    static std::unique_ptr<Pcon> from_fastq(const Context &ctx, const std::string &path, int k) {
        fasta::Chunk reads;
        fastq::Reader r(path);
        while (r.read_chunk(reads, 1u << 20)) {}
        return from_fasta(ctx, reads, k);
    }

    // Pcon::from_csv (src/set/pcon.rs:27-45, cargo feature `csv`): Solid::new(k), then
    // set.set(seq2bit(record[0]), true) for every data record (the first record is the header)
    static std::unique_ptr<Pcon> from_csv(const Context &ctx, const std::string &path, int k) {
        std::unique_ptr<Pcon> out = new_(ctx, k);
        for_each_csv_batch(path, k, [&](const std::vector<uint64_t> &kmers) { out->set(kmers); });
        return out;
    }

    // The same over a stream of files (src/main.rs:72-78: count_fasta(inputs, 8192) reads the records chunk by
    // chunk): chunks of ~chunk_bases bases are uploaded, partitioned by table-index range and dropped; the set
    // is counted over all partitions at once.  Neither the host nor the device ever holds all the reads.
    static std::unique_ptr<Pcon> from_count_stream(const Context &ctx, const std::vector<std::string> &paths, int k,
                                                   int abundance, AbundanceSelection selection, double percent = 0.0,
                                                   size_t chunk_bases = 1u << 30) {
        k = k - (!(k & 1) & 1);
        int sel = BRGPU_ABUNDANCE_EXPLICIT;
        if (abundance < 0) {
            switch (selection) {
            case AbundanceSelection::FirstMinimum: sel = BRGPU_ABUNDANCE_FIRST_MINIMUM; break;
            case AbundanceSelection::Rarefaction: sel = BRGPU_ABUNDANCE_RAREFACTION; break;
            case AbundanceSelection::PercentMost: sel = BRGPU_ABUNDANCE_PERCENT_AT_MOST; break;
            case AbundanceSelection::PercentLeast: sel = BRGPU_ABUNDANCE_PERCENT_AT_LEAST; break;
            case AbundanceSelection::None: break;
            }
        }
        brgpu_set *h = nullptr;
        if (k < 15) { // small tables: the literal counter accumulates chunk by chunk
            brgpu_counts *c = nullptr;
            ctx.check(brgpu_counts_create(ctx.handle(), k, &c));
            try {
                for_each_chunk(paths, chunk_bases, [&](const fasta::Chunk &ch) {
                    brgpu_reads *r = nullptr;
                    ctx.check(brgpu_reads_upload(ctx.handle(), ch.seq.data(), ch.offsets.data(), ch.size(), &r));
                    int st = brgpu_counts_add_reads(c, r);
                    brgpu_reads_free(r);
                    ctx.check(st);
                });
                if (abundance < 0) {
                    if (sel == BRGPU_ABUNDANCE_EXPLICIT) ctx.check(BRGPU_E_NEED_ABUNDANCE);
                    uint64_t hist[256];
                    ctx.check(brgpu_counts_spectrum(c, hist));
                    abundance = brgpu_spectrum_threshold(hist, sel, percent);
                    if (abundance < 0) ctx.check(BRGPU_E_NO_THRESHOLD);
                }
                ctx.check(brgpu_set_from_counts(c, abundance, &h));
            } catch (...) {
                brgpu_counts_free(c);
                throw;
            }
            brgpu_counts_free(c);
            return std::make_unique<Pcon>(ctx, h);
        }
        std::vector<brgpu_kmers *> parts;
        auto drop = [&]() {
            for (auto p : parts) brgpu_kmers_free(p);
        };
        try {
            for_each_chunk(paths, chunk_bases, [&](const fasta::Chunk &ch) {
                brgpu_reads *r = nullptr;
                ctx.check(brgpu_reads_upload(ctx.handle(), ch.seq.data(), ch.offsets.data(), ch.size(), &r));
                brgpu_kmers *km = nullptr;
                int st = brgpu_kmers_create(ctx.handle(), k, r, &km);
                brgpu_reads_free(r);
                ctx.check(st);
                parts.push_back(km);
            });
            if (parts.empty()) { // no record at all: an empty chunk still gives the (all-zero) spectrum and set
                fasta::Chunk none;
                none.offsets.assign(1, 0);
                drop();
                return from_count(ctx, none, k, abundance, selection, percent);
            }
            ctx.check(brgpu_set_from_kmers(ctx.handle(), parts.data(), (int)parts.size(), abundance, sel, percent, &h));
        } catch (...) {
            drop();
            throw;
        }
        drop();
        return std::make_unique<Pcon>(ctx, h);
    }

    // The `count` sub-command (src/main.rs:59-70): pcon Counter::from_stream + count2solid.  pcon is not vendored
    // and no fixture of the reference holds a count file, so the container is restated as recalled from pcon
    // @0184ae77 (PARITY UNPINNED): one raw byte k, then the 2^(2k-1) u8 counters as a (multi-member) gzip stream;
    // a file that is gzip as a whole (niffler sniffing, src/cli.rs:400-420) is unwrapped first, and raw
    // counters after the k byte are accepted too.
    static std::unique_ptr<Pcon> from_pcon_count(const Context &ctx, const std::string &path, int abundance,
                                                 AbundanceSelection selection, double percent = 0.0) {
        std::vector<uint8_t> file = read_maybe_gzip(path);
        if (file.empty()) throw std::runtime_error("empty count file " + path);
        const int k = file[0];
        if (k < 3 || k > 19 || !(k & 1)) throw std::runtime_error("count file: k must be odd and in 3..=19");
        const uint64_t n = 1ULL << (2 * k - 1);
        std::vector<uint8_t> counts;
        if (file.size() >= 3 && file[1] == 0x1f && file[2] == 0x8b)
            counts = gunzip_members(file.data() + 1, file.size() - 1, n);
        else
            counts.assign(file.begin() + 1, file.end());
        if (counts.size() != n) throw std::runtime_error("count file: expected 2^(2k-1) counters");
        int sel = BRGPU_ABUNDANCE_EXPLICIT;
        if (abundance < 0) {
            switch (selection) {
            case AbundanceSelection::FirstMinimum: sel = BRGPU_ABUNDANCE_FIRST_MINIMUM; break;
            case AbundanceSelection::Rarefaction: sel = BRGPU_ABUNDANCE_RAREFACTION; break;
            case AbundanceSelection::PercentMost: sel = BRGPU_ABUNDANCE_PERCENT_AT_MOST; break;
            case AbundanceSelection::PercentLeast: sel = BRGPU_ABUNDANCE_PERCENT_AT_LEAST; break;
            case AbundanceSelection::None: break;
            }
        }
        brgpu_counts *c = nullptr;
        brgpu_set *h = nullptr;
        ctx.check(brgpu_counts_create(ctx.handle(), k, &c));
        try {
            ctx.check(brgpu_counts_upload(c, counts.data(), n));
            if (abundance < 0) { // count2solid's match (src/main.rs:95-110)
                if (sel == BRGPU_ABUNDANCE_EXPLICIT) ctx.check(BRGPU_E_NEED_ABUNDANCE);
                uint64_t hist[256];
                ctx.check(brgpu_counts_spectrum(c, hist));
                abundance = brgpu_spectrum_threshold(hist, sel, percent);
                if (abundance < 0) ctx.check(BRGPU_E_NO_THRESHOLD);
            }
            ctx.check(brgpu_set_from_counts(c, abundance, &h));
        } catch (...) {
            brgpu_counts_free(c);
            throw;
        }
        brgpu_counts_free(c);
        return std::make_unique<Pcon>(ctx, h);
    }

    // the whole file; gzip as a whole is unwrapped (gzread reads plain files transparently)
    static std::vector<uint8_t> read_maybe_gzip(const std::string &path) {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) throw std::runtime_error("can't open " + path);
        std::vector<uint8_t> data, buf(1u << 20);
        int n;
        while ((n = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) data.insert(data.end(), buf.begin(), buf.begin() + n);
        gzclose(gz);
        if (n < 0) throw std::runtime_error("read error in " + path);
        return data;
    }
    // concatenated gzip members -> at most `limit` bytes (flate2's MultiGzDecoder)
    static std::vector<uint8_t> gunzip_members(const uint8_t *src, size_t len, uint64_t limit) {
        std::vector<uint8_t> out;
        out.reserve((size_t)limit);
        z_stream z;
        std::memset(&z, 0, sizeof(z));
        if (inflateInit2(&z, 15 + 16) != Z_OK) throw std::runtime_error("zlib init failed");
        z.next_in = const_cast<uint8_t *>(src);
        z.avail_in = (uInt)std::min<size_t>(len, 0xffffffffu);
        size_t consumed_base = 0;
        std::vector<uint8_t> buf(1u << 20);
        for (;;) {
            z.next_out = buf.data();
            z.avail_out = (uInt)buf.size();
            int rc = inflate(&z, Z_NO_FLUSH);
            out.insert(out.end(), buf.data(), buf.data() + (buf.size() - z.avail_out));
            if (out.size() > limit) break;
            if (rc == Z_STREAM_END) { // next member, if any
                const size_t used = consumed_base + (z.next_in - (src + consumed_base));
                if (used >= len) break;
                consumed_base = used;
                inflateReset(&z);
                z.next_in = const_cast<uint8_t *>(src + used);
                z.avail_in = (uInt)std::min<size_t>(len - used, 0xffffffffu);
                continue;
            }
            if (rc == Z_BUF_ERROR && z.avail_in == 0) {
                const size_t used = (size_t)(z.next_in - src);
                if (used >= len) break;
                z.avail_in = (uInt)std::min<size_t>(len - used, 0xffffffffu);
                continue;
            }
            if (rc != Z_OK) {
                inflateEnd(&z);
                throw std::runtime_error("count file: corrupt gzip stream");
            }
        }
        inflateEnd(&z);
        return out;
    }

    // Pcon::get (src/set/pcon.rs:189-191), forward k-mers accepted.  Served from a host mirror of
    // the bitfield (exported once): a GPU round trip per k-mer would be useless.  Batches go
    // through get_batch.
    bool get(uint64_t kmer) const override {
        if (mirror_.empty()) mirror_ = bitfield();
        const int kk = k();
        const uint64_t mask = (kk < 32) ? ((1ULL << (2 * kk)) - 1ULL) : ~0ULL;
        const uint64_t idx = kmer::canonical(kmer & mask, kk) >> 1;
        return (mirror_[idx >> 3] >> (idx & 7)) & 1;
    }
    // Solid::set(kmer, true) on a batch (canonicalises, like the reference's unit tests rely on)
    void set(const std::vector<uint64_t> &kmers) {
        ctx_->check(brgpu_set_insert_batch(h_, kmers.data(), kmers.size()));
        mirror_.clear();
    }
    // `for kmer in Tokenizer::new(seq, k) { data.set(kmer, true) }` of the reference's tests
    void set_all_kmers(const std::string &seq) {
        const int kk = k();
        if ((int)seq.size() < kk) return;
        const uint64_t mask = (1ULL << (2 * kk)) - 1ULL;
        std::vector<uint64_t> kmers;
        uint64_t v = kmer::seq2bit((const uint8_t *)seq.data(), (size_t)kk - 1);
        for (size_t i = (size_t)kk - 1; i < seq.size(); i++) {
            v = ((v << 2) & mask) | kmer::nuc2bit((uint8_t)seq[i]);
            kmers.push_back(v);
        }
        set(kmers);
    }

    int abundance() const { return brgpu_set_abundance(h_); }
    std::vector<uint64_t> spectrum() const {
        std::vector<uint64_t> h(256);
        ctx_->check(brgpu_set_spectrum(h_, h.data()));
        return h;
    }
    std::vector<uint8_t> bitfield() const {
        std::vector<uint8_t> out(brgpu_set_bitfield_bytes(h_));
        ctx_->check(brgpu_set_export_bitfield(h_, out.data(), out.size()));
        return out;
    }
    // pcon's `.solid` container: gzip(u8 k || bitfield) — what from_pcon_solid reads back
    void write_solid(const std::string &path) const {
        std::vector<uint8_t> bits = bitfield();
        gzFile gz = gzopen(path.c_str(), "wb6");
        if (!gz) throw std::runtime_error("can't create " + path);
        uint8_t kk = k();
        bool ok = gzwrite(gz, &kk, 1) == 1;
        for (size_t p = 0; ok && p < bits.size();) {
            unsigned n = (unsigned)std::min<size_t>(bits.size() - p, 1u << 30);
            ok = gzwrite(gz, bits.data() + p, n) == (int)n;
            p += n;
        }
        ok = (gzclose(gz) == Z_OK) && ok;
        if (!ok) throw std::runtime_error("write error in " + path);
    }

  private:
    mutable std::vector<uint8_t> mirror_;
};

// set::Hash (src/set/hash.rs:14-186): canonical k-mers of any k <= 31 in a device hash table — the set
// behind the `large-kmer` sub-command (src/main.rs:147-163)
class Hash : public DeviceSet {
  public:
    Hash(const Context &ctx, brgpu_set *h) : DeviceSet(ctx, h) {}

    // Hash::from_fasta (src/set/hash.rs:41-100): presence of every canonical k-mer of every record with
    // len >= k, read chunk by chunk like the reference's 8192-record loop
    static std::unique_ptr<Hash> from_fasta(const Context &ctx, const std::vector<std::string> &paths, int k,
                                            size_t chunk_bases = 1u << 28) {
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_hash_new(ctx.handle(), k, 1u << 20, &h));
        auto out = std::make_unique<Hash>(ctx, h);
        for_each_chunk(paths, chunk_bases, [&](const fasta::Chunk &ch) {
            brgpu_reads *r = nullptr;
            ctx.check(brgpu_reads_upload(ctx.handle(), ch.seq.data(), ch.offsets.data(), ch.size(), &r));
            int st = brgpu_set_hash_add_reads(h, r);
            brgpu_reads_free(r);
            ctx.check(st);
        });
        return out;
    }
    // Hash::from_fastq (src/set/hash.rs:102-175, cargo feature `fastq`).  The reference's non-parallel build reads the
    // stream with the FASTA reader (hash.rs:110); its parallel build and Pcon::from_fastq use the FASTQ reader — that is
    // what this does.
    static std::unique_ptr<Hash> from_fastq(const Context &ctx, const std::vector<std::string> &paths, int k,
                                            size_t chunk_bases = 1u << 28) {
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_hash_new(ctx.handle(), k, 1u << 20, &h));
        auto out = std::make_unique<Hash>(ctx, h);
        for_each_chunk<fastq::Reader>(paths, chunk_bases, [&](const fasta::Chunk &ch) {
            brgpu_reads *r = nullptr;
            ctx.check(brgpu_reads_upload(ctx.handle(), ch.seq.data(), ch.offsets.data(), ch.size(), &r));
            int st = brgpu_set_hash_add_reads(h, r);
            brgpu_reads_free(r);
            ctx.check(st);
        });
        return out;
    }
    // Hash::from_csv (src/set/hash.rs:20-39, cargo feature `csv`): canonical(seq2bit(record[0]), k) of every data record
    static std::unique_ptr<Hash> from_csv(const Context &ctx, const std::string &path, int k) {
        std::unique_ptr<Hash> out = new_(ctx, k);
        for_each_csv_batch(path, k, [&](const std::vector<uint64_t> &kmers) { out->insert(kmers); });
        return out;
    }
    // an empty Hash and insertion of (forward or canonical) k-mers, for the unit KATs
    static std::unique_ptr<Hash> new_(const Context &ctx, int k) {
        brgpu_set *h = nullptr;
        ctx.check(brgpu_set_hash_new(ctx.handle(), k, 0, &h));
        return std::make_unique<Hash>(ctx, h);
    }
    void insert(const std::vector<uint64_t> &kmers) { ctx_->check(brgpu_set_insert_batch(h_, kmers.data(), kmers.size())); }
    // Hash::get (src/set/hash.rs:179-181); one k-mer per call is a GPU round trip: tests only
    bool get(uint64_t kmer) const override { return get_batch({kmer})[0] != 0; }
    uint64_t size() const { return brgpu_set_hash_size(h_); }
};

} // namespace set

namespace cli {
// cli::CorrectionMethod in declaration order (src/cli.rs:11-17) == the ABI's method ids
enum class CorrectionMethod : uint8_t { One = BRGPU_ONE, Two = BRGPU_TWO, Graph = BRGPU_GRAPH, Greedy = BRGPU_GREEDY, GapSize = BRGPU_GAP_SIZE };
} // namespace cli

namespace correct {

// src/correct/mod.rs:44-108.  A corrector is a (method, parameters, set) description; the scan
// itself runs in libbrgpu.so.  correct() is the unit-test entry point (one read per call);
// run_correction hands whole chunks to the library.
class Corrector {
  public:
    virtual ~Corrector() = default;
    const set::DeviceSet &valid_kmer() const { return *set_; }
    uint8_t k() const { return set_->k(); }
    cli::CorrectionMethod method() const { return method_; }
    int confirm() const { return confirm_; }
    int max_search() const { return max_search_; }

    std::vector<uint8_t> correct(const uint8_t *seq, size_t len) const {
        std::vector<uint8_t> out(2 * len + 256);
        for (;;) {
            uint64_t n = 0;
            int st = brgpu_correct_one(set_->context().handle(), set_->handle(), (int)method_, confirm_, max_search_, seq,
                                       len, out.data(), out.size(), &n);
            if (st == BRGPU_E_OVERFLOW && n > out.size()) {
                out.resize(n);
                continue;
            }
            set_->context().check(st);
            out.resize(n);
            return out;
        }
    }
    std::vector<uint8_t> correct(const std::string &seq) const { return correct((const uint8_t *)seq.data(), seq.size()); }

  protected:
    Corrector(const set::DeviceSet &s, cli::CorrectionMethod m, int confirm, int max_search)
        : set_(&s), method_(m), confirm_(confirm), max_search_(max_search) {}

  private:
    const set::DeviceSet *set_;
    cli::CorrectionMethod method_;
    int confirm_, max_search_;
};

struct One : Corrector { // src/correct/exist/one.rs:74
    One(const set::DeviceSet &s, uint8_t c) : Corrector(s, cli::CorrectionMethod::One, c, 7) {}
};
struct Two : Corrector { // src/correct/exist/two.rs:328
    Two(const set::DeviceSet &s, uint8_t c) : Corrector(s, cli::CorrectionMethod::Two, c, 7) {}
};
struct Graph : Corrector { // src/correct/graph.rs:29-37
    explicit Graph(const set::DeviceSet &s) : Corrector(s, cli::CorrectionMethod::Graph, 5, 7) {}
};
struct Greedy : Corrector { // src/correct/greedy.rs:41-54
    Greedy(const set::DeviceSet &s, uint8_t max_search, uint8_t nb_validate)
        : Corrector(s, cli::CorrectionMethod::Greedy, nb_validate, max_search) {}
};
struct GapSize : Corrector { // src/correct/gap_size.rs:29-42
    GapSize(const set::DeviceSet &s, uint8_t c) : Corrector(s, cli::CorrectionMethod::GapSize, c, 7) {}
};

} // namespace correct

using Methods = std::vector<std::unique_ptr<correct::Corrector>>;

// src/lib.rs:141-164 — same argument mapping
inline Methods build_methods(const std::vector<cli::CorrectionMethod> &params, const set::DeviceSet &solid, uint8_t confirm,
                             uint8_t max_search) {
    Methods methods;
    for (auto m : params) {
        switch (m) {
        case cli::CorrectionMethod::One: methods.emplace_back(new correct::One(solid, confirm)); break;
        case cli::CorrectionMethod::Two: methods.emplace_back(new correct::Two(solid, confirm)); break;
        case cli::CorrectionMethod::Graph: methods.emplace_back(new correct::Graph(solid)); break;
        case cli::CorrectionMethod::Greedy: methods.emplace_back(new correct::Greedy(solid, max_search, confirm)); break;
        case cli::CorrectionMethod::GapSize: methods.emplace_back(new correct::GapSize(solid, confirm)); break;
        }
    }
    return methods;
}

// the per-chunk body of run_correction (src/lib.rs:93-128) in one library call
struct Corrected {
    fasta::Bytes seq; // not value-initialised on resize: the library fills it
    std::vector<uint64_t> offsets;
};

enum class Transport { Ascii, Packed };

// a corrected chunk still in its 2-bit transport form (what comes back over PCIe)
struct CorrectedPacked {
    fasta::Bytes bases;
    std::vector<uint64_t> offsets, exc_pos;
    std::vector<uint8_t> exc_byte;
    uint64_t n_bases = 0, n_exc = 0;
};

// device leg of the packed transport: upload the packed chunk, correct, download packed
inline void correct_packed(const set::DeviceSet &solid, const std::vector<uint8_t> &ids, int confirm, int max_search, bool two_side,
                           const fasta::Chunk &in, const fasta::Packed &pk, CorrectedPacked &out) {
    const Context &ctx = solid.context();
    const size_t n = in.size(), total = in.seq.size();
    out.offsets.assign(n + 1, 0);
    brgpu_reads *r = nullptr, *c = nullptr;
    ctx.check(brgpu_reads_upload_packed(ctx.handle(), pk.bases.data(), in.offsets.data(), n, pk.exc_pos.data(), pk.exc_byte.data(),
                                        pk.exc_pos.size(), &r));
    int st = brgpu_correct_reads(ctx.handle(), solid.handle(), ids.data(), ids.size(), confirm, max_search, two_side ? 1 : 0, r, &c);
    brgpu_reads_free(r);
    ctx.check(st);
    out.bases.resize((total + total / 8 + 64 * n + 64) / 4 + 8);
    out.exc_pos.resize(pk.exc_pos.size() + 1);
    out.exc_byte.resize(pk.exc_pos.size() + 1);
    uint64_t counts[2] = {0, 0};
    for (;;) {
        st = brgpu_reads_download_packed(c, out.bases.data(), out.bases.size(), out.offsets.data(), out.exc_pos.data(),
                                         out.exc_byte.data(), pk.exc_pos.size(), counts);
        if (st == BRGPU_E_OVERFLOW && (counts[0] + 3) / 4 > out.bases.size()) {
            out.bases.resize((counts[0] + 3) / 4 + 8);
            continue;
        }
        break;
    }
    brgpu_reads_free(c);
    ctx.check(st);
    out.n_bases = counts[0];
    out.n_exc = counts[1];
}

inline void unpack_corrected(const CorrectedPacked &cp, fasta::Bytes &seq) {
    seq.resize(cp.n_bases);
    fasta::unpack(cp.bases.data(), cp.n_bases, cp.exc_pos.data(), cp.exc_byte.data(), cp.n_exc, seq.data());
}

// the (confirm, max_search) pair and the method ids of a chain, as the C ABI takes them
inline const set::DeviceSet &chain_parameters(const Methods &methods, std::vector<uint8_t> &ids, int &confirm, int &max_search) {
    const set::DeviceSet &solid = methods[0]->valid_kmer();
    // the C ABI takes one (confirm, max_search) pair for the chain, as br's command line does (build_methods
    // passes one -C and one -M, src/lib.rs:141-164); a hand-built chain that disagrees is refused
    confirm = -1, max_search = -1;
    ids.clear();
    for (auto &m : methods) {
        ids.push_back((uint8_t)m->method());
        if (&m->valid_kmer() != &solid) throw std::invalid_argument("all methods of a chain must share the set");
        if (m->method() != cli::CorrectionMethod::Graph) {
            if (confirm >= 0 && confirm != m->confirm()) throw std::invalid_argument("all methods of a chain must share confirm");
            confirm = m->confirm();
        }
        if (m->method() == cli::CorrectionMethod::Greedy) {
            if (max_search >= 0 && max_search != m->max_search()) throw std::invalid_argument("all Greedy methods of a chain must share max_search");
            max_search = m->max_search();
        }
    }
    if (confirm < 0) confirm = 5;
    if (max_search < 0) max_search = 7;
    return solid;
}

inline void correct_chunk(const Methods &methods, bool two_side, const fasta::Chunk &in, Corrected &out,
                          Transport transport = Transport::Packed) {
    const size_t n = in.size();
    out.offsets.assign(n + 1, 0);
    if (methods.empty()) { // fold over no method: records pass through
        out.seq = in.seq;
        out.offsets = in.offsets;
        return;
    }
    std::vector<uint8_t> ids;
    int confirm, max_search;
    const set::DeviceSet &solid = chain_parameters(methods, ids, confirm, max_search);
    const size_t total = in.seq.size();
    if (transport == Transport::Packed) {
        // 2-bit transport: the chunk crosses PCIe at 2 bits per base (+ the exception list) in both
        // directions; device-resident handles in between
        fasta::Packed pk;
        fasta::pack(in.seq.data(), total, pk);
        CorrectedPacked cp;
        correct_packed(solid, ids, confirm, max_search, two_side, in, pk, cp);
        out.offsets.swap(cp.offsets);
        unpack_corrected(cp, out.seq);
        return;
    }
    out.seq.resize(total + total / 8 + 64 * n + 64);
    for (;;) {
        uint64_t need = 0;
        int st = brgpu_correct_batch(solid.context().handle(), solid.handle(), ids.data(), ids.size(), confirm, max_search,
                                     two_side ? 1 : 0, in.seq.data(), in.offsets.data(), n, out.seq.data(),
                                     out.seq.size(), out.offsets.data(), &need);
        if (st == BRGPU_E_OVERFLOW && need > out.seq.size()) {
            out.seq.resize(need);
            continue;
        }
        solid.context().check(st);
        out.seq.resize(need);
        return;
    }
}

constexpr size_t CHUNK_RECORDS = 8192; // hard-coded in src/lib.rs:90

// src/lib.rs:72-139: for each (input, output) pair read FASTA records, correct them chunk by
// chunk (8192 records), write FASTA in input order (the serial path's order, src/lib.rs:21-69).
// `record_buffer_len` is accepted and, like in the reference, only a capacity hint.  The parse of
// chunk c+1 and the formatting of chunk c-1 run on host threads while the GPU corrects chunk c.
inline void run_correction(const std::vector<std::string> &inputs, const std::vector<std::string> &outputs,
                           const Methods &methods, bool two_side, uint64_t record_buffer_len = CHUNK_RECORDS,
                           Transport transport = Transport::Packed) {
    (void)record_buffer_len;
    const size_t pairs = inputs.size() < outputs.size() ? inputs.size() : outputs.size(); // zip (src/lib.rs:79)
    const bool packed = transport == Transport::Packed && !methods.empty();
    std::vector<uint8_t> ids;
    int confirm = 5, max_search = 7;
    const set::DeviceSet *solid = methods.empty() ? nullptr : &chain_parameters(methods, ids, confirm, max_search);
    for (size_t p = 0; p < pairs; p++) {
        fasta::Reader reader(inputs[p]);
        fasta::Writer writer(outputs[p]);
        // Three stages, two buffers each.  Reader thread: parse chunk c+1 and (2-bit transport) pack it.  This thread:
        // upload, correct, download chunk c — nothing but the library calls, so the GPU is not left waiting for host
        // passes over the bases.  Writer thread: unpack and format chunk c-1, in input order.
        fasta::Chunk chunks[2];
        fasta::Packed packs[2];
        CorrectedPacked cps[2];
        Corrected results[2];
        std::vector<std::string> written_defs[2];
        auto read_into = [&reader, packed](fasta::Chunk *c, fasta::Packed *pk) {
            c->clear();
            const bool more = reader.read_chunk(*c, CHUNK_RECORDS);
            if (packed && c->size()) fasta::pack(c->seq.data(), c->seq.size(), *pk);
            return more;
        };
        std::future<bool> next = std::async(std::launch::async, read_into, &chunks[0], &packs[0]);
        std::shared_future<void> writes[2]; // writes[b]: the write that still reads results[b] / cps[b]
        for (int cur = 0;; cur ^= 1) {
            const bool more = next.get();
            fasta::Chunk &c = chunks[cur];
            // chunks[cur ^ 1] and packs[cur ^ 1] went through the (synchronous) device calls one turn ago: both are free
            if (more) next = std::async(std::launch::async, read_into, &chunks[cur ^ 1], &packs[cur ^ 1]);
            if (c.size()) {
                if (writes[cur].valid()) writes[cur].wait(); // results[cur] / cps[cur] are free again
                if (packed) correct_packed(*solid, ids, confirm, max_search, two_side, c, packs[cur], cps[cur]);
                else correct_chunk(methods, two_side, c, results[cur], transport);
                written_defs[cur].swap(c.definitions);
                Corrected *r = &results[cur];
                CorrectedPacked *cp = &cps[cur];
                std::vector<std::string> *d = &written_defs[cur];
                std::shared_future<void> before = writes[cur ^ 1]; // keep the records in input order
                writes[cur] = std::async(std::launch::async, [&writer, r, cp, d, before, packed]() {
                                  if (packed) {
                                      unpack_corrected(*cp, r->seq);
                                      r->offsets.swap(cp->offsets);
                                  }
                                  if (before.valid()) before.wait();
                                  writer.write(*d, r->seq.data(), r->offsets.data());
                              }).share();
            }
            if (!more) break;
        }
        for (auto &w : writes)
            if (w.valid()) w.get();
    }
}

} // namespace br
