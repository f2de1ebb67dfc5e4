// fasta.hpp — host-side FASTA record stream for run_correction and the set builders.
//
// Stands in for the pieces of noodles-fasta 0.38 / niffler 2.5 that br's hot path is fed by
// (src/lib.rs:30-35, :80-91, :168-188; src/cli.rs:265-274, :400-420).  Those crates are not
// vendored in the reference tree, so the framing is the published FASTA convention:
//   reader: a record is a `>` definition line followed by sequence lines up to the next `>`
//           (line ends stripped, `\r\n` tolerated); gzip input is sniffed like niffler does;
//   writer: `>` + definition + `\n`, then the sequence wrapped at 80 columns (noodles' default).
// Parity tests compare sequences per record, not file bytes ("FASTA framing").
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace br {
namespace fasta {

constexpr size_t LINE_BASES = 80;

// Growable byte buffer without value initialisation (std::vector<uint8_t>::resize zero-fills, which
// doubles the memory traffic of a parser that only appends).
class Bytes {
  public:
    Bytes() = default;
    Bytes(const Bytes &o) { append(o.p_, o.n_); }
    Bytes &operator=(const Bytes &o) {
        if (this != &o) {
            n_ = 0;
            append(o.p_, o.n_);
        }
        return *this;
    }
    Bytes(Bytes &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr, o.n_ = o.cap_ = 0; }
    Bytes &operator=(Bytes &&o) noexcept {
        if (this != &o) {
            std::free(p_);
            p_ = o.p_, n_ = o.n_, cap_ = o.cap_;
            o.p_ = nullptr, o.n_ = o.cap_ = 0;
        }
        return *this;
    }
    ~Bytes() { std::free(p_); }
    uint8_t *data() { return p_; }
    const uint8_t *data() const { return p_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    void clear() { n_ = 0; }
    void reserve(size_t cap) {
        if (cap <= cap_) return;
        size_t want = cap_ ? cap_ : (size_t)1 << 16;
        while (want < cap) want += want / 2 + 4096;
        void *q = std::realloc(p_, want);
        if (!q) throw std::bad_alloc();
        p_ = static_cast<uint8_t *>(q);
        cap_ = want;
    }
    void resize(size_t n) { // new bytes are NOT initialised
        reserve(n);
        n_ = n;
    }
    void append(const void *src, size_t n) {
        if (!n) return;
        reserve(n_ + n);
        std::memcpy(p_ + n_, src, n);
        n_ += n;
    }
    void push_back(uint8_t b) {
        reserve(n_ + 1);
        p_[n_++] = b;
    }
    uint8_t back() const { return p_[n_ - 1]; }
    void pop_back() { n_--; }

  private:
    uint8_t *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

// One buffer of records (populate_buffer, src/lib.rs:168-188): sequences concatenated, with
// prefix-sum offsets — exactly what brgpu_correct_batch / brgpu_reads_upload take.
struct Chunk {
    std::vector<std::string> definitions; // without the leading '>' and the line end
    Bytes seq;
    std::vector<uint64_t> offsets{0};
    size_t size() const { return definitions.size(); }
    void clear() {
        definitions.clear();
        seq.clear();
        offsets.assign(1, 0);
    }
};

// ------------------------------------------------------------------------------------------
// 2-bit transport form of a chunk (include/brgpu.h, "2-bit transport"): four bases per byte, first base
// in the two high bits, code = (byte >> 1) & 3, plus the exception list of every byte that is not the
// upper-case letter of its own code (lower case, N, ...).  A quarter of the PCIe bytes; the host pays one
// pass over the bases, split over `threads` threads at 4-base-aligned cuts.
// ------------------------------------------------------------------------------------------
struct Packed {
    Bytes bases;                   // ceil(n / 4) bytes
    std::vector<uint64_t> exc_pos; // base positions in the concatenation
    std::vector<uint8_t> exc_byte;
    uint64_t n_bases = 0;
};

inline void pack_range(const uint8_t *seq, uint64_t begin, uint64_t end, uint8_t *packed, std::vector<uint64_t> &exc_pos,
                       std::vector<uint8_t> &exc_byte) {
    static const uint8_t LETTER[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t t = begin; t < end; t += 4) { // begin is a multiple of 4
        uint32_t out = 0;
        const uint64_t m = end - t < 4 ? end - t : 4;
        for (uint64_t j = 0; j < m; j++) {
            const uint8_t c = seq[t + j];
            const uint32_t code = (c >> 1) & 3u;
            out |= code << (2 * (3 - j));
            if (c != LETTER[code]) {
                exc_pos.push_back(t + j);
                exc_byte.push_back(c);
            }
        }
        packed[t >> 2] = (uint8_t)out;
    }
}

inline void pack(const uint8_t *seq, uint64_t n, Packed &out, unsigned threads = 4) {
    out.n_bases = n;
    out.bases.resize((size_t)((n + 3) >> 2));
    out.exc_pos.clear();
    out.exc_byte.clear();
    if (threads < 2 || n < (1u << 20)) {
        pack_range(seq, 0, n, out.bases.data(), out.exc_pos, out.exc_byte);
        return;
    }
    std::vector<std::vector<uint64_t>> ep(threads);
    std::vector<std::vector<uint8_t>> eb(threads);
    std::vector<std::thread> pool;
    const uint64_t per = ((n / threads) + 3) & ~3ULL;
    for (unsigned w = 0; w < threads; w++) {
        const uint64_t b = std::min<uint64_t>(n, per * w), e = w + 1 == threads ? n : std::min<uint64_t>(n, per * (w + 1));
        pool.emplace_back([&, w, b, e]() { pack_range(seq, b, e, out.bases.data(), ep[w], eb[w]); });
    }
    for (auto &t : pool) t.join();
    for (unsigned w = 0; w < threads; w++) {
        out.exc_pos.insert(out.exc_pos.end(), ep[w].begin(), ep[w].end());
        out.exc_byte.insert(out.exc_byte.end(), eb[w].begin(), eb[w].end());
    }
}

// inverse: n bases of ASCII, exceptions written back (any order)
inline void unpack(const uint8_t *packed, uint64_t n, const uint64_t *exc_pos, const uint8_t *exc_byte, uint64_t n_exc,
                   uint8_t *seq_out, unsigned threads = 4) {
    auto range = [&](uint64_t b, uint64_t e) {
        static const uint8_t LETTER[4] = {'A', 'C', 'T', 'G'};
        for (uint64_t t = b; t < e; t++) seq_out[t] = LETTER[(packed[t >> 2] >> (2 * (3 - (t & 3)))) & 3u];
    };
    if (threads < 2 || n < (1u << 20)) {
        range(0, n);
    } else {
        std::vector<std::thread> pool;
        const uint64_t per = n / threads + 1;
        for (unsigned w = 0; w < threads; w++) pool.emplace_back(range, std::min<uint64_t>(n, per * w), std::min<uint64_t>(n, per * (w + 1)));
        for (auto &t : pool) t.join();
    }
    for (uint64_t i = 0; i < n_exc; i++)
        if (exc_pos[i] < n) seq_out[exc_pos[i]] = exc_byte[i];
}

class Reader {
  public:
    // path == "-" or empty: stdin.  gzopen reads plain files transparently.
    explicit Reader(const std::string &path) {
        if (path.empty() || path == "-")
            gz_ = gzdopen(0, "rb");
        else
            gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) throw std::runtime_error("can't open " + (path.empty() ? std::string("stdin") : path));
        gzbuffer(gz_, 1u << 20);
        buf_.resize(1u << 22);
    }
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    ~Reader() {
        if (gz_) gzclose(gz_);
    }

    // Appends up to max_records records to `out`; returns false once the stream is exhausted
    // (the last call may still have appended records, like populate_buffer's `false`).
    bool read_chunk(Chunk &out, size_t max_records) {
        size_t got = 0;
        while (got < max_records) {
            if (!next_record(out)) return false;
            got++;
        }
        return true;
    }

  private:
    bool fill() {
        if (eof_) return false;
        int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (n < 0) throw std::runtime_error("read error in FASTA input");
        if (n == 0) {
            eof_ = true;
            return false;
        }
        pos_ = 0;
        end_ = (size_t)n;
        return true;
    }
    int peek() {
        if (pos_ == end_ && !fill()) return -1;
        return (unsigned char)buf_[pos_];
    }
    // the rest of the current line (without the line end) -> dst (a std::string: definition lines)
    void take_line(std::string &dst) {
        for (;;) {
            if (pos_ == end_ && !fill()) break;
            const char *p = buf_.data() + pos_;
            const char *nl = (const char *)memchr(p, '\n', end_ - pos_);
            size_t n = nl ? (size_t)(nl - p) : end_ - pos_;
            dst.append(p, n);
            pos_ += n;
            if (nl) {
                pos_++;
                break;
            }
        }
        if (!dst.empty() && dst.back() == '\r') dst.pop_back();
    }
    // all sequence lines up to the next '>' (or the end of the stream) -> dst, line ends dropped:
    // one memchr and one memcpy per line, straight out of the read buffer
    void take_sequence(Bytes &dst) {
        bool line_start = true;
        size_t line_len = 0; // bytes of the current line already appended (a line may span two buffer fills)
        for (;;) {
            if (pos_ == end_ && !fill()) return;
            const char *p = buf_.data() + pos_;
            if (line_start && *p == '>') return;
            const char *nl = (const char *)memchr(p, '\n', end_ - pos_);
            size_t n = nl ? (size_t)(nl - p) : end_ - pos_;
            dst.append(p, n);
            pos_ += n;
            line_len += n;
            line_start = false;
            if (nl) {
                pos_++;
                if (line_len && dst.back() == '\r') dst.pop_back();
                line_start = true;
                line_len = 0;
            }
        }
    }
    bool next_record(Chunk &out) {
        // skip anything before the first '>' (blank lines)
        int c;
        while ((c = peek()) >= 0 && c != '>') {
            std::string junk;
            take_line(junk);
        }
        if (c < 0) return false;
        pos_++; // '>'
        std::string def;
        take_line(def);
        out.definitions.push_back(std::move(def));
        take_sequence(out.seq);
        out.offsets.push_back(out.seq.size());
        return true;
    }

    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

class Writer {
  public:
    explicit Writer(const std::string &path) {
        if (path.empty() || path == "-") {
            f_ = stdout;
        } else {
            f_ = fopen(path.c_str(), "wb");
            own_ = true;
        }
        if (!f_) throw std::runtime_error("can't create " + path);
        setvbuf(f_, nullptr, _IONBF, 0); // one fwrite per chunk: no second copy through stdio's buffer
    }
    Writer(const Writer &) = delete;
    Writer &operator=(const Writer &) = delete;
    ~Writer() {
        if (f_) {
            fflush(f_);
            if (own_) fclose(f_);
        }
    }

    // one chunk: definitions[i] with seq[offsets[i], offsets[i+1])
    void write(const std::vector<std::string> &definitions, const uint8_t *seq, const uint64_t *offsets) {
        size_t total = 0;
        for (size_t i = 0; i < definitions.size(); i++) {
            const size_t n = (size_t)(offsets[i + 1] - offsets[i]);
            total += 2 + definitions[i].size() + n + (n + LINE_BASES - 1) / LINE_BASES;
        }
        line_.resize(total);
        uint8_t *w = line_.data();
        for (size_t i = 0; i < definitions.size(); i++) {
            *w++ = '>';
            std::memcpy(w, definitions[i].data(), definitions[i].size());
            w += definitions[i].size();
            *w++ = '\n';
            const uint8_t *s = seq + offsets[i];
            const size_t n = (size_t)(offsets[i + 1] - offsets[i]);
            for (size_t p = 0; p < n; p += LINE_BASES) {
                const size_t m = n - p < LINE_BASES ? n - p : LINE_BASES;
                std::memcpy(w, s + p, m);
                w += m;
                *w++ = '\n';
            }
        }
        if (total && fwrite(line_.data(), 1, total, f_) != total) throw std::runtime_error("write error in FASTA output");
    }

  private:
    FILE *f_ = nullptr;
    bool own_ = false;
    Bytes line_;
};

} // namespace fasta
} // namespace br
