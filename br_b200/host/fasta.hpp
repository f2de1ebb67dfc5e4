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
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace br {
namespace fasta {

constexpr size_t LINE_BASES = 80;

// One buffer of records (populate_buffer, src/lib.rs:168-188): sequences concatenated, with
// prefix-sum offsets — exactly what brgpu_correct_batch / brgpu_reads_upload take.
struct Chunk {
    std::vector<std::string> definitions; // without the leading '>' and the line end
    std::vector<uint8_t> seq;
    std::vector<uint64_t> offsets{0};
    size_t size() const { return definitions.size(); }
    void clear() {
        definitions.clear();
        seq.clear();
        offsets.assign(1, 0);
    }
};

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
    // appends the rest of the current line (without the line end) to dst
    template <class Sink> void take_line(Sink &dst) {
        for (;;) {
            if (pos_ == end_ && !fill()) break;
            const char *p = buf_.data() + pos_;
            const char *nl = (const char *)memchr(p, '\n', end_ - pos_);
            size_t n = nl ? (size_t)(nl - p) : end_ - pos_;
            dst.insert(dst.end(), p, p + n);
            pos_ += n;
            if (nl) {
                pos_++;
                break;
            }
        }
        if (!dst.empty() && dst.back() == '\r') dst.pop_back();
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
        while ((c = peek()) >= 0 && c != '>') {
            size_t before = out.seq.size();
            take_line(out.seq);
            (void)before;
        }
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
        setvbuf(f_, nullptr, _IOFBF, 1u << 22);
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
        line_.clear();
        for (size_t i = 0; i < definitions.size(); i++) {
            line_.push_back('>');
            line_.insert(line_.end(), definitions[i].begin(), definitions[i].end());
            line_.push_back('\n');
            const uint8_t *s = seq + offsets[i];
            const size_t n = (size_t)(offsets[i + 1] - offsets[i]);
            for (size_t p = 0; p < n; p += LINE_BASES) {
                size_t m = n - p < LINE_BASES ? n - p : LINE_BASES;
                line_.insert(line_.end(), s + p, s + p + m);
                line_.push_back('\n');
            }
        }
        if (!line_.empty() && fwrite(line_.data(), 1, line_.size(), f_) != line_.size())
            throw std::runtime_error("write error in FASTA output");
    }

  private:
    FILE *f_ = nullptr;
    bool own_ = false;
    std::vector<char> line_;
};

} // namespace fasta
} // namespace br
