// formats.hpp — the other inputs a solid set can be built from: FASTQ records and a CSV of k-mers.
//
// br reads them behind its optional `fastq` and `csv` cargo features (Cargo.toml:53-58):
//   solid      -f fastq|csv -k K   set::Pcon::from_fastq / from_csv   src/set/pcon.rs:27-45, :114-181, src/main.rs:117-145
//   large-kmer -f fastq|csv -k K   set::Hash::from_fastq / from_csv   src/set/hash.rs:20-39, :102-175, src/main.rs:147-163
// through noodles-fastq and the `csv` crate (neither vendored in the reference tree), so the framing is the
// published convention of each format, restated here:
//   FASTQ: four lines per record — `@` name [description], the sequence on ONE line, a `+` line, the qualities.
//          The reference's loops are `while let Some(Ok(record)) = records.next()` (pcon.rs:122, hash.rs:112): the
//          first malformed or truncated record silently ends the input; what was read before it counts.
//   CSV:   csv::Reader::from_reader defaults — `,` delimiter, `"` quote with `""` as the escaped quote, records ended
//          by `\n`, `\r\n` or `\r`, empty lines skipped, the FIRST record is a header and is not data, and a record with
//          another number of fields than the first is an error (`result?`, pcon.rs:35-36).  The k-mer is field 0.
// Gzip input is sniffed by zlib (niffler in the reference, src/cli.rs:400-420).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fasta.hpp"

namespace br {
namespace formats {

// a (possibly gzip) byte stream, block by block
class ByteStream {
  public:
    explicit ByteStream(const std::string &path) {
        gz_ = (path.empty() || path == "-") ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
        if (!gz_) throw std::runtime_error("can't open " + (path.empty() ? std::string("stdin") : path));
        gzbuffer(gz_, 1u << 20);
        buf_.resize(1u << 20);
    }
    ByteStream(const ByteStream &) = delete;
    ByteStream &operator=(const ByteStream &) = delete;
    ~ByteStream() { gzclose(gz_); }
    int get() { // next byte, or -1 at the end
        if (pos_ == end_ && !fill()) return -1;
        return (unsigned char)buf_[pos_++];
    }
    int peek() {
        if (pos_ == end_ && !fill()) return -1;
        return (unsigned char)buf_[pos_];
    }
    // the rest of the current line without its end (`\n` or `\r\n`) is appended to dst; false when the stream had
    // ended before the call (no line at all)
    template <class Dst> bool line(Dst &dst, size_t *appended = nullptr) {
        if (pos_ == end_ && !fill()) return false;
        size_t total = 0;
        bool cr = false; // the line so far ends in '\r' (still in dst)
        for (;;) {
            const char *p = buf_.data() + pos_;
            const char *nl = (const char *)memchr(p, '\n', end_ - pos_);
            const size_t n = nl ? (size_t)(nl - p) : end_ - pos_;
            if (n) {
                dst.append(p, n);
                cr = p[n - 1] == '\r';
                total += n;
            }
            pos_ += n;
            if (nl) {
                pos_++;
                break;
            }
            if (!fill()) break;
        }
        if (cr) {
            drop_last(dst);
            total--;
        }
        if (appended) *appended = total;
        return true;
    }

  private:
    static void drop_last(std::string &s) { s.pop_back(); }
    static void drop_last(fasta::Bytes &b) { b.pop_back(); }
    bool fill() {
        if (eof_) return false;
        const int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (n < 0) throw std::runtime_error("read error in input");
        if (n == 0) {
            eof_ = true;
            return false;
        }
        pos_ = 0;
        end_ = (size_t)n;
        return true;
    }
    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

} // namespace formats

namespace fastq {

// noodles::fastq::Reader::records() into the chunk form the set builders upload (same Chunk as FASTA)
class Reader {
  public:
    explicit Reader(const std::string &path) : in_(path) {}

    // Appends up to max_records records; false once the input is exhausted — or at the first record that is not
    // `@..` / sequence / `+..` / qualities (the reference's `while let Some(Ok(..))` stops there without an error).
    bool read_chunk(fasta::Chunk &out, size_t max_records) {
        if (done_) return false;
        for (size_t got = 0; got < max_records; got++) {
            if (!next_record(out)) {
                done_ = true;
                return false;
            }
        }
        return true;
    }
    bool stopped_on_malformed_record() const { return malformed_; }

  private:
    bool next_record(fasta::Chunk &out) {
        const int c = in_.peek();
        if (c < 0) return false;
        if (c != '@') return bad();
        std::string def;
        in_.line(def);
        def.erase(0, 1);
        const size_t seq_before = out.seq.size();
        if (!in_.line(out.seq)) return bad(); // the three other lines must exist
        std::string plus, qual;
        if (!in_.line(plus) || plus.empty() || plus[0] != '+' || !in_.line(qual)) {
            out.seq.resize(seq_before);
            return bad();
        }
        out.definitions.push_back(std::move(def));
        out.offsets.push_back(out.seq.size());
        return true;
    }
    bool bad() {
        malformed_ = true;
        return false;
    }
    formats::ByteStream in_;
    bool done_ = false, malformed_ = false;
};

} // namespace fastq

namespace csv {

// csv::Reader::from_reader(input).byte_records() reduced to what from_csv uses: field 0 of every data record
class FirstColumn {
  public:
    explicit FirstColumn(const std::string &path) : in_(path) {}

    // the next data record's first field -> field; false at the end of the input.  The header record is skipped;
    // a record whose field count differs from the header's throws (csv::ErrorKind::UnequalLengths).
    bool next(std::string &field) {
        if (!header_done_) {
            header_done_ = true;
            std::string h;
            size_t n = 0;
            if (!record(h, n)) return false;
            fields_ = n;
        }
        size_t n = 0;
        if (!record(field, n)) return false;
        if (n != fields_)
            throw std::runtime_error("CSV error: record " + std::to_string(records_ - 1) + ": found record with " + std::to_string(n) +
                                     " fields, but the previous record has " + std::to_string(fields_) + " fields");
        return true;
    }

  private:
    // one record: its first field and its number of fields; empty lines are not records
    bool record(std::string &first, size_t &n_fields) {
        first.clear();
        int c;
        while ((c = in_.peek()) == '\n' || c == '\r') in_.get(); // skip empty lines
        if (c < 0) return false;
        n_fields = 1;
        bool in_quotes = false, field_start = true;
        for (;;) {
            c = in_.get();
            if (c < 0) break; // the last record needs no terminator (an unclosed quote ends with the input too)
            if (in_quotes) {
                if (c == '"') {
                    if (in_.peek() == '"') { // "" inside quotes = one quote
                        in_.get();
                        if (n_fields == 1) first.push_back('"');
                    } else {
                        in_quotes = false;
                    }
                } else if (n_fields == 1) {
                    first.push_back((char)c);
                }
                continue;
            }
            if (c == '"' && field_start) {
                in_quotes = true;
                field_start = false;
                continue;
            }
            field_start = false;
            if (c == ',') {
                n_fields++;
                field_start = true;
                continue;
            }
            if (c == '\n') break;
            if (c == '\r') {
                if (in_.peek() == '\n') in_.get();
                break;
            }
            if (n_fields == 1) first.push_back((char)c);
        }
        records_++;
        return true;
    }
    formats::ByteStream in_;
    bool header_done_ = false;
    size_t fields_ = 0, records_ = 0;
};

} // namespace csv
} // namespace br
