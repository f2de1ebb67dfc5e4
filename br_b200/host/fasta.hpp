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
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace br {
namespace fasta {

constexpr size_t LINE_BASES = 80;

// host threads the reader / writer / 2-bit packer fan out over (br's -t sizes a rayon pool, src/main.rs:30-33)
inline unsigned default_threads() {
    unsigned h = std::thread::hardware_concurrency();
    return h == 0 ? 4 : (h > 8 ? 8 : h);
}

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
    uint64_t t = begin; // a multiple of 4
    // eight bases per step on 64-bit words (little endian: byte j of the word is base t + j).  codes = (byte >> 1) & 3
    // in every byte; the letter a code stands for is 'A' + 2 b0 + 19 b1 - 15 (b0 & b1) (A C T G = 0x41 0x43 0x54 0x47),
    // so "every byte is the upper-case letter of its code" is one XOR; four codes are gathered into one output byte by
    // a multiplication whose partial products do not overlap (c0 << 30 | c1 << 28 | c2 << 26 | c3 << 24 after it).
    constexpr uint64_t ONES = 0x0101010101010101ULL, GATHER = (1ULL << 30) | (1ULL << 20) | (1ULL << 10) | 1ULL;
    for (; t + 8 <= end; t += 8) {
        uint64_t w;
        std::memcpy(&w, seq + t, 8);
        const uint64_t x = (w >> 1) & (3 * ONES);
        const uint64_t b0 = x & ONES, b1 = (x >> 1) & ONES;
        const uint64_t expect = 0x41 * ONES + 2 * b0 + 19 * b1 - 15 * (b0 & b1);
        if (w != expect) {
            for (int j = 0; j < 8; j++) {
                const uint8_t c = seq[t + j];
                if (c != LETTER[(c >> 1) & 3u]) {
                    exc_pos.push_back(t + j);
                    exc_byte.push_back(c);
                }
            }
        }
        packed[t >> 2] = (uint8_t)((((x & 0xffffffffULL) * GATHER) >> 24) & 0xffu);
        packed[(t >> 2) + 1] = (uint8_t)((((x >> 32) * GATHER) >> 24) & 0xffu);
    }
    for (; t < end; t += 4) {
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
    static const uint8_t LETTER[4] = {'A', 'C', 'T', 'G'};
    struct Lut { // the four letters of every packed byte, in memory order
        uint8_t v[256][4];
        Lut() {
            for (int b = 0; b < 256; b++)
                for (int j = 0; j < 4; j++) v[b][j] = LETTER[(b >> (2 * (3 - j))) & 3];
        }
    };
    static const Lut lut;
    auto range = [&](uint64_t b, uint64_t e) {
        uint64_t t = b;
        for (; t < e && (t & 3); t++) seq_out[t] = LETTER[(packed[t >> 2] >> (2 * (3 - (t & 3)))) & 3u];
        for (; t + 4 <= e; t += 4) std::memcpy(seq_out + t, lut.v[packed[t >> 2]], 4);
        for (; t < e; t++) seq_out[t] = LETTER[(packed[t >> 2] >> (2 * (3 - (t & 3)))) & 3u];
    };
    if (threads < 2 || n < (1u << 20)) {
        range(0, n);
    } else {
        std::vector<std::thread> pool;
        const uint64_t per = (n / threads + 4) & ~3ULL;
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
        if (path.empty() || path == "-") {
            gz_ = gzdopen(0, "rb");
        } else {
            // an uncompressed file is read in large blocks and parsed by several threads (the reference fans its
            // records over rayon workers, src/lib.rs:93-128); gzip input and stdin go through zlib's serial stream
            FILE *f = fopen(path.c_str(), "rb");
            if (!f) throw std::runtime_error("can't open " + path);
            unsigned char magic[2] = {0, 0};
            const size_t got = fread(magic, 1, 2, f);
            if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
                rewind(f);
                plain_ = f;
                // a regular file is mapped: the parser threads read the page cache directly (no read() copy, no
                // buffer compaction); pipes and special files keep the block reads
                struct stat st;
                if (fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
                    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(f), 0);
                    if (m != MAP_FAILED) {
                        map_ = static_cast<const char *>(m);
                        map_len_ = (size_t)st.st_size;
                        madvise(m, map_len_, MADV_SEQUENTIAL);
                    }
                }
                return;
            }
            fclose(f);
            gz_ = gzopen(path.c_str(), "rb");
        }
        if (!gz_) throw std::runtime_error("can't open " + (path.empty() ? std::string("stdin") : path));
        gzbuffer(gz_, 1u << 20);
        buf_.resize(1u << 22);
    }
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    ~Reader() {
        if (map_) munmap(const_cast<char *>(map_), map_len_);
        if (gz_) gzclose(gz_);
        if (plain_) fclose(plain_);
    }

    void set_threads(unsigned t) { threads_ = t ? t : 1; }

    // Appends up to max_records records to `out`; returns false once the stream is exhausted
    // (the last call may still have appended records, like populate_buffer's `false`).
    bool read_chunk(Chunk &out, size_t max_records) {
        if (map_) return read_chunk_mapped(out, max_records);
        if (plain_) return read_chunk_parallel(out, max_records);
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
            if (pos_ == end_ && !fill()) { // end of the stream inside a line: a lone '\r' is still a line end
                if (line_len && dst.back() == '\r') dst.pop_back();
                return;
            }
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

    // ---- plain files: block reads + parallel parse ----
    // raw_[rpos_, rend_) holds file bytes not yet handed out; a record starts at a '>' that begins a line
    bool refill_raw() {
        if (reof_) return false;
        if (rpos_ > 0 && rpos_ == rend_) rpos_ = rend_ = 0;
        if (rpos_ > (raw_.size() >> 1)) { // compact
            std::memmove(raw_.data(), raw_.data() + rpos_, rend_ - rpos_);
            rend_ -= rpos_;
            rpos_ = 0;
        }
        const size_t block = (size_t)1 << 25;
        if (raw_.size() < rend_ + block) raw_.resize(rend_ + block);
        const size_t n = fread(raw_.data() + rend_, 1, block, plain_);
        if (n == 0) {
            if (ferror(plain_)) throw std::runtime_error("read error in FASTA input");
            reof_ = true;
            return false;
        }
        rend_ += n;
        return true;
    }
    static void parse_records(const char *base, const std::vector<size_t> &starts, size_t lo, size_t hi, Chunk &dst) {
        for (size_t r = lo; r < hi; r++) {
            const char *p = base + starts[r] + 1, *end = base + starts[r + 1]; // after '>'
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *dend = nl ? nl : end;
            size_t dl = (size_t)(dend - p);
            if (dl && p[dl - 1] == '\r') dl--;
            dst.definitions.emplace_back(p, dl);
            p = nl ? nl + 1 : end;
            while (p < end) {
                nl = (const char *)memchr(p, '\n', (size_t)(end - p));
                size_t n = nl ? (size_t)(nl - p) : (size_t)(end - p);
                const char *next = nl ? nl + 1 : end;
                if (n && p[n - 1] == '\r') n--;
                dst.seq.append(p, n);
                p = next;
            }
            dst.offsets.push_back(dst.seq.size());
        }
    }
    bool read_chunk_parallel(Chunk &out, size_t max_records) {
        // record starts inside the buffered bytes; more bytes are read until max_records + 1 starts are known
        // (the extra one delimits the last record) or the file ends
        std::vector<size_t> starts;
        size_t scan = rpos_;
        bool at_line_start = true; // rpos_ always sits at the start of a line
        for (;;) {
            const char *b = raw_.data();
            while (scan < rend_ && starts.size() <= max_records) {
                if (at_line_start && b[scan] == '>') starts.push_back(scan);
                const char *nl = (const char *)memchr(b + scan, '\n', rend_ - scan);
                if (!nl) {
                    scan = rend_;
                    at_line_start = false;
                    break;
                }
                scan = (size_t)(nl - b) + 1;
                at_line_start = true;
            }
            if (starts.size() > max_records) break;
            const size_t old_pos = rpos_;
            if (!refill_raw()) break;
            if (rpos_ != old_pos) { // the buffer was compacted: shift what we know
                const size_t shift = old_pos - rpos_;
                for (auto &x : starts) x -= shift;
                scan -= shift;
            }
        }
        const bool more = starts.size() > max_records; // the (max_records + 1)-th start belongs to the next call
        const size_t n_rec = more ? max_records : starts.size();
        const size_t stop = more ? starts[max_records] : rend_;
        if (n_rec == 0) {
            rpos_ = stop;
            return more;
        }
        starts.resize(n_rec);
        starts.push_back(stop);
        parse_parallel(raw_.data(), starts, n_rec, out);
        rpos_ = stop;
        return more;
    }
    // Mapped file: record starts are the '>' that begin a line.  '>' is rare (definition lines only), so memchr finds
    // them at memory speed — the line-by-line scan of the block reader costs as much as parsing.  The file is scanned
    // one window ahead of the records handed out, the window split over the threads; starts found beyond the chunk
    // are kept for the next call.
    void scan_window() {
        const size_t lo = scanned_, hi = std::min(map_len_, scanned_ + ((size_t)64 << 20));
        const unsigned T = (hi - lo) > (1u << 22) ? threads_ : 1;
        std::vector<std::vector<size_t>> found(T);
        auto scan = [&](unsigned t) {
            size_t at = lo + (hi - lo) / T * t;
            const size_t end = t + 1 == T ? hi : lo + (hi - lo) / T * (t + 1);
            while (at < end) {
                const char *gt = (const char *)memchr(map_ + at, '>', end - at);
                if (!gt) break;
                at = (size_t)(gt - map_);
                if (at == 0 || map_[at - 1] == '\n') found[t].push_back(at);
                at++;
            }
        };
        if (T < 2) {
            scan(0);
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < T; t++) pool.emplace_back(scan, t);
            for (auto &th : pool) th.join();
        }
        for (auto &f : found) known_.insert(known_.end(), f.begin(), f.end());
        scanned_ = hi;
    }
    bool read_chunk_mapped(Chunk &out, size_t max_records) {
        // known_[next_ ...] are the record starts at or after mpos_ found so far
        while (known_.size() - next_ <= max_records && scanned_ < map_len_) scan_window();
        const size_t have = known_.size() - next_;
        const bool more = have > max_records;
        const size_t n_rec = more ? max_records : have;
        const size_t stop = more ? known_[next_ + max_records] : map_len_;
        if (n_rec) {
            std::vector<size_t> starts(known_.begin() + (long)next_, known_.begin() + (long)(next_ + n_rec));
            starts.push_back(stop);
            parse_parallel(map_, starts, n_rec, out);
        }
        next_ += n_rec;
        if (next_ > ((size_t)1 << 20)) { // drop the starts already handed out
            known_.erase(known_.begin(), known_.begin() + (long)next_);
            next_ = 0;
        }
        mpos_ = stop;
        return more;
    }
    // records [0, n_rec) of `starts` (starts[n_rec] = end of the last one) -> out, parsed by up to threads_ threads
    void parse_parallel(const char *base, const std::vector<size_t> &starts, size_t n_rec, Chunk &out) {
        const size_t stop = starts[n_rec];
        const unsigned T = (stop - starts[0]) > (1u << 22) ? threads_ : 1;
        if (T < 2) {
            parse_records(base, starts, 0, n_rec, out);
        } else {
            // the per-thread pieces live in the reader: their buffers are paged in once, not once per chunk
            if (parts_.size() < T) parts_.resize(T);
            std::vector<Chunk> &parts = parts_;
            for (unsigned t = 0; t < T; t++) parts[t].clear();
            std::vector<std::thread> pool;
            size_t lo = 0;
            const size_t bytes = stop - starts[0];
            for (unsigned t = 0; t < T; t++) { // cut by input bytes, at record boundaries
                size_t hi = lo;
                const size_t want = starts[0] + bytes / T * (t + 1);
                while (hi < n_rec && (t + 1 == T || starts[hi + 1] <= want)) hi++;
                pool.emplace_back(parse_records, base, std::cref(starts), lo, hi, std::ref(parts[t]));
                lo = hi;
            }
            for (auto &th : pool) th.join();
            size_t total = out.seq.size();
            std::vector<size_t> at(T);
            for (unsigned t = 0; t < T; t++) {
                at[t] = total;
                total += parts[t].seq.size();
            }
            out.seq.resize(total);
            pool.clear();
            for (unsigned t = 0; t < T; t++)
                pool.emplace_back([&, t]() { if (parts[t].seq.size()) std::memcpy(out.seq.data() + at[t], parts[t].seq.data(), parts[t].seq.size()); });
            for (unsigned t = 0; t < T; t++) {
                for (auto &d : parts[t].definitions) out.definitions.push_back(std::move(d));
                for (size_t i = 1; i < parts[t].offsets.size(); i++) out.offsets.push_back(at[t] + parts[t].offsets[i]);
            }
            for (auto &th : pool) th.join();
        }
    }

    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    FILE *plain_ = nullptr;
    std::vector<char> raw_;
    size_t rpos_ = 0, rend_ = 0;
    bool reof_ = false;
    const char *map_ = nullptr; // regular plain file: the whole file, mapped
    size_t map_len_ = 0, mpos_ = 0;
    size_t scanned_ = 0;        // the file has been searched for record starts up to here
    std::vector<size_t> known_; // record starts found, in file order; known_[next_] is the next one to hand out
    size_t next_ = 0;
    std::vector<Chunk> parts_;
    unsigned threads_ = default_threads();
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
        // every record's place in the output is known up front, so the records are formatted by `threads_`
        // threads (the reference formats inside its rayon workers too, src/lib.rs:112-127) and leave in one write
        start_.resize(definitions.size() + 1);
        size_t at = 0;
        for (size_t i = 0; i < definitions.size(); i++) {
            start_[i] = at;
            const size_t n = (size_t)(offsets[i + 1] - offsets[i]);
            at += 2 + definitions[i].size() + n + (n + LINE_BASES - 1) / LINE_BASES;
        }
        start_[definitions.size()] = at;
        auto format = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                uint8_t *w = line_.data() + start_[i];
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
        };
        const unsigned T = total > (1u << 22) ? threads_ : 1;
        if (T < 2) {
            format(0, definitions.size());
        } else {
            std::vector<std::thread> pool;
            size_t lo = 0;
            for (unsigned t = 0; t < T; t++) { // cut by output bytes
                size_t hi = lo;
                const size_t want = total / T * (t + 1);
                while (hi < definitions.size() && (t + 1 == T || start_[hi + 1] <= want)) hi++;
                pool.emplace_back(format, lo, hi);
                lo = hi;
            }
            for (auto &th : pool) th.join();
        }
        if (total && fwrite(line_.data(), 1, total, f_) != total) throw std::runtime_error("write error in FASTA output");
    }

    void set_threads(unsigned t) { threads_ = t ? t : 1; }

  private:
    FILE *f_ = nullptr;
    bool own_ = false;
    Bytes line_;
    std::vector<size_t> start_;
    unsigned threads_ = default_threads();
};

} // namespace fasta
} // namespace br
