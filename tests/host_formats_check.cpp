// Test helper (tests/test_host_formats_cpu.py): runs br_b200/host/formats.hpp's FASTQ record reader and CSV
// first-column reader (+ br.hpp's for_each_csv_batch) over one input and dumps what they produced.  No GPU.
//   host_formats_check fastq INPUT PREFIX CHUNK_RECORDS     -> PREFIX.seq / .off / .defs, prints "malformed" when the
//                                                              reader stopped at a bad record
//   host_formats_check csv INPUT K BATCH                    -> one line per k-mer (decimal u64), "ERROR <what>" on a throw
//   host_formats_check csvfields INPUT - -                  -> the first field of every data record in hex
#include "br.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>

static void dump(const std::string &path, const void *p, size_t n) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || (n && fwrite(p, 1, n, f) != n)) {
        fprintf(stderr, "can't write %s\n", path.c_str());
        exit(2);
    }
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 5) return 1;
    const std::string mode = argv[1], in = argv[2];
    if (mode == "fastq") {
        const std::string prefix = argv[3];
        const size_t chunk_records = (size_t)atoll(argv[4]);
        br::fastq::Reader rd(in);
        br::fasta::Chunk all;
        std::string defs;
        bool more = true;
        while (more) {
            br::fasta::Chunk c;
            more = rd.read_chunk(c, chunk_records);
            const uint64_t base = all.seq.size();
            all.seq.append(c.seq.data(), c.seq.size());
            for (size_t i = 0; i < c.size(); i++) {
                defs += c.definitions[i];
                defs += '\n';
                all.offsets.push_back(base + c.offsets[i + 1]);
            }
        }
        dump(prefix + ".seq", all.seq.data(), all.seq.size());
        dump(prefix + ".off", all.offsets.data(), all.offsets.size() * 8);
        dump(prefix + ".defs", defs.data(), defs.size());
        if (rd.stopped_on_malformed_record()) puts("malformed");
        return 0;
    }
    if (mode == "csv") {
        const int k = atoi(argv[3]);
        const size_t batch = (size_t)atoll(argv[4]);
        try {
            br::set::for_each_csv_batch(in, k, [](const std::vector<uint64_t> &kmers) {
                for (uint64_t v : kmers) printf("%llu\n", (unsigned long long)v);
                puts("-");
            }, batch);
        } catch (const std::exception &e) {
            printf("ERROR %s\n", e.what());
        }
        return 0;
    }
    if (mode == "csvfields") { // the first field of every data record, hex, one per line (no k-mer conversion)
        try {
            br::csv::FirstColumn rows(in);
            std::string f;
            while (rows.next(f)) {
                for (unsigned char ch : f) printf("%02x", ch);
                puts("");
            }
        } catch (const std::exception &e) {
            printf("ERROR %s\n", e.what());
        }
        return 0;
    }
    return 1;
}
