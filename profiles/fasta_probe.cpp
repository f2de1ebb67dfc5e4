// Host-side throughput of the FASTA reader / 2-bit packer / writer (br_b200/host/fasta.hpp), no GPU involved:
//   g++ -O2 -std=c++17 -pthread -Ibr_b200/host profiles/fasta_probe.cpp -lz -o /tmp/fasta_probe && /tmp/fasta_probe reads.fa [threads]
#include "fasta.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }
int main(int argc, char **argv) {
    if (argc < 2) return 1;
    const unsigned threads = argc > 2 ? (unsigned)atoi(argv[2]) : br::fasta::default_threads();
    for (int rep = 0; rep < 3; rep++) {
        br::fasta::Reader rd(argv[1]);
        rd.set_threads(threads);
        br::fasta::Writer wr("/dev/null");
        wr.set_threads(threads);
        double t_read = 0, t_pack = 0, t_unpack = 0, t_write = 0;
        uint64_t bases = 0, records = 0;
        bool more = true;
        br::fasta::Chunk chunks[2]; // reused in turn, as br::run_correction does
        br::fasta::Packed p;
        br::fasta::Bytes back;
        for (int cur = 0; more; cur ^= 1) {
            br::fasta::Chunk &c = chunks[cur];
            c.clear();
            auto t0 = clk::now();
            more = rd.read_chunk(c, 8192);
            auto t1 = clk::now();
            if (!c.size()) break;
            br::fasta::pack(c.seq.data(), c.offsets.back(), p, threads);
            auto t2 = clk::now();
            back.resize(c.offsets.back());
            br::fasta::unpack(p.bases.data(), p.n_bases, p.exc_pos.data(), p.exc_byte.data(), p.exc_pos.size(), back.data(), threads);
            auto t3 = clk::now();
            wr.write(c.definitions, back.data(), c.offsets.data());
            auto t4 = clk::now();
            t_read += secs(t0, t1); t_pack += secs(t1, t2); t_unpack += secs(t2, t3); t_write += secs(t3, t4);
            bases += c.offsets.back(); records += c.size();
        }
        printf("threads %u: %llu records, %.1f Mbases: read+parse %.2f GB/s, pack %.2f, unpack %.2f, format+write %.2f (bases per second of each stage)\n",
               threads, (unsigned long long)records, bases / 1e6, bases / t_read / 1e9, bases / t_pack / 1e9, bases / t_unpack / 1e9, bases / t_write / 1e9);
    }
    return 0;
}
