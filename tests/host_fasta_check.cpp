// Test helper (tests/test_host_fasta_cpu.py): runs br_b200/host/fasta.hpp's reader, 2-bit packer / unpacker and
// writer over one input and dumps what they produced.  No GPU, no libbrgpu.
//   host_fasta_check INPUT PREFIX THREADS CHUNK_RECORDS
#include "fasta.hpp"

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
    const std::string in = argv[1], prefix = argv[2];
    const unsigned threads = (unsigned)atoi(argv[3]);
    const size_t chunk_records = (size_t)atoll(argv[4]);
    br::fasta::Reader rd(in);
    rd.set_threads(threads);
    br::fasta::Chunk all;
    std::string defs;
    {
        br::fasta::Writer wr(prefix + ".fa");
        wr.set_threads(threads);
        bool more = true;
        while (more) {
            br::fasta::Chunk c;
            more = rd.read_chunk(c, chunk_records);
            if (!c.size()) continue;
            wr.write(c.definitions, c.seq.data(), c.offsets.data());
            const uint64_t base = all.seq.size();
            all.seq.append((const char *)c.seq.data(), c.seq.size());
            for (size_t i = 0; i < c.size(); i++) {
                defs += c.definitions[i];
                defs += '\n';
                all.offsets.push_back(base + c.offsets[i + 1]);
            }
        }
    }
    const uint64_t n = all.seq.size();
    br::fasta::Packed pk;
    br::fasta::pack(all.seq.data(), n, pk, threads);
    br::fasta::Bytes back;
    back.resize(n);
    br::fasta::unpack(pk.bases.data(), pk.n_bases, pk.exc_pos.data(), pk.exc_byte.data(), pk.exc_pos.size(), back.data(), threads);
    if (pk.n_bases != n || (n && memcmp(back.data(), all.seq.data(), n) != 0)) {
        fprintf(stderr, "unpack(pack(x)) != x\n");
        return 3;
    }
    dump(prefix + ".seq", all.seq.data(), n);
    dump(prefix + ".off", all.offsets.data(), all.offsets.size() * 8);
    dump(prefix + ".defs", defs.data(), defs.size());
    dump(prefix + ".packed", pk.bases.data(), pk.bases.size());
    dump(prefix + ".excpos", pk.exc_pos.data(), pk.exc_pos.size() * 8);
    dump(prefix + ".excbyte", pk.exc_byte.data(), pk.exc_byte.size());
    return 0;
}
