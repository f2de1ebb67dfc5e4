// kat_runner — runs known-answer tests through the C++ mirror of br's interface (br.hpp), the way
// the reference's own #[test] functions are written: build a pcon Solid from Tokenizer(refe, k)
// (+ extra k-mers), construct a corrector, assert correct(read) == expected
// (e.g. src/correct/exist/one.rs:89-107).  The vectors come on stdin, one KAT per block:
//
//   KAT <name> <k> <One|Two|Graph|Greedy|GapSize> <confirm> <max_search>
//   ALL <sequence>            every k-mer of the sequence is set (Tokenizer loop)
//   KMER <kmer>               one more k-mer is set
//   CASE <input> <expected>   assert corrector.correct(input) == expected
//   GET <kmer> <0|1>          assert set.get(kmer) == value (KmerSet::get, forward k-mers accepted)
//   END
//
// tests/test_host_cli.py feeds it tests/golden/kats.json.  Exit status 0 iff every assert held.
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

#include "br.hpp"

int main() {
    int failed = 0, asserts = 0, kats = 0;
    try {
        br::Context ctx(0);
        std::string line, name;
        std::unique_ptr<br::set::Pcon> solid;
        std::unique_ptr<br::correct::Corrector> corrector;
        std::string method;
        int k = 0, confirm = 0, max_search = 0;
        auto make_corrector = [&]() {
            if (corrector) return;
            if (method == "One") corrector.reset(new br::correct::One(*solid, (uint8_t)confirm));
            else if (method == "Two") corrector.reset(new br::correct::Two(*solid, (uint8_t)confirm));
            else if (method == "Graph") corrector.reset(new br::correct::Graph(*solid));
            else if (method == "Greedy") corrector.reset(new br::correct::Greedy(*solid, (uint8_t)max_search, (uint8_t)confirm));
            else if (method == "GapSize") corrector.reset(new br::correct::GapSize(*solid, (uint8_t)confirm));
            else throw std::runtime_error("unknown method " + method);
        };
        while (std::getline(std::cin, line)) {
            std::istringstream ss(line);
            std::string tag;
            if (!(ss >> tag)) continue;
            if (tag == "KAT") {
                ss >> name >> k >> method >> confirm >> max_search;
                corrector.reset();
                solid = br::set::Pcon::new_(ctx, k);
                kats++;
            } else if (tag == "ALL") {
                std::string s;
                ss >> s;
                solid->set_all_kmers(s);
            } else if (tag == "KMER") {
                std::string s;
                ss >> s;
                solid->set({br::kmer::seq2bit((const uint8_t *)s.data(), s.size())});
            } else if (tag == "CASE") {
                std::string in, expected;
                ss >> in >> expected;
                make_corrector();
                std::vector<uint8_t> got = corrector->correct(in);
                asserts++;
                if (std::string(got.begin(), got.end()) != expected) {
                    failed++;
                    std::printf("FAIL %s: correct(%s) = %s, expected %s\n", name.c_str(), in.c_str(),
                                std::string(got.begin(), got.end()).c_str(), expected.c_str());
                }
            } else if (tag == "GET") {
                std::string s;
                int want;
                ss >> s >> want;
                asserts++;
                bool got = solid->get(br::kmer::seq2bit((const uint8_t *)s.data(), s.size()));
                if ((int)got != want) {
                    failed++;
                    std::printf("FAIL %s: get(%s) = %d, expected %d\n", name.c_str(), s.c_str(), (int)got, want);
                }
            } else if (tag == "END") {
                corrector.reset();
                solid.reset();
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 2;
    }
    std::printf("%d KATs, %d asserts, %d failed\n", kats, asserts, failed);
    return failed ? 1 : 0;
}
