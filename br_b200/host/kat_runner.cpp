// kat_runner — runs known-answer tests through the C++ mirror of br's interface (br.hpp), the way
// the reference's own #[test] functions are written: build a pcon Solid from Tokenizer(refe, k)
// (+ extra k-mers), construct a corrector, assert correct(read) == expected
// (e.g. src/correct/exist/one.rs:89-107).  The vectors come on stdin, one KAT per block:
//
//   KAT <name> <k> <One|Two|Graph|Greedy|GapSize> <confirm> <max_search>      the set is a br::set::Pcon
//   HASHKAT <name> <k> <method> <confirm> <max_search>                         the set is a br::set::Hash (src/set/hash.rs)
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
        std::unique_ptr<br::set::DeviceSet> solid; // Pcon or Hash: the correctors only see KmerSet
        br::set::Pcon *dense = nullptr;
        br::set::Hash *hash = nullptr;
        auto insert = [&](const std::vector<uint64_t> &kmers) { // Solid::set / FxHashSet::insert (both canonicalise)
            if (dense) dense->set(kmers);
            else hash->insert(kmers);
        };
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
                auto p = br::set::Pcon::new_(ctx, k);
                dense = p.get(), hash = nullptr;
                solid = std::move(p);
                kats++;
            } else if (tag == "HASHKAT") {
                ss >> name >> k >> method >> confirm >> max_search;
                corrector.reset();
                auto p = br::set::Hash::new_(ctx, k);
                hash = p.get(), dense = nullptr;
                solid = std::move(p);
                kats++;
            } else if (tag == "ALL") {
                std::string s;
                ss >> s;
                if (dense) {
                    dense->set_all_kmers(s);
                } else if ((int)s.size() >= k) { // Tokenizer(seq, k)
                    std::vector<uint64_t> kmers;
                    for (size_t i = 0; i + (size_t)k <= s.size(); i++) kmers.push_back(br::kmer::seq2bit((const uint8_t *)s.data() + i, (size_t)k));
                    insert(kmers);
                }
            } else if (tag == "KMER") {
                std::string s;
                ss >> s;
                insert({br::kmer::seq2bit((const uint8_t *)s.data(), s.size())});
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
                dense = nullptr, hash = nullptr;
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 2;
    }
    std::printf("%d KATs, %d asserts, %d failed\n", kats, asserts, failed);
    return failed ? 1 : 0;
}
