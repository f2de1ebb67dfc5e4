// brgpu-cli — br's command line (src/cli.rs, src/main.rs:17-58) in front of libbrgpu.so.
//
//   brgpu-cli [-i IN..] [-o OUT..] [-s] [-c METHOD..] [-C CONFIRM] [-M MAX_SEARCH] [-b N] [-t N] [-d DEVICE] [-q] [-v..]
//             fasta -i READS.. -k K [-a N] [first-minimum | rarefaction P | percent-most P | percent-least P]   src/main.rs:72-115
//           | solid -i FILE -f solid|fasta|fastq|csv [-k K]            src/main.rs:117-145
//           | large-kmer -i FILE -f fasta|fastq|csv -k K   (set::Hash, 3 <= K <= 31; src/main.rs:147-163)
//           | count -i FILE [-a N | method]      (pcon count file, src/main.rs:59-70; container restated as recalled)
//
// Same flag names, defaults and quirks as the reference: -s *disables* the
// reversed pass; -b is accepted and does not change the 8192-record chunk; `fasta -k` decrements an
// even k; omitting both -a and a selection method is an error; nothing is printed on success
// (tests/br.rs:28-30 demands an empty stderr).  -t is accepted and ignored (the GPU is the pool).
// One extra sub-command, `echo`, copies records through the FASTA reader and writer without
// touching the GPU (I/O self-check).  `--write-solid PATH` additionally stores the set it built.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "br.hpp"

namespace {

struct Args {
    std::vector<std::string> inputs, outputs;
    bool two_side = false;
    std::vector<br::cli::CorrectionMethod> corrections;
    bool corrections_given = false;
    int confirm = 5, max_search = 7; // src/cli.rs:135-142
    uint64_t record_buffer = 8192;
    int device = 0;
    std::vector<int> devices; // more than one: single-process multi-GPU
    std::string write_solid;
    bool packed = true;
    size_t chunk_bases = 1u << 30; // bases per chunk while the set is built (the reads are streamed, not held)
    // sub-command
    std::string sub;
    std::vector<std::string> sub_inputs;
    int k = -1;
    int abundance = -1;
    std::string selection;
    double percent = 0.0;
    std::string format;
};

[[noreturn]] void usage_error(const std::string &msg) {
    std::fprintf(stderr, "error: %s\n", msg.c_str());
    std::exit(2);
}

const char *USAGE =
    "Usage: brgpu-cli [OPTIONS] <COMMAND>\n"
    "\n"
    "br's command line (src/cli.rs) in front of libbrgpu.so: the solid k-mer set is built and the reads are corrected on the GPU.\n"
    "\n"
    "Commands:\n"
    "  fasta       count the k-mers of FASTA files: -i FILE.. -k K [-a N | first-minimum | rarefaction P | percent-most P | percent-least P]\n"
    "  solid       load or build a dense set:        -i FILE -f solid|fasta|fastq|csv [-k K]\n"
    "  count       load a pcon count file:            -i FILE [-a N | <method>]\n"
    "  large-kmer  hash set for k <= 31:              -i FILE -f fasta|fastq|csv -k K\n"
    "  echo        FASTA reader -> writer, no GPU\n"
    "\n"
    "Options:\n"
    "  -i, --inputs <FILE>...       reads to correct (default: stdin)\n"
    "  -o, --outputs <FILE>...      corrected reads, paired with the inputs (default: stdout)\n"
    "  -c, --corrections <M>...     one, two, graph, greedy, gap-size (default: all five, in this order)\n"
    "  -C, --confirm <N>            solid k-mers that must follow a correction [default: 5]\n"
    "  -M, --max-search <N>         greedy: bases explored [default: 7]\n"
    "  -s, --two-side               do NOT run the reversed pass (br's flag, br's meaning)\n"
    "  -b, --record_buffer <N>      accepted; the chunk is 8192 records like br's\n"
    "  -t, --threads <N>            accepted; the GPU is the pool\n"
    "  -d, --device <D[,D...]>      CUDA device, or a list: one process drives all of them\n"
    "      --transport <T>          packed (2 bits per base + exceptions, default) | ascii\n"
    "      --chunk-bases <N>        bases per chunk while a set is built from FASTA / FASTQ\n"
    "      --write-solid <FILE>     also store the dense set as a pcon .solid file\n"
    "  -h, --help                   print this\n"
    "  -V, --version                print the library version\n";

bool is_sub(const std::string &s) {
    return s == "fasta" || s == "solid" || s == "large-kmer" || s == "count" || s == "echo";
}

br::cli::CorrectionMethod parse_method(const std::string &s) { // clap ValueEnum: kebab-case
    if (s == "one") return br::cli::CorrectionMethod::One;
    if (s == "two") return br::cli::CorrectionMethod::Two;
    if (s == "graph") return br::cli::CorrectionMethod::Graph;
    if (s == "greedy") return br::cli::CorrectionMethod::Greedy;
    if (s == "gap-size" || s == "gap_size") return br::cli::CorrectionMethod::GapSize;
    usage_error("invalid value '" + s + "' for '--corrections': one, two, graph, greedy, gap-size");
}

long parse_int(const std::string &flag, const std::string &v, long lo, long hi) {
    char *end = nullptr;
    long x = std::strtol(v.c_str(), &end, 10);
    if (v.empty() || *end || x < lo || x > hi) usage_error("invalid value '" + v + "' for '" + flag + "'");
    return x;
}

Args parse(int argc, char **argv) {
    Args a;
    int i = 1;
    auto value = [&](const std::string &flag) -> std::string {
        if (i + 1 >= argc) usage_error("a value is required for '" + flag + "'");
        return argv[++i];
    };
    // a multi-valued option takes every following token up to the next option / sub-command
    auto values = [&](const std::string &flag, std::vector<std::string> &dst, bool stop_at_sub) {
        size_t before = dst.size();
        auto is_method = [](const std::string &v) { // the abundance sub-sub-commands end a value list too (clap)
            return v == "first-minimum" || v == "rarefaction" || v == "percent-most" || v == "percent-least";
        };
        while (i + 1 < argc && argv[i + 1][0] != '-' && !(stop_at_sub && is_sub(argv[i + 1])) && !(!stop_at_sub && is_method(argv[i + 1])))
            dst.push_back(argv[++i]);
        if (dst.size() == before) usage_error("a value is required for '" + flag + "'");
    };
    for (; i < argc; i++) {
        std::string t = argv[i];
        if (is_sub(t)) {
            a.sub = t;
            i++;
            break;
        }
        if (t == "-i" || t == "--inputs") values(t, a.inputs, true);
        else if (t == "-o" || t == "--outputs") values(t, a.outputs, true);
        else if (t == "-s" || t == "--two-side") a.two_side = true;
        else if (t == "-c" || t == "--corrections") {
            std::vector<std::string> ms;
            values(t, ms, true);
            for (auto &m : ms) a.corrections.push_back(parse_method(m));
            a.corrections_given = true;
        } else if (t == "-C" || t == "--confirm") a.confirm = (int)parse_int(t, value(t), 0, 255);
        else if (t == "-M" || t == "--max-search") a.max_search = (int)parse_int(t, value(t), 0, 255);
        else if (t == "-b" || t == "--record_buffer") a.record_buffer = (uint64_t)parse_int(t, value(t), 0, 1L << 40);
        else if (t == "-t" || t == "--threads") (void)parse_int(t, value(t), 0, 1 << 20);
        else if (t == "-d" || t == "--device") { // one device, or a comma-separated list: the group path (brgpu_group_*)
            std::string v = value(t);
            a.devices.clear();
            size_t p = 0;
            while (p <= v.size()) {
                const size_t q = v.find(',', p);
                a.devices.push_back((int)parse_int(t, v.substr(p, q == std::string::npos ? q : q - p), 0, 1023));
                if (q == std::string::npos) break;
                p = q + 1;
            }
            a.device = a.devices[0];
        }
        else if (t == "--write-solid") a.write_solid = value(t);
        else if (t == "--transport") { // how a chunk crosses PCIe: packed (2 bits per base + exceptions, default) | ascii
            const std::string v = value(t);
            if (v != "packed" && v != "ascii") usage_error("invalid value '" + v + "' for '--transport': packed, ascii");
            a.packed = v == "packed";
        } else if (t == "--chunk-bases") a.chunk_bases = (size_t)parse_int(t, value(t), 1, 1LL << 40); // set construction streams chunks of this size
        else if (t == "-h" || t == "--help") {
            std::fputs(USAGE, stdout);
            std::exit(0);
        } else if (t == "-V" || t == "--version") {
            std::printf("brgpu-cli (%s)\n", brgpu_version());
            std::exit(0);
        }
        else if (t == "-q" || t == "--quiet") {}
        else if (t.rfind("-v", 0) == 0 || t == "--verbosity") {}
        else if (t == "-T" || t == "--timestamp") (void)value(t);
        else usage_error("unexpected argument '" + t + "'");
    }
    if (a.sub.empty()) usage_error("a sub-command is required: fasta, solid, large-kmer, count");
    for (; i < argc; i++) {
        std::string t = argv[i];
        if (t == "-i" || t == "--inputs" || t == "--input") values(t, a.sub_inputs, false);
        else if (t == "-k" || t == "--kmer-size") a.k = (int)parse_int(t, value(t), 1, 255);
        else if (t == "-a" || t == "--abundance") a.abundance = (int)parse_int(t, value(t), 0, 255);
        else if (t == "-f" || t == "--format") a.format = value(t);
        else if (t == "first-minimum") a.selection = t;
        else if (t == "rarefaction" || t == "percent-most" || t == "percent-least") {
            a.selection = t;
            const std::string v = value(t);
            char *end = nullptr;
            a.percent = std::strtod(v.c_str(), &end);
            if (v.empty() || *end) usage_error("invalid value '" + v + "' for '<PERCENT>'");
        } else usage_error("unexpected argument '" + t + "' for '" + a.sub + "'");
    }
    if (!a.corrections_given) // src/cli.rs:121-131
        a.corrections = {br::cli::CorrectionMethod::One, br::cli::CorrectionMethod::Two, br::cli::CorrectionMethod::Graph,
                         br::cli::CorrectionMethod::Greedy, br::cli::CorrectionMethod::GapSize};
    if (a.inputs.empty()) a.inputs.push_back("-");   // default stdin  (src/cli.rs:80-81)
    if (a.outputs.empty()) a.outputs.push_back("-"); // default stdout (src/cli.rs:99-103)
    return a;
}

// Fasta::inputs / Solid::input: every file chained into one record stream (src/cli.rs:265-274)
void read_all(const std::vector<std::string> &paths, br::fasta::Chunk &all) {
    for (auto &p : paths) {
        br::fasta::Reader r(p);
        while (r.read_chunk(all, 1u << 20)) {}
    }
}

std::unique_ptr<br::set::DeviceSet> build_set(const br::Context &ctx, const Args &a) {
    using br::set::AbundanceSelection;
    using br::set::Pcon;
    if (a.sub_inputs.empty()) usage_error("the following required arguments were not provided: --inputs");
    if (a.sub == "fasta") { // src/main.rs:72-85
        if (a.k < 0) usage_error("the following required arguments were not provided: --kmer-size");
        AbundanceSelection sel = AbundanceSelection::None;
        if (a.selection == "first-minimum") sel = AbundanceSelection::FirstMinimum;
        else if (a.selection == "rarefaction") sel = AbundanceSelection::Rarefaction;
        else if (a.selection == "percent-most") sel = AbundanceSelection::PercentMost;
        else if (a.selection == "percent-least") sel = AbundanceSelection::PercentLeast;
        // count_fasta(inputs, 8192) streams the records (src/main.rs:74): so does this, 1 Gbase per chunk
        return Pcon::from_count_stream(ctx, a.sub_inputs, a.k, a.abundance, sel, a.percent, a.chunk_bases);
    }
    if (a.sub == "count") { // src/main.rs:59-70 (pcon's count-file container: restated as recalled, see Pcon::from_pcon_count)
        AbundanceSelection sel = AbundanceSelection::None;
        if (a.selection == "first-minimum") sel = AbundanceSelection::FirstMinimum;
        else if (a.selection == "rarefaction") sel = AbundanceSelection::Rarefaction;
        else if (a.selection == "percent-most") sel = AbundanceSelection::PercentMost;
        else if (a.selection == "percent-least") sel = AbundanceSelection::PercentLeast;
        return Pcon::from_pcon_count(ctx, a.sub_inputs[0], a.abundance, sel, a.percent);
    }
    if (a.sub == "solid") { // src/main.rs:117-145
        if (a.format == "solid") return Pcon::from_pcon_solid(ctx, a.sub_inputs[0]);
        if (a.format == "fasta") {
            if (a.k < 0) throw std::runtime_error("Solid input in this format require a kmer size"); // Error::SolidRequireKmerSize
            br::fasta::Chunk reads;
            read_all({a.sub_inputs[0]}, reads);
            return Pcon::from_fasta(ctx, reads, a.k);
        }
        if (a.format == "fastq" || a.format == "csv") { // br's cargo features `fastq` / `csv` (src/main.rs:120-139)
            if (a.k < 0) throw std::runtime_error("Solid input in this format require a kmer size");
            return a.format == "fastq" ? Pcon::from_fastq(ctx, a.sub_inputs[0], a.k) : Pcon::from_csv(ctx, a.sub_inputs[0], a.k);
        }
        usage_error("invalid value '" + a.format + "' for '--format': solid, fasta, fastq, csv");
    }
    if (a.sub == "large-kmer") { // src/main.rs:147-163: set::Hash == presence-only set of canonical k-mers
        if (a.k < 0) usage_error("the following required arguments were not provided: --kmer-size");
        if (a.format != "fasta" && a.format != "fastq" && a.format != "csv")
            usage_error("invalid value '" + a.format + "' for '--format': fasta, fastq, csv");
        if (a.k < 3 || a.k > 31) throw std::runtime_error("large-kmer: k must be in 3..=31 (k-mers are 2-bit packed in 64 bits)");
        if (a.format == "fastq") return br::set::Hash::from_fastq(ctx, {a.sub_inputs[0]}, a.k, a.chunk_bases);
        if (a.format == "csv") return br::set::Hash::from_csv(ctx, a.sub_inputs[0], a.k);
        return br::set::Hash::from_fasta(ctx, {a.sub_inputs[0]}, a.k, a.chunk_bases);
    }
    throw std::runtime_error("sub-command '" + a.sub + "' is not supported by brgpu");
}

// `-d 0,1,...`: one process owns several GPUs (brgpu_group_*): the `fasta` sub-command shards the k-mer
// counting over them, every device holds a replica of the set and corrects its share of each chunk.
// Limit of this mode: the set-construction input is read whole (one host buffer, one device chunk per GPU:
// about 3.5 Gbases per device); the single-device path streams it chunk by chunk (Pcon::from_count_stream).
int run_group(const Args &a) {
    if (a.sub != "fasta") throw std::runtime_error("-d with several devices supports the fasta sub-command");
    if (a.k < 0) usage_error("the following required arguments were not provided: --kmer-size");
    if (a.sub_inputs.empty()) usage_error("the following required arguments were not provided: --inputs");
    brgpu_group *g = nullptr;
    int st = brgpu_group_create(a.devices.data(), (int)a.devices.size(), &g);
    if (st != BRGPU_OK) throw std::runtime_error(st == BRGPU_E_NO_DEVICE ? "no CUDA device / no peer access between the devices (brgpu has no CPU path)" : "can't create the device group");
    const int n = brgpu_group_size(g);
    std::vector<brgpu_set *> sets((size_t)n, nullptr);
    auto check = [&](int s) {
        if (s == BRGPU_OK) return;
        std::string msg = brgpu_group_last_error(g);
        if (s == BRGPU_E_NEED_ABUNDANCE) msg = "need an abundance threshold or an abundance method";
        brgpu_group_sets_free(g, sets.data());
        brgpu_group_destroy(g);
        throw std::runtime_error(msg);
    };
    int sel = BRGPU_ABUNDANCE_EXPLICIT;
    if (a.abundance < 0) {
        if (a.selection == "first-minimum") sel = BRGPU_ABUNDANCE_FIRST_MINIMUM;
        else if (a.selection == "rarefaction") sel = BRGPU_ABUNDANCE_RAREFACTION;
        else if (a.selection == "percent-most") sel = BRGPU_ABUNDANCE_PERCENT_AT_MOST;
        else if (a.selection == "percent-least") sel = BRGPU_ABUNDANCE_PERCENT_AT_LEAST;
    }
    {
        br::fasta::Chunk reads;
        read_all(a.sub_inputs, reads);
        const int k = a.k - (!(a.k & 1) & 1);
        check(brgpu_group_set_from_host_reads(g, k, a.abundance, sel, a.percent, reads.seq.data(), reads.offsets.data(),
                                              reads.size(), sets.data()));
    }
    std::vector<uint8_t> ids;
    for (auto m : a.corrections) ids.push_back((uint8_t)m);
    const size_t pairs = std::min(a.inputs.size(), a.outputs.size());
    for (size_t p = 0; p < pairs; p++) {
        br::fasta::Reader reader(a.inputs[p]);
        br::fasta::Writer writer(a.outputs[p]);
        br::fasta::Chunk c;
        br::fasta::Bytes out;
        std::vector<uint64_t> out_off;
        bool more = true;
        while (more) {
            c.clear();
            more = reader.read_chunk(c, br::CHUNK_RECORDS);
            if (!c.size()) continue;
            out.resize(c.seq.size() + c.seq.size() / 8 + 64 * c.size() + 64);
            out_off.assign(c.size() + 1, 0);
            for (;;) {
                uint64_t need = 0;
                st = brgpu_group_correct_batch(g, sets.data(), ids.data(), ids.size(), a.confirm, a.max_search, a.two_side ? 1 : 0,
                                               c.seq.data(), c.offsets.data(), c.size(), out.data(), out.size(), out_off.data(), &need);
                if (st == BRGPU_E_OVERFLOW && need > out.size()) {
                    out.resize(need);
                    continue;
                }
                break;
            }
            check(st);
            writer.write(c.definitions, out.data(), out_off.data());
        }
    }
    brgpu_group_sets_free(g, sets.data());
    brgpu_group_destroy(g);
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    Args a = parse(argc, argv);
    try {
        if (a.sub == "echo") { // FASTA reader -> writer, no GPU
            const size_t pairs = std::min(a.inputs.size(), a.outputs.size());
            for (size_t p = 0; p < pairs; p++) {
                br::fasta::Reader r(a.inputs[p]);
                br::fasta::Writer w(a.outputs[p]);
                br::fasta::Chunk c;
                bool more = true;
                while (more) {
                    c.clear();
                    more = r.read_chunk(c, br::CHUNK_RECORDS);
                    w.write(c.definitions, c.seq.data(), c.offsets.data());
                }
            }
            return 0;
        }
        if (a.devices.size() > 1) return run_group(a);
        br::Context ctx(a.device);
        std::unique_ptr<br::set::DeviceSet> kmer_set = build_set(ctx, a);
        if (!a.write_solid.empty()) {
            auto *dense = dynamic_cast<br::set::Pcon *>(kmer_set.get());
            if (!dense) throw std::runtime_error("--write-solid needs a dense set (fasta / solid sub-commands)");
            dense->write_solid(a.write_solid);
        }
        br::Methods methods = br::build_methods(a.corrections, *kmer_set, (uint8_t)a.confirm, (uint8_t)a.max_search);
        br::run_correction(a.inputs, a.outputs, methods, a.two_side, a.record_buffer,
                           a.packed ? br::Transport::Packed : br::Transport::Ascii);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Error: %s\n", e.what()); // anyhow's top-level report
        return 1;
    }
    return 0;
}
