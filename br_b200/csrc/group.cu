// group.cu — the sharded hot path behind the C ABI for ONE process that owns several GPUs
// (SURVEY §8b "Threading": one owner of all GPUs; a Rust or C++ host calls these and gets every GPU of
// the box, no torch.distributed, no CUDA IPC).
//
// A group is N contexts, one per device, with peer access enabled between every pair.  The protocol is
// the one br_b200/dist.py runs with one process per GPU:
//   part 1  every device partitions the k-mers of its shard by table-index range; device i counts bucket
//           range i over ALL partitions — the peers' residues of that range are pulled over NVLink by
//           bulk peer-to-peer copies — thresholds it into its slice of the bitfield (+ summary) and
//           pushes the slice to every peer, so that every device ends up with the whole set;
//           small k (< 15): private count tables, saturating merge of slice i over peer memory.
//   part 2  each device corrects its own reads against its replica; records keep their input order.
// One host thread per device drives a phase; joining the threads (every phase ends with its stream
// drained) is the barrier between phases.  min(255, sum) is associative and commutative, so the set is
// bit-identical to the single-GPU one.
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

using namespace brgpu;

struct brgpu_group {
    std::vector<brgpu_ctx *> ctx;
    std::string err;
};

static const int GROUP_BUCKET_BITS = 15; // must match BUCKET_BITS in set_kernels.cu

static int group_fail(brgpu_group *g, int code, const std::string &what) {
    if (g) g->err = what;
    return code;
}

// run f(i) for every device on its own host thread; returns the first non-OK status
template <class F> static int for_each_device(brgpu_group *g, F f) {
    const int n = (int)g->ctx.size();
    std::vector<int> st(n, BRGPU_OK);
    std::vector<std::thread> pool;
    for (int i = 0; i < n; i++)
        pool.emplace_back([&, i]() {
            cudaSetDevice(g->ctx[i]->device);
            st[i] = f(i);
            if (st[i] == BRGPU_OK && cudaStreamSynchronize(g->ctx[i]->stream) != cudaSuccess) {
                cudaGetLastError();
                st[i] = BRGPU_E_CUDA;
            }
        });
    for (auto &t : pool) t.join();
    for (int i = 0; i < n; i++)
        if (st[i] != BRGPU_OK) return group_fail(g, st[i], std::string("device ") + std::to_string(i) + ": " + brgpu_last_error(g->ctx[i]));
    return BRGPU_OK;
}

extern "C" int brgpu_group_create(const int *devices, int n, brgpu_group **out) {
    if (!devices || n < 1 || n > 16 || !out) return BRGPU_E_INVALID;
    *out = nullptr;
    brgpu_group *g = new (std::nothrow) brgpu_group;
    if (!g) return BRGPU_E_NOMEM;
    for (int i = 0; i < n; i++) {
        brgpu_ctx *c = nullptr;
        int st = brgpu_ctx_create(devices[i], nullptr, &c);
        if (st != BRGPU_OK) {
            for (auto p : g->ctx) brgpu_ctx_destroy(p);
            delete g;
            return st;
        }
        g->ctx.push_back(c);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j || devices[i] == devices[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (!can) {
                for (auto p : g->ctx) brgpu_ctx_destroy(p);
                delete g;
                return BRGPU_E_NO_DEVICE; // no peer access between two devices of the group
            }
            cudaSetDevice(devices[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                for (auto p : g->ctx) brgpu_ctx_destroy(p);
                delete g;
                return BRGPU_E_CUDA;
            }
            cudaGetLastError();
        }
    *out = g;
    return BRGPU_OK;
}

extern "C" void brgpu_group_destroy(brgpu_group *g) {
    if (!g) return;
    for (auto c : g->ctx) brgpu_ctx_destroy(c);
    delete g;
}

extern "C" int brgpu_group_size(const brgpu_group *g) { return g ? (int)g->ctx.size() : 0; }
extern "C" brgpu_ctx *brgpu_group_ctx(brgpu_group *g, int i) { return g && i >= 0 && i < (int)g->ctx.size() ? g->ctx[i] : nullptr; }
extern "C" const char *brgpu_group_last_error(const brgpu_group *g) { return g ? g->err.c_str() : "no group"; }

// contiguous record ranges balanced by bases (prefix sum over lengths): cuts[i] .. cuts[i+1] is device i's
static std::vector<uint64_t> shard_cuts(const uint64_t *off, uint64_t n_reads, int n) {
    std::vector<uint64_t> cuts(n + 1, n_reads);
    cuts[0] = 0;
    const uint64_t total = n_reads ? off[n_reads] - off[0] : 0;
    for (int i = 1; i < n; i++) {
        const uint64_t want = off[0] + total / (uint64_t)n * (uint64_t)i;
        cuts[i] = (uint64_t)(std::lower_bound(off, off + n_reads + 1, want) - off);
        if (cuts[i] < cuts[i - 1]) cuts[i] = cuts[i - 1];
        if (cuts[i] > n_reads) cuts[i] = n_reads;
    }
    return cuts;
}

static void bucket_range(uint64_t n_buckets, int n, int i, uint64_t *b0, uint64_t *b1) {
    const uint64_t per = n_buckets / (uint64_t)n;
    *b0 = per * (uint64_t)i;
    *b1 = i == n - 1 ? n_buckets : per * (uint64_t)(i + 1);
}

// slice [begin, end) of the index space cut into n 1024-aligned pieces (table protocol)
static void slice_range(uint64_t len, int n, int i, uint64_t *b, uint64_t *e) {
    const uint64_t per = (len / (uint64_t)n) & ~1023ULL;
    if (per == 0) {
        *b = i == 0 ? 0 : len;
        *e = len;
        return;
    }
    *b = per * (uint64_t)i;
    *e = i == n - 1 ? len : per * (uint64_t)(i + 1);
}

extern "C" int brgpu_group_set_from_reads(brgpu_group *g, int k, int abundance, int selection, double percent,
                                          brgpu_reads *const *reads, brgpu_set **out_sets) {
    if (!g || !reads || !out_sets) return BRGPU_E_INVALID;
    const int n = (int)g->ctx.size();
    for (int i = 0; i < n; i++) {
        out_sets[i] = nullptr;
        if (!reads[i] || reads[i]->ctx != g->ctx[i]) return group_fail(g, BRGPU_E_INVALID, "reads[i] must live in the group's i-th context");
    }
    if (k < 3 || k > 19 || !(k & 1)) return group_fail(g, BRGPU_E_INVALID, "k must be odd and in 3..=19");
    if (selection == BRGPU_ABUNDANCE_EXPLICIT && abundance < 0) return group_fail(g, BRGPU_E_NEED_ABUNDANCE, "need an abundance threshold or an abundance method");
    if (selection < BRGPU_ABUNDANCE_EXPLICIT || selection > BRGPU_ABUNDANCE_PERCENT_AT_LEAST || abundance > 255)
        return group_fail(g, BRGPU_E_INVALID, "bad abundance selection");
    if (n == 1) {
        int st = brgpu_set_from_reads_ex(g->ctx[0], k, abundance, selection, percent, reads[0], &out_sets[0]);
        return st == BRGPU_OK ? st : group_fail(g, st, brgpu_last_error(g->ctx[0]));
    }
    std::vector<brgpu_set *> sets(n, nullptr);
    auto drop_sets = [&]() {
        for (auto s : sets) brgpu_set_free(s);
    };
    int st;
    const uint64_t table = 1ULL << (2 * k - 1);
    if (k >= 15) {
        // ---- k-mer protocol ----
        std::vector<brgpu_kmers *> parts(n, nullptr);
        auto drop_parts = [&]() {
            for (auto p : parts) brgpu_kmers_free(p);
        };
        const uint64_t n_buckets = table >> GROUP_BUCKET_BITS;
        std::vector<std::vector<uint64_t>> cut_off(n, std::vector<uint64_t>(n + 1)); // cut_off[j][i]: partition j's residue offset at rank i's first bucket
        st = for_each_device(g, [&](int i) {
            int s = brgpu_kmers_create(g->ctx[i], k, reads[i], &parts[i]);
            if (s == BRGPU_OK) s = brgpu_set_new_sliced(g->ctx[i], k, &sets[i]);
            if (s != BRGPU_OK) return s;
            std::vector<uint64_t> cuts(n + 1);
            for (int r = 0; r < n; r++) {
                uint64_t b0, b1;
                bucket_range(n_buckets, n, r, &b0, &b1);
                cuts[r] = b0;
            }
            cuts[n] = n_buckets;
            return brgpu_kmers_offsets_at(parts[i], cuts.data(), (uint64_t)(n + 1), cut_off[i].data());
        });
        auto count_phase = [&](int ab, bool with_set, std::vector<uint64_t> *hist_sum) {
            std::vector<std::vector<uint64_t>> hist(n, std::vector<uint64_t>(256, 0));
            int s = for_each_device(g, [&](int i) {
                uint64_t b0, b1;
                bucket_range(n_buckets, n, i, &b0, &b1);
                std::vector<void *> res, off;
                std::vector<uint64_t> first, last;
                for (int j = 0; j < n; j++) {
                    if (j == i) continue;
                    res.push_back(parts[j]->d_res);
                    off.push_back(parts[j]->d_base);
                    first.push_back(cut_off[j][i]);
                    last.push_back(cut_off[j][i + 1]);
                }
                return brgpu_kmers_count_parts(g->ctx[i], &parts[i], 1, res.data(), off.data(), first.data(), last.data(), n - 1,
                                               b0, b1, ab, with_set ? sets[i] : nullptr, hist_sum ? hist[i].data() : nullptr);
            });
            if (s == BRGPU_OK && hist_sum) {
                hist_sum->assign(256, 0);
                for (int i = 0; i < n; i++)
                    for (int t = 0; t < 256; t++) (*hist_sum)[t] += hist[i][t];
            }
            return s;
        };
        std::vector<uint64_t> hist;
        if (st == BRGPU_OK && selection != BRGPU_ABUNDANCE_EXPLICIT) { // threshold from the global spectrum
            st = count_phase(0, false, &hist);
            if (st == BRGPU_OK) {
                abundance = brgpu_spectrum_threshold(hist.data(), selection, percent);
                if (abundance < 0) st = group_fail(g, BRGPU_E_NO_THRESHOLD, "can't compute the abundance threshold");
            }
        }
        if (st == BRGPU_OK) st = count_phase(abundance, true, &hist);
        // A sparse set travels in rank-compacted form (the same decision as dist.py's _exchange_compacted): every
        // device compacts its slice, and if the blocks of all slices take less than half the bitfield each device
        // pulls all slices into its replica's block array (brgpu_set_compact_pull) and pushes its summary slice to the peers;
        // the replicas then only build the rank directory.  Otherwise the bitfield slices themselves are pushed.
        std::vector<void *> my_blocks(n, nullptr), all_blocks(n, nullptr);
        std::vector<uint64_t> n_blocks(n, 0);
        bool compacted = false;
        if (st == BRGPU_OK) {
            st = for_each_device(g, [&](int i) {
                uint64_t b0, b1;
                bucket_range(n_buckets, n, i, &b0, &b1);
                return brgpu_set_slice_compact(sets[i], b0 << GROUP_BUCKET_BITS, b1 << GROUP_BUCKET_BITS, &my_blocks[i], &n_blocks[i]);
            });
            uint64_t total = 0;
            compacted = st == BRGPU_OK;
            for (int i = 0; i < n; i++) {
                compacted = compacted && my_blocks[i] != nullptr;
                total += n_blocks[i];
            }
            compacted = compacted && total * 8 <= (table >> 3) / 2;
            if (compacted)
                st = for_each_device(g, [&](int i) { return brgpu_set_compact_alloc(sets[i], total, &all_blocks[i]); });
        }
        if (st == BRGPU_OK)
            st = for_each_device(g, [&](int i) {
                uint64_t b0, b1;
                bucket_range(n_buckets, n, i, &b0, &b1);
                const uint64_t byte0 = (b0 << GROUP_BUCKET_BITS) >> 3, bytes = ((b1 - b0) << GROUP_BUCKET_BITS) >> 3;
                if (compacted) { // every replica pulls all slices into its block array: one kernel, all peers in flight
                    std::vector<void *> src(n);
                    for (int j = 0; j < n; j++) src[j] = j == i ? nullptr : my_blocks[j];
                    int s = brgpu_set_compact_pull(sets[i], src.data(), n_blocks.data(), n);
                    if (s != BRGPU_OK) return s;
                }
                for (int j = 0; j < n; j++) {
                    if (!compacted && j != i && cudaMemcpyAsync(sets[j]->d_bits + byte0, sets[i]->d_bits + byte0, bytes, cudaMemcpyDefault,
                                                                g->ctx[i]->stream) != cudaSuccess)
                        return fail(g->ctx[i], BRGPU_E_CUDA, "bitfield slice copy", cudaGetLastError());
                    if (j != i && sets[i]->d_summary && sets[j]->d_summary &&
                        cudaMemcpyAsync((uint8_t *)sets[j]->d_summary + (byte0 >> 6), (uint8_t *)sets[i]->d_summary + (byte0 >> 6), bytes >> 6,
                                        cudaMemcpyDefault, g->ctx[i]->stream) != cudaSuccess)
                        return fail(g->ctx[i], BRGPU_E_CUDA, "summary slice copy", cudaGetLastError());
                }
                return (int)BRGPU_OK;
            });
        if (st == BRGPU_OK)
            st = for_each_device(g, [&](int i) {
                memcpy(sets[i]->hist, hist.data(), 256 * sizeof(uint64_t));
                sets[i]->abundance = abundance;
                return compacted ? brgpu_set_compact_commit(sets[i]) : brgpu_set_commit_slices(sets[i], sets[i]->d_summary != nullptr ? 1 : 0);
            });
        drop_parts();
    } else {
        // ---- table protocol (small k: the tables are at most 128 MiB) ----
        std::vector<brgpu_counts *> tabs(n, nullptr);
        auto drop_tabs = [&]() {
            for (auto c : tabs) brgpu_counts_free(c);
        };
        st = for_each_device(g, [&](int i) {
            int s = brgpu_counts_create(g->ctx[i], k, &tabs[i]);
            if (s == BRGPU_OK) s = brgpu_counts_add_reads(tabs[i], reads[i]);
            if (s == BRGPU_OK) s = brgpu_set_new(g->ctx[i], k, &sets[i]);
            return s;
        });
        if (st == BRGPU_OK)
            st = for_each_device(g, [&](int i) {
                uint64_t b, e;
                slice_range(table, n, i, &b, &e);
                std::vector<void *> peers;
                for (int j = 0; j < n; j++)
                    if (j != i) peers.push_back(tabs[j]->d_counts);
                return brgpu_counts_merge_slice(tabs[i], peers.data(), n - 1, b, e);
            });
        std::vector<uint64_t> hist(256, 0);
        std::vector<std::vector<uint64_t>> hs(n, std::vector<uint64_t>(256, 0));
        if (st == BRGPU_OK)
            st = for_each_device(g, [&](int i) {
                uint64_t b, e;
                slice_range(table, n, i, &b, &e);
                return b < e ? brgpu_counts_spectrum_slice(tabs[i], b, e, hs[i].data()) : (int)BRGPU_OK;
            });
        if (st == BRGPU_OK) {
            for (int i = 0; i < n; i++)
                for (int t = 0; t < 256; t++) hist[t] += hs[i][t];
            if (selection != BRGPU_ABUNDANCE_EXPLICIT) {
                abundance = brgpu_spectrum_threshold(hist.data(), selection, percent);
                if (abundance < 0) st = group_fail(g, BRGPU_E_NO_THRESHOLD, "can't compute the abundance threshold");
            }
        }
        if (st == BRGPU_OK)
            st = for_each_device(g, [&](int i) {
                uint64_t b, e;
                slice_range(table, n, i, &b, &e);
                int s = b < e ? brgpu_set_threshold_slice(sets[i], tabs[i], abundance, b, e) : (int)BRGPU_OK;
                if (s != BRGPU_OK || b >= e) return s;
                for (int j = 0; j < n; j++) {
                    if (j == i) continue;
                    if (cudaMemcpyAsync(sets[j]->d_bits + (b >> 3), sets[i]->d_bits + (b >> 3), (e - b) >> 3, cudaMemcpyDefault,
                                        g->ctx[i]->stream) != cudaSuccess)
                        return fail(g->ctx[i], BRGPU_E_CUDA, "bitfield slice copy", cudaGetLastError());
                }
                return (int)BRGPU_OK;
            });
        if (st == BRGPU_OK)
            for (int i = 0; i < n; i++) {
                memcpy(sets[i]->hist, hist.data(), 256 * sizeof(uint64_t));
                sets[i]->abundance = abundance;
                sets[i]->summary_valid = false;
            }
        drop_tabs();
    }
    if (st != BRGPU_OK) {
        drop_sets();
        return st;
    }
    for (int i = 0; i < n; i++) out_sets[i] = sets[i];
    return BRGPU_OK;
}

extern "C" int brgpu_group_set_from_host_reads(brgpu_group *g, int k, int abundance, int selection, double percent,
                                               const uint8_t *seq_host, const uint64_t *offsets_host, uint64_t n_reads,
                                               brgpu_set **out_sets) {
    if (!g || !offsets_host || !out_sets) return BRGPU_E_INVALID;
    const int n = (int)g->ctx.size();
    const std::vector<uint64_t> cuts = shard_cuts(offsets_host, n_reads, n);
    std::vector<brgpu_reads *> reads(n, nullptr);
    int st = for_each_device(g, [&](int i) {
        return brgpu_reads_upload(g->ctx[i], seq_host, offsets_host + cuts[i], cuts[i + 1] - cuts[i], &reads[i]);
    });
    if (st == BRGPU_OK) st = brgpu_group_set_from_reads(g, k, abundance, selection, percent, reads.data(), out_sets);
    for (auto r : reads) brgpu_reads_free(r);
    return st;
}

extern "C" void brgpu_group_sets_free(brgpu_group *g, brgpu_set **sets) {
    if (!g || !sets) return;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        brgpu_set_free(sets[i]);
        sets[i] = nullptr;
    }
}

// run_correction's chunk body over the group: the records are sharded by bases, every device corrects its
// shard against its replica (one host thread per device), the corrected records come back in input order.
extern "C" int brgpu_group_correct_batch(brgpu_group *g, brgpu_set *const *sets, const uint8_t *methods, uint64_t n_methods,
                                         int confirm, int max_search, int two_side, const uint8_t *seq_host,
                                         const uint64_t *offsets_host, uint64_t n_reads, uint8_t *out_host, uint64_t out_cap,
                                         uint64_t *out_offsets_host, uint64_t *required) {
    if (!g || !sets || !offsets_host || !out_offsets_host) return BRGPU_E_INVALID;
    const int n = (int)g->ctx.size();
    const std::vector<uint64_t> cuts = shard_cuts(offsets_host, n_reads, n);
    std::vector<brgpu_reads *> done(n, nullptr);
    std::vector<uint64_t> bases(n, 0);
    int st = for_each_device(g, [&](int i) {
        brgpu_reads *r = nullptr;
        int s = brgpu_reads_upload(g->ctx[i], seq_host, offsets_host + cuts[i], cuts[i + 1] - cuts[i], &r);
        if (s == BRGPU_OK) s = brgpu_correct_reads(g->ctx[i], sets[i], methods, n_methods, confirm, max_search, two_side, r, &done[i]);
        brgpu_reads_free(r);
        if (s == BRGPU_OK) bases[i] = brgpu_reads_bases(done[i]);
        return s;
    });
    uint64_t total = 0;
    std::vector<uint64_t> at(n + 1, 0);
    for (int i = 0; i < n; i++) {
        at[i] = total;
        total += bases[i];
    }
    at[n] = total;
    if (required) *required = total;
    if (st == BRGPU_OK && (total > out_cap || (!out_host && total))) st = group_fail(g, BRGPU_E_OVERFLOW, "output buffer too small");
    if (st == BRGPU_OK)
        st = for_each_device(g, [&](int i) {
            uint64_t req = 0;
            std::vector<uint64_t> off(cuts[i + 1] - cuts[i] + 1);
            int s = brgpu_reads_download(done[i], out_host + at[i], out_cap - at[i], off.data(), &req);
            if (s == BRGPU_OK)
                for (size_t r = 0; r < off.size(); r++) out_offsets_host[cuts[i] + r] = at[i] + off[r];
            return s;
        });
    for (auto r : done) brgpu_reads_free(r);
    return st;
}
