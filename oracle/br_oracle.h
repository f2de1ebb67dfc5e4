/*
 * br_oracle — CPU restatement of natir/br's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libbrgpu.so, br_b200/) never links,
 * imports or calls it and has no CPU fallback.
 *
 * Why a restatement: the reference is Rust and no Rust toolchain exists in the build
 * image; its arithmetic lives in un-vendored crates (pcon @0184ae77, cocktail @f63f0ba9,
 * bio 1.6.0).  Each function below cites the reference file:line it follows.
 *
 * Parity status:
 *   pinned   — 2-bit code, canonical-by-parity, index>>1, LSB-first bitfield, `.solid`
 *              container, count>abundance: against tests/golden/br_reads.k11.a2.solid
 *              regenerated from tests/golden/br_reads.fa.gz byte-for-byte;
 *              correctors One/Two/Graph/GapSize and the scan loop: against the 55+1
 *              known-answer tests in tests/golden/kats.json (from the reference's
 *              #[test] functions).
 *   PARITY UNPINNED — u8 saturation at 255, Spectrum/FirstMinimum, and the
 *              bio-1.6.0 traceback tie-breaking that Greedy's returned offsets depend
 *              on (the reference's active Greedy KATs only pin "read unchanged").
 */
#ifndef BR_ORACLE_H
#define BR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* correction method ids — order of cli::CorrectionMethod (src/cli.rs:121-131) */
enum { BRO_ONE = 0, BRO_TWO = 1, BRO_GRAPH = 2, BRO_GREEDY = 3, BRO_GAP_SIZE = 4 };

/* ---- k-mer primitives (cocktail::kmer) ---- */
uint64_t bro_nuc2bit(uint8_t b);
uint8_t bro_bit2nuc(uint64_t x);
uint64_t bro_seq2bit(const uint8_t *seq, size_t len);
uint64_t bro_revcomp(uint64_t kmer, int k);
uint64_t bro_canonical(uint64_t kmer, int k);

/* ---- pcon::solid::Solid ---- */
typedef struct bro_set bro_set;
bro_set *bro_set_new(int k);                                            /* Solid::new */
bro_set *bro_set_from_bitfield(int k, const uint8_t *bits, size_t n);   /* body of a .solid payload */
/* set::Hash (src/set/hash.rs): the same handle type, backed by an exact hash set of canonical k-mers */
bro_set *bro_hash_new(int k);
void bro_hash_add_reads(bro_set *, const uint8_t *seq, const uint64_t *offsets, size_t n_reads); /* Hash::from_fasta */
size_t bro_hash_size(const bro_set *);
void bro_set_free(bro_set *);
int bro_set_k(const bro_set *);
void bro_set_set(bro_set *, uint64_t kmer, int value);                  /* Solid::set (canonicalises) */
int bro_set_get(const bro_set *, uint64_t kmer);                        /* Solid::get / Pcon::get src/set/pcon.rs:189 */
const uint8_t *bro_set_bits(const bro_set *, size_t *nbytes);
void bro_set_get_batch(const bro_set *, const uint64_t *kmers, size_t n, uint8_t *out);
void bro_set_insert_all_kmers(bro_set *, const uint8_t *seq, size_t len); /* Tokenizer loop of the KATs */

/* ---- pcon::counter::Counter<u8> + Spectrum + Solid::from_count (src/main.rs:72-115) ---- */
typedef struct bro_counter bro_counter;
bro_counter *bro_counter_new(int k);
void bro_counter_free(bro_counter *);
/* count every canonical k-mer of every read with len >= k; threads<=1 → serial */
void bro_counter_count(bro_counter *, const uint8_t *seq, const uint64_t *offsets, size_t n_reads, int threads);
const uint8_t *bro_counter_raw(const bro_counter *, size_t *n);
void bro_spectrum(const bro_counter *, uint64_t hist[256], int threads);
int bro_first_minimum(const uint64_t hist[256]);                         /* -1 == None */
int bro_spectrum_threshold(const uint64_t hist[256], int method, double percent); /* 2 rarefaction, 3 at most, 4 at least */
bro_set *bro_solid_from_count(const bro_counter *, int abundance, int threads);

/* ---- correct module ---- */
/* alt_nucs / next_nucs (src/correct/mod.rs:114-128): returns count, fills out[4] */
int bro_alt_nucs(const bro_set *, uint64_t kmer, uint64_t out[4]);
int bro_next_nucs(const bro_set *, uint64_t kmer, uint64_t out[4]);

/* Corrector::correct_error for one method; returns number of emitted bases, or -1 for None.
 * *offset receives the read offset.  out must hold `cap` bytes (emission is truncated to
 * cap, return value is the untruncated length). */
long bro_correct_error(const bro_set *, int method, int confirm, int max_search, uint64_t kmer,
                       const uint8_t *seq, size_t len, uint8_t *out, size_t cap, size_t *offset);

/* Corrector::correct (src/correct/mod.rs:53-107) for one method. Returns output length;
 * writes min(length, cap) bytes. */
size_t bro_correct(const bro_set *, int method, int confirm, int max_search, const uint8_t *seq,
                   size_t len, uint8_t *out, size_t cap);

/* run_correction's per-record body (src/lib.rs:44-55): fold methods, reversed pass unless
 * two_side.  Batch over reads, input order preserved (serial path semantics,
 * src/lib.rs:21-69); threads>1 parallelises over reads with OpenMP (the rayon analogue). */
typedef struct bro_result bro_result;
bro_result *bro_run_correction(const bro_set *, const uint8_t *methods, size_t n_methods, int confirm,
                               int max_search, int two_side, const uint8_t *seq, const uint64_t *offsets,
                               size_t n_reads, int threads);
const uint8_t *bro_result_data(const bro_result *);
const uint64_t *bro_result_offsets(const bro_result *); /* n_reads + 1 */
void bro_result_free(bro_result *);

/* bio 1.6.0 pairwise::Aligner::global restatement (gap_open=-1, gap_extend=-1, +1/-1).
 * ops: 0 Match, 1 Subst, 2 Del, 3 Ins.  Returns number of operations (<= cap). */
size_t bro_bio_global(const uint8_t *x, size_t m, const uint8_t *y, size_t n, uint8_t *ops, size_t cap);
/* Greedy::match_alignement (src/correct/greedy.rs:56-89); returns 1 and *off if Some */
int bro_match_alignement(const uint8_t *before, size_t nb, const uint8_t *read, size_t nr,
                         const uint8_t *corr, size_t nc, long *off);

int bro_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
