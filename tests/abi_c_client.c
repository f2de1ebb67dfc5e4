/* A host written in plain C against include/brgpu.h (tests/test_abi.py): the whole hot path through the C ABI the way
 * a cgo / Rust-FFI / C caller would drive it — set from host reads, one correction batch with the caller-owned output
 * buffer grown on BRGPU_E_OVERFLOW, lookups, teardown.  Without a CUDA device it must stop at brgpu_ctx_create with
 * BRGPU_E_NO_DEVICE (exit status 2): there is no CPU path to fall back to.  With one, it prints a digest of the
 * corrected bytes (exit status 0) that the GPU test compares with the oracle's.
 *   abi_c_client K ABUNDANCE < reads   (stdin: one read per line) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "brgpu.h"

int main(int argc, char **argv) {
    int k = argc > 1 ? atoi(argv[1]) : 11, abundance = argc > 2 ? atoi(argv[2]) : 2;
    size_t cap = 1 << 20, len = 0, n = 0, ocap = 1 << 16;
    uint8_t *seq = malloc(cap);
    uint64_t *off = malloc(ocap * sizeof(uint64_t));
    char *line = NULL;
    size_t lcap = 0;
    ssize_t got;
    brgpu_ctx *ctx = NULL;
    brgpu_set *set = NULL;
    int st;

    off[0] = 0;
    while ((got = getline(&line, &lcap, stdin)) >= 0) {
        while (got && (line[got - 1] == '\n' || line[got - 1] == '\r')) got--;
        if (len + (size_t)got > cap) seq = realloc(seq, cap = 2 * (len + (size_t)got));
        if (n + 2 > ocap) off = realloc(off, (ocap *= 2) * sizeof(uint64_t));
        memcpy(seq + len, line, (size_t)got);
        len += (size_t)got;
        off[++n] = len;
    }
    st = brgpu_ctx_create(0, NULL, &ctx);
    if (st != BRGPU_OK) {
        printf("brgpu_ctx_create: status %d\n", st);
        return st == BRGPU_E_NO_DEVICE ? 2 : 1;
    }
    st = brgpu_set_from_host_reads(ctx, k, abundance, BRGPU_ABUNDANCE_EXPLICIT, seq, off, n, &set);
    if (st != BRGPU_OK) {
        printf("brgpu_set_from_host_reads: status %d (%s)\n", st, brgpu_last_error(ctx));
        return 1;
    }
    {
        const uint8_t methods[2] = {BRGPU_ONE, BRGPU_TWO};
        uint64_t out_cap = len / 2 + 16, need = 0, i, h = 1469598103934665603ULL; /* too small on purpose: E_OVERFLOW first */
        uint8_t *out = malloc(out_cap);
        uint64_t *out_off = malloc((n + 1) * sizeof(uint64_t));
        int calls = 0;
        for (;;) {
            st = brgpu_correct_batch(ctx, set, methods, 2, 5, 7, 0, seq, off, n, out, out_cap, out_off, &need);
            calls++;
            if (st == BRGPU_E_OVERFLOW && need > out_cap) {
                out = realloc(out, out_cap = need);
                continue;
            }
            break;
        }
        if (st != BRGPU_OK) {
            printf("brgpu_correct_batch: status %d (%s)\n", st, brgpu_last_error(ctx));
            return 1;
        }
        for (i = 0; i < need; i++) h = (h ^ out[i]) * 1099511628211ULL; /* FNV-1a of the corrected bytes */
        printf("reads %llu bases_in %llu bases_out %llu calls %d fnv1a %016llx k %d\n", (unsigned long long)n, (unsigned long long)len,
               (unsigned long long)out_off[n], calls, (unsigned long long)h, brgpu_set_k(set));
        free(out);
        free(out_off);
    }
    brgpu_set_free(set);
    brgpu_ctx_destroy(ctx);
    free(seq);
    free(off);
    free(line);
    return 0;
}
