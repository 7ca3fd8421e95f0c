/* abi_smoke.c — the C ABI of include/clipdb.h used from plain C (no Python, no torch).
 *
 *   gcc -O2 -I include examples/abi_smoke.c -o abi_smoke -L clip_database_b200 -lclipdb_b200 \
 *       -Wl,-rpath,$PWD/clip_database_b200 -lm
 *   ./abi_smoke            (needs a B200; exits non-zero on any mismatch)
 *
 * Loads 20,000 pseudo-random 1152-d rows, plants two copies of the query, searches (k = 8) and
 * checks the result against a brute-force loop written the way the reference's per-row
 * arithmetic is (three float accumulators in index order, double finish — oracle/oracle_ref.c),
 * then does the same through the sign-code search.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clipdb.h"

#define N 20000
#define D 1152
#define K 8

static uint32_t lcg_state = 12345u;
static float lcg_unit(void) {
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)((lcg_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

static float ref_distance(const float *a, const float *b) {
    float dot = 0.f, am = 0.f, bm = 0.f;
    for (int i = 0; i < D; i++) {
        dot += a[i] * b[i];
        am += a[i] * a[i];
        bm += b[i] * b[i];
    }
    return (float)(1.0 - (double)dot / (sqrt((double)am) * sqrt((double)bm)));
}

#define CHECK(rc)                                                                      \
    do {                                                                               \
        int rc__ = (rc);                                                               \
        if (rc__ != CLIPDB_OK) {                                                       \
            fprintf(stderr, "%s -> %d: %s\n", #rc, rc__, clipdb_last_error(ctx));      \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

int main(void) {
    clipdb_ctx *ctx = NULL;
    if (clipdb_create(0, &ctx) != CLIPDB_OK) {
        fprintf(stderr, "clipdb_create failed: no usable CUDA device (there is no CPU fallback)\n");
        return 2;
    }
    float *rows = malloc(sizeof(float) * N * D), *q = malloc(sizeof(float) * D);
    uint8_t *codes = malloc((size_t)N * D), qcode[D];
    for (size_t i = 0; i < (size_t)N * D; i++) rows[i] = lcg_unit();
    for (int i = 0; i < D; i++) q[i] = lcg_unit();
    memcpy(rows + (size_t)777 * D, q, sizeof(float) * D);        /* exact copies: a tie, rowid order */
    memcpy(rows + (size_t)15000 * D, q, sizeof(float) * D);
    CHECK(clipdb_load_rows(ctx, rows, NULL, N, D));

    int64_t ids[K];
    float dist[K];
    int32_t n = 0;
    int64_t nan = 0;
    CHECK(clipdb_search(ctx, q, 1, K, CLIPDB_METRIC_COSINE, 0, ids, dist, &n, &nan));
    if (n != K || nan != 0 || ids[0] != 777 || ids[1] != 15000) {
        fprintf(stderr, "unexpected head of the result: n=%d ids %lld %lld\n", n, (long long)ids[0], (long long)ids[1]);
        return 1;
    }
    /* brute force on the host: the K smallest (distance, position) pairs */
    float best_d[K];
    int64_t best_i[K];
    int found = 0;
    for (int64_t r = 0; r < N; r++) {
        const float d = ref_distance(rows + r * D, q);
        int at = found;
        while (at > 0 && best_d[at - 1] > d) at--;
        if (at >= K) continue;
        for (int j = (found < K ? found : K - 1); j > at; j--) {
            best_d[j] = best_d[j - 1];
            best_i[j] = best_i[j - 1];
        }
        best_d[at] = d;
        best_i[at] = r;
        if (found < K) found++;
    }
    for (int j = 0; j < K; j++) {
        const double tol = 1e-5 * fmax(fabs((double)best_d[j]), 1.0);
        if (fabs((double)dist[j] - (double)best_d[j]) > tol || (ids[j] != best_i[j] && fabs((double)best_d[j] - (double)ref_distance(rows + ids[j] * D, q)) > tol)) {
            fprintf(stderr, "rank %d: got (%lld, %.9g), expected (%lld, %.9g)\n", j, (long long)ids[j], dist[j],
                    (long long)best_i[j], best_d[j]);
            return 1;
        }
    }

    /* sign-code fallback: score = popcount(q AND row) (plain popcount mode) */
    for (size_t i = 0; i < (size_t)N * D; i++) codes[i] = rows[i] >= 0.f;
    for (int i = 0; i < D; i++) qcode[i] = q[i] >= 0.f;
    CHECK(clipdb_load_codes(ctx, codes, NULL, N, D));
    int64_t bid[K];
    int32_t score[K];
    CHECK(clipdb_binary_search(ctx, qcode, K, CLIPDB_SCORE_POPCOUNT, 0, bid, score, &n));
    int ones = 0;
    for (int i = 0; i < D; i++) ones += qcode[i];
    if (n != K || bid[0] != 777 || bid[1] != 15000 || score[0] != ones || score[1] != ones || score[2] > ones) {
        fprintf(stderr, "sign-code search: unexpected result (%lld %lld, scores %d %d, |q| = %d)\n", (long long)bid[0],
                (long long)bid[1], score[0], score[1], ones);
        return 1;
    }
    printf("abi_smoke ok: %d results, best distance %.3g, %lld kernel launches, ABI v%d\n", K, dist[0],
           (long long)clipdb_launch_count(ctx), clipdb_abi_version());
    clipdb_destroy(ctx);
    free(rows);
    free(q);
    free(codes);
    return 0;
}
