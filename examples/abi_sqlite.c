/* abi_sqlite.c — from a CLIP-database SQLite file to answers, through the C ABI only (no Python, no torch):
 * what a C / cgo / JNI host of the reference would do instead of `sqlite3.connect` + `sqlite_vec.load` + the
 * search statement (image_database.py:1475-1583).
 *
 *   gcc -O2 -I include examples/abi_sqlite.c -o abi_sqlite -L clip_database_b200 -lclipdb_b200 \
 *       -Wl,-rpath,$PWD/clip_database_b200
 *   ./abi_sqlite images.db query.f32 [k]      (query.f32 = 1152 raw little-endian float32)
 *
 * Reserves a store, streams vec0 JOIN image_embeddings JOIN images into HBM with the native reader
 * (clipdb_append_sqlite: SQLite's C library -> pinned slots -> DMA), keeps rowid -> file_path on the host from
 * the chunk callback, searches, and prints "similarity<TAB>file_path" lines best first — the reference's result
 * list (image_database.py:1586-1589).  Exits 2 with "no CPU fallback" when there is no CUDA device.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clipdb.h"

#define D 1152

struct meta {
    int64_t n, cap;
    int64_t *rowids;   /* ascending: the scan order */
    char **paths;
};

/* clipdb_sqlite_chunk_fn: remember what the host needs per row */
static int on_chunk(void *user, int64_t n, const int64_t *rowids, const int64_t *image_ids, const double *last_modified,
                    const char *paths, int64_t paths_bytes) {
    struct meta *m = (struct meta *)user;
    (void)image_ids;
    (void)last_modified;
    (void)paths_bytes;
    if (m->n + n > m->cap) {
        m->cap = (m->n + n) * 2;
        m->rowids = (int64_t *)realloc(m->rowids, (size_t)m->cap * sizeof(int64_t));
        m->paths = (char **)realloc(m->paths, (size_t)m->cap * sizeof(char *));
        if (!m->rowids || !m->paths) return 1;
    }
    const char *p = paths;
    for (int64_t i = 0; i < n; i++) {
        m->rowids[m->n] = rowids[i];
        m->paths[m->n] = strdup(p);
        p += strlen(p) + 1;
        m->n++;
    }
    return 0;
}

static const char *path_of(const struct meta *m, int64_t rowid) {
    int64_t lo = 0, hi = m->n - 1;
    while (lo <= hi) {
        const int64_t mid = (lo + hi) / 2;
        if (m->rowids[mid] == rowid) return m->paths[mid];
        if (m->rowids[mid] < rowid) lo = mid + 1;
        else hi = mid - 1;
    }
    return "?";
}

#define CHECK(call)                                                                  \
    do {                                                                             \
        const int rc_ = (call);                                                      \
        if (rc_ != CLIPDB_OK) {                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, clipdb_last_error(ctx)); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s images.db query.f32 [k]\n", argv[0]);
        return 64;
    }
    const int k = argc > 3 ? atoi(argv[3]) : 10;
    float query[D];
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(query, sizeof(float), D, f) != D) {
        fprintf(stderr, "cannot read %d float32 from %s\n", D, argv[2]);
        return 64;
    }
    fclose(f);

    clipdb_ctx *ctx = NULL;
    if (clipdb_create(0, &ctx) != CLIPDB_OK) {
        fprintf(stderr, "clipdb_create failed: no usable CUDA device (there is no CPU fallback)\n");
        return 2;
    }
    struct meta m = {0, 0, NULL, NULL};
    int64_t vec0_rows = 0, joined = 0;
    CHECK(clipdb_reserve_rows(ctx, 1024, D, /*explicit_rowids=*/1, CLIPDB_PLACE_DEVICE));   /* appends grow it */
    CHECK(clipdb_append_sqlite(ctx, argv[1], INT64_MIN, INT64_MAX, 8192, on_chunk, &m, &vec0_rows, &joined));
    if (joined != m.n || clipdb_num_rows(ctx) != joined) {
        fprintf(stderr, "loaded %lld rows, callback saw %lld\n", (long long)joined, (long long)m.n);
        return 1;
    }
    fprintf(stderr, "loaded %lld of %lld vec0 rows\n", (long long)joined, (long long)vec0_rows);

    int64_t *ids = (int64_t *)malloc((size_t)k * sizeof(int64_t));
    float *dist = (float *)malloc((size_t)k * sizeof(float));
    int32_t found = 0;
    int64_t nan_rows = 0;
    CHECK(clipdb_search(ctx, query, 1, k, CLIPDB_METRIC_COSINE, /*use_mask=*/0, ids, dist, &found, &nan_rows));
    for (int i = 0; i < found; i++) printf("%.6f\t%s\n", 1.0 - (double)dist[i], path_of(&m, ids[i]));
    clipdb_destroy(ctx);
    return 0;
}
