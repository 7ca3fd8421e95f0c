/*
 * vec_shim.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A loadable-extension-shaped shared object that registers the scalar SQL
 * function `vec_distance_cosine(blob, blob) -> REAL` on one SQLite connection,
 * so the reference's search statement (image_database.py:1564-1574) can be run
 * verbatim by the real SQLite in this image when the third-party `sqlite-vec`
 * extension is absent (it is: SURVEY.md §0.5).  The arithmetic is
 * oracle_ref.c's restatement of sqlite-vec's scalar float32 cosine distance;
 * PARITY UNPINNED applies here exactly as it does there.
 *
 * There are no SQLite development headers in the image, so the handful of
 * public prototypes used are declared here; they resolve against the
 * libsqlite3.so.0 that Python's _sqlite3 module already has loaded (the
 * Makefile links it explicitly).  Stands in for `sqlite_vec.load(conn)`
 * (image_database.py:264, 1481).
 */
#include <stdint.h>
#include <stddef.h>

typedef struct sqlite3 sqlite3;
typedef struct sqlite3_context sqlite3_context;
typedef struct sqlite3_value sqlite3_value;

extern int sqlite3_create_function_v2(sqlite3 *, const char *, int, int, void *,
                                      void (*)(sqlite3_context *, int, sqlite3_value **),
                                      void (*)(sqlite3_context *, int, sqlite3_value **),
                                      void (*)(sqlite3_context *), void (*)(void *));
extern const void *sqlite3_value_blob(sqlite3_value *);
extern int sqlite3_value_bytes(sqlite3_value *);
extern int sqlite3_value_type(sqlite3_value *);
extern void sqlite3_result_double(sqlite3_context *, double);
extern void sqlite3_result_error(sqlite3_context *, const char *, int);

#define SHIM_SQLITE_UTF8 1
#define SHIM_SQLITE_DETERMINISTIC 0x000000800
#define SHIM_SQLITE_BLOB 4

extern float oracle_cosine_distance_f32(const float *a, const float *b, int64_t dim);

static void shim_distance_cosine(sqlite3_context *ctx, int argc, sqlite3_value **argv)
{
    (void)argc;
    if (sqlite3_value_type(argv[0]) != SHIM_SQLITE_BLOB ||
        sqlite3_value_type(argv[1]) != SHIM_SQLITE_BLOB) {
        sqlite3_result_error(ctx, "vec_distance_cosine: both arguments must be float32 blobs", -1);
        return;
    }
    int na = sqlite3_value_bytes(argv[0]);
    int nb = sqlite3_value_bytes(argv[1]);
    if (na % 4 != 0 || nb % 4 != 0 || na == 0 || nb == 0) {
        sqlite3_result_error(ctx, "vec_distance_cosine: invalid float32 vector BLOB length", -1);
        return;
    }
    if (na != nb) {
        sqlite3_result_error(ctx, "vec_distance_cosine: vector dimension mismatch", -1);
        return;
    }
    const float *a = (const float *)sqlite3_value_blob(argv[0]);
    const float *b = (const float *)sqlite3_value_blob(argv[1]);
    /* a NaN result becomes SQL NULL inside sqlite3_result_double, as it does
     * for the real extension */
    sqlite3_result_double(ctx, (double)oracle_cosine_distance_f32(a, b, na / 4));
}

int sqlite3_extension_init(sqlite3 *db, char **err, const void *api)
{
    (void)err; (void)api;
    return sqlite3_create_function_v2(db, "vec_distance_cosine", 2,
                                      SHIM_SQLITE_UTF8 | SHIM_SQLITE_DETERMINISTIC, NULL,
                                      shim_distance_cosine, NULL, NULL, NULL);
}
