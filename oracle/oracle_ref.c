/*
 * oracle_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's brute-force KNN path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under clip_database_b200/ may link or call it.
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party `sqlite-vec`
 * extension (requirements.txt:6, `sqlite-vec>=0.0.1`, no lockfile), which is
 * absent from /root/reference and from this image; the reference ships no
 * tests or golden vectors for this path.  What follows
 * restates sqlite-vec's published scalar `vec_distance_cosine` on two float32
 * blobs, anchored on the reference's own call site:
 *
 *   image_database.py:1564-1574   SELECT ... vec_distance_cosine(vec0.embedding, ?)
 *                                 ... ORDER BY distance ASC LIMIT ?
 *   image_database.py:1582-1589   execute, fetchall, similarity = 1.0 - distance
 *
 * Restated semantics:
 *   distance: three float32 accumulators (dot, aMag, bMag) updated in index
 *             order by one scalar loop; result 1 - dot/(sqrt(aMag)*sqrt(bMag))
 *             with the roots, product, divide and subtract in double, narrowed
 *             to float32, then widened to the double SQLite sees.
 *   top-k:    SQLite's ORDER BY ... LIMIT sorter keeps the k smallest
 *             (distance, scan-sequence) pairs; exact ties come back in scan
 *             (= rowid) order and a later row that ties the current k-th is
 *             not admitted (probed on SQLite 3.45.1).
 *   NaN:      sqlite3_result_double(NaN) stores NULL; NULL sorts first in ASC
 *             order.  The oracle reports such rows separately (n_nan) and
 *             leaves policy to the caller (tests cover both policies).
 *
 * Build: see oracle/Makefile.  Compiled with -ffp-contract=off so the scalar
 * loop is not fused into FMAs (the x86-64 sqlite-vec wheels are built without
 * -mfma).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* sqlite-vec `distance_cosine_float` restated (a = stored row, b = query;
 * image_database.py:1567 passes the vec0 column first). */
float oracle_cosine_distance_f32(const float *a, const float *b, int64_t dim)
{
    float dot = 0.0f, a_mag = 0.0f, b_mag = 0.0f;
    for (int64_t i = 0; i < dim; i++) {
        dot += a[i] * b[i];
        a_mag += a[i] * a[i];
        b_mag += b[i] * b[i];
    }
    return (float)(1.0 - ((double)dot / (sqrt((double)a_mag) * sqrt((double)b_mag))));
}

/* All N distances, scan order. */
void oracle_distances(const float *rows, int64_t n, int64_t dim,
                      const float *query, float *out)
{
    for (int64_t r = 0; r < n; r++)
        out[r] = oracle_cosine_distance_f32(rows + r * dim, query, dim);
}

/* Multi-threaded variant of the same per-row arithmetic (pthreads over
 * contiguous row ranges; the image's gcc has no libgomp).  The reference
 * itself is single threaded; this exists so the benchmark can also quote
 * "what if the CPU path used every core". */
typedef struct {
    const float *rows; const float *query; float *out;
    int64_t lo, hi, dim;
} dist_job_t;

static void *dist_worker(void *p)
{
    dist_job_t *j = (dist_job_t *)p;
    for (int64_t r = j->lo; r < j->hi; r++)
        j->out[r] = oracle_cosine_distance_f32(j->rows + r * j->dim, j->query, j->dim);
    return NULL;
}

void oracle_distances_mt(const float *rows, int64_t n, int64_t dim,
                         const float *query, float *out, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    dist_job_t job[256];
    int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        int64_t lo = t * per, hi = lo + per;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        job[t] = (dist_job_t){ rows, query, out, lo, hi, dim };
        pthread_create(&tid[t], NULL, dist_worker, &job[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
}

typedef struct {
    double dist;   /* what SQLite compares: the float32 widened to double */
    int64_t seq;   /* scan sequence = position in rowid order */
} entry_t;

static int entry_less(const entry_t *x, const entry_t *y)
{
    if (x->dist < y->dist) return 1;
    if (x->dist > y->dist) return 0;
    return x->seq < y->seq;
}

/* max-heap on (dist, seq): root = current k-th (worst kept) */
static void sift_down(entry_t *h, int64_t n, int64_t i)
{
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && entry_less(&h[m], &h[l])) m = l;
        if (r < n && entry_less(&h[m], &h[r])) m = r;
        if (m == i) return;
        entry_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

static void sift_up(entry_t *h, int64_t i)
{
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!entry_less(&h[p], &h[i])) return;
        entry_t t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}

static int cmp_entry(const void *a, const void *b)
{
    const entry_t *x = (const entry_t *)a, *y = (const entry_t *)b;
    if (entry_less(x, y)) return -1;
    if (entry_less(y, x)) return 1;
    return 0;
}

/*
 * The whole statement: scan rows in order, score each, keep the k smallest
 * (distance, sequence).  `mask` (nullable, one byte per row, non-zero = row
 * passes the WHERE clause / joins) restates image_database.py:1569-1571.
 * k < 0 means no LIMIT (SQLite semantics).  Returns the number of result rows
 * written (<= min(k, rows passing)); rows whose distance is NaN are counted in
 * *n_nan and not written (the caller applies the NULL-sorts-first policy).
 */
int64_t oracle_knn(const float *rows, const int64_t *rowids, int64_t n, int64_t dim,
                   const float *query, const uint8_t *mask, int64_t k,
                   int64_t *out_rowids, float *out_dist, int64_t *out_seq,
                   int64_t *n_nan)
{
    int64_t cap = (k < 0 || k > n) ? n : k;
    int64_t nan_rows = 0, size = 0;
    if (n_nan) *n_nan = 0;
    entry_t *heap = (entry_t *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(entry_t));
    if (!heap) return -1;
    for (int64_t r = 0; r < n; r++) {
        if (mask && !mask[r]) continue;
        float d = oracle_cosine_distance_f32(rows + r * dim, query, dim);
        if (d != d) { nan_rows++; continue; }
        if (cap == 0) continue;
        entry_t e = { (double)d, r };
        if (size < cap) {
            heap[size] = e;
            sift_up(heap, size);
            size++;
        } else if (entry_less(&e, &heap[0])) {
            heap[0] = e;
            sift_down(heap, size, 0);
        }
    }
    qsort(heap, (size_t)size, sizeof(entry_t), cmp_entry);
    for (int64_t i = 0; i < size; i++) {
        if (out_rowids) out_rowids[i] = rowids ? rowids[heap[i].seq] : heap[i].seq;
        if (out_dist) out_dist[i] = (float)heap[i].dist;
        if (out_seq) out_seq[i] = heap[i].seq;
    }
    free(heap);
    if (n_nan) *n_nan = nan_rows;
    return size;
}

/* fp64 ground truth for the tolerance statement (oracle_exact):
 * every sum in double, no float32 narrowing. */
void oracle_distances_f64(const float *rows, int64_t n, int64_t dim,
                          const float *query, double *out)
{
    double b_mag = 0.0;
    for (int64_t i = 0; i < dim; i++) b_mag += (double)query[i] * (double)query[i];
    for (int64_t r = 0; r < n; r++) {
        const float *a = rows + r * dim;
        double dot = 0.0, a_mag = 0.0;
        for (int64_t i = 0; i < dim; i++) {
            dot += (double)a[i] * (double)query[i];
            a_mag += (double)a[i] * (double)a[i];
        }
        out[r] = 1.0 - dot / (sqrt(a_mag) * sqrt(b_mag));
    }
}

/* Cheap deterministic generator so the CPU baseline can build a bounded sample
 * without paying numpy's Gaussian sampler: approximately normal entries
 * (sum of 4 uniforms, centred), rows L2-normalised in float32.  Test/bench
 * data only; its statistics do not affect the loop being timed. */
static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_fill_unit_rows(float *rows, int64_t n, int64_t dim, uint64_t seed)
{
    for (int64_t r = 0; r < n; r++) {
        uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(r + 1));
        float *v = rows + r * dim;
        double ss = 0.0;
        for (int64_t i = 0; i < dim; i++) {
            uint64_t z = splitmix64(&s);
            float u = (float)(z & 0xFFFF) + (float)((z >> 16) & 0xFFFF) +
                      (float)((z >> 32) & 0xFFFF) + (float)((z >> 48) & 0xFFFF);
            float x = u * (1.0f / 65536.0f) - 2.0f;
            v[i] = x;
            ss += (double)x * x;
        }
        float inv = (float)(1.0 / sqrt(ss > 0 ? ss : 1.0));
        for (int64_t i = 0; i < dim; i++) v[i] *= inv;
    }
}
