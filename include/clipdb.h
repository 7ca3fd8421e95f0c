/*
 * clipdb.h — C ABI of the B200-native brute-force KNN path for CLIP-database.
 *
 * This is the drop-in boundary: everything the reference's search path gets
 * from the third-party sqlite-vec extension + SQLite's ORDER BY/LIMIT sorter,
 * plus the numpy query arithmetic that precedes it, behind plain C entry
 * points (extern "C", pointers and sizes only, no torch / C++ types).
 * `idb` = /root/reference/image_database.py.
 *
 * Conventions
 *   - every function returns CLIPDB_OK (0) or a CLIPDB_ERR_* code; the message
 *     for the last failure on a context is clipdb_last_error(ctx).  No C++
 *     exception and no sticky CUDA error crosses this boundary.
 *   - a context owns one GPU's resident row store, workspaces and a stream.
 *     Calls on one context are serialised on its stream; several contexts
 *     (one per GPU, or several per GPU) coexist; there is no global state.
 *   - "host" entry points are synchronous like the reference (results are on
 *     the host when they return); "*_device" entry points take device
 *     pointers, enqueue on the context's stream and return without syncing.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     CLIPDB_ERR_CUDA.
 */
#ifndef CLIPDB_H
#define CLIPDB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clipdb_ctx clipdb_ctx;

#define CLIPDB_OK               0
#define CLIPDB_ERR_INVALID      1   /* bad argument */
#define CLIPDB_ERR_CUDA         2   /* CUDA runtime / launch failure, or no device */
#define CLIPDB_ERR_NOMEM        3   /* device or pinned-host allocation failed */
#define CLIPDB_ERR_STATE        4   /* e.g. search before any rows were loaded */
#define CLIPDB_ERR_UNSUPPORTED  5

/* Distance kinds.  COSINE is what the reference asks for
 * (`vec_distance_cosine(vec0.embedding, ?)`, idb:1567).  L2 is the metric a
 * `vec0(embedding float[1152])` table declares by default (idb:290-294), which
 * only sqlite-vec's MATCH form would use; shipped as a variant of the same scan. */
#define CLIPDB_METRIC_COSINE    0
#define CLIPDB_METRIC_L2        1

/* bit flags reported by clipdb_blend (which fallback of idb:1390-1395 /
 * idb:558-570 / idb:591-603 fired) */
#define CLIPDB_BLEND_POSITIVE_ZERO_NORM  1  /* blend had zero norm -> first query used (idb:1393-1395) */
#define CLIPDB_BLEND_NEGATIVE_ZERO_NORM  2  /* zero after negatives -> original restored (idb:560-570) */

/* ABI version of this header (bumped on any signature change). */
int clipdb_abi_version(void);

/* Hex SHA-256 of the sources (csrc/ and this header) the library was compiled from, so a binding can
 * refuse to run kernels older than the code next to it (the built library is not under version control). */
const char *clipdb_source_hash(void);

/* ---- context -------------------------------------------------------------
 * Replaces: sqlite3.connect + sqlite_vec.load(conn) per search (idb:1475-1484);
 * here the store is loaded once and stays resident in HBM. */
int  clipdb_create(int device, clipdb_ctx **out);
void clipdb_destroy(clipdb_ctx *ctx);
const char *clipdb_last_error(const clipdb_ctx *ctx);

/* Run on a caller-provided CUDA stream (cudaStream_t passed as void*), e.g.
 * torch's current stream so torch.cuda.Event brackets the kernels.  NULL is
 * CUDA's legacy default stream (torch's default).  clipdb_use_own_stream
 * returns to the non-blocking stream the context created for itself. */
int clipdb_set_stream(clipdb_ctx *ctx, void *cuda_stream);
int clipdb_use_own_stream(clipdb_ctx *ctx);
int clipdb_synchronize(clipdb_ctx *ctx);

/* Tuning knobs for experiments ("scan_variant", "scan_ctas", ...).  Unknown
 * names fail with CLIPDB_ERR_INVALID.  Read-only names for clipdb_get_option:
 * "sm_count", "device_free_bytes", "device_total_bytes" (cudaMemGetInfo of the
 * context's device: what a loader needs to choose between an all-HBM and a
 * tiered store, see clipdb_reserve_rows). */
int clipdb_set_option(clipdb_ctx *ctx, const char *name, int64_t value);
int clipdb_get_option(clipdb_ctx *ctx, const char *name, int64_t *value);

/* Kernels launched by this context since creation (bench.py "gpu_launches"). */
int64_t clipdb_launch_count(const clipdb_ctx *ctx);

/* ---- resident row store ---------------------------------------------------
 * Replaces: the vec0 virtual table's storage of one float32[dim] blob per
 * rowid (DDL idb:290-294, writer idb:1164-1181).  Rows are given in scan order
 * (ascending rowid); `rowids` may be NULL (rowid = rowid_base + position).
 *
 * clipdb_load_rows   copies n x dim float32 from a host OR device pointer into
 *                    context-owned HBM (replaces any previous store).
 * clipdb_append_rows appends (the scanner only ever INSERTs new rowids or
 *                    UPDATEs in place, idb:1165-1175).
 * clipdb_update_row  overwrites one stored row by position (UPDATE vec0 ...,
 *                    idb:1165-1167).
 * clipdb_attach_rows borrows caller-owned DEVICE memory without copying
 *                    (dim % 4 == 0, 16-byte aligned); the caller keeps it
 *                    alive until detach/destroy. */
int clipdb_load_rows(clipdb_ctx *ctx, const float *rows, const int64_t *rowids,
                     int64_t n, int32_t dim);
int clipdb_append_rows(clipdb_ctx *ctx, const float *rows, const int64_t *rowids, int64_t n);
int clipdb_update_row(clipdb_ctx *ctx, int64_t position, const float *row);
int clipdb_attach_rows(clipdb_ctx *ctx, const float *d_rows, const int64_t *d_rowids,
                       int64_t n, int32_t dim, int64_t rowid_base);

/* clipdb_reserve_rows creates an EMPTY owned store with room for `capacity` rows, to be filled with
 * clipdb_append_rows chunk by chunk (a loader that streams a large SQLite file never holds more than one
 * chunk on the host; appends beyond the capacity still grow it).  `explicit_rowids` != 0: every append
 * must pass rowids.  `placement` says where the float32 rows live:
 *   CLIPDB_PLACE_DEVICE  HBM (the default of clipdb_load_rows): 4 * dim bytes per row resident.
 *   CLIPDB_PLACE_HOST    pinned, device-mapped HOST memory.  Meant to be combined with
 *                        clipdb_enable_batch + option "batch_min_nq" = 1 (a "bf16-primary" store): the
 *                        bf16 copy (2 * dim bytes per row) is the only thing resident in HBM and every
 *                        search pre-selects on it with the tensor-core path, while the exact float32
 *                        re-rank reads its few hundred candidate rows from host memory over PCIe
 *                        (zero-copy).  Results are identical to the all-HBM store; a 50M-row shard
 *                        (230 GB of float32) fits one B200 this way (115 GB of bf16).  The exact scan
 *                        still works on such a store but streams the rows over PCIe.
 * clipdb_stage_buffer returns a context-owned pinned host buffer of at least `bytes` bytes (valid until
 * the next call asking for more, or destroy): rows written there reach HBM by one DMA per
 * clipdb_append_rows call instead of a pageable-memory copy. */
#define CLIPDB_PLACE_DEVICE 0
#define CLIPDB_PLACE_HOST   1
int clipdb_reserve_rows(clipdb_ctx *ctx, int64_t capacity, int32_t dim, int32_t explicit_rowids, int32_t placement);
int clipdb_stage_buffer(clipdb_ctx *ctx, int64_t bytes, void **out_host);
/* Native loader: append to the (reserved) store the rows the reference's search statement scans — `vec0` in rowid
 * order, INNER JOINed to `image_embeddings` (vec0.rowid = ie.rowid) and `images` (ie.image_id = i.id), idb:1564-1571
 * — for vec0 rowids in (min_rowid, max_rowid], read from the SQLite file with SQLite's own C library (libsqlite3.so.0,
 * resolved with dlopen).  The rowid range is cut into stripes of `chunk_rows` consecutive rowid values; up to
 * option "sqlite_readers" (default 4) connections, each inside its own read transaction, fill stripes ahead of one
 * committer that appends them in rowid order.  Every float32 blob is copied once, from SQLite's page buffer into a
 * pinned staging slot, and leaves for the GPU by DMA; vec0 rows without both partners are dropped, as the INNER
 * JOINs drop them (idb:1569-1570).  A sparse rowid space (mostly gaps) is read in one ordered pass instead.  After every chunk of up to `chunk_rows`
 * rows `on_chunk` (nullable) receives what the host keeps per row: the rowids, image ids, `images.last_modified`
 * and the file paths (`n` NUL-terminated UTF-8 strings back to back, `paths_bytes` in total); a non-zero return
 * stops the load.  Only a plain table named vec0 is read here (CLIPDB_ERR_UNSUPPORTED otherwise, and when
 * libsqlite3 is not available): sqlite-vec's virtual table is read by the binding's own reader.
 * out_vec0_rows = COUNT(*) of vec0 in the range before the joins (the reference's guard, idb:1532-1540). */
typedef int (*clipdb_sqlite_chunk_fn)(void *user, int64_t n, const int64_t *rowids, const int64_t *image_ids,
                                      const double *last_modified, const char *paths, int64_t paths_bytes);
int clipdb_append_sqlite(clipdb_ctx *ctx, const char *db_path, int64_t min_rowid, int64_t max_rowid,
                         int64_t chunk_rows, clipdb_sqlite_chunk_fn on_chunk, void *user,
                         int64_t *out_vec0_rows, int64_t *out_joined_rows);
int64_t clipdb_num_rows(const clipdb_ctx *ctx);
int32_t clipdb_dim(const clipdb_ctx *ctx);

/* Per-row admission bitset: bit (r & 31) of word (r >> 5) set = row r takes
 * part in the search.  Restates the WHERE (file_path LIKE ...) pre-filter that
 * is applied before the top-k (idb:1509-1530, 1571).  `words` is a host or
 * device pointer to ceil(n/32) uint32; it is copied.  clipdb_clear_mask
 * returns to "all rows". */
int clipdb_set_mask(clipdb_ctx *ctx, const uint32_t *words, int64_t n_words);
int clipdb_clear_mask(clipdb_ctx *ctx);

/* ---- query arithmetic (K3) ------------------------------------------------
 * Replaces the numpy float32 arithmetic of idb:1378-1398 (weighted blend +
 * L2 normalise), idb:545-571 (one negative) and idb:573-604 (several
 * negatives, subtracted in list order), including their zero-norm fallbacks.
 *   e2 NULL           -> single positive query (no blend, idb:1397-1398)
 *   w0, w1            -> the caller's raw weights; normalised as idb:1379-1383
 *   negs              -> n_neg x dim float32, neg_w -> n_neg weights
 * All pointers are HOST pointers; `out` receives dim floats; `out_flags`
 * (nullable) receives CLIPDB_BLEND_* bits. */
int clipdb_blend(clipdb_ctx *ctx, const float *e1, const float *e2, double w0, double w1,
                 const float *negs, const double *neg_w, int32_t n_neg, int32_t dim,
                 float *out, int32_t *out_flags);

/* Batched device form: `batch` independent queries laid out contiguously
 * (e1/e2: batch x dim; negs: batch x n_neg x dim; w: batch x 2 NORMALISED
 * float32 weights; neg_w: batch x n_neg float32).  Async on the ctx stream. */
int clipdb_blend_device(clipdb_ctx *ctx, const float *d_e1, const float *d_e2, const float *d_w,
                        const float *d_negs, const float *d_neg_w, int32_t n_neg, int32_t dim,
                        int32_t batch, float *d_out, int32_t *d_out_flags);

/* ---- search (K1 + K2) -----------------------------------------------------
 * Replaces the statement at idb:1564-1574 as executed at idb:1582-1583:
 *     SELECT ..., vec_distance_cosine(vec0.embedding, ?) AS distance FROM vec0 ...
 *     ORDER BY distance ASC LIMIT ?
 * For each of the nq queries (nq x dim float32, HOST pointer) writes up to k
 * results sorted by (distance ascending, scan position ascending) — SQLite's
 * sorter order, exact ties in rowid order — into out_rowids / out_dist (each
 * nq x k, row-major; entries past out_n[q] are untouched).
 *   out_n[q]   = number of results = min(k, rows admitted with a non-NaN distance)
 *   out_nan[q] = admitted rows whose distance was NaN (zero-norm row or query);
 *                SQLite would turn those into NULLs that sort FIRST and make
 *                the reference's `1.0 - distance` raise (idb:1588, 1637-1640);
 *                they are excluded here and counted so the caller can apply
 *                either policy.  Nullable.
 * use_mask != 0 applies the bitset installed by clipdb_set_mask.
 * k <= 0 -> out_n = 0 (SQLite: LIMIT 0).  k may exceed the row count. */
int clipdb_search(clipdb_ctx *ctx, const float *queries, int32_t nq, int32_t k,
                  int32_t metric, int32_t use_mask,
                  int64_t *out_rowids, float *out_dist, int32_t *out_n, int64_t *out_nan);

/* Same, device pointers in and out, enqueued on the ctx stream, no sync. */
int clipdb_search_device(clipdb_ctx *ctx, const float *d_queries, int32_t nq, int32_t k,
                         int32_t metric, int32_t use_mask,
                         int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n,
                         int64_t *d_out_nan);

/* blend + search in one host call (config 4: text+image 0.7/0.3 blend with a
 * negative prompt, k=20); the blended query never leaves the GPU.  `out_query`
 * (nullable) receives the blended vector. */
int clipdb_blend_search(clipdb_ctx *ctx, const float *e1, const float *e2, double w0, double w1,
                        const float *negs, const double *neg_w, int32_t n_neg,
                        int32_t k, int32_t metric, int32_t use_mask,
                        int64_t *out_rowids, float *out_dist, int32_t *out_n, int64_t *out_nan,
                        float *out_query, int32_t *out_flags);

/* ---- batched queries (K4, BASELINE configs[2]) ------------------------------
 * No reference equivalent (the reference answers one query at a time,
 * idb:2070-2299).  clipdb_enable_batch(1) builds a bf16 copy of the resident
 * store (2 bytes per element next to the 4-byte one) so that batches of queries
 * can be pre-selected by one tcgen05 tensor-core contraction and re-ranked with
 * the exact single-query arithmetic; results are identical to clipdb_search.
 * Requirements: dim == 1152, cosine, 1 <= k <= 128 (and <= rows); with a mask
 * (clipdb_set_mask) additionally a store without zero-norm rows.  Rows appended or updated
 * after clipdb_enable_batch are picked up: the bf16 copy is rebuilt before the next batched search.
 * With the batch store enabled, clipdb_search uses this path by itself for
 * nq >= 2 (option "batch_min_nq") and transparently re-runs flagged queries
 * through the exact scan.  The contraction runs 64, 128 or 256 queries per pass
 * (the smallest that holds the batch); below 256 it streams the bf16 copy at HBM
 * rate, so with option "batch_min_nq" = 1 even a single query costs half the
 * float32 scan's bytes, with identical results.
 *
 * clipdb_search_batch_device: device pointers, async, nq <= 256 per call.
 * d_flags[q] != 0 (CLIPDB_BATCH_*) marks queries whose result is NOT valid
 * (candidate overflow, zero-norm query): the caller must re-run those with
 * clipdb_search_device. */
#define CLIPDB_BATCH_OVERFLOW   1
#define CLIPDB_BATCH_BAD_QUERY  2
int clipdb_enable_batch(clipdb_ctx *ctx, int32_t enable);
int clipdb_search_batch_device(clipdb_ctx *ctx, const float *d_queries, int32_t nq, int32_t k, int32_t use_mask,
                               int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n,
                               int64_t *d_out_nan, int32_t *d_flags);

/* Measurement: per query slot (256 entries each, host pointers, nullable) of the LAST batched
 * pass — rows the tensor-core filter kept, and rows that also passed the second threshold and
 * were re-ranked in float32.  Synchronises the stream. */
int clipdb_batch_stats(clipdb_ctx *ctx, uint32_t *cand_counts, uint32_t *surv_counts);

/* ---- sign-code search (the reference's binary fallback) ----------------------------
 * Replaces idb:1591-1629, the path search() takes when `vec0` is empty but
 * `binary_embeddings` is not: every stored code is dim bytes of 0/1
 * (`(embedding >= 0).astype(np.uint8)`, idb:1189-1190), the query is quantised the same
 * way (idb:1593), score(row) = np.dot(query_code, row_code) (idb:1621) and the k rows
 * with the largest score come back, ties in scan order (Python's stable sort, idb:1627-1628).
 *   CLIPDB_SCORE_REFERENCE_UINT8  numpy evaluates that dot product in uint8, so the
 *                                 reference's score is popcount(q AND row) MODULO 256;
 *                                 this mode reproduces it.
 *   CLIPDB_SCORE_POPCOUNT         the plain AND-popcount (what the reference meant).
 * clipdb_load_codes copies n x dim bytes (host or device pointer; dim must be 1152; any
 * byte other than 0/1 is refused) and bit-packs them on the device: 144 B per row
 * resident.  `ids` (nullable) are returned instead of scan positions.  The code store is
 * independent of the float32 row store and has its own admission bitset
 * (clipdb_set_code_mask: the WHERE clause applies to this statement too, idb:1597-1611).
 * `order` (nullable, n uint32, a permutation of 0..n-1): the tie-break sequence of each row
 * while this mask is in use.  SQLite 3.45 answers the FILTERED statement by walking the
 * UNIQUE index on images.file_path (rows arrive in file_path order) and the unfiltered one by
 * scanning binary_embeddings (rowid order) [PROBED, EXPLAIN QUERY PLAN]; Python's stable sort
 * keeps that arrival order among equal scores, so the caller passes each row's file_path
 * rank here to get the reference's tie order.
 * clipdb_binary_search: host query code (dim bytes of 0/1) in, out_ids / out_scores
 * (k entries each, score descending, scan position ascending) and out_n out; synchronous.
 * similarity = score / dim is left to the caller (idb:1624). */
#define CLIPDB_SCORE_REFERENCE_UINT8  0
#define CLIPDB_SCORE_POPCOUNT         1
int clipdb_load_codes(clipdb_ctx *ctx, const uint8_t *codes, const int64_t *ids, int64_t n, int32_t dim);
int64_t clipdb_num_codes(const clipdb_ctx *ctx);
int clipdb_set_code_mask(clipdb_ctx *ctx, const uint32_t *words, int64_t n_words, const uint32_t *order);
int clipdb_clear_code_mask(clipdb_ctx *ctx);
int clipdb_binary_search(clipdb_ctx *ctx, const uint8_t *query_code, int32_t k, int32_t score_mode,
                         int32_t use_mask, int64_t *out_ids, int32_t *out_scores, int32_t *out_n);
/* device form: the query already bit-packed (36 words, bit (i & 31) of word (i >> 5) =
 * element i); enqueued on the ctx stream, no sync. */
int clipdb_binary_search_device(clipdb_ctx *ctx, const uint32_t *d_query_words, int32_t k, int32_t score_mode,
                                int32_t use_mask, int64_t *d_out_ids, int32_t *d_out_scores, int32_t *d_out_n);

/* ---- shard merge (multi-GPU) --------------------------------
 * Merges `lists` per-shard result lists (each k entries, already sorted, shard
 * order = rowid order; counts[l] valid entries in list l) into the global
 * top-k with the comparator (distance, shard, position) == (distance, rowid).
 * Device pointers (e.g. the output of an NCCL all-gather); async. */
int clipdb_merge_device(clipdb_ctx *ctx, const float *d_dist, const int64_t *d_rowids,
                        const int32_t *d_counts, int32_t lists, int32_t k,
                        float *d_out_dist, int64_t *d_out_rowids, int32_t *d_out_n);

/* Same with explicit byte strides between consecutive lists, so the lists can
 * sit inside packed per-rank records (one NCCL all-gather instead of three). */
int clipdb_merge_strided_device(clipdb_ctx *ctx, const void *d_dist, int64_t dist_stride,
                                const void *d_rowids, int64_t rowid_stride,
                                const void *d_counts, int64_t count_stride,
                                int32_t lists, int32_t k,
                                float *d_out_dist, int64_t *d_out_rowids, int32_t *d_out_n);

/* Batched form: nq queries per list; consecutive queries of one list are
 * *_query_stride bytes apart.  Outputs are nq x k (row-major) and nq counts. */
int clipdb_merge_batch_device(clipdb_ctx *ctx, const void *d_dist, int64_t dist_stride, int64_t dist_query_stride,
                              const void *d_rowids, int64_t rowid_stride, int64_t rowid_query_stride,
                              const void *d_counts, int64_t count_stride, int64_t count_query_stride,
                              int32_t lists, int32_t nq, int32_t k,
                              float *d_out_dist, int64_t *d_out_rowids, int32_t *d_out_n);

/* ---- fused shard exchange (multi-GPU, one launch per query) ---------------------------------
 * No reference equivalent.  With the exchange connected, clipdb_search_sharded_device answers a
 * query over ALL shards in one kernel launch per GPU: the scan kernel's last CTA stores its
 * shard's k best (distance, rowid) pairs directly into every peer GPU's inbox over NVLink
 * (peer memory mapped with CUDA IPC), signals with a system-scope release, waits for the other
 * shards' records with acquire loads and merges — no NCCL call and no host round trip between
 * the scan and the merged result.  Every rank ends up with the same answer, identical to the
 * unsharded one (shards are contiguous rowid ranges, rank order = rowid order).
 *
 *   clipdb_exchange_init      allocates this rank's inbox; out_ipc_handle (nullable) receives
 *                             CLIPDB_IPC_HANDLE_BYTES bytes to hand to the other ranks (e.g. with
 *                             torch.distributed.all_gather_object), out_inbox (nullable) the raw
 *                             device pointer (for ranks living in the same process).
 *   clipdb_exchange_connect   world x CLIPDB_IPC_HANDLE_BYTES handle bytes in rank order
 *                             (one process per GPU; the own slot is ignored).
 *   clipdb_exchange_connect_pointers  same-process variant: inbox pointers (and their CUDA
 *                             devices, nullable = same device) in rank order.
 *   clipdb_search_sharded_device      device pointers, async; EVERY rank must issue the same
 *                             sequence of calls.  Needs dim 1152, 1 <= k <= 128, a non-empty shard.
 *                             If a peer does not deliver within option "xchg_timeout_ms"
 *                             (default 10 s) *d_out_n is set to -1 instead of hanging the GPU; -2 if
 *                             the ranks passed different k. */
#define CLIPDB_IPC_HANDLE_BYTES 64
int clipdb_exchange_init(clipdb_ctx *ctx, int32_t world, int32_t rank, void *out_ipc_handle, void **out_inbox);
int clipdb_exchange_connect(clipdb_ctx *ctx, const void *ipc_handles);
int clipdb_exchange_connect_pointers(clipdb_ctx *ctx, void *const *inboxes, const int32_t *devices);
int clipdb_search_sharded_device(clipdb_ctx *ctx, const float *d_query, int32_t k, int32_t metric, int32_t use_mask,
                                 int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan);
/* Batched form (clipdb_enable_batch on every rank, nq <= 256 per call, every rank the same nq and k):
 * the last kernel of the batched path exchanges each query's k candidates with the peer GPUs the
 * same way, so the outputs hold the answers over ALL shards.  d_flags[q] != 0: SOME shard could
 * not answer query q through the batched path; re-run it with clipdb_search_sharded_device (on
 * every rank: all ranks see the same flags).  d_out_n[q] = -1 / -2 as above. */
int clipdb_search_batch_sharded_device(clipdb_ctx *ctx, const float *d_queries, int32_t nq, int32_t k, int32_t use_mask,
                                       int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n,
                                       int64_t *d_out_nan, int32_t *d_flags);

/* Host-side control of the in-kernel exchange.
 *   clipdb_exchange_abort      abort != 0: every exchange of this context that is (or will be) waiting for
 *                              its peers gives up at once with *d_out_n = -1 instead of spinning until
 *                              "xchg_timeout_ms" (a watchdog thread can call this while another thread
 *                              waits on the stream: it takes no lock); abort == 0 re-arms.
 *   clipdb_exchange_set_epoch  the sequence numbers the NEXT single-query / batched exchange will use are
 *                              single_epoch + 1 / batch_epoch + 1.  All ranks must be idle and pass the same
 *                              values, larger than any epoch used since clipdb_exchange_init: this is how
 *                              ranks that fell out of step (one of them failed to enqueue a search) are
 *                              brought back in step without re-creating the inboxes.
 *   clipdb_exchange_stats      per-launch timeline of clipdb_search_sharded_device, summed on the device
 *                              from %globaltimer stamps taken by the scan kernel's last CTA since the last
 *                              reset: out_ns[0] first CTA started -> last CTA finished its rows (scan),
 *                              [1] -> per-SM lists merged, [2] -> record stored into every peer's inbox and
 *                              published, [3] -> every peer's record arrived (exchange wait: includes the
 *                              time the other GPUs were still scanning), [4] -> merged and decoded.
 *                              `enable` != 0 turns the stamping on (off by default), `reset` != 0 zeroes the
 *                              sums after reading.  Synchronises the stream. */
int clipdb_exchange_abort(clipdb_ctx *ctx, int32_t abort);
int clipdb_exchange_set_epoch(clipdb_ctx *ctx, uint32_t single_epoch, uint32_t batch_epoch);
int clipdb_exchange_stats(clipdb_ctx *ctx, int32_t enable, int32_t reset, double *out_ns, int64_t *out_launches);

/* ---- measurement ------------------------------------------------------------
 * With profiling enabled every scan kernel (the dominant, HBM-bound launch) is
 * bracketed by CUDA events on the launching stream; clipdb_profile_read syncs
 * the stream and returns the summed scan-kernel time and the number of scans
 * since profiling was (re)enabled.  bench.py derives roofline.achieved from it. */
int clipdb_profile(clipdb_ctx *ctx, int32_t enable);
int clipdb_profile_read(clipdb_ctx *ctx, double *scan_ms_total, int64_t *scans);
/* The SM clock the batched path's contraction actually ran at: while profiling is enabled, CTA 0 of every
 * filter-pass launch (batch_gemm_pair_kernel<FILTER>) reads %clock64 and %globaltimer at entry and exit.  Returns
 * the sums since the last call (and clears them): sm_cycles / nanoseconds = GHz.  The tensor pipe's peak scales
 * with that clock, and under the board's power cap it sits far below the 1965 MHz the datasheet peak assumes. */
int clipdb_profile_clock(clipdb_ctx *ctx, double *sm_cycles, double *nanoseconds, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* CLIPDB_H */
