// clipdb.cu — the C ABI declared in include/clipdb.h (host orchestration).
//
// One context = one GPU's resident float32 row store + workspaces + a stream.
// The search entry points replace the statement the reference executes at
// image_database.py:1564-1583; clipdb_blend replaces the numpy arithmetic at
// image_database.py:1378-1398 / 545-604.  No CPU fallback anywhere: a missing
// device or a failed launch is an error code.
#include <cuda_runtime.h>

#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/clipdb.h"
#include "batch.cuh"
#include "binary.cuh"
#include "blend.cuh"
#include "merge.cuh"
#include "scan_topk.cuh"
#include "sqlite_reader.cuh"

using namespace clipdb;

namespace {

constexpr int ABI_VERSION = 5;
#ifndef CLIPDB_SOURCE_HASH
#define CLIPDB_SOURCE_HASH "unknown"
#endif
constexpr int FUSED_K_MAX = 128;        // largest k served by the register-resident lists
constexpr int MERGE_SHARD_MAX_KEYS = 16384;

struct Buffer {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct clipdb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    int64_t launches = 0;

    // resident store
    float *rows = nullptr;           // HBM tier: rows [0, min(n, split))
    float *rows_host = nullptr;      // tiered store (clipdb_reserve_rows): pinned host memory, rows [split, cap)
    float *rows_host_dev = nullptr;  // the same memory as the device addresses it
    int64_t split = LLONG_MAX;       // first row position that lives in rows_host (LLONG_MAX: all rows in HBM)
    int64_t *rowids = nullptr;
    bool owns_rows = false;
    int64_t n = 0, cap = 0;
    int32_t dim = 0, ld = 0;
    int64_t rowid_base = 0;
    uint32_t *mask = nullptr;
    int64_t mask_words = 0;

    // sign-code store (binary fallback search, image_database.py:1591-1629)
    uint32_t *codes = nullptr;       // [n_codes][36] bit-packed
    int64_t *code_ids = nullptr;     // nullable: id returned for a row (else its scan position)
    int64_t n_codes = 0;
    uint32_t *code_mask = nullptr;
    int64_t code_mask_words = 0;
    uint32_t *code_seq = nullptr;        // tie-break sequence installed with the mask (nullable)
    int64_t *code_ids_by_seq = nullptr;  // ids indexed by that sequence
    Buffer d_code_query, code_stage, code_bad;

    // fused shard exchange (merge.cuh): own inbox + every rank's inbox as mapped in this process
    ExchangeSlot *xchg_inbox = nullptr;
    ExchangeSlot *xchg_peer[XCHG_MAX_WORLD] = {};
    bool xchg_ipc[XCHG_MAX_WORLD] = {};       // peer pointer came from cudaIpcOpenMemHandle
    int xchg_world = 0, xchg_rank = 0;
    bool xchg_connected = false;
    uint32_t xchg_epoch = 0, xchg_batch_epoch = 0;   // single-query and batched passes count separately
    int64_t xchg_timeout_ms = 10000;
    int *xchg_abort_host = nullptr;   // mapped host flag the waiting kernels poll (clipdb_exchange_abort)
    int *xchg_abort_dev = nullptr;
    Buffer xchg_stats;                // XCHG_STATS_WORDS uint64: per-launch timeline sums (clipdb_exchange_stats)
    bool xchg_stats_on = false;

    // workspaces (grown on demand)
    Buffer cand_a, cand_b, sync_buf, all_keys_a, all_keys_b, cub_tmp;
    bool sync_dirty = true;   // the scan kernels' counters may be non-zero
    Buffer d_query, d_results;   // host entry points: staged queries, one packed result block
    int64_t *r_nan = nullptr, *r_ids = nullptr;
    float *r_dist = nullptr;
    int32_t *r_n = nullptr;
    size_t r_bytes = 0;
    Buffer d_blend_in, d_blend_flags;
    Buffer pinned;      // host staging (inputs, then results)
    Buffer stage;       // pinned host buffer lent to the caller (clipdb_stage_buffer)
    std::vector<void *> sql_slots;     // pinned staging slots of the native SQLite loader (clipdb_append_sqlite)
    size_t sql_slot_bytes = 0;
    Buffer pinned_aux;  // host staging for the blended query read-back
    Buffer pinned_flags; // host staging for the batched path's per-query flags

    // batched path (K4): bf16 copy of the store + workspaces
    bool batch_enabled = false;
    bool batch_dirty = false;            // the bf16 copy must be rebuilt from scratch before its next use
    int64_t bf16_built = 0;              // rows [0, bf16_built) of the bf16 copy are current
    int64_t bf16_cap_rows = 0;           // rows the bf16 buffer has room for (a multiple of 128)
    unsigned long long batch_bad_rows = 0;  // zero-norm / non-finite rows found when the copy was built
    Buffer bf16_rows, bad_rows, bq_queries, bq_qnorm, bq_scores, bq_thr, bq_flags, bq_count, bq_cand, bq_parts, bq_qerr, row_err;
    Buffer bq_cand_u, bq_margin, bq_surv, bq_surv_count, bq_tilectr;
    Buffer bq_clock;                     // 3 x uint64: SM cycles, ns, launches of the profiled contraction launches
    CUtensorMap map_rows, map_q, map_qhalf[3];   // query-half boxes for 64 / 128 / 256 queries per pass
    int64_t bq_sample_groups_used = 0;  // groups the last pass A wrote
    int64_t batch_min_nq = 2;       // clipdb_search switches to the batched path from this nq (one batch
                                    // costs about one single-query scan, whatever its size)
    int64_t batch_cand_cap = 65536; // candidate rows kept per query (8x the ~8,000 the default sampling admits)
    int64_t batch_cta_pair = 1;     // 1: cta_group::2 contraction (CTA pairs), 0: single-CTA kernel
    int64_t batch_sample_stride = 0; // pass A samples 1/s of the rows; 0 = auto (see batch_sample_stride())
    int64_t batch_refine = 1;       // 1: second threshold from the candidates' own scores before the re-rank
    int64_t batch_static_tiles = 0; // 1: static tile interleave in the CTA-pair kernel (for A/B measurements)
    int64_t batch_npass = 0;        // 0: auto (64 / 128 / 256 queries per pass by batch size); else force >= that

    // scan-kernel event timing (clipdb_profile)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
    size_t ev_used = 0;                 // events handed out since the last read
    double prof_ms = 0.0;
    int64_t prof_scans = 0;

    // options
    int64_t scan_variant = 0;  // 0 auto (TMA ring when dim == 1152), 1 TMA ring, 2 direct loads
    int64_t scan_ctas = 0;     // 0 = one per SM (TMA) / ldg_ctas_per_sm per SM (direct)
    int64_t ldg_ctas_per_sm = 4;
    int64_t evict_first = 1;
    int64_t scan_cfg = 0;      // ring shape (ScanCfg0..4); shapes other than 0 exist for k <= 32 cosine only
    int64_t scan_assign = 2;   // SCAN_ASSIGN_* (dynamic: +5 % over static interleaving, profiles/r01_sweep2.json)
    int64_t scan_chunk = 4;    // tiles per atomicAdd (dynamic assignment)
    int64_t fuse_tail = 1;     // 1: the scan's last CTA merges and decodes (one launch per query); 0: merge-tree kernels
    int64_t host_tier_from_row = 0;  // clipdb_reserve_rows(CLIPDB_PLACE_HOST): rows below this position stay in HBM
    int64_t sqlite_readers = 4;      // clipdb_append_sqlite: connections reading rowid stripes in parallel
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fail(clipdb_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CU_TRY(c, expr)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail((c), e__ == cudaErrorMemoryAllocation ? CLIPDB_ERR_NOMEM            \
                                                              : CLIPDB_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,  \
                        __LINE__);                                                          \
    } while (0)

#define RC_TRY(expr)                   \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != CLIPDB_OK) return rc__; \
    } while (0)

int ensure_device(clipdb_ctx *c, Buffer &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return CLIPDB_OK;
    if (b.p) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    CU_TRY(c, cudaMalloc(&b.p, want));
    b.bytes = want;
    return CLIPDB_OK;
}

int ensure_pinned(clipdb_ctx *c, size_t bytes) {
    Buffer &b = c->pinned;
    if (b.bytes >= bytes && b.p) return CLIPDB_OK;
    if (b.p) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, cudaFreeHost(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes < 65536 ? 65536 : bytes;
    CU_TRY(c, cudaMallocHost(&b.p, want));
    b.bytes = want;
    return CLIPDB_OK;
}

void free_buffer(Buffer &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

bool is_device_pointer(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

void release_store(clipdb_ctx *c) {
    if (c->owns_rows) {
        if (c->rows) cudaFree(c->rows);
        if (c->rowids) cudaFree(c->rowids);
    }
    if (c->rows_host) cudaFreeHost(c->rows_host);
    c->rows = nullptr;
    c->rows_host = c->rows_host_dev = nullptr;
    c->split = LLONG_MAX;
    c->rowids = nullptr;
    c->owns_rows = false;
    c->n = c->cap = 0;
    c->dim = c->ld = 0;
    c->rowid_base = 0;
    if (c->mask) cudaFree(c->mask);
    c->mask = nullptr;
    c->mask_words = 0;
    c->batch_enabled = false;   // the bf16 copy described the old rows
    c->bf16_built = c->bf16_cap_rows = 0;
    free_buffer(c->bf16_rows);
}

// where row `pos` of the store lives, as the HOST addresses it (cudaMemcpy destinations)
float *store_row_host_view(clipdb_ctx *c, int64_t pos) {
    return pos < c->split ? c->rows + pos * c->ld : c->rows_host + (pos - c->split) * c->ld;
}
// biased base of the host tier for the kernels: rows_hi + pos * ld addresses row pos >= split
const float *store_rows_hi(const clipdb_ctx *c) {
    return c->rows_host_dev ? c->rows_host_dev - c->split * c->ld : nullptr;
}

void release_exchange(clipdb_ctx *c) {
    for (int r = 0; r < XCHG_MAX_WORLD; r++) {
        if (c->xchg_ipc[r] && c->xchg_peer[r]) cudaIpcCloseMemHandle(c->xchg_peer[r]);
        c->xchg_peer[r] = nullptr;
        c->xchg_ipc[r] = false;
    }
    if (c->xchg_inbox) cudaFree(c->xchg_inbox);
    c->xchg_inbox = nullptr;
    c->xchg_world = 0;
    c->xchg_connected = false;
    if (c->xchg_abort_host) cudaFreeHost(c->xchg_abort_host);
    c->xchg_abort_host = c->xchg_abort_dev = nullptr;
    cudaGetLastError();
}

void release_codes(clipdb_ctx *c) {
    if (c->codes) cudaFree(c->codes);
    if (c->code_ids) cudaFree(c->code_ids);
    if (c->code_mask) cudaFree(c->code_mask);
    if (c->code_seq) cudaFree(c->code_seq);
    if (c->code_ids_by_seq) cudaFree(c->code_ids_by_seq);
    c->codes = nullptr;
    c->code_ids = nullptr;
    c->code_mask = nullptr;
    c->code_seq = nullptr;
    c->code_ids_by_seq = nullptr;
    c->n_codes = 0;
    c->code_mask_words = 0;
}

// copy `m` rows of `dim` floats (host or device) into the store at row `at`; a range that crosses the
// tier boundary of a tiered store is copied in two pieces
int copy_rows_in(clipdb_ctx *c, const float *src, const int64_t *src_ids, int64_t at, int64_t m) {
    if (m == 0) return CLIPDB_OK;
    for (int64_t lo = at; lo < at + m;) {
        const int64_t hi = (lo < c->split && at + m > c->split) ? c->split : at + m;
        float *dst = store_row_host_view(c, lo);
        const float *from = src + (lo - at) * c->dim;
        const size_t rows_here = static_cast<size_t>(hi - lo);
        if (c->ld == c->dim) {
            CU_TRY(c, cudaMemcpyAsync(dst, from, rows_here * c->dim * sizeof(float), cudaMemcpyDefault, c->stream));
        } else {
            if (lo < c->split) CU_TRY(c, cudaMemsetAsync(dst, 0, rows_here * c->ld * sizeof(float), c->stream));
            else memset(dst, 0, rows_here * c->ld * sizeof(float));
            CU_TRY(c, cudaMemcpy2DAsync(dst, c->ld * sizeof(float), from, c->dim * sizeof(float),
                                        c->dim * sizeof(float), rows_here, cudaMemcpyDefault, c->stream));
        }
        lo = hi;
    }
    if (c->rowids) {
        if (!src_ids) return fail(c, CLIPDB_ERR_INVALID, "store has explicit rowids; rowids required");
        CU_TRY(c, cudaMemcpyAsync(c->rowids + at, src_ids, static_cast<size_t>(m) * sizeof(int64_t),
                                  cudaMemcpyDefault, c->stream));
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));  // caller's buffer may be freed on return
    return CLIPDB_OK;
}

// device_rows < 0 (or >= cap): every row in HBM; otherwise rows [0, device_rows rounded down to 128) in HBM
// and the rest in pinned, device-mapped host memory (the tiered store of clipdb_reserve_rows)
int alloc_store(clipdb_ctx *c, int64_t cap, int32_t dim, bool with_ids, int64_t device_rows = -1) {
    release_store(c);
    c->dim = dim;
    c->ld = (dim + 3) & ~3;
    c->cap = cap > 0 ? cap : 1;
    c->owns_rows = true;
    int64_t dev_cap = c->cap;
    if (device_rows >= 0 && device_rows < c->cap) dev_cap = device_rows & ~static_cast<int64_t>(127);
    CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->rows),
                         static_cast<size_t>(dev_cap > 0 ? dev_cap : 1) * c->ld * sizeof(float)));
    if (dev_cap < c->cap) {
        CU_TRY(c, cudaHostAlloc(reinterpret_cast<void **>(&c->rows_host),
                                static_cast<size_t>(c->cap - dev_cap) * c->ld * sizeof(float),
                                cudaHostAllocMapped | cudaHostAllocPortable));
        CU_TRY(c, cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->rows_host_dev), c->rows_host, 0));
        c->split = dev_cap;
    }
    if (with_ids)
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->rowids),
                             static_cast<size_t>(c->cap) * sizeof(int64_t)));
    return CLIPDB_OK;
}

// ---- scan-kernel event timing -----------------------------------------------------

constexpr size_t PROFILE_MAX_PENDING = 4096;  // event pairs kept before folding into the totals

int profile_fold(clipdb_ctx *c) {
    if (c->ev_used == 0) return CLIPDB_OK;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0.f;
        CU_TRY(c, cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
        c->prof_ms += ms;
        c->prof_scans++;
    }
    c->ev_used = 0;
    return CLIPDB_OK;
}

int profile_mark(clipdb_ctx *c, bool start) {
    if (!c->profiling) return CLIPDB_OK;
    if (start && c->ev_used + 2 > 2 * PROFILE_MAX_PENDING) RC_TRY(profile_fold(c));
    if (c->ev_used >= c->ev_pool.size()) {
        cudaEvent_t e;
        CU_TRY(c, cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    CU_TRY(c, cudaEventRecord(c->ev_pool[c->ev_used++], c->stream));
    return CLIPDB_OK;
}

// ---- kernel dispatch ------------------------------------------------------------

template <typename CFG, int KPL, int METRIC, bool ALL>
int launch_scan_tma_cfg(clipdb_ctx *c, const ScanArgs &a, int grid) {
    auto kern = scan_tma_kernel<CFG, KPL, METRIC, ALL>;
    static thread_local int configured_device = -1;  // attribute is per device & per instantiation
    if (configured_device != c->device) {
        CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM_BYTES));
        configured_device = c->device;
    }
    kern<<<grid, CFG::THREADS, CFG::SMEM_BYTES, c->stream>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

template <int KPL, int METRIC, bool ALL>
int launch_scan_tma(clipdb_ctx *c, const ScanArgs &a, int grid) {
    if (KPL == 1 && METRIC == METRIC_COSINE && !ALL) {  // experiment shapes
        switch (c->scan_cfg) {
            case 1: return launch_scan_tma_cfg<ScanCfg1, 1, METRIC_COSINE, false>(c, a, grid);
            case 2: return launch_scan_tma_cfg<ScanCfg2, 1, METRIC_COSINE, false>(c, a, grid);
            case 3: return launch_scan_tma_cfg<ScanCfg3, 1, METRIC_COSINE, false>(c, a, grid);
            case 4: return launch_scan_tma_cfg<ScanCfg4, 1, METRIC_COSINE, false>(c, a, grid);
            default: break;
        }
    }
    return launch_scan_tma_cfg<ScanCfg0, KPL, METRIC, ALL>(c, a, grid);
}

template <int DIM_T, int KPL, int METRIC, bool ALL>
int launch_scan_ldg(clipdb_ctx *c, const ScanArgs &a, int grid) {
    auto kern = scan_ldg_kernel<DIM_T, KPL, METRIC, ALL>;
    const size_t smem = static_cast<size_t>(next_pow2(LDG_WARPS * 32 * KPL)) * sizeof(uint64_t) +
                        static_cast<size_t>(a.ld) * sizeof(float);
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        static thread_local size_t configured_bytes = 0;
        if (configured_device != c->device || configured_bytes < smem) {
            CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
            configured_device = c->device;
            configured_bytes = smem;
        }
    }
    kern<<<grid, LDG_THREADS, smem, c->stream>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

template <int KPL, bool ALL>
int launch_scan_metric_raw(clipdb_ctx *c, const ScanArgs &a, int metric, bool tma, int grid);

template <int KPL, bool ALL>
int launch_scan_metric(clipdb_ctx *c, const ScanArgs &a, int metric, bool tma, int grid) {
    RC_TRY(profile_mark(c, true));
    RC_TRY((launch_scan_metric_raw<KPL, ALL>(c, a, metric, tma, grid)));
    return profile_mark(c, false);
}

template <int KPL, bool ALL>
int launch_scan_metric_raw(clipdb_ctx *c, const ScanArgs &a, int metric, bool tma, int grid) {
    if (tma) {
        return metric == CLIPDB_METRIC_COSINE ? launch_scan_tma<KPL, METRIC_COSINE, ALL>(c, a, grid)
                                              : launch_scan_tma<KPL, METRIC_L2, ALL>(c, a, grid);
    }
    if (a.dim == SCAN_DIM) {
        return metric == CLIPDB_METRIC_COSINE
                   ? launch_scan_ldg<SCAN_DIM, KPL, METRIC_COSINE, ALL>(c, a, grid)
                   : launch_scan_ldg<SCAN_DIM, KPL, METRIC_L2, ALL>(c, a, grid);
    }
    return metric == CLIPDB_METRIC_COSINE ? launch_scan_ldg<0, KPL, METRIC_COSINE, ALL>(c, a, grid)
                                          : launch_scan_ldg<0, KPL, METRIC_L2, ALL>(c, a, grid);
}

bool use_tma_variant(const clipdb_ctx *c) {
    if (c->scan_variant == 2) return false;
    return c->dim == SCAN_DIM && c->ld == SCAN_DIM;
}

int scan_grid(const clipdb_ctx *c, bool tma) {
    if (c->scan_ctas > 0) return static_cast<int>(c->scan_ctas);
    return tma ? c->sm_count : c->sm_count * static_cast<int>(c->ldg_ctas_per_sm);
}

// merge tree over the per-CTA candidate lists in cand_a: lists -> ceil(lists / per_cta) -> ... -> 1 (decoded)
int merge_cta_lists(clipdb_ctx *c, int grid, int stride, const DecodeArgs &dec) {
    const int per_cta = MERGE_KEYS / stride;
    int lists = grid;
    const uint64_t *in = static_cast<const uint64_t *>(c->cand_a.p);
    RC_TRY(ensure_device(c, c->cand_b,
                         static_cast<size_t>((grid + per_cta - 1) / per_cta) * stride * sizeof(uint64_t) * 2));
    uint64_t *ping = static_cast<uint64_t *>(c->cand_b.p);
    uint64_t *pong = ping + static_cast<size_t>((grid + per_cta - 1) / per_cta) * stride;
    for (;;) {
        const int ctas = (lists + per_cta - 1) / per_cta;
        const int last = ctas == 1;
        reduce_lists_kernel<<<ctas, MERGE_THREADS, 0, c->stream>>>(in, lists, stride, per_cta, ping, dec, last);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        if (last) break;
        in = ping;
        lists = ctas;
        uint64_t *t = ping;
        ping = pong;
        pong = t;
    }
    return CLIPDB_OK;
}

// general k: all_keys_a holds one key per row; radix sort, decode the first kk
int sort_all_keys_and_decode(clipdb_ctx *c, int64_t n, int64_t kk, const DecodeArgs &dec) {
    uint64_t *keys = static_cast<uint64_t *>(c->all_keys_a.p);
    size_t tmp_bytes = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, static_cast<uint64_t *>(c->all_keys_b.p), n, 0,
                                             64, c->stream));
    RC_TRY(ensure_device(c, c->cub_tmp, tmp_bytes));
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tmp_bytes, keys, static_cast<uint64_t *>(c->all_keys_b.p), n,
                                             0, 64, c->stream));
    c->launches++;  // counted once; CUB's passes are library kernels, not ours
    CU_TRY(c, cudaMemsetAsync(dec.out_n, 0, sizeof(int32_t), c->stream));
    const int threads = 256;
    const int blocks = static_cast<int>((kk + threads - 1) / threads);
    decode_sorted_kernel<<<blocks, threads, 0, c->stream>>>(static_cast<const uint64_t *>(c->all_keys_b.p), n, dec);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

// The scan kernels' shared counters (merge.cuh ScanSync).  A fused launch leaves them zero (its
// last CTA resets them); any other use marks them dirty and the next user zeroes them first.
int prepare_sync(clipdb_ctx *c, ScanSync **out) {
    const bool fresh = c->sync_buf.p == nullptr;
    RC_TRY(ensure_device(c, c->sync_buf, sizeof(ScanSync)));
    if (fresh || c->sync_dirty) CU_TRY(c, cudaMemsetAsync(c->sync_buf.p, 0, sizeof(ScanSync), c->stream));
    c->sync_dirty = true;   // until a fused launch has been enqueued successfully
    *out = static_cast<ScanSync *>(c->sync_buf.p);
    return CLIPDB_OK;
}

// one query: ONE launch (TMA scan with the fused merge/decode tail, k <= FUSED_K_MAX), or
// scan + merge tree (direct-load kernel), or scan + radix sort (any k)
int search_one(clipdb_ctx *c, const float *d_query, int k, int metric, bool use_mask,
               int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan,
               const ExchangeArgs *xa = nullptr) {
    const int64_t kk = k < c->n ? k : c->n;
    if (xa && (kk <= 0 || k > FUSED_K_MAX))
        return fail(c, CLIPDB_ERR_UNSUPPORTED, "sharded search: needs 1 <= k <= %d and a non-empty shard", FUSED_K_MAX);
    if (kk <= 0) {
        CU_TRY(c, cudaMemsetAsync(d_out_n, 0, sizeof(int32_t), c->stream));
        if (d_out_nan) CU_TRY(c, cudaMemsetAsync(d_out_nan, 0, sizeof(int64_t), c->stream));
        return CLIPDB_OK;
    }
    const bool tma = use_tma_variant(c);
    const int grid = scan_grid(c, tma);

    ScanArgs a{};
    a.rows = c->rows;
    a.rows_hi = store_rows_hi(c);
    a.split = c->split;
    a.query = d_query;
    a.mask = use_mask ? c->mask : nullptr;
    RC_TRY(prepare_sync(c, &a.sync));
    a.n = c->n;
    a.dim = c->dim;
    a.ld = c->ld;
    a.k = static_cast<int>(kk);
    a.evict_first = static_cast<int>(c->evict_first);
    a.assign = static_cast<int>(c->scan_assign);
    a.chunk_tiles = static_cast<int>(c->scan_chunk);

    DecodeArgs dec{};
    dec.rowids = c->rowids;
    dec.rowid_base = c->rowid_base;
    dec.out_rowids = d_out_rowids;
    dec.out_dist = d_out_dist;
    dec.out_n = d_out_n;
    dec.out_nan = d_out_nan;
    dec.nan_rows = &a.sync->nan_rows;
    dec.k = static_cast<int>(kk);

    if (kk <= FUSED_K_MAX) {
        const int kpl = kk <= 32 ? 1 : (kk <= 64 ? 2 : 4);
        const int stride = 32 * kpl;
        a.cand_stride = stride;
        RC_TRY(ensure_device(c, c->cand_a, static_cast<size_t>(grid) * stride * sizeof(uint64_t)));
        a.cand = static_cast<uint64_t *>(c->cand_a.p);
        // every CTA's list must fit the last CTA's shared memory (the smallest ring among the shapes)
        const size_t ring_bytes = c->scan_cfg == 0 ? ScanCfg0::STAGES * ScanCfg0::STAGE_BYTES
                                                   : ScanCfg3::STAGES * ScanCfg3::STAGE_BYTES;
        a.fuse_tail = tma && c->fuse_tail && static_cast<size_t>(grid) * stride * sizeof(uint64_t) <= ring_bytes;
        a.dec = dec;
        if (xa) {
            if (!a.fuse_tail)
                return fail(c, CLIPDB_ERR_UNSUPPORTED, "sharded search: needs the TMA scan kernel with its fused tail "
                                                       "(dim 1152, option fuse_tail = 1)");
            a.xchg = *xa;
        }
        switch (kpl) {
            case 1: RC_TRY((launch_scan_metric<1, false>(c, a, metric, tma, grid))); break;
            case 2: RC_TRY((launch_scan_metric<2, false>(c, a, metric, tma, grid))); break;
            default: RC_TRY((launch_scan_metric<4, false>(c, a, metric, tma, grid))); break;
        }
        if (a.fuse_tail) {
            c->sync_dirty = false;
            return CLIPDB_OK;
        }
        return merge_cta_lists(c, grid, stride, dec);
    }

    // general k: every key to HBM (8 B/row next to the 4*dim B/row just read), radix sort, decode
    const size_t key_bytes = static_cast<size_t>(c->n) * sizeof(uint64_t);
    RC_TRY(ensure_device(c, c->all_keys_a, key_bytes));
    RC_TRY(ensure_device(c, c->all_keys_b, key_bytes));
    a.all_keys = static_cast<uint64_t *>(c->all_keys_a.p);
    a.cand_stride = 32;
    a.k = 0;
    RC_TRY((launch_scan_metric<1, true>(c, a, metric, tma, grid)));
    return sort_all_keys_and_decode(c, c->n, kk, dec);
}

int search_device_locked(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k,
                         int32_t metric, int32_t use_mask, int64_t *d_out_rowids,
                         float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan) {
    if (!c->rows || c->dim == 0) return fail(c, CLIPDB_ERR_STATE, "no rows loaded");
    if (nq < 0 || !d_queries || !d_out_n || (k > 0 && (!d_out_rowids || !d_out_dist)))
        return fail(c, CLIPDB_ERR_INVALID, "search: null pointer or negative nq");
    if (metric != CLIPDB_METRIC_COSINE && metric != CLIPDB_METRIC_L2)
        return fail(c, CLIPDB_ERR_INVALID, "search: unknown metric %d", metric);
    if (use_mask && !c->mask) return fail(c, CLIPDB_ERR_STATE, "search: use_mask set but no mask installed");
    if (c->n >= (1ll << 32)) return fail(c, CLIPDB_ERR_UNSUPPORTED, "more than 2^32-1 rows per context");
    const int64_t kcols = k > 0 ? k : 0;
    for (int32_t q = 0; q < nq; q++) {
        RC_TRY(search_one(c, d_queries + static_cast<size_t>(q) * c->dim, k, metric, use_mask != 0,
                          d_out_rowids + q * kcols, d_out_dist + q * kcols, d_out_n + q,
                          d_out_nan ? d_out_nan + q : nullptr));
    }
    return CLIPDB_OK;
}

int blend_device_locked(clipdb_ctx *c, const float *d_e1, const float *d_e2, const float *d_w,
                        const float *d_negs, const float *d_neg_w, int32_t n_neg, int32_t dim,
                        int32_t batch, float *d_out, int32_t *d_flags) {
    if (!d_e1 || !d_out || dim <= 0 || batch < 0 || n_neg < 0)
        return fail(c, CLIPDB_ERR_INVALID, "blend: bad argument");
    if (d_e2 && !d_w) return fail(c, CLIPDB_ERR_INVALID, "blend: weights required with a second query");
    if (n_neg > 0 && (!d_negs || !d_neg_w))
        return fail(c, CLIPDB_ERR_INVALID, "blend: negatives given without data or weights");
    if (batch == 0) return CLIPDB_OK;
    const size_t smem = static_cast<size_t>(dim) * sizeof(float);
    if (smem > 48 * 1024) return fail(c, CLIPDB_ERR_UNSUPPORTED, "blend: dim %d too large", dim);
    BlendArgs a{};
    a.e1 = d_e1;
    a.e2 = d_e2;
    a.w = d_w;
    a.negs = n_neg > 0 ? d_negs : nullptr;
    a.neg_w = d_neg_w;
    a.out = d_out;
    a.flags = d_flags;
    a.n_neg = n_neg;
    a.dim = dim;
    blend_kernel<<<batch, BLEND_THREADS, smem, c->stream>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

// stage the host inputs of one blend into device memory; returns device pointers
struct BlendStaged {
    float *e1, *e2, *w, *negs, *neg_w, *out;
    int32_t *flags;
};

int stage_blend(clipdb_ctx *c, const float *e1, const float *e2, double w0, double w1,
                const float *negs, const double *neg_w, int32_t n_neg, int32_t dim,
                BlendStaged *st) {
    if (!e1 || dim <= 0 || n_neg < 0 || (n_neg > 0 && (!negs || !neg_w)))
        return fail(c, CLIPDB_ERR_INVALID, "blend: bad argument");
    // float layout: e1[dim] e2[dim] w[2] neg_w[n_neg] negs[n_neg*dim] | out[dim]
    const size_t n_in = static_cast<size_t>(dim) * 2 + 2 + n_neg + static_cast<size_t>(n_neg) * dim;
    const size_t n_in_pad = (n_in + 3) & ~static_cast<size_t>(3);
    RC_TRY(ensure_pinned(c, (n_in_pad + dim) * sizeof(float) + 64));
    RC_TRY(ensure_device(c, c->d_blend_in, (n_in_pad + dim) * sizeof(float)));
    RC_TRY(ensure_device(c, c->d_blend_flags, sizeof(int32_t)));
    float *h = static_cast<float *>(c->pinned.p);
    memcpy(h, e1, dim * sizeof(float));
    if (e2) memcpy(h + dim, e2, dim * sizeof(float));
    else memset(h + dim, 0, dim * sizeof(float));
    // weight normalisation exactly as image_database.py:1379-1383 (Python floats = doubles),
    // then the weak-scalar cast to float32 numpy applies when multiplying a float32 array
    double total = w0 + w1;
    if (total == 0.0) {
        w0 = 0.5;
        w1 = 0.5;
        total = 1.0;
    }
    h[2 * dim + 0] = static_cast<float>(w0 / total);
    h[2 * dim + 1] = static_cast<float>(w1 / total);
    for (int j = 0; j < n_neg; j++) h[2 * dim + 2 + j] = static_cast<float>(neg_w[j]);
    if (n_neg > 0) memcpy(h + 2 * dim + 2 + n_neg, negs, static_cast<size_t>(n_neg) * dim * sizeof(float));
    float *d = static_cast<float *>(c->d_blend_in.p);
    CU_TRY(c, cudaMemcpyAsync(d, h, n_in * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    st->e1 = d;
    st->e2 = e2 ? d + dim : nullptr;
    st->w = d + 2 * dim;
    st->neg_w = d + 2 * dim + 2;
    st->negs = d + 2 * dim + 2 + n_neg;
    st->out = d + n_in_pad;
    st->flags = static_cast<int32_t *>(c->d_blend_flags.p);
    return CLIPDB_OK;
}


// ---- batched path (K4) ----------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int encode_bf16_map(clipdb_ctx *c, CUtensorMap *map, void *base, uint64_t rows, uint32_t box_rows,
                    uint32_t row_elems = SCAN_DIM) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU_TRY(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess)
            return fail(c, CLIPDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_elems), rows};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_elems) * 2};
    const cuuint32_t box[2] = {BQ_BLOCK_K, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, CLIPDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return CLIPDB_OK;
}


// Pass A's sampling ratio s: it visits one tile (single-CTA kernel: 128 rows; CTA-pair kernel: 256
// rows) out of every s, i.e. 1/s of the rows.  Sparser sampling makes pass A cheaper and the
// thresholds looser: pass B then keeps about 2.5 * k * s candidates per query (measured on unit
// random data), which must stay well inside the candidate capacity and, past ~8,000 per query,
// starts to cost pass B more than pass A saves (profiles/r01v4_ab_tests.txt).  Auto: min(capacity,
// 32768) / (8k), at most 256, never fewer than 8k sampled 32-row groups; a power of two.
int batch_sample_stride(const clipdb_ctx *c, int64_t tiles, int k) {
    int64_t s = c->batch_sample_stride;
    if (s <= 0) {
        const int64_t kk = k > 0 ? k : 1;
        const int64_t budget = c->batch_cand_cap < 32768 ? c->batch_cand_cap : 32768;   // keep 2x+ headroom
        s = budget / (8 * kk);
        if (s > tiles / (2 * kk)) s = tiles / (2 * kk);   // at least 8k group maxima to pick the k-th from
    }
    if (s > BQ_SAMPLE_STRIDE_MAX) s = BQ_SAMPLE_STRIDE_MAX;
    int p = 1;
    while (p * 2 <= s) p *= 2;
    return p;
}

// bf16 buffer with room for `cap_rows` rows (whole 128-row tiles, zero filled); rows already converted
// survive a growth.  The TMA map covers the whole buffer, so it only changes when the buffer does.
int batch_ensure_store(clipdb_ctx *c, int64_t cap_rows) {
    const int64_t tiles = (cap_rows + BQ_M - 1) / BQ_M;
    const int64_t want_rows = (tiles > 0 ? tiles : 1) * BQ_M;
    if (c->bf16_rows.p && c->bf16_cap_rows >= want_rows) return CLIPDB_OK;
    const size_t bytes = static_cast<size_t>(want_rows) * SCAN_DIM * 2;
    void *fresh = nullptr;
    CU_TRY(c, cudaMalloc(&fresh, bytes));
    const size_t keep = c->bf16_rows.p ? static_cast<size_t>((c->bf16_built + BQ_M - 1) / BQ_M) * BQ_M * SCAN_DIM * 2 : 0;
    if (keep) CU_TRY(c, cudaMemcpyAsync(fresh, c->bf16_rows.p, keep, cudaMemcpyDeviceToDevice, c->stream));
    CU_TRY(c, cudaMemsetAsync(static_cast<uint8_t *>(fresh) + keep, 0, bytes - keep, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    free_buffer(c->bf16_rows);
    c->bf16_rows.p = fresh;
    c->bf16_rows.bytes = bytes;
    c->bf16_cap_rows = want_rows;
    return encode_bf16_map(c, &c->map_rows, c->bf16_rows.p, static_cast<uint64_t>(want_rows / BQ_M) * BQ_K_BLOCKS * BQ_M,
                           BQ_M, BQ_BLOCK_K);
}

// convert rows [begin, end) into the bf16 copy.  `src` (nullable) = a DEVICE-addressable buffer holding exactly
// those rows (the caller's chunk being appended: saves re-reading it from the store, which for the host tier
// of a tiered store means a second trip over PCIe); otherwise they are read from the store itself.
int batch_convert_rows(clipdb_ctx *c, int64_t begin, int64_t end, const float *src) {
    auto launch = [&](const float *biased_base, int64_t lo, int64_t hi) -> int {
        if (hi <= lo) return CLIPDB_OK;
        const int64_t warps_wanted = hi - lo;
        int64_t blocks = (warps_wanted + 7) / 8;
        if (blocks > static_cast<int64_t>(c->sm_count) * 8) blocks = static_cast<int64_t>(c->sm_count) * 8;
        build_bf16_store_kernel<<<static_cast<unsigned>(blocks), 256, 0, c->stream>>>(
            biased_base, lo, hi, static_cast<__nv_bfloat16 *>(c->bf16_rows.p),
            static_cast<unsigned long long *>(c->bad_rows.p), static_cast<unsigned int *>(c->row_err.p));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        return CLIPDB_OK;
    };
    if (src) {
        RC_TRY(launch(src - begin * SCAN_DIM, begin, end));
    } else {
        int64_t mid = begin > c->split ? begin : c->split;   // first row of the range that lives in the host tier
        if (mid > end) mid = end;
        RC_TRY(launch(c->rows, begin, mid));
        RC_TRY(launch(store_rows_hi(c), mid, end));
    }
    CU_TRY(c, cudaMemcpyAsync(&c->batch_bad_rows, c->bad_rows.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                              c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (end > c->bf16_built) c->bf16_built = end;
    return CLIPDB_OK;
}

// (re)build whatever part of the bf16 copy is missing: everything after clipdb_enable_batch or when the copy
// was invalidated, otherwise only the rows appended since (each append converts its own rows, so this is
// normally a no-op).
int batch_build_locked(clipdb_ctx *c) {
    if (!c->rows || c->dim != SCAN_DIM || c->ld != SCAN_DIM)
        return fail(c, CLIPDB_ERR_UNSUPPORTED, "batched path needs a loaded (or reserved) store with dim == %d", SCAN_DIM);
    const bool first = !c->batch_enabled;
    RC_TRY(ensure_device(c, c->bad_rows, sizeof(unsigned long long)));
    RC_TRY(ensure_device(c, c->bq_queries, static_cast<size_t>(BQ_N) * SCAN_DIM * 2));
    RC_TRY(ensure_device(c, c->bq_qnorm, BQ_N * sizeof(float)));
    RC_TRY(ensure_device(c, c->bq_qerr, BQ_N * sizeof(float)));
    RC_TRY(ensure_device(c, c->row_err, sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_thr, BQ_N * sizeof(float)));
    RC_TRY(ensure_device(c, c->bq_flags, BQ_N * sizeof(int32_t)));
    RC_TRY(ensure_device(c, c->bq_count, BQ_N * sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_cand, static_cast<size_t>(BQ_N) * c->batch_cand_cap * sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_cand_u, static_cast<size_t>(BQ_N) * c->batch_cand_cap * sizeof(float)));
    RC_TRY(ensure_device(c, c->bq_margin, BQ_N * sizeof(float)));
    RC_TRY(ensure_device(c, c->bq_surv, static_cast<size_t>(BQ_N) * c->batch_cand_cap * sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_surv_count, BQ_N * sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_tilectr, 2 * sizeof(unsigned int)));
    RC_TRY(ensure_device(c, c->bq_parts, static_cast<size_t>(BQ_N) * 8 * 128 * sizeof(uint64_t)));
    if (first || c->batch_dirty) {
        c->bf16_built = 0;
        CU_TRY(c, cudaMemsetAsync(c->row_err.p, 0, sizeof(unsigned int), c->stream));
        CU_TRY(c, cudaMemsetAsync(c->bad_rows.p, 0, sizeof(unsigned long long), c->stream));
        c->batch_bad_rows = 0;
    }
    RC_TRY(batch_ensure_store(c, c->cap > c->n ? c->cap : c->n));
    if (first) {
        RC_TRY(encode_bf16_map(c, &c->map_q, c->bq_queries.p, BQ_N, BQ_N));
        RC_TRY(encode_bf16_map(c, &c->map_qhalf[0], c->bq_queries.p, BQ_N, 64 / 2));
        RC_TRY(encode_bf16_map(c, &c->map_qhalf[1], c->bq_queries.p, BQ_N, 128 / 2));
        RC_TRY(encode_bf16_map(c, &c->map_qhalf[2], c->bq_queries.p, BQ_N, 256 / 2));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<64>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<64>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<128>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<128>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_pair_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SMEM_BYTES));
        CU_TRY(c, cudaFuncSetAttribute(batch_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SMEM_BYTES));
    }
    if (c->bf16_built < c->n) RC_TRY(batch_convert_rows(c, c->bf16_built, c->n, nullptr));
    c->batch_enabled = true;
    c->batch_dirty = false;
    return CLIPDB_OK;
}

// A mask is honoured by both passes' epilogues; the NaN count reported by the batched path is
// the store-wide one taken when the bf16 copy was built, so a masked search over a store that
// holds zero-norm rows goes through the exact scan (which counts admitted NaN rows only).
bool batch_eligible(const clipdb_ctx *c, int32_t nq, int32_t k, int32_t metric, int32_t use_mask) {
    return c->batch_enabled && metric == CLIPDB_METRIC_COSINE && (!use_mask || (c->mask && c->batch_bad_rows == 0)) &&
           k >= 1 && k <= FUSED_K_MAX && k <= c->n && nq >= 1 && nq <= BQ_N;
}

template <bool DUMP>
void launch_pair_gemm(clipdb_ctx *c, int npass, int grid, const BatchGemmArgs &g) {
    switch (npass) {
        case 64:
            batch_gemm_pair_kernel<DUMP, 64><<<grid, BQ_THREADS, PairCfg<64>::SMEM_BYTES, c->stream>>>(
                c->map_rows, c->map_qhalf[0], g);
            break;
        case 128:
            batch_gemm_pair_kernel<DUMP, 128><<<grid, BQ_THREADS, PairCfg<128>::SMEM_BYTES, c->stream>>>(
                c->map_rows, c->map_qhalf[1], g);
            break;
        default:
            batch_gemm_pair_kernel<DUMP, 256><<<grid, BQ_THREADS, PairCfg<256>::SMEM_BYTES, c->stream>>>(
                c->map_rows, c->map_qhalf[2], g);
            break;
    }
}

template <int KPL>
int batch_launch_tail(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k, int64_t *d_out_rowids,
                      float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan, int32_t *d_flags, bool threshold,
                      const ExchangeArgs *bx = nullptr) {
    uint64_t *parts = static_cast<uint64_t *>(c->bq_parts.p);
    if (threshold) {
        batch_threshold_kernel<<<BQ_N / BQ_THR_QPC, BQ_THR_THREADS, 0, c->stream>>>(
            static_cast<const float *>(c->bq_scores.p), c->bq_sample_groups_used,
            static_cast<const float *>(c->bq_qnorm.p), static_cast<const float *>(c->bq_qerr.p),
            static_cast<const unsigned int *>(c->row_err.p), nq, k, static_cast<float *>(c->bq_thr.p),
            static_cast<float *>(c->bq_margin.p), d_flags);
    } else {
        batch_refine_kernel<<<nq, BQ_SELECT_THREADS, 0, c->stream>>>(
            static_cast<const float *>(c->bq_cand_u.p), static_cast<const unsigned int *>(c->bq_cand.p),
            static_cast<const unsigned int *>(c->bq_count.p), static_cast<int>(c->batch_cand_cap), k,
            static_cast<int>(c->batch_refine), static_cast<const float *>(c->bq_margin.p),
            static_cast<unsigned int *>(c->bq_surv.p), static_cast<unsigned int *>(c->bq_surv_count.p));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        RerankArgs r{};
        r.rows = c->rows;
        r.rows_hi = store_rows_hi(c);
        r.split = c->split;
        r.queries = d_queries;
        r.cand_count = static_cast<const unsigned int *>(c->bq_count.p);
        r.surv_rows = static_cast<const unsigned int *>(c->bq_surv.p);
        r.surv_count = static_cast<const unsigned int *>(c->bq_surv_count.p);
        r.part_keys = parts;
        r.cand_cap = static_cast<int>(c->batch_cand_cap);
        r.k = k;
        r.flags = d_flags;
        r.dec.rowids = c->rowids;
        r.dec.rowid_base = c->rowid_base;
        r.dec.out_rowids = d_out_rowids;
        r.dec.out_dist = d_out_dist;
        r.dec.out_n = d_out_n;
        r.dec.out_nan = d_out_nan;
        r.dec.k = k;
        r.bad_rows = static_cast<const unsigned long long *>(c->bad_rows.p);
        if (bx) r.xchg = *bx;
        batch_rerank_kernel<KPL><<<dim3(nq, BQ_RERANK_PARTS), BQ_SEL_THREADS, 0, c->stream>>>(r);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        batch_rerank_finish_kernel<<<nq, 256, 0, c->stream>>>(r, BQ_RERANK_PARTS, 32 * KPL);
    }
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

int batch_search_device_locked(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k, int32_t use_mask,
                               int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan,
                               int32_t *d_flags, const ExchangeArgs *bx = nullptr) {
    if (!d_queries || !d_out_rowids || !d_out_dist || !d_out_n || !d_flags)
        return fail(c, CLIPDB_ERR_INVALID, "search_batch: null pointer");
    if (c->batch_enabled && c->batch_dirty) RC_TRY(batch_build_locked(c));   // rows were appended / updated
    if (!batch_eligible(c, nq, k, CLIPDB_METRIC_COSINE, use_mask))
        return fail(c, CLIPDB_ERR_STATE, "search_batch: batch store not enabled or arguments out of range "
                                         "(1 <= nq <= 256, 1 <= k <= 128 and <= rows; a mask needs a store without "
                                         "zero-norm rows)");
    const int kpl = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    const int tiles = static_cast<int>((c->n + BQ_M - 1) / BQ_M);
    prep_queries_kernel<<<BQ_N, 128, 0, c->stream>>>(d_queries, nq, static_cast<__nv_bfloat16 *>(c->bq_queries.p),
                                                     static_cast<float *>(c->bq_qnorm.p), static_cast<float *>(c->bq_qerr.p));
    CU_TRY(c, cudaGetLastError());
    c->launches++;

    BatchGemmArgs g{};
    g.thr = static_cast<const float *>(c->bq_thr.p);
    g.cand_count = static_cast<unsigned int *>(c->bq_count.p);
    g.cand_rows = static_cast<unsigned int *>(c->bq_cand.p);
    g.cand_u = static_cast<float *>(c->bq_cand_u.p);
    g.mask = use_mask ? c->mask : nullptr;
    g.n = c->n;
    g.total_tiles = tiles;
    g.cand_cap = static_cast<int>(c->batch_cand_cap);

    // dynamic tile scheduler counters of the CTA-pair kernel: [0] pass A, [1] pass B
    CU_TRY(c, cudaMemsetAsync(c->bq_tilectr.p, 0, 2 * sizeof(unsigned int), c->stream));
    g.tile_counter = static_cast<unsigned int *>(c->bq_tilectr.p);
    g.static_tiles = static_cast<int>(c->batch_static_tiles);
    // pass A: group maxima over a tile sample -> per-query thresholds
    const bool pair = c->batch_cta_pair != 0 && (c->sm_count % 2 == 0);
    const int pair_grid = c->sm_count & ~1;
    // queries per pass (UMMA N): the smallest instantiation that holds the batch
    int npass = nq <= 64 ? 64 : (nq <= 128 ? 128 : 256);
    if (c->batch_npass == 64 || c->batch_npass == 128 || c->batch_npass == 256)
        npass = c->batch_npass >= npass ? static_cast<int>(c->batch_npass) : npass;
    const int sstride = batch_sample_stride(c, tiles, k);
    // one maximum per (sampled 128-row tile, epilogue warp)
    g.tile_stride = sstride;   // in the kernel's own tiles (pair kernel: 256-row pair tiles): 1/sstride of the rows
    const int eff = ((pair ? (tiles + 1) / 2 : tiles) + g.tile_stride - 1) / g.tile_stride;
    g.sample_groups = static_cast<long long>(eff) * (pair ? 8 : 4);
    RC_TRY(ensure_device(c, c->bq_scores, static_cast<size_t>(BQ_N) * g.sample_groups * sizeof(float)));
    g.scores = static_cast<float *>(c->bq_scores.p);
    if (pair) {
        const int grid = 2 * eff < pair_grid ? 2 * eff : pair_grid;
        launch_pair_gemm<true>(c, npass, grid, g);
    } else {
        batch_gemm_kernel<true><<<eff < c->sm_count ? eff : c->sm_count, BQ_THREADS, BQ_SMEM_BYTES, c->stream>>>(
            c->map_rows, c->map_q, g);
    }
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    c->bq_sample_groups_used = g.sample_groups;
    switch (kpl) {
        case 1: RC_TRY((batch_launch_tail<1>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, true))); break;
        case 2: RC_TRY((batch_launch_tail<2>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, true))); break;
        default: RC_TRY((batch_launch_tail<4>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, true))); break;
    }
    // pass B: all tiles, keep (q,row) above the threshold
    CU_TRY(c, cudaMemsetAsync(c->bq_count.p, 0, BQ_N * sizeof(unsigned int), c->stream));
    g.tile_stride = 1;
    g.tile_counter = static_cast<unsigned int *>(c->bq_tilectr.p) + 1;
    if (c->profiling) {
        const bool fresh = c->bq_clock.p == nullptr;
        RC_TRY(ensure_device(c, c->bq_clock, 3 * sizeof(unsigned long long)));
        if (fresh) CU_TRY(c, cudaMemsetAsync(c->bq_clock.p, 0, 3 * sizeof(unsigned long long), c->stream));
        g.clock = static_cast<unsigned long long *>(c->bq_clock.p);
    }
    RC_TRY(profile_mark(c, true));
    if (pair) {
        const int ptiles = (tiles + 1) / 2;
        const int grid = 2 * ptiles < pair_grid ? 2 * ptiles : pair_grid;
        launch_pair_gemm<false>(c, npass, grid, g);
    } else {
        batch_gemm_kernel<false><<<tiles < c->sm_count ? tiles : c->sm_count, BQ_THREADS, BQ_SMEM_BYTES, c->stream>>>(
            c->map_rows, c->map_q, g);
    }
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    RC_TRY(profile_mark(c, false));
    // exact re-rank
    switch (kpl) {
        case 1: return batch_launch_tail<1>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, false, bx);
        case 2: return batch_launch_tail<2>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, false, bx);
        default: return batch_launch_tail<4>(c, d_queries, nq, k, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags, false, bx);
    }
}

// ---- sign-code search (binary.cuh) -----------------------------------------------------

template <int KPL, bool ALL>
int launch_binary_scan(clipdb_ctx *c, const BinaryArgs &a, int grid) {
    auto kern = binary_scan_kernel<KPL, ALL>;
    static thread_local int configured_device = -1;
    if (configured_device != c->device) {
        CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM_BYTES));
        configured_device = c->device;
    }
    RC_TRY(profile_mark(c, true));
    kern<<<grid, BIN_THREADS, BIN_SMEM_BYTES, c->stream>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return profile_mark(c, false);
}

int binary_search_device_locked(clipdb_ctx *c, const uint32_t *d_query_words, int32_t k, int32_t score_mode,
                                int32_t use_mask, int64_t *d_out_ids, int32_t *d_out_scores, int32_t *d_out_n) {
    if (!c->codes || c->n_codes == 0) return fail(c, CLIPDB_ERR_STATE, "binary_search: no codes loaded");
    if (!d_query_words || !d_out_n || (k > 0 && (!d_out_ids || !d_out_scores)))
        return fail(c, CLIPDB_ERR_INVALID, "binary_search: null pointer");
    if (score_mode != CLIPDB_SCORE_REFERENCE_UINT8 && score_mode != CLIPDB_SCORE_POPCOUNT)
        return fail(c, CLIPDB_ERR_INVALID, "binary_search: unknown score mode %d", score_mode);
    if (use_mask && !c->code_mask) return fail(c, CLIPDB_ERR_STATE, "binary_search: use_mask set but no code mask installed");
    if (c->n_codes >= (1ll << 32)) return fail(c, CLIPDB_ERR_UNSUPPORTED, "more than 2^32-1 codes per context");
    const int64_t kk = k < c->n_codes ? k : c->n_codes;
    if (kk <= 0) {
        CU_TRY(c, cudaMemsetAsync(d_out_n, 0, sizeof(int32_t), c->stream));
        return CLIPDB_OK;
    }
    const int grid = c->scan_ctas > 0 ? static_cast<int>(c->scan_ctas) : c->sm_count;

    BinaryArgs a{};
    a.codes = c->codes;
    a.query = d_query_words;
    a.mask = use_mask ? c->code_mask : nullptr;
    a.seq = use_mask ? c->code_seq : nullptr;
    RC_TRY(prepare_sync(c, &a.sync));
    a.n = c->n_codes;
    a.k = static_cast<int>(kk);
    a.chunk_tiles = static_cast<int>(c->scan_chunk);
    a.score_mask = score_mode == CLIPDB_SCORE_REFERENCE_UINT8 ? 0xFFu : 0xFFFFu;
    a.score_max = score_mode == CLIPDB_SCORE_REFERENCE_UINT8 ? 255u : static_cast<uint32_t>(SCAN_DIM);

    DecodeArgs dec{};
    dec.rowids = a.seq ? c->code_ids_by_seq : c->code_ids;
    dec.rowid_base = 0;
    dec.out_rowids = d_out_ids;
    dec.out_dist = reinterpret_cast<float *>(d_out_scores);   // int32 scores (score_max > 0)
    dec.out_n = d_out_n;
    dec.k = static_cast<int>(kk);
    dec.score_max = static_cast<int>(a.score_max);

    if (kk <= FUSED_K_MAX) {
        const int kpl = kk <= 32 ? 1 : (kk <= 64 ? 2 : 4);
        a.cand_stride = 32 * kpl;
        RC_TRY(ensure_device(c, c->cand_a, static_cast<size_t>(grid) * a.cand_stride * sizeof(uint64_t)));
        a.cand = static_cast<uint64_t *>(c->cand_a.p);
        a.fuse_tail = c->fuse_tail &&
                      static_cast<size_t>(grid) * a.cand_stride * sizeof(uint64_t) <= BIN_STAGES * BIN_STAGE_BYTES;
        a.dec = dec;
        switch (kpl) {
            case 1: RC_TRY((launch_binary_scan<1, false>(c, a, grid))); break;
            case 2: RC_TRY((launch_binary_scan<2, false>(c, a, grid))); break;
            default: RC_TRY((launch_binary_scan<4, false>(c, a, grid))); break;
        }
        if (a.fuse_tail) {
            c->sync_dirty = false;
            return CLIPDB_OK;
        }
        return merge_cta_lists(c, grid, a.cand_stride, dec);
    }
    const size_t key_bytes = static_cast<size_t>(c->n_codes) * sizeof(uint64_t);
    RC_TRY(ensure_device(c, c->all_keys_a, key_bytes));
    RC_TRY(ensure_device(c, c->all_keys_b, key_bytes));
    a.all_keys = static_cast<uint64_t *>(c->all_keys_a.p);
    a.cand_stride = 32;
    a.k = 0;
    RC_TRY((launch_binary_scan<1, true>(c, a, grid)));
    return sort_all_keys_and_decode(c, c->n_codes, kk, dec);
}

}  // namespace

// ================================ C ABI ==========================================

extern "C" {

int clipdb_abi_version(void) { return ABI_VERSION; }

// the marker lets a binding read the hash out of the file without loading it (a stale library must not be
// dlopen'ed in the process that is about to rebuild and load the fresh one)
static const char SOURCE_HASH_MARKER[] = "clipdb-source-hash:" CLIPDB_SOURCE_HASH;
const char *clipdb_source_hash(void) { return SOURCE_HASH_MARKER + 19; }

int clipdb_create(int device, clipdb_ctx **out) {
    if (!out) return CLIPDB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return CLIPDB_ERR_CUDA;  // no CPU fallback
    }
    if (device < 0 || device >= count) return CLIPDB_ERR_INVALID;
    clipdb_ctx *c = new (std::nothrow) clipdb_ctx();
    if (!c) return CLIPDB_ERR_NOMEM;
    c->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (!g.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return CLIPDB_ERR_CUDA;
    }
    if (prop.major < 10) {
        cudaStreamDestroy(c->own_stream);
        delete c;
        return CLIPDB_ERR_UNSUPPORTED;  // sm_100a code only
    }
    c->sm_count = prop.multiProcessorCount;
    c->stream = c->own_stream;
    *out = c;
    return CLIPDB_OK;
}

void clipdb_destroy(clipdb_ctx *c) {
    if (!c) return;
    {
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
        release_store(c);
        release_codes(c);
        release_exchange(c);
        Buffer *bufs[] = {&c->d_code_query, &c->code_stage, &c->code_bad, &c->cand_a, &c->cand_b, &c->sync_buf, &c->all_keys_a, &c->all_keys_b,
                          &c->cub_tmp, &c->d_query, &c->d_results, &c->d_blend_in, &c->d_blend_flags, &c->bf16_rows,
                          &c->bad_rows, &c->bq_queries, &c->bq_qnorm, &c->bq_scores, &c->bq_thr, &c->bq_flags,
                          &c->bq_count, &c->bq_cand, &c->bq_parts, &c->bq_qerr, &c->row_err, &c->bq_cand_u, &c->bq_margin,
                          &c->bq_surv, &c->bq_surv_count, &c->bq_tilectr, &c->xchg_stats, &c->bq_clock};
        for (Buffer *b : bufs) free_buffer(*b);
        if (c->pinned.p) cudaFreeHost(c->pinned.p);
        if (c->stage.p) cudaFreeHost(c->stage.p);
        for (void *p : c->sql_slots) cudaFreeHost(p);
        if (c->pinned_aux.p) cudaFreeHost(c->pinned_aux.p);
        if (c->pinned_flags.p) cudaFreeHost(c->pinned_flags.p);
        for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        cudaGetLastError();
    }
    delete c;
}

const char *clipdb_last_error(const clipdb_ctx *c) { return c ? c->err.c_str() : "null context"; }

int clipdb_set_stream(clipdb_ctx *c, void *cuda_stream) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    RC_TRY(profile_fold(c));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    return CLIPDB_OK;
}

int clipdb_use_own_stream(clipdb_ctx *c) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    RC_TRY(profile_fold(c));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stream = c->own_stream;
    return CLIPDB_OK;
}

int clipdb_synchronize(clipdb_ctx *c) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return CLIPDB_OK;
}

static int64_t *option_slot(clipdb_ctx *c, const char *name) {
    if (!name) return nullptr;
    if (!strcmp(name, "scan_variant")) return &c->scan_variant;
    if (!strcmp(name, "scan_ctas")) return &c->scan_ctas;
    if (!strcmp(name, "ldg_ctas_per_sm")) return &c->ldg_ctas_per_sm;
    if (!strcmp(name, "evict_first")) return &c->evict_first;
    if (!strcmp(name, "scan_cfg")) return &c->scan_cfg;
    if (!strcmp(name, "scan_assign")) return &c->scan_assign;
    if (!strcmp(name, "scan_chunk")) return &c->scan_chunk;
    if (!strcmp(name, "fuse_tail")) return &c->fuse_tail;
    if (!strcmp(name, "host_tier_from_row")) return &c->host_tier_from_row;
    if (!strcmp(name, "sqlite_readers")) return &c->sqlite_readers;
    if (!strcmp(name, "xchg_timeout_ms")) return &c->xchg_timeout_ms;
    if (!strcmp(name, "batch_min_nq")) return &c->batch_min_nq;
    if (!strcmp(name, "batch_cand_cap")) return &c->batch_cand_cap;
    if (!strcmp(name, "batch_cta_pair")) return &c->batch_cta_pair;
    if (!strcmp(name, "batch_sample_stride")) return &c->batch_sample_stride;
    if (!strcmp(name, "batch_refine")) return &c->batch_refine;
    if (!strcmp(name, "batch_npass")) return &c->batch_npass;
    if (!strcmp(name, "batch_static_tiles")) return &c->batch_static_tiles;
    return nullptr;
}

int clipdb_set_option(clipdb_ctx *c, const char *name, int64_t value) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    int64_t *slot = option_slot(c, name);
    if (!slot) return fail(c, CLIPDB_ERR_INVALID, "unknown option '%s'", name ? name : "(null)");
    if (value < 0 || (value > (1 << 20) && slot != &c->host_tier_from_row))
        return fail(c, CLIPDB_ERR_INVALID, "option '%s' out of range", name);
    if (slot == &c->ldg_ctas_per_sm && value == 0) value = 1;
    if (slot == &c->batch_cand_cap) {
        if (value < 256) return fail(c, CLIPDB_ERR_INVALID, "option 'batch_cand_cap' must be at least 256");
        if (c->batch_enabled && value != *slot) c->batch_dirty = true;   // candidate buffers are sized at build time
    }
    *slot = value;
    return CLIPDB_OK;
}

int clipdb_get_option(clipdb_ctx *c, const char *name, int64_t *value) {
    if (!c || !value) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (name && !strcmp(name, "sm_count")) {
        *value = c->sm_count;
        return CLIPDB_OK;
    }
    if (name && (!strcmp(name, "device_free_bytes") || !strcmp(name, "device_total_bytes"))) {
        // read-only: what a loader needs to decide between an all-HBM store and a tiered one
        DeviceGuard g(c->device);
        size_t free_b = 0, total_b = 0;
        CU_TRY(c, cudaMemGetInfo(&free_b, &total_b));
        *value = static_cast<int64_t>(name[7] == 'f' ? free_b : total_b);
        return CLIPDB_OK;
    }
    int64_t *slot = option_slot(c, name);
    if (!slot) return fail(c, CLIPDB_ERR_INVALID, "unknown option '%s'", name ? name : "(null)");
    *value = *slot;
    return CLIPDB_OK;
}

int64_t clipdb_launch_count(const clipdb_ctx *c) { return c ? c->launches : 0; }

int clipdb_load_rows(clipdb_ctx *c, const float *rows, const int64_t *rowids, int64_t n, int32_t dim) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n < 0 || dim <= 0 || (n > 0 && !rows)) return fail(c, CLIPDB_ERR_INVALID, "load_rows: bad argument");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    RC_TRY(alloc_store(c, n, dim, rowids != nullptr));
    RC_TRY(copy_rows_in(c, rows, rowids, 0, n));
    c->n = n;
    return CLIPDB_OK;
}

static int append_rows_locked(clipdb_ctx *c, const float *rows, const int64_t *rowids, int64_t m);

int clipdb_append_rows(clipdb_ctx *c, const float *rows, const int64_t *rowids, int64_t m) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    return append_rows_locked(c, rows, rowids, m);
}

static int append_rows_locked(clipdb_ctx *c, const float *rows, const int64_t *rowids, int64_t m) {
    if (!c->rows || !c->owns_rows) return fail(c, CLIPDB_ERR_STATE, "append_rows: no owned store (load first)");
    if (m < 0 || (m > 0 && !rows)) return fail(c, CLIPDB_ERR_INVALID, "append_rows: bad argument");
    if ((c->rowids != nullptr) != (rowids != nullptr) && m > 0)
        return fail(c, CLIPDB_ERR_INVALID, "append_rows: rowids must match the store (explicit vs implicit)");
    DeviceGuard g(c->device);
    if (c->n + m > c->cap) {
        int64_t cap = c->cap + c->cap / 2;
        if (cap < c->n + m) cap = c->n + m;
        int64_t *ni = nullptr;
        if (c->rows_host) {
            // tiered store: the HBM tier keeps its size, the host tier grows
            float *nh = nullptr, *nh_dev = nullptr;
            CU_TRY(c, cudaHostAlloc(reinterpret_cast<void **>(&nh), static_cast<size_t>(cap - c->split) * c->ld * sizeof(float),
                                    cudaHostAllocMapped | cudaHostAllocPortable));
            CU_TRY(c, cudaHostGetDevicePointer(reinterpret_cast<void **>(&nh_dev), nh, 0));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            if (c->n > c->split) memcpy(nh, c->rows_host, static_cast<size_t>(c->n - c->split) * c->ld * sizeof(float));
            cudaFreeHost(c->rows_host);
            c->rows_host = nh;
            c->rows_host_dev = nh_dev;
        } else {
            float *nr = nullptr;
            CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&nr), static_cast<size_t>(cap) * c->ld * sizeof(float)));
            CU_TRY(c, cudaMemcpyAsync(nr, c->rows, static_cast<size_t>(c->n) * c->ld * sizeof(float),
                                      cudaMemcpyDeviceToDevice, c->stream));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->rows);
            c->rows = nr;
        }
        if (c->rowids) {
            CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&ni), static_cast<size_t>(cap) * sizeof(int64_t)));
            CU_TRY(c, cudaMemcpyAsync(ni, c->rowids, static_cast<size_t>(c->n) * sizeof(int64_t),
                                      cudaMemcpyDeviceToDevice, c->stream));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->rowids);
            c->rowids = ni;
        }
        c->cap = cap;
    }
    RC_TRY(copy_rows_in(c, rows, rowids, c->n, m));
    const int64_t old_n = c->n;
    c->n += m;
    if (m > 0 && c->batch_enabled && !c->batch_dirty) {
        // keep the bf16 copy current: convert just the new rows (from the caller's buffer when the GPU can
        // address it, else from where they landed in the store)
        RC_TRY(batch_ensure_store(c, c->cap));
        RC_TRY(batch_convert_rows(c, old_n, c->n, is_device_pointer(rows) ? rows : nullptr));
    }
    if (c->mask) {  // a mask sized for the old row count no longer applies
        cudaFree(c->mask);
        c->mask = nullptr;
        c->mask_words = 0;
    }
    return CLIPDB_OK;
}

int clipdb_update_row(clipdb_ctx *c, int64_t position, const float *row) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->rows || !c->owns_rows) return fail(c, CLIPDB_ERR_STATE, "update_row: no owned store");
    if (position < 0 || position >= c->n || !row) return fail(c, CLIPDB_ERR_INVALID, "update_row: bad argument");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaMemcpyAsync(store_row_host_view(c, position), row, c->dim * sizeof(float), cudaMemcpyDefault,
                              c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->batch_enabled && !c->batch_dirty) {
        // re-convert the one row; the store-wide count of zero-norm rows can only be kept exact that way
        // while it is zero (a repaired bad row would stay counted), otherwise rebuild before the next use
        if (c->batch_bad_rows == 0) RC_TRY(batch_convert_rows(c, position, position + 1, nullptr));
        else c->batch_dirty = true;
    }
    return CLIPDB_OK;
}

int clipdb_attach_rows(clipdb_ctx *c, const float *d_rows, const int64_t *d_rowids, int64_t n,
                       int32_t dim, int64_t rowid_base) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n < 0 || dim <= 0 || !d_rows) return fail(c, CLIPDB_ERR_INVALID, "attach_rows: bad argument");
    if (dim % 4 != 0 || (reinterpret_cast<uintptr_t>(d_rows) & 15))
        return fail(c, CLIPDB_ERR_INVALID, "attach_rows: dim must be a multiple of 4 and rows 16-byte aligned");
    DeviceGuard g(c->device);
    if (!is_device_pointer(d_rows) || (d_rowids && !is_device_pointer(d_rowids)))
        return fail(c, CLIPDB_ERR_INVALID, "attach_rows: device pointers required");
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    release_store(c);
    c->rows = const_cast<float *>(d_rows);
    c->rowids = const_cast<int64_t *>(d_rowids);
    c->owns_rows = false;
    c->n = c->cap = n;
    c->dim = c->ld = dim;
    c->rowid_base = rowid_base;
    return CLIPDB_OK;
}

int clipdb_reserve_rows(clipdb_ctx *c, int64_t capacity, int32_t dim, int32_t explicit_rowids, int32_t placement) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (capacity < 0 || dim <= 0) return fail(c, CLIPDB_ERR_INVALID, "reserve_rows: bad argument");
    if (placement != CLIPDB_PLACE_DEVICE && placement != CLIPDB_PLACE_HOST)
        return fail(c, CLIPDB_ERR_INVALID, "reserve_rows: unknown placement %d", placement);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    // option "host_tier_from_row": with CLIPDB_PLACE_HOST, rows below it still go to HBM (what is left of it)
    const int64_t device_rows = placement == CLIPDB_PLACE_DEVICE ? -1 : c->host_tier_from_row;
    const int rc = alloc_store(c, capacity, dim, explicit_rowids != 0, device_rows);
    if (rc != CLIPDB_OK) release_store(c);
    return rc;
}

int clipdb_stage_buffer(clipdb_ctx *c, int64_t bytes, void **out_host) {
    if (!c || !out_host) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (bytes <= 0) return fail(c, CLIPDB_ERR_INVALID, "stage_buffer: bytes must be positive");
    DeviceGuard g(c->device);
    Buffer &b = c->stage;
    if (!b.p || b.bytes < static_cast<size_t>(bytes)) {
        if (b.p) {
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            CU_TRY(c, cudaFreeHost(b.p));
            b.p = nullptr;
            b.bytes = 0;
        }
        CU_TRY(c, cudaMallocHost(&b.p, static_cast<size_t>(bytes)));
        b.bytes = static_cast<size_t>(bytes);
    }
    *out_host = b.p;
    return CLIPDB_OK;
}

// The statement of the native loader: the row set and order of the reference's search statement
// (image_database.py:1564-1571) — vec0 scanned in rowid order, INNER JOINed to image_embeddings and images — with the
// join order pinned by CROSS JOIN (SQLite keeps the written nesting), so rows arrive in rowid order without a sorter.
static const char *const SQL_LOAD_ROWS =
    "SELECT v.rowid, ie.image_id, i.last_modified, i.file_path, v.embedding "
    "FROM vec0 AS v CROSS JOIN image_embeddings AS ie ON ie.rowid = v.rowid "
    "CROSS JOIN images AS i ON i.id = ie.image_id "
    "WHERE v.rowid > ?1 AND v.rowid <= ?2 ORDER BY v.rowid";

namespace {

// one staging slot of the native loader: a pinned row buffer + what the host keeps per row
struct SqlSlot {
    std::vector<int64_t> ids, image_ids;
    std::vector<double> mtimes;
    std::string paths;
    int64_t m = 0;
};

// an open read-only connection inside one read transaction, with the row statement prepared
struct SqlReader {
    const SqliteApi *sq = nullptr;
    void *db = nullptr, *st_rows = nullptr;
    std::string err;

    bool open(const SqliteApi *api, const char *db_path) {
        sq = api;
        const std::string uri = std::string("file:") + db_path + "?mode=ro";
        if (sq->open_v2(uri.c_str(), &db, SQLITE_OPEN_READONLY_ | SQLITE_OPEN_URI_, nullptr) != SQLITE_OK_) {
            err = std::string("cannot open the database: ") + (db ? sq->errmsg(db) : "out of memory");
            return false;
        }
        sq->busy_timeout(db, 30000);
        if (sq->exec(db, "BEGIN", nullptr, nullptr, nullptr) != SQLITE_OK_ ||
            sq->prepare_v2(db, SQL_LOAD_ROWS, -1, &st_rows, nullptr) != SQLITE_OK_) {
            err = std::string("cannot prepare the row statement: ") + sq->errmsg(db);
            return false;
        }
        return true;
    }

    // rows with rowid in (lo, hi] into `slot` / `pinned`, at most `cap` of them; false (and err) on failure
    bool read_range(int64_t lo, int64_t hi, int32_t dim, int64_t cap, uint8_t *pinned, SqlSlot &slot) {
        const size_t row_bytes = static_cast<size_t>(dim) * sizeof(float);
        slot.m = 0;
        slot.paths.clear();
        sq->reset(st_rows);
        sq->bind_int64(st_rows, 1, lo);
        sq->bind_int64(st_rows, 2, hi);
        int rc;
        while ((rc = sq->step(st_rows)) == SQLITE_ROW_) {
            const int nbytes = sq->column_bytes(st_rows, 4);
            const void *blob = sq->column_blob(st_rows, 4);
            if (static_cast<size_t>(nbytes) != row_bytes || !blob) {
                char msg[160];
                snprintf(msg, sizeof msg, "vec0 rowid %lld: %d bytes, expected float32[%d]", sq->column_int64(st_rows, 0),
                         nbytes, dim);
                err = msg;
                return false;
            }
            if (slot.m >= cap) {
                err = "more rows in a rowid stripe than rowids (corrupt table?)";
                return false;
            }
            memcpy(pinned + static_cast<size_t>(slot.m) * row_bytes, blob, row_bytes);
            slot.ids[slot.m] = sq->column_int64(st_rows, 0);
            slot.image_ids[slot.m] = sq->column_int64(st_rows, 1);
            slot.mtimes[slot.m] = sq->column_double(st_rows, 2);
            const unsigned char *path = sq->column_text(st_rows, 3);
            if (path) slot.paths.append(reinterpret_cast<const char *>(path));
            slot.paths.push_back('\0');
            slot.m++;
        }
        if (rc != SQLITE_DONE_) {
            err = std::string("reading vec0 failed: ") + sq->errmsg(db);
            return false;
        }
        return true;
    }

    ~SqlReader() {
        if (st_rows) sq->finalize(st_rows);
        if (db) {
            sq->exec(db, "ROLLBACK", nullptr, nullptr, nullptr);
            sq->close_v2(db);
        }
    }
};

bool sql_scalar_i64(const SqliteApi *sq, void *db, const char *sql, int64_t a, int64_t b, int64_t *out, bool *is_null) {
    void *st = nullptr;
    if (sq->prepare_v2(db, sql, -1, &st, nullptr) != SQLITE_OK_) return false;
    sq->bind_int64(st, 1, a);
    sq->bind_int64(st, 2, b);
    bool ok = false;
    if (sq->step(st) == SQLITE_ROW_) {
        ok = true;
        *is_null = sq->column_bytes(st, 0) == 0 && sq->column_int64(st, 0) == 0 && sq->column_text(st, 0) == nullptr;
        *out = sq->column_int64(st, 0);
    }
    sq->finalize(st);
    return ok;
}

}  // namespace

int clipdb_append_sqlite(clipdb_ctx *c, const char *db_path, int64_t min_rowid, int64_t max_rowid, int64_t chunk_rows,
                         clipdb_sqlite_chunk_fn on_chunk, void *user, int64_t *out_vec0_rows, int64_t *out_joined_rows) {
    if (!c || !db_path) return CLIPDB_ERR_INVALID;
    if (out_vec0_rows) *out_vec0_rows = 0;
    if (out_joined_rows) *out_joined_rows = 0;
    const SqliteApi *sq = sqlite_api();
    if (!sq) {
        std::lock_guard<std::mutex> lk(c->mu);
        return fail(c, CLIPDB_ERR_UNSUPPORTED, "append_sqlite: libsqlite3.so.0 is not available to the native loader");
    }
    auto report = [&](int code, const std::string &what) {
        std::lock_guard<std::mutex> lk(c->mu);
        return fail(c, code, "append_sqlite: %s", what.c_str());
    };
    int32_t dim = 0;
    int readers = 1;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->rows || !c->owns_rows || c->dim <= 0)
            return fail(c, CLIPDB_ERR_STATE, "append_sqlite: reserve (or load) an owned store first");
        if (!c->rowids && c->n > 0)
            return fail(c, CLIPDB_ERR_STATE, "append_sqlite: the store must carry explicit rowids");
        dim = c->dim;
        readers = static_cast<int>(c->sqlite_readers < 1 ? 1 : (c->sqlite_readers > 16 ? 16 : c->sqlite_readers));
    }
    if (chunk_rows <= 0) chunk_rows = 8192;

    // ---- the control connection: what kind of table, how many rows, which rowids
    SqlReader control;
    if (!control.open(sq, db_path)) {
        // a database without vec0 / image_embeddings / images cannot even prepare the statement
        return report(strstr(control.err.c_str(), "prepare") ? CLIPDB_ERR_UNSUPPORTED : CLIPDB_ERR_INVALID, control.err);
    }
    {
        void *st = nullptr;
        bool plain = false;
        if (sq->prepare_v2(control.db, "SELECT type, sql FROM sqlite_master WHERE name = 'vec0'", -1, &st, nullptr) == SQLITE_OK_ &&
            sq->step(st) == SQLITE_ROW_) {
            const char *type = reinterpret_cast<const char *>(sq->column_text(st, 0));
            const char *sql = reinterpret_cast<const char *>(sq->column_text(st, 1));
            plain = type && !strcmp(type, "table") && !(sql && (strstr(sql, "VIRTUAL") || strstr(sql, "virtual")));
        }
        if (st) sq->finalize(st);
        // only a plain table named vec0 is read here; the virtual table / its shadow tables go through the binding's reader
        if (!plain) return report(CLIPDB_ERR_UNSUPPORTED, "vec0 is not a plain table (use the Python reader)");
    }
    int64_t vec0_rows = 0, first = 0, last = 0;
    bool null_a = false, null_b = false, null_c = false;
    if (!sql_scalar_i64(sq, control.db, "SELECT COUNT(*) FROM vec0 WHERE rowid > ?1 AND rowid <= ?2", min_rowid, max_rowid,
                        &vec0_rows, &null_a) ||
        !sql_scalar_i64(sq, control.db, "SELECT MIN(rowid) FROM vec0 WHERE rowid > ?1 AND rowid <= ?2", min_rowid, max_rowid,
                        &first, &null_b) ||
        !sql_scalar_i64(sq, control.db, "SELECT MAX(rowid) FROM vec0 WHERE rowid > ?1 AND rowid <= ?2", min_rowid, max_rowid,
                        &last, &null_c))
        return report(CLIPDB_ERR_INVALID, std::string("cannot count vec0: ") + sq->errmsg(control.db));
    if (out_vec0_rows) *out_vec0_rows = vec0_rows;
    if (vec0_rows == 0) return CLIPDB_OK;

    // ---- stripes of `chunk_rows` consecutive rowid VALUES: a stripe holds at most chunk_rows rows, so it fits one
    // staging slot whatever the joins drop, and stripes committed in order are rows appended in rowid order.  Several
    // readers (own connection, own statement) fill stripes ahead of the committer.  A sparse rowid space (many more
    // stripes than rows warrant) is read by one reader as a single stripe sequence instead.
    const uint64_t span = static_cast<uint64_t>(last) - static_cast<uint64_t>(first) + 1;
    uint64_t n_stripes = (span + static_cast<uint64_t>(chunk_rows) - 1) / static_cast<uint64_t>(chunk_rows);
    const uint64_t dense = static_cast<uint64_t>(vec0_rows) / static_cast<uint64_t>(chunk_rows) + 1;
    const bool striped = n_stripes <= 4 * dense;
    if (!striped) readers = 1;
    // a small load does not repay the extra connections and pinned slots (38 MB each): one more reader per 8 stripes
    if (static_cast<uint64_t>(readers) > 1 + n_stripes / 8) readers = static_cast<int>(1 + n_stripes / 8);
    const int n_slots = striped ? 2 * readers : 2;

    const size_t slot_bytes = static_cast<size_t>(chunk_rows) * dim * sizeof(float);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        DeviceGuard g(c->device);
        if (c->sql_slot_bytes < slot_bytes || static_cast<int>(c->sql_slots.size()) < n_slots) {
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            for (void *p : c->sql_slots) cudaFreeHost(p);
            c->sql_slots.clear();
            c->sql_slot_bytes = 0;
            for (int i = 0; i < n_slots; i++) {
                void *p = nullptr;
                CU_TRY(c, cudaMallocHost(&p, slot_bytes));
                c->sql_slots.push_back(p);
            }
            c->sql_slot_bytes = slot_bytes;
        }
    }
    std::vector<SqlSlot> slots(n_slots);
    for (SqlSlot &sl : slots) {
        sl.ids.resize(chunk_rows);
        sl.image_ids.resize(chunk_rows);
        sl.mtimes.resize(chunk_rows);
    }
    int64_t joined = 0;
    // commit one filled slot: append (DMA from pinned memory), wait for the copy, hand the metadata to the host
    auto commit = [&](int slot_index) -> int {
        SqlSlot &sl = slots[slot_index];
        if (sl.m == 0) return CLIPDB_OK;
        {
            std::lock_guard<std::mutex> lk(c->mu);
            DeviceGuard g(c->device);
            RC_TRY(append_rows_locked(c, static_cast<const float *>(c->sql_slots[slot_index]), sl.ids.data(), sl.m));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
        }
        if (on_chunk && on_chunk(user, sl.m, sl.ids.data(), sl.image_ids.data(), sl.mtimes.data(), sl.paths.data(),
                                 static_cast<int64_t>(sl.paths.size())) != 0)
            return report(CLIPDB_ERR_INVALID, "the chunk callback asked to stop");
        joined += sl.m;
        return CLIPDB_OK;
    };

    int rc = CLIPDB_OK;
    if (!striped) {
        // one reader, one pass: chunks of chunk_rows ROWS cut from a single ordered statement, two slots alternating
        const SqliteApi *api = sq;
        api->reset(control.st_rows);
        api->bind_int64(control.st_rows, 1, min_rowid);
        api->bind_int64(control.st_rows, 2, max_rowid);
        const size_t row_bytes = static_cast<size_t>(dim) * sizeof(float);
        int buf = 0, step_rc;
        slots[0].m = slots[1].m = 0;
        while (rc == CLIPDB_OK && (step_rc = api->step(control.st_rows)) == SQLITE_ROW_) {
            SqlSlot &sl = slots[buf];
            const int nbytes = api->column_bytes(control.st_rows, 4);
            const void *blob = api->column_blob(control.st_rows, 4);
            if (static_cast<size_t>(nbytes) != row_bytes || !blob) {
                char msg[160];
                snprintf(msg, sizeof msg, "vec0 rowid %lld: %d bytes, expected float32[%d]",
                         api->column_int64(control.st_rows, 0), nbytes, dim);
                rc = report(CLIPDB_ERR_INVALID, msg);
                break;
            }
            memcpy(static_cast<uint8_t *>(c->sql_slots[buf]) + static_cast<size_t>(sl.m) * row_bytes, blob, row_bytes);
            sl.ids[sl.m] = api->column_int64(control.st_rows, 0);
            sl.image_ids[sl.m] = api->column_int64(control.st_rows, 1);
            sl.mtimes[sl.m] = api->column_double(control.st_rows, 2);
            const unsigned char *path = api->column_text(control.st_rows, 3);
            if (path) sl.paths.append(reinterpret_cast<const char *>(path));
            sl.paths.push_back('\0');
            if (++sl.m == chunk_rows) {
                rc = commit(buf);
                sl.m = 0;
                sl.paths.clear();
            }
        }
        if (rc == CLIPDB_OK && step_rc != SQLITE_DONE_)
            rc = report(CLIPDB_ERR_INVALID, std::string("reading vec0 failed: ") + api->errmsg(control.db));
        if (rc == CLIPDB_OK) rc = commit(buf);
    } else {
        std::mutex mx;
        std::condition_variable cv;
        std::vector<uint8_t> filled(n_stripes, 0);
        uint64_t committed = 0, next = 0;
        bool stop = false;
        std::string worker_err;
        auto worker = [&](SqlReader *rd) {
            for (;;) {
                uint64_t s;
                {
                    std::unique_lock<std::mutex> lk(mx);
                    s = next++;
                    if (s >= n_stripes) return;
                    // the slot of stripe s was last used by stripe s - n_slots: wait until that one is committed
                    cv.wait(lk, [&] { return stop || s < committed + static_cast<uint64_t>(n_slots); });
                    if (stop) return;
                }
                const int slot_index = static_cast<int>(s % static_cast<uint64_t>(n_slots));
                // rowids in (lo, hi] with lo = first - 1 + s * chunk_rows, clipped to the caller's range
                const uint64_t off = s * static_cast<uint64_t>(chunk_rows);
                const int64_t lo = static_cast<int64_t>(static_cast<uint64_t>(first) + off) - 1;
                int64_t hi = (span - off <= static_cast<uint64_t>(chunk_rows)) ? last
                                                                               : static_cast<int64_t>(static_cast<uint64_t>(lo) + static_cast<uint64_t>(chunk_rows));
                const bool ok = rd->read_range(lo < min_rowid ? min_rowid : lo, hi, dim, chunk_rows,
                                               static_cast<uint8_t *>(c->sql_slots[slot_index]), slots[slot_index]);
                std::lock_guard<std::mutex> lk(mx);
                if (!ok) {
                    if (worker_err.empty()) worker_err = rd->err;
                    stop = true;
                } else {
                    filled[s] = 1;
                }
                cv.notify_all();
                if (!ok) return;
            }
        };
        // the control connection is reader 0; the others open their own (a connection is used by one thread only)
        std::vector<std::unique_ptr<SqlReader>> extra;
        std::vector<std::thread> threads;
        bool opened = true;
        for (int r = 1; r < readers; r++) {
            extra.emplace_back(new SqlReader());
            if (!extra.back()->open(sq, db_path)) {
                worker_err = extra.back()->err;
                opened = false;
                break;
            }
        }
        if (!opened) {
            rc = report(CLIPDB_ERR_INVALID, worker_err);
        } else {
            threads.emplace_back(worker, &control);
            for (auto &rd : extra) threads.emplace_back(worker, rd.get());
            for (uint64_t s = 0; s < n_stripes && rc == CLIPDB_OK; s++) {
                {
                    std::unique_lock<std::mutex> lk(mx);
                    cv.wait(lk, [&] { return stop || filled[s]; });
                    if (!filled[s]) break;     // a reader failed
                }
                rc = commit(static_cast<int>(s % static_cast<uint64_t>(n_slots)));
                std::lock_guard<std::mutex> lk(mx);
                committed = s + 1;
                if (rc != CLIPDB_OK) stop = true;
                cv.notify_all();
            }
            {
                std::lock_guard<std::mutex> lk(mx);
                if (rc != CLIPDB_OK) stop = true;
                cv.notify_all();
            }
            for (std::thread &t : threads) t.join();
            if (rc == CLIPDB_OK && !worker_err.empty()) rc = report(CLIPDB_ERR_INVALID, worker_err);
        }
    }
    {
        std::lock_guard<std::mutex> lk(c->mu);
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
    }
    if (out_joined_rows) *out_joined_rows = joined;
    return rc;
}

int64_t clipdb_num_rows(const clipdb_ctx *c) { return c ? c->n : 0; }
int32_t clipdb_dim(const clipdb_ctx *c) { return c ? c->dim : 0; }

int clipdb_set_mask(clipdb_ctx *c, const uint32_t *words, int64_t n_words) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->rows) return fail(c, CLIPDB_ERR_STATE, "set_mask: no rows loaded");
    const int64_t need = (c->n + 31) / 32;
    if (!words || n_words < need) return fail(c, CLIPDB_ERR_INVALID, "set_mask: need %lld words", (long long)need);
    DeviceGuard g(c->device);
    if (c->mask_words < need) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->mask) cudaFree(c->mask);
        c->mask = nullptr;
        c->mask_words = 0;
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->mask), static_cast<size_t>(need > 0 ? need : 1) * 4));
        c->mask_words = need;
    }
    CU_TRY(c, cudaMemcpyAsync(c->mask, words, static_cast<size_t>(need) * 4, cudaMemcpyDefault, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return CLIPDB_OK;
}

int clipdb_clear_mask(clipdb_ctx *c) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->mask) cudaFree(c->mask);
    c->mask = nullptr;
    c->mask_words = 0;
    return CLIPDB_OK;
}

int clipdb_blend_device(clipdb_ctx *c, const float *d_e1, const float *d_e2, const float *d_w,
                        const float *d_negs, const float *d_neg_w, int32_t n_neg, int32_t dim,
                        int32_t batch, float *d_out, int32_t *d_out_flags) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return blend_device_locked(c, d_e1, d_e2, d_w, d_negs, d_neg_w, n_neg, dim, batch, d_out, d_out_flags);
}

int clipdb_blend(clipdb_ctx *c, const float *e1, const float *e2, double w0, double w1,
                 const float *negs, const double *neg_w, int32_t n_neg, int32_t dim, float *out,
                 int32_t *out_flags) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!out) return fail(c, CLIPDB_ERR_INVALID, "blend: out is null");
    DeviceGuard g(c->device);
    BlendStaged st;
    RC_TRY(stage_blend(c, e1, e2, w0, w1, negs, neg_w, n_neg, dim, &st));
    RC_TRY(blend_device_locked(c, st.e1, st.e2, st.w, st.negs, st.neg_w, n_neg, dim, 1, st.out, st.flags));
    float *h = static_cast<float *>(c->pinned.p);
    // results land behind the staged inputs in the pinned buffer
    const size_t off = static_cast<size_t>(st.out - static_cast<float *>(c->d_blend_in.p));
    CU_TRY(c, cudaMemcpyAsync(h + off, st.out, dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    int32_t *hflags = reinterpret_cast<int32_t *>(h + off + dim);
    CU_TRY(c, cudaMemcpyAsync(hflags, st.flags, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    memcpy(out, h + off, dim * sizeof(float));
    if (out_flags) *out_flags = *hflags;
    return CLIPDB_OK;
}

int clipdb_search_device(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k, int32_t metric,
                         int32_t use_mask, int64_t *d_out_rowids, float *d_out_dist,
                         int32_t *d_out_n, int64_t *d_out_nan) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return search_device_locked(c, d_queries, nq, k, metric, use_mask, d_out_rowids, d_out_dist,
                                d_out_n, d_out_nan);
}

// Host entry points write their results into ONE device block laid out
//     nan int64[nq] | rowids int64[nq][k] | dist float32[nq][k] | n int32[nq]
// so that they reach the host with a single D2H copy (small stores are latency-bound: the four
// separate copies of the first version cost ~10 us of a 130 us query).
static int ensure_result_buffers(clipdb_ctx *c, int32_t nq, int64_t kcols) {
    const size_t q = static_cast<size_t>(nq > 0 ? nq : 1), kc = static_cast<size_t>(kcols > 0 ? kcols : 0);
    const size_t b_nan = q * sizeof(int64_t), b_ids = q * kc * sizeof(int64_t), b_dist = q * kc * sizeof(float);
    const size_t b_n = q * sizeof(int32_t);
    RC_TRY(ensure_device(c, c->d_results, b_nan + b_ids + b_dist + b_n + 64));
    uint8_t *d = static_cast<uint8_t *>(c->d_results.p);
    c->r_nan = reinterpret_cast<int64_t *>(d);
    c->r_ids = reinterpret_cast<int64_t *>(d + b_nan);
    c->r_dist = reinterpret_cast<float *>(d + b_nan + b_ids);
    c->r_n = reinterpret_cast<int32_t *>(d + b_nan + b_ids + b_dist);
    c->r_bytes = b_nan + b_ids + b_dist + b_n;
    return CLIPDB_OK;
}

// shared tail of the host search entry points: one D2H of the result block + unpack
static int fetch_results(clipdb_ctx *c, int32_t nq, int64_t kcols, int64_t *out_rowids,
                         float *out_dist, int32_t *out_n, int64_t *out_nan) {
    RC_TRY(ensure_pinned(c, c->r_bytes + 64));
    uint8_t *h = static_cast<uint8_t *>(c->pinned.p);
    CU_TRY(c, cudaMemcpyAsync(h, c->d_results.p, c->r_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const uint8_t *d = static_cast<const uint8_t *>(c->d_results.p);
    const uint8_t *h_nan = h;
    const uint8_t *h_ids = h + (reinterpret_cast<const uint8_t *>(c->r_ids) - d);
    const uint8_t *h_dist = h + (reinterpret_cast<const uint8_t *>(c->r_dist) - d);
    const uint8_t *h_n = h + (reinterpret_cast<const uint8_t *>(c->r_n) - d);
    memcpy(out_n, h_n, static_cast<size_t>(nq) * sizeof(int32_t));
    if (out_nan) memcpy(out_nan, h_nan, static_cast<size_t>(nq) * sizeof(int64_t));
    for (int32_t q = 0; q < nq; q++) {
        const int32_t m = reinterpret_cast<const int32_t *>(h_n)[q];
        if (m > 0) {
            memcpy(out_rowids + q * kcols, h_ids + static_cast<size_t>(q) * kcols * sizeof(int64_t),
                   static_cast<size_t>(m) * sizeof(int64_t));
            memcpy(out_dist + q * kcols, h_dist + static_cast<size_t>(q) * kcols * sizeof(float),
                   static_cast<size_t>(m) * sizeof(float));
        }
    }
    return CLIPDB_OK;
}

// Queries are in device memory, the result block is allocated: pick the path, run, fetch.
// Batched path (bf16 pre-selection + exact re-rank) when the batch store is enabled and the
// request qualifies; queries it flags (candidate overflow, zero norm) are re-run through the
// exact scan.  The flags travel to the host with the results: one synchronisation per call.
static int run_and_fetch(clipdb_ctx *c, const float *dq, int32_t nq, int32_t k, int32_t metric, int32_t use_mask,
                         int64_t *out_rowids, float *out_dist, int32_t *out_n, int64_t *out_nan) {
    const int64_t kcols = k > 0 ? k : 0;
    int64_t *o_ids = c->r_ids;
    float *o_dist = c->r_dist;
    int32_t *o_n = c->r_n;
    int64_t *o_nan = c->r_nan;
    if (nq >= c->batch_min_nq && c->batch_enabled && c->batch_dirty) RC_TRY(batch_build_locked(c));
    if (nq >= c->batch_min_nq && batch_eligible(c, nq < BQ_N ? nq : BQ_N, k, metric, use_mask)) {
        RC_TRY(ensure_device(c, c->bq_flags, static_cast<size_t>(nq > BQ_N ? nq : BQ_N) * sizeof(int32_t)));
        int32_t *d_flags = static_cast<int32_t *>(c->bq_flags.p);
        for (int32_t q0 = 0; q0 < nq; q0 += BQ_N) {
            const int32_t m = nq - q0 < BQ_N ? nq - q0 : BQ_N;
            RC_TRY(batch_search_device_locked(c, dq + static_cast<size_t>(q0) * c->dim, m, k, use_mask, o_ids + q0 * kcols,
                                              o_dist + q0 * kcols, o_n + q0, o_nan + q0, d_flags + q0));
        }
        const size_t fbytes = static_cast<size_t>(nq) * sizeof(int32_t);
        if (c->pinned_flags.bytes < fbytes) {
            if (c->pinned_flags.p) {
                CU_TRY(c, cudaStreamSynchronize(c->stream));
                CU_TRY(c, cudaFreeHost(c->pinned_flags.p));
                c->pinned_flags.p = nullptr;
                c->pinned_flags.bytes = 0;
            }
            const size_t want = fbytes < 4096 ? 4096 : fbytes;
            CU_TRY(c, cudaMallocHost(&c->pinned_flags.p, want));
            c->pinned_flags.bytes = want;
        }
        int32_t *flags = static_cast<int32_t *>(c->pinned_flags.p);
        CU_TRY(c, cudaMemcpyAsync(flags, d_flags, fbytes, cudaMemcpyDeviceToHost, c->stream));
        RC_TRY(fetch_results(c, nq, kcols, out_rowids, out_dist, out_n, out_nan));   // synchronises
        bool rerun = false;
        for (int32_t q = 0; q < nq; q++) {
            if (!flags[q]) continue;
            rerun = true;
            RC_TRY(search_device_locked(c, dq + static_cast<size_t>(q) * c->dim, 1, k, metric, use_mask,
                                        o_ids + q * kcols, o_dist + q * kcols, o_n + q, o_nan + q));
        }
        if (rerun) return fetch_results(c, nq, kcols, out_rowids, out_dist, out_n, out_nan);
        return CLIPDB_OK;
    }
    RC_TRY(search_device_locked(c, dq, nq, k, metric, use_mask, o_ids, o_dist, o_n, o_nan));
    return fetch_results(c, nq, kcols, out_rowids, out_dist, out_n, out_nan);
}

int clipdb_search(clipdb_ctx *c, const float *queries, int32_t nq, int32_t k, int32_t metric,
                  int32_t use_mask, int64_t *out_rowids, float *out_dist, int32_t *out_n,
                  int64_t *out_nan) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->rows || c->dim == 0) return fail(c, CLIPDB_ERR_STATE, "no rows loaded");
    if (nq < 0 || !queries || !out_n || (k > 0 && (!out_rowids || !out_dist)))
        return fail(c, CLIPDB_ERR_INVALID, "search: null pointer or negative nq");
    if (nq == 0) return CLIPDB_OK;
    DeviceGuard g(c->device);
    // effective columns: the caller's arrays are nq x k, ours too
    const int64_t kcols = k > 0 ? k : 0;
    const size_t qbytes = static_cast<size_t>(nq) * c->dim * sizeof(float);
    RC_TRY(ensure_pinned(c, qbytes));
    RC_TRY(ensure_device(c, c->d_query, qbytes));
    RC_TRY(ensure_result_buffers(c, nq, kcols));
    memcpy(c->pinned.p, queries, qbytes);
    CU_TRY(c, cudaMemcpyAsync(c->d_query.p, c->pinned.p, qbytes, cudaMemcpyHostToDevice, c->stream));
    // the pinned buffer is reused for results: the H2D above is ordered before them on the stream
    const float *dq = static_cast<const float *>(c->d_query.p);
    return run_and_fetch(c, dq, nq, k, metric, use_mask, out_rowids, out_dist, out_n, out_nan);
}

int clipdb_blend_search(clipdb_ctx *c, const float *e1, const float *e2, double w0, double w1,
                        const float *negs, const double *neg_w, int32_t n_neg, int32_t k,
                        int32_t metric, int32_t use_mask, int64_t *out_rowids, float *out_dist,
                        int32_t *out_n, int64_t *out_nan, float *out_query, int32_t *out_flags) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->rows || c->dim == 0) return fail(c, CLIPDB_ERR_STATE, "no rows loaded");
    if (!out_n || (k > 0 && (!out_rowids || !out_dist)))
        return fail(c, CLIPDB_ERR_INVALID, "blend_search: null output pointer");
    DeviceGuard g(c->device);
    const int32_t dim = c->dim;
    const int64_t kcols = k > 0 ? k : 0;
    BlendStaged st;
    RC_TRY(stage_blend(c, e1, e2, w0, w1, negs, neg_w, n_neg, dim, &st));
    RC_TRY(ensure_result_buffers(c, 1, kcols));
    RC_TRY(ensure_device(c, c->d_query, static_cast<size_t>(dim) * sizeof(float) + sizeof(int32_t)));
    float *d_q = static_cast<float *>(c->d_query.p);
    RC_TRY(blend_device_locked(c, st.e1, st.e2, st.w, st.negs, st.neg_w, n_neg, dim, 1, d_q, st.flags));
    // optional read-back of the blended query and the fallback flags, through their own
    // pinned area so the result staging below can reuse the main one
    float *h_aux = nullptr;
    if (out_query || out_flags) {
        const size_t aux_bytes = static_cast<size_t>(dim) * sizeof(float) + sizeof(int32_t);
        if (c->pinned_aux.bytes < aux_bytes) {
            if (c->pinned_aux.p) {
                CU_TRY(c, cudaStreamSynchronize(c->stream));
                CU_TRY(c, cudaFreeHost(c->pinned_aux.p));
                c->pinned_aux.p = nullptr;
                c->pinned_aux.bytes = 0;
            }
            CU_TRY(c, cudaMallocHost(&c->pinned_aux.p, aux_bytes));
            c->pinned_aux.bytes = aux_bytes;
        }
        h_aux = static_cast<float *>(c->pinned_aux.p);
        CU_TRY(c, cudaMemcpyAsync(h_aux, d_q, dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaMemcpyAsync(h_aux + dim, st.flags, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    // scan (or, with the batch store and batch_min_nq = 1, bf16 pre-selection + re-rank); syncs the stream
    RC_TRY(run_and_fetch(c, d_q, 1, k, metric, use_mask, out_rowids, out_dist, out_n, out_nan));
    if (out_query) memcpy(out_query, h_aux, dim * sizeof(float));
    if (out_flags) memcpy(out_flags, h_aux + dim, sizeof(int32_t));
    return CLIPDB_OK;
}

int clipdb_enable_batch(clipdb_ctx *c, int32_t enable) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (!enable) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        c->batch_enabled = false;
        c->bf16_built = c->bf16_cap_rows = 0;
        Buffer *bufs[] = {&c->bf16_rows, &c->bq_scores, &c->bq_cand, &c->bq_cand_u, &c->bq_surv};
        for (Buffer *b : bufs) free_buffer(*b);
        return CLIPDB_OK;
    }
    return batch_build_locked(c);
}

int clipdb_search_batch_device(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k, int32_t use_mask,
                               int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n,
                               int64_t *d_out_nan, int32_t *d_flags) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (use_mask && !c->mask) return fail(c, CLIPDB_ERR_STATE, "search_batch: use_mask set but no mask installed");
    return batch_search_device_locked(c, d_queries, nq, k, use_mask, d_out_rowids, d_out_dist, d_out_n, d_out_nan,
                                      d_flags);
}

int clipdb_batch_stats(clipdb_ctx *c, uint32_t *cand_counts, uint32_t *surv_counts) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->batch_enabled || !c->bq_count.p || !c->bq_surv_count.p)
        return fail(c, CLIPDB_ERR_STATE, "batch_stats: batch store not enabled");
    DeviceGuard g(c->device);
    if (cand_counts)
        CU_TRY(c, cudaMemcpyAsync(cand_counts, c->bq_count.p, BQ_N * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (surv_counts)
        CU_TRY(c, cudaMemcpyAsync(surv_counts, c->bq_surv_count.p, BQ_N * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return CLIPDB_OK;
}

int clipdb_load_codes(clipdb_ctx *c, const uint8_t *codes, const int64_t *ids, int64_t n, int32_t dim) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n < 0 || (n > 0 && !codes)) return fail(c, CLIPDB_ERR_INVALID, "load_codes: bad argument");
    if (dim != SCAN_DIM) return fail(c, CLIPDB_ERR_UNSUPPORTED, "load_codes: sign codes are %d elements wide", SCAN_DIM);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    release_codes(c);
    if (n == 0) return CLIPDB_OK;
    CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->codes), static_cast<size_t>(n) * BIN_ROW_BYTES));
    if (ids) {
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->code_ids), static_cast<size_t>(n) * sizeof(int64_t)));
        CU_TRY(c, cudaMemcpyAsync(c->code_ids, ids, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyDefault, c->stream));
    }
    // bytes -> bits on the device, through a bounded staging buffer
    const int64_t chunk_rows = 32768;   // 37.7 MB of bytes per chunk
    RC_TRY(ensure_device(c, c->code_stage, static_cast<size_t>(chunk_rows < n ? chunk_rows : n) * SCAN_DIM));
    RC_TRY(ensure_device(c, c->code_bad, sizeof(unsigned int)));
    CU_TRY(c, cudaMemsetAsync(c->code_bad.p, 0, sizeof(unsigned int), c->stream));
    for (int64_t lo = 0; lo < n; lo += chunk_rows) {
        const int64_t m = n - lo < chunk_rows ? n - lo : chunk_rows;
        CU_TRY(c, cudaMemcpyAsync(c->code_stage.p, codes + lo * SCAN_DIM, static_cast<size_t>(m) * SCAN_DIM,
                                  cudaMemcpyDefault, c->stream));
        const long long words = m * BIN_WORDS;
        pack_codes_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, c->stream>>>(
            static_cast<const uint8_t *>(c->code_stage.p), words, c->codes + lo * BIN_WORDS,
            static_cast<unsigned int *>(c->code_bad.p));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        CU_TRY(c, cudaStreamSynchronize(c->stream));   // the staging buffer (and a pageable source) is reused
    }
    unsigned int bad = 0;
    CU_TRY(c, cudaMemcpy(&bad, c->code_bad.p, sizeof bad, cudaMemcpyDeviceToHost));
    if (bad) {
        release_codes(c);
        return fail(c, CLIPDB_ERR_INVALID, "load_codes: %u words hold bytes other than 0/1 (not sign codes)", bad);
    }
    c->n_codes = n;
    return CLIPDB_OK;
}

int64_t clipdb_num_codes(const clipdb_ctx *c) { return c ? c->n_codes : 0; }

int clipdb_set_code_mask(clipdb_ctx *c, const uint32_t *words, int64_t n_words, const uint32_t *order) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->codes) return fail(c, CLIPDB_ERR_STATE, "set_code_mask: no codes loaded");
    const int64_t need = (c->n_codes + 31) / 32;
    if (!words || n_words < need) return fail(c, CLIPDB_ERR_INVALID, "set_code_mask: need %lld words", (long long)need);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->code_mask_words < need) {
        if (c->code_mask) cudaFree(c->code_mask);
        c->code_mask = nullptr;
        c->code_mask_words = 0;
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->code_mask), static_cast<size_t>(need) * 4));
        c->code_mask_words = need;
    }
    CU_TRY(c, cudaMemcpyAsync(c->code_mask, words, static_cast<size_t>(need) * 4, cudaMemcpyDefault, c->stream));
    if (c->code_seq) cudaFree(c->code_seq);
    if (c->code_ids_by_seq) cudaFree(c->code_ids_by_seq);
    c->code_seq = nullptr;
    c->code_ids_by_seq = nullptr;
    if (order) {
        const size_t n = static_cast<size_t>(c->n_codes);
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->code_seq), n * sizeof(uint32_t)));
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->code_ids_by_seq), n * sizeof(int64_t)));
        CU_TRY(c, cudaMemcpyAsync(c->code_seq, order, n * sizeof(uint32_t), cudaMemcpyDefault, c->stream));
        CU_TRY(c, cudaMemsetAsync(c->code_ids_by_seq, 0xFF, n * sizeof(int64_t), c->stream));
        RC_TRY(ensure_device(c, c->code_bad, sizeof(unsigned int)));
        CU_TRY(c, cudaMemsetAsync(c->code_bad.p, 0, sizeof(unsigned int), c->stream));
        ids_by_sequence_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(
            c->code_seq, c->code_ids, c->n_codes, c->code_ids_by_seq, static_cast<unsigned int *>(c->code_bad.p));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        unsigned int bad = 0;
        CU_TRY(c, cudaMemcpyAsync(&bad, c->code_bad.p, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (bad) {
            cudaFree(c->code_seq);
            cudaFree(c->code_ids_by_seq);
            c->code_seq = nullptr;
            c->code_ids_by_seq = nullptr;
            return fail(c, CLIPDB_ERR_INVALID, "set_code_mask: order holds %u values >= the number of codes", bad);
        }
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return CLIPDB_OK;
}

int clipdb_clear_code_mask(clipdb_ctx *c) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->code_mask) cudaFree(c->code_mask);
    if (c->code_seq) cudaFree(c->code_seq);
    if (c->code_ids_by_seq) cudaFree(c->code_ids_by_seq);
    c->code_mask = nullptr;
    c->code_seq = nullptr;
    c->code_ids_by_seq = nullptr;
    c->code_mask_words = 0;
    return CLIPDB_OK;
}

int clipdb_binary_search_device(clipdb_ctx *c, const uint32_t *d_query_words, int32_t k, int32_t score_mode,
                                int32_t use_mask, int64_t *d_out_ids, int32_t *d_out_scores, int32_t *d_out_n) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return binary_search_device_locked(c, d_query_words, k, score_mode, use_mask, d_out_ids, d_out_scores, d_out_n);
}

int clipdb_binary_search(clipdb_ctx *c, const uint8_t *query_code, int32_t k, int32_t score_mode, int32_t use_mask,
                         int64_t *out_ids, int32_t *out_scores, int32_t *out_n) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!query_code || !out_n || (k > 0 && (!out_ids || !out_scores)))
        return fail(c, CLIPDB_ERR_INVALID, "binary_search: null pointer");
    DeviceGuard g(c->device);
    const int64_t kcols = k > 0 ? k : 0;
    RC_TRY(ensure_pinned(c, BIN_ROW_BYTES));
    RC_TRY(ensure_device(c, c->d_code_query, BIN_ROW_BYTES));
    RC_TRY(ensure_result_buffers(c, 1, kcols));
    uint32_t *hq = static_cast<uint32_t *>(c->pinned.p);
    for (int w = 0; w < BIN_WORDS; w++) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; b++) {
            const uint8_t v = query_code[w * 32 + b];
            if (v > 1) return fail(c, CLIPDB_ERR_INVALID, "binary_search: query code bytes must be 0 or 1");
            bits |= static_cast<uint32_t>(v) << b;
        }
        hq[w] = bits;
    }
    CU_TRY(c, cudaMemcpyAsync(c->d_code_query.p, hq, BIN_ROW_BYTES, cudaMemcpyHostToDevice, c->stream));
    RC_TRY(binary_search_device_locked(c, static_cast<const uint32_t *>(c->d_code_query.p), k, score_mode, use_mask,
                                       c->r_ids, reinterpret_cast<int32_t *>(c->r_dist), c->r_n));
    CU_TRY(c, cudaMemsetAsync(c->r_nan, 0, sizeof(int64_t), c->stream));
    return fetch_results(c, 1, kcols, out_ids, reinterpret_cast<float *>(out_scores), out_n, nullptr);
}

// ---- fused shard exchange -----------------------------------------------------------------

int clipdb_exchange_init(clipdb_ctx *c, int32_t world, int32_t rank, void *out_ipc_handle, void **out_inbox) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world)
        return fail(c, CLIPDB_ERR_INVALID, "exchange_init: world must be 1..%d and 0 <= rank < world", XCHG_MAX_WORLD);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    release_exchange(c);
    const size_t bytes = exchange_inbox_slots(world) * sizeof(ExchangeSlot);   // single-query + batched slots
    CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&c->xchg_inbox), bytes));
    CU_TRY(c, cudaMemset(c->xchg_inbox, 0, bytes));
    c->xchg_world = world;
    c->xchg_rank = rank;
    c->xchg_epoch = 0;
    c->xchg_batch_epoch = 0;
    c->xchg_peer[rank] = c->xchg_inbox;
    c->xchg_connected = world == 1;
    CU_TRY(c, cudaHostAlloc(reinterpret_cast<void **>(&c->xchg_abort_host), sizeof(int), cudaHostAllocMapped));
    *c->xchg_abort_host = 0;
    CU_TRY(c, cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->xchg_abort_dev), c->xchg_abort_host, 0));
    if (out_ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == CLIPDB_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        CU_TRY(c, cudaIpcGetMemHandle(&h, c->xchg_inbox));
        memcpy(out_ipc_handle, &h, sizeof h);
    }
    if (out_inbox) *out_inbox = c->xchg_inbox;
    return CLIPDB_OK;
}

int clipdb_exchange_connect(clipdb_ctx *c, const void *ipc_handles) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg_inbox) return fail(c, CLIPDB_ERR_STATE, "exchange_connect: call exchange_init first");
    if (!ipc_handles) return fail(c, CLIPDB_ERR_INVALID, "exchange_connect: null handles");
    DeviceGuard g(c->device);
    for (int r = 0; r < c->xchg_world; r++) {
        if (r == c->xchg_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t *>(ipc_handles) + static_cast<size_t>(r) * sizeof h, sizeof h);
        void *p = nullptr;
        CU_TRY(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->xchg_peer[r] = static_cast<ExchangeSlot *>(p);
        c->xchg_ipc[r] = true;
    }
    c->xchg_connected = true;
    return CLIPDB_OK;
}

int clipdb_exchange_connect_pointers(clipdb_ctx *c, void *const *inboxes, const int32_t *devices) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg_inbox) return fail(c, CLIPDB_ERR_STATE, "exchange_connect: call exchange_init first");
    if (!inboxes) return fail(c, CLIPDB_ERR_INVALID, "exchange_connect: null pointers");
    DeviceGuard g(c->device);
    for (int r = 0; r < c->xchg_world; r++) {
        if (r == c->xchg_rank) continue;
        if (!inboxes[r]) return fail(c, CLIPDB_ERR_INVALID, "exchange_connect: inbox %d is null", r);
        if (devices && devices[r] != c->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(c, CLIPDB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", devices[r],
                            cudaGetErrorString(e));
            cudaGetLastError();
        }
        c->xchg_peer[r] = static_cast<ExchangeSlot *>(inboxes[r]);
        c->xchg_ipc[r] = false;
    }
    c->xchg_connected = true;
    return CLIPDB_OK;
}

// The epoch after `cur`: never 0 ("slot never written"), and consecutive epochs always differ in parity (the
// two inbox banks alternate), so 0xFFFFFFFF is followed by 2, not by 1.
static uint32_t next_epoch(uint32_t cur) {
    const uint32_t e = cur + 1;
    return e == 0 ? 2 : e;
}

static void fill_exchange_args(const clipdb_ctx *c, ExchangeArgs &xa, int32_t k, uint32_t epoch, bool with_stats) {
    for (int r = 0; r < c->xchg_world; r++) xa.inbox[r] = c->xchg_peer[r];
    xa.world = c->xchg_world;
    xa.rank = c->xchg_rank;
    xa.k = k;
    xa.epoch = epoch;
    xa.timeout_ns = static_cast<unsigned long long>(c->xchg_timeout_ms) * 1000000ull;
    xa.abort_flag = c->xchg_abort_dev;
    xa.stats = with_stats && c->xchg_stats_on ? static_cast<unsigned long long *>(c->xchg_stats.p) : nullptr;
}

int clipdb_exchange_abort(clipdb_ctx *c, int32_t abort) {
    // no lock: meant to be called while another thread sits in a stream synchronisation
    if (!c || !c->xchg_abort_host) return CLIPDB_ERR_INVALID;
    *static_cast<volatile int *>(c->xchg_abort_host) = abort != 0;
    return CLIPDB_OK;
}

int clipdb_exchange_set_epoch(clipdb_ctx *c, uint32_t single_epoch, uint32_t batch_epoch) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg_inbox) return fail(c, CLIPDB_ERR_STATE, "exchange_set_epoch: call exchange_init first");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->xchg_epoch = single_epoch;
    c->xchg_batch_epoch = batch_epoch;
    return CLIPDB_OK;
}

int clipdb_exchange_stats(clipdb_ctx *c, int32_t enable, int32_t reset, double *out_ns, int64_t *out_launches) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    const bool fresh = c->xchg_stats.p == nullptr;
    RC_TRY(ensure_device(c, c->xchg_stats, XCHG_STATS_WORDS * sizeof(unsigned long long)));
    if (fresh) CU_TRY(c, cudaMemsetAsync(c->xchg_stats.p, 0, XCHG_STATS_WORDS * sizeof(unsigned long long), c->stream));
    unsigned long long h[XCHG_STATS_WORDS] = {};
    CU_TRY(c, cudaMemcpyAsync(h, c->xchg_stats.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    if (reset) CU_TRY(c, cudaMemsetAsync(c->xchg_stats.p, 0, sizeof h, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (out_ns)
        for (int i = 0; i < XCHG_STATS_WORDS - 1; i++) out_ns[i] = static_cast<double>(h[i]);
    if (out_launches) *out_launches = static_cast<int64_t>(h[XCHG_STATS_WORDS - 1]);
    c->xchg_stats_on = enable != 0;
    return CLIPDB_OK;
}

int clipdb_search_sharded_device(clipdb_ctx *c, const float *d_query, int32_t k, int32_t metric, int32_t use_mask,
                                 int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg_connected) return fail(c, CLIPDB_ERR_STATE, "sharded search: exchange not connected");
    if (!c->rows || c->dim == 0) return fail(c, CLIPDB_ERR_STATE, "no rows loaded");
    if (!d_query || !d_out_rowids || !d_out_dist || !d_out_n) return fail(c, CLIPDB_ERR_INVALID, "sharded search: null pointer");
    if (metric != CLIPDB_METRIC_COSINE && metric != CLIPDB_METRIC_L2)
        return fail(c, CLIPDB_ERR_INVALID, "search: unknown metric %d", metric);
    if (use_mask && !c->mask) return fail(c, CLIPDB_ERR_STATE, "search: use_mask set but no mask installed");
    DeviceGuard g(c->device);
    ExchangeArgs xa{};
    const uint32_t epoch = next_epoch(c->xchg_epoch);
    fill_exchange_args(c, xa, k, epoch, true);
    if (xa.world == 1) return search_one(c, d_query, k, metric, use_mask != 0, d_out_rowids, d_out_dist, d_out_n, d_out_nan);
    RC_TRY(search_one(c, d_query, k, metric, use_mask != 0, d_out_rowids, d_out_dist, d_out_n, d_out_nan, &xa));
    c->xchg_epoch = epoch;   // only a launch that was enqueued consumes an epoch: the ranks stay in step
    return CLIPDB_OK;
}

int clipdb_search_batch_sharded_device(clipdb_ctx *c, const float *d_queries, int32_t nq, int32_t k, int32_t use_mask,
                                       int64_t *d_out_rowids, float *d_out_dist, int32_t *d_out_n, int64_t *d_out_nan,
                                       int32_t *d_flags) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg_connected) return fail(c, CLIPDB_ERR_STATE, "sharded batch search: exchange not connected");
    if (use_mask && !c->mask) return fail(c, CLIPDB_ERR_STATE, "search_batch: use_mask set but no mask installed");
    DeviceGuard g(c->device);
    if (c->xchg_world == 1)
        return batch_search_device_locked(c, d_queries, nq, k, use_mask, d_out_rowids, d_out_dist, d_out_n, d_out_nan, d_flags);
    ExchangeArgs xa{};
    const uint32_t epoch = next_epoch(c->xchg_batch_epoch);
    fill_exchange_args(c, xa, k, epoch, false);
    RC_TRY(batch_search_device_locked(c, d_queries, nq, k, use_mask, d_out_rowids, d_out_dist, d_out_n, d_out_nan,
                                      d_flags, &xa));
    c->xchg_batch_epoch = epoch;
    return CLIPDB_OK;
}

int clipdb_merge_batch_device(clipdb_ctx *c, const void *d_dist, int64_t dist_stride, int64_t dist_qstride,
                              const void *d_rowids, int64_t rowid_stride, int64_t rowid_qstride,
                              const void *d_counts, int64_t count_stride, int64_t count_qstride, int32_t lists,
                              int32_t nq, int32_t k, float *d_out_dist, int64_t *d_out_rowids, int32_t *d_out_n) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!d_dist || !d_rowids || !d_counts || !d_out_dist || !d_out_rowids || !d_out_n || lists <= 0 || k < 0 || nq < 0)
        return fail(c, CLIPDB_ERR_INVALID, "merge: bad argument");
    if (dist_stride % 4 || rowid_stride % 8 || count_stride % 4 || dist_qstride % 4 || rowid_qstride % 8 ||
        count_qstride % 4 || (reinterpret_cast<uintptr_t>(d_rowids) & 7) ||
        (reinterpret_cast<uintptr_t>(d_dist) & 3) || (reinterpret_cast<uintptr_t>(d_counts) & 3))
        return fail(c, CLIPDB_ERR_INVALID, "merge: misaligned list layout");
    DeviceGuard g(c->device);
    if (nq == 0) return CLIPDB_OK;
    if (k == 0) {
        CU_TRY(c, cudaMemsetAsync(d_out_n, 0, static_cast<size_t>(nq) * sizeof(int32_t), c->stream));
        return CLIPDB_OK;
    }
    const int64_t total = static_cast<int64_t>(lists) * k;
    if (total > MERGE_SHARD_MAX_KEYS)
        return fail(c, CLIPDB_ERR_UNSUPPORTED, "merge: lists*k = %lld exceeds %d", (long long)total,
                    MERGE_SHARD_MAX_KEYS);
    const size_t smem = static_cast<size_t>(next_pow2(static_cast<int>(total))) * sizeof(uint64_t);
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        if (configured_device != c->device) {
            CU_TRY(c, cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           MERGE_SHARD_MAX_KEYS * static_cast<int>(sizeof(uint64_t))));
            configured_device = c->device;
        }
    }
    ShardLists in{};
    in.dist = static_cast<const uint8_t *>(d_dist);
    in.rowids = static_cast<const uint8_t *>(d_rowids);
    in.counts = static_cast<const uint8_t *>(d_counts);
    in.dist_stride = dist_stride;
    in.rowid_stride = rowid_stride;
    in.count_stride = count_stride;
    in.dist_qstride = dist_qstride;
    in.rowid_qstride = rowid_qstride;
    in.count_qstride = count_qstride;
    merge_shards_kernel<<<nq, MERGE_THREADS, smem, c->stream>>>(in, lists, k, d_out_dist, d_out_rowids, d_out_n);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return CLIPDB_OK;
}

int clipdb_merge_strided_device(clipdb_ctx *c, const void *d_dist, int64_t dist_stride,
                                const void *d_rowids, int64_t rowid_stride, const void *d_counts,
                                int64_t count_stride, int32_t lists, int32_t k, float *d_out_dist,
                                int64_t *d_out_rowids, int32_t *d_out_n) {
    return clipdb_merge_batch_device(c, d_dist, dist_stride, 0, d_rowids, rowid_stride, 0, d_counts, count_stride, 0,
                                     lists, 1, k, d_out_dist, d_out_rowids, d_out_n);
}

int clipdb_merge_device(clipdb_ctx *c, const float *d_dist, const int64_t *d_rowids,
                        const int32_t *d_counts, int32_t lists, int32_t k, float *d_out_dist,
                        int64_t *d_out_rowids, int32_t *d_out_n) {
    return clipdb_merge_strided_device(c, d_dist, static_cast<int64_t>(k) * sizeof(float), d_rowids,
                                       static_cast<int64_t>(k) * sizeof(int64_t), d_counts,
                                       sizeof(int32_t), lists, k, d_out_dist, d_out_rowids, d_out_n);
}

int clipdb_profile(clipdb_ctx *c, int32_t enable) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    RC_TRY(profile_fold(c));
    c->profiling = enable != 0;
    c->prof_ms = 0.0;
    c->prof_scans = 0;
    return CLIPDB_OK;
}

int clipdb_profile_read(clipdb_ctx *c, double *scan_ms_total, int64_t *scans) {
    if (!c || !scan_ms_total || !scans) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    RC_TRY(profile_fold(c));
    *scan_ms_total = c->prof_ms;
    *scans = c->prof_scans;
    return CLIPDB_OK;
}

int clipdb_profile_clock(clipdb_ctx *c, double *sm_cycles, double *nanoseconds, int64_t *launches) {
    if (!c) return CLIPDB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    unsigned long long h[3] = {0, 0, 0};
    if (c->bq_clock.p) {
        CU_TRY(c, cudaMemcpyAsync(h, c->bq_clock.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaMemsetAsync(c->bq_clock.p, 0, sizeof h, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
    if (sm_cycles) *sm_cycles = static_cast<double>(h[0]);
    if (nanoseconds) *nanoseconds = static_cast<double>(h[1]);
    if (launches) *launches = static_cast<int64_t>(h[2]);
    return CLIPDB_OK;
}

}  // extern "C"
