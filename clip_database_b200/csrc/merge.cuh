// merge.cuh — reduce per-CTA candidate lists to the final top-k and decode it.
//
// Integer work only: the keys (common.cuh) already encode SQLite's
// (distance, scan sequence) order (image_database.py:1572-1573), so merging is
// a sort of 64-bit keys and the result is bit-exact whatever the list layout.
#pragma once

#include "common.cuh"

namespace clipdb {

constexpr int MERGE_KEYS = 2048;   // keys one CTA sorts in shared memory (16 KB)
constexpr int MERGE_THREADS = 1024;

struct DecodeArgs {
    const int64_t *rowids;               // nullable: rowid = rowid_base + position
    int64_t rowid_base;
    int64_t *out_rowids;                 // [k]
    float *out_dist;                     // [k]
    int32_t *out_n;                      // [1]
    int64_t *out_nan;                    // nullable [1]
    const unsigned long long *nan_rows;  // counter the scan accumulated
    int k;
    int score_max;                       // > 0: keys carry (score_max - integer score) instead of a float32
                                         // distance (binary.cuh); out_dist then receives int32 scores
};

__device__ __forceinline__ void decode_one(const DecodeArgs &d, int i, uint64_t key) {
    const uint32_t pos = static_cast<uint32_t>(key & 0xFFFFFFFFull);
    if (d.score_max > 0)
        reinterpret_cast<int32_t *>(d.out_dist)[i] = d.score_max - static_cast<int32_t>(key >> 32);
    else
        d.out_dist[i] = orderable_f32(static_cast<uint32_t>(key >> 32));
    d.out_rowids[i] = d.rowids ? d.rowids[pos] : d.rowid_base + static_cast<int64_t>(pos);
}

// ---- fused tail: the CTA that finishes last merges every CTA's list and decodes the answer ----
// Each CTA has written one ascending list of L keys.  The last one to finish (a ticket from a
// global counter, after a __threadfence so the lists are visible) pulls all lists into shared
// memory and runs a tournament: per round, lists p and p+half are combined with the bitonic
// half-cleaner C[i] = min(A[i], B[L-1-i]) — the L smallest of the union, as a bitonic sequence —
// followed by log2(L) compare-exchange steps.  ceil(log2(lists)) rounds of 1 + log2(L) steps:
// ~50 block-wide steps for 148 lists of 32 keys (a few microseconds) instead of two more kernel
// launches that each sort 2048 keys from scratch, so one query is ONE launch.
__device__ __forceinline__ bool last_cta_done(unsigned int *done_counter, int tid) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    return s_last != 0;
}

template <int L>
__device__ __forceinline__ void merge_sorted_lists_tournament(uint64_t *m, int cnt, int tid, int nthreads) {
    while (cnt > 1) {
        const int half = (cnt + 1) >> 1, pairs = cnt >> 1;
        for (int w = tid; w < pairs * L; w += nthreads) {
            const int p = w / L, i = w - p * L;
            const uint64_t b = m[(p + half) * L + (L - 1 - i)];
            if (b < m[p * L + i]) m[p * L + i] = b;
        }
        __syncthreads();
#pragma unroll 1
        for (int s = L / 2; s >= 1; s >>= 1) {
            for (int w = tid; w < pairs * (L / 2); w += nthreads) {
                const int p = w / (L / 2), j = w - p * (L / 2);
                const int lo = p * L + 2 * j - (j & (s - 1)), hi = lo + s;
                const uint64_t a = m[lo], b = m[hi];
                if (a > b) {
                    m[lo] = b;
                    m[hi] = a;
                }
            }
            __syncthreads();
        }
        cnt = half;
    }
}

// Counters shared by the CTAs of one scan launch.  All zero between launches: the last CTA
// resets them, so a query needs no memset nodes.
struct ScanSync {
    unsigned int tile_counter;        // dynamic tile scheduler
    unsigned int done_counter;        // CTAs finished
    unsigned long long nan_rows;      // admitted rows whose distance was NaN
    unsigned long long t_start;       // %globaltimer of the first CTA to start (exchange statistics only)
};

// ---- fused shard exchange (multi-GPU) --------------------------------------------
// Row-sharded search needs one exchange step: every GPU's k best (distance, rowid) pairs must
// reach every GPU and be merged.  Instead of ending the kernel, calling an NCCL all-gather and
// launching a merge, the scan kernel's last CTA does it itself over NVLink peer memory:
//   1. it decodes its shard's merged list and STORES it straight into slot [my rank] of every
//      peer's inbox (the inboxes are cudaMalloc'ed buffers mapped into every process with CUDA
//      IPC; 8 x 252 B at k = 20),
//   2. __threadfence_system, then st.release.sys of the query's epoch number into each slot,
//   3. ld.acquire.sys-spins on the G slots of its OWN inbox until they carry this epoch
//      (bounded by a timeout: a missing peer becomes an error code, never a hung GPU),
//   4. sorts the G x k keys (distance, shard, position) = (distance, rowid) and decodes.
// So a sharded query is ONE launch per GPU with no host round trip and no collective call.
// Two inbox banks alternate by epoch parity: a peer can be at most one query ahead (it needs
// this rank's record of query e to finish e), so bank (e+1)&1 is never read while written.
constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_K = 128;   // largest k of the fused path

constexpr int XCHG_BATCH = 256;   // queries per batched pass

struct ExchangeSlot {
    uint32_t epoch;
    int32_t count;      // valid entries | requested k << 16
    long long nan;
    int32_t flags;      // batched path: this shard could not answer the query (overflow / bad query)
    int32_t pad;
    float dist[XCHG_K];
    long long rowid[XCHG_K];
};

// One rank's inbox = [2 banks][world] slots for single queries, followed by
// [2 banks][world][XCHG_BATCH] slots for batched passes.  Slot of (bank, sender r[, query q]):
//   single   bank * world + r
//   batched  2 * world + (bank * world + r) * XCHG_BATCH + q
struct ExchangeArgs {
    ExchangeSlot *inbox[XCHG_MAX_WORLD];   // inbox[r] = rank r's inbox, as mapped here
    int world;                             // 0: no exchange
    int rank;
    int k;                                 // requested (global) k
    uint32_t epoch;                        // same on every rank for the same query / batch, never 0
    unsigned long long timeout_ns;
    const volatile int *abort_flag;        // nullable; mapped host memory: != 0 = stop waiting for the peers now
    unsigned long long *stats;             // nullable; [0..4] summed segment durations (ns), [5] launches
};

constexpr int XCHG_STATS_WORDS = 6;

__host__ __device__ inline size_t exchange_inbox_slots(int world) {
    return static_cast<size_t>(2) * world * (1 + XCHG_BATCH);
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// scratch[0 .. L) holds this shard's merged ascending list for one query.  `slot0` / `slot_stride`
// address the slots of this query: sender r's slot is slot0 + r * slot_stride in every inbox.
// `nan_local` / `flags_local`: this shard's NaN-row count and failure flags for the query.
// On return every thread knows the outcome: >= 0 results written, -1 a peer never delivered,
// -2 the ranks asked different questions.  `out_flags` (nullable) receives the OR of all shards' flags.
template <int L>
__device__ __forceinline__ int exchange_merge_decode(uint64_t *scratch, const DecodeArgs &dec, const ExchangeArgs &xa,
                                                     int slot0, int slot_stride, long long nan_local,
                                                     int flags_local, int32_t *out_flags, int tid, int nthreads,
                                                     unsigned long long *stamps = nullptr) {
    const int k = xa.k, G = xa.world;
    int mine = 0;
    for (int base = 0; base < k; base += nthreads) {
        const int i = base + tid;
        mine += __syncthreads_count(i < k && i < L && scratch[i] != KEY_EMPTY);
    }
    // 1. this shard's record into its slot of every inbox (peer stores go over NVLink)
    for (int w = tid; w < G * k; w += nthreads) {
        const int r = w / k, j = w - r * k;
        if (j < mine) {
            ExchangeSlot *slot = xa.inbox[r] + slot0 + xa.rank * slot_stride;
            const uint64_t key = scratch[j];
            const uint32_t pos = static_cast<uint32_t>(key & 0xFFFFFFFFull);
            slot->dist[j] = orderable_f32(static_cast<uint32_t>(key >> 32));
            slot->rowid[j] = dec.rowids ? dec.rowids[pos] : dec.rowid_base + static_cast<int64_t>(pos);
        }
    }
    if (tid < G) {
        ExchangeSlot *slot = xa.inbox[tid] + slot0 + xa.rank * slot_stride;
        slot->count = mine | (k << 16);   // the requested k rides along: every rank must ask the same question
        slot->nan = nan_local;
        slot->flags = flags_local;
    }
    // 2. publish
    __threadfence_system();
    __syncthreads();
    if (tid < G) st_release_sys(&(xa.inbox[tid] + slot0 + xa.rank * slot_stride)->epoch, xa.epoch);
    if (stamps && tid == 0) stamps[0] = global_timer_ns();   // published
    // 3. wait for every shard's record in the own inbox
    const ExchangeSlot *own = xa.inbox[xa.rank] + slot0;
    bool late = false;
    if (tid < G) {
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(&own[tid * slot_stride].epoch) != xa.epoch) {
            // the abort flag lives in host memory (a PCIe round trip): looked at every 32nd spin only
            if (global_timer_ns() - t0 > xa.timeout_ns || ((++spins & 31u) == 0 && xa.abort_flag && *xa.abort_flag)) {
                late = true;
                break;
            }
            __nanosleep(200);
        }
    }
    if (__syncthreads_count(late)) {   // a peer never delivered (or the host gave up): an error code, not a hang
        if (tid == 0) *dec.out_n = -1;
        return -1;
    }
    if (stamps && tid == 0) stamps[1] = global_timer_ns();   // every record arrived
    // a peer that answered a different request (another k) would make the merge meaningless
    if (__syncthreads_count(tid < G && (__ldcg(&own[(tid < G ? tid : 0) * slot_stride].count) >> 16) != k)) {
        if (tid == 0) *dec.out_n = -2;
        return -2;
    }
    // 4. merge: key = (distance, shard, position in the shard's list) = (distance, rowid)
    const int total = G * k, padded = next_pow2(total);
    for (int i = tid; i < padded; i += nthreads) {
        uint64_t key = KEY_EMPTY;
        if (i < total) {
            const int l = i / k, p = i - l * k;
            const ExchangeSlot *src = own + l * slot_stride;
            if (p < (__ldcg(&src->count) & 0xFFFF)) key = make_key(__ldcg(&src->dist[p]), static_cast<uint32_t>(i));
        }
        scratch[i] = key;
    }
    block_bitonic_sort(scratch, padded, tid, nthreads);
    int found = 0;
    for (int base = 0; base < k; base += nthreads) {
        const int i = base + tid;
        const bool valid = i < k && i < padded && scratch[i] != KEY_EMPTY;
        if (valid) {
            const int src = static_cast<int>(scratch[i] & 0xFFFFFFFFull);
            const int l = src / k, p = src - l * k;
            dec.out_dist[i] = __ldcg(&own[l * slot_stride].dist[p]);
            dec.out_rowids[i] = __ldcg(&own[l * slot_stride].rowid[p]);
        }
        found += __syncthreads_count(valid);
    }
    if (tid == 0) {
        *dec.out_n = found;
        long long nan = 0;
        int flags = 0;
        for (int l = 0; l < G; l++) {
            nan += __ldcg(&own[l * slot_stride].nan);
            flags |= __ldcg(&own[l * slot_stride].flags);
        }
        if (dec.out_nan) *dec.out_nan = nan;
        if (out_flags) *out_flags = flags;
    }
    return found;
}

// lists: [n_lists][L] in global memory (written by other CTAs: read through L2), scratch: shared
template <int L>
__device__ __forceinline__ void merge_decode_reset(const uint64_t *lists, int n_lists, uint64_t *scratch,
                                                   const DecodeArgs &dec, const ExchangeArgs &xa, ScanSync *sync,
                                                   int tid, int nthreads) {
    __shared__ unsigned long long s_t[4];   // exchange statistics: scanned | merged | published | received
    const bool timing = xa.world > 1 && xa.stats != nullptr;
    if (timing && tid == 0) s_t[0] = global_timer_ns();
    for (int i = tid; i < n_lists * L; i += nthreads) scratch[i] = __ldcg(lists + i);
    __syncthreads();
    merge_sorted_lists_tournament<L>(scratch, n_lists, tid, nthreads);
    if (xa.world > 1) {
        __shared__ long long s_nan;
        if (tid == 0) {
            s_nan = static_cast<long long>(atomicExch(&sync->nan_rows, 0ull));
            if (timing) s_t[1] = global_timer_ns();
        }
        __syncthreads();
        const int rc = exchange_merge_decode<L>(scratch, dec, xa, static_cast<int>(xa.epoch & 1u) * xa.world, 1, s_nan,
                                                0, nullptr, tid, nthreads, timing ? s_t + 2 : nullptr);
        if (tid == 0) {
            if (timing && rc >= 0) {   // one writer per launch, launches of a context are serialised on its stream
                const unsigned long long t_end = global_timer_ns(), t0 = sync->t_start;
                xa.stats[0] += s_t[0] - (t0 ? t0 : s_t[0]);
                xa.stats[1] += s_t[1] - s_t[0];
                xa.stats[2] += s_t[2] - s_t[1];
                xa.stats[3] += s_t[3] - s_t[2];
                xa.stats[4] += t_end - s_t[3];
                xa.stats[5] += 1;
            }
            sync->t_start = 0;
            sync->tile_counter = 0;
            sync->done_counter = 0;
        }
        return;
    }
    int found = 0;
    for (int base = 0; base < dec.k; base += nthreads) {
        const int i = base + tid;
        const bool valid = i < dec.k && i < L && scratch[i] != KEY_EMPTY;
        if (valid) decode_one(dec, i, scratch[i]);
        found += __syncthreads_count(valid);
    }
    if (tid == 0) {
        *dec.out_n = found;
        const unsigned long long nan = atomicExch(&sync->nan_rows, 0ull);
        if (dec.out_nan) *dec.out_nan = static_cast<int64_t>(nan);
        sync->t_start = 0;
        sync->tile_counter = 0;
        sync->done_counter = 0;
    }
}

// One level of the reduction tree: CTA b merges lists [b*lists_per_cta, ...) of
// `stride` ascending keys each into one ascending list of `stride` keys.  When
// do_decode != 0 (single CTA, last level) the first k keys are decoded into the
// caller's result arrays instead.
__global__ void __launch_bounds__(MERGE_THREADS) reduce_lists_kernel(
    const uint64_t *__restrict__ in, int n_lists, int stride, int lists_per_cta,
    uint64_t *__restrict__ out, DecodeArgs dec, int do_decode) {
    __shared__ uint64_t s[MERGE_KEYS];
    const int tid = threadIdx.x;
    const int l0 = blockIdx.x * lists_per_cta;
    int nl = n_lists - l0;
    if (nl > lists_per_cta) nl = lists_per_cta;
    const int total = nl * stride;
    const int padded = next_pow2(total);
    const uint64_t *src = in + static_cast<size_t>(l0) * stride;
    for (int i = tid; i < padded; i += MERGE_THREADS) s[i] = i < total ? src[i] : KEY_EMPTY;
    block_bitonic_sort(s, padded, tid, MERGE_THREADS);
    if (!do_decode) {
        uint64_t *dst = out + static_cast<size_t>(blockIdx.x) * stride;
        for (int i = tid; i < stride; i += MERGE_THREADS) dst[i] = i < padded ? s[i] : KEY_EMPTY;
        return;
    }
    int found = 0;
    for (int base = 0; base < dec.k; base += MERGE_THREADS) {
        const int i = base + tid;
        const bool valid = i < dec.k && i < padded && s[i] != KEY_EMPTY;
        if (valid) decode_one(dec, i, s[i]);
        found += __syncthreads_count(valid);
    }
    if (tid == 0) {
        *dec.out_n = found;
        if (dec.out_nan) *dec.out_nan = dec.nan_rows ? static_cast<int64_t>(*dec.nan_rows) : 0;
    }
}

// General-k path: `keys` is the fully sorted key vector (n entries); decode the
// first k.  out_n must be zeroed beforehand.
__global__ void decode_sorted_kernel(const uint64_t *__restrict__ keys, long long n, DecodeArgs dec) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool valid = i < dec.k && i < n && keys[i] != KEY_EMPTY;
    if (valid) decode_one(dec, static_cast<int>(i), keys[i]);
    const unsigned votes = __ballot_sync(FULL_MASK, valid);
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(dec.out_n, __popc(votes));
    if (i == 0 && dec.out_nan) *dec.out_nan = dec.nan_rows ? static_cast<int64_t>(*dec.nan_rows) : 0;
}

// Shard merge (multi-GPU): `lists` sorted result lists of k (distance, rowid)
// pairs, shard order = rowid order.  Key = (distance, list, position), which is
// (distance, rowid) because shards are contiguous rowid ranges.  Each list l
// lives at base + l * stride (bytes) for the three arrays, so the same kernel
// reads separate [lists][k] arrays or the packed per-rank records an NCCL
// all-gather delivers.
struct ShardLists {
    const uint8_t *dist;    // float32[k] per (list, query)
    const uint8_t *rowids;  // int64[k] per (list, query)
    const uint8_t *counts;  // int32 per (list, query)
    long long dist_stride, rowid_stride, count_stride;     // bytes between consecutive lists
    long long dist_qstride, rowid_qstride, count_qstride;  // bytes between consecutive queries
};

__device__ __forceinline__ const float *shard_dist(const ShardLists &s, int l, int q) {
    return reinterpret_cast<const float *>(s.dist + static_cast<long long>(l) * s.dist_stride +
                                           static_cast<long long>(q) * s.dist_qstride);
}
__device__ __forceinline__ const int64_t *shard_rowids(const ShardLists &s, int l, int q) {
    return reinterpret_cast<const int64_t *>(s.rowids + static_cast<long long>(l) * s.rowid_stride +
                                             static_cast<long long>(q) * s.rowid_qstride);
}
__device__ __forceinline__ int shard_count(const ShardLists &s, int l, int q) {
    return *reinterpret_cast<const int32_t *>(s.counts + static_cast<long long>(l) * s.count_stride +
                                              static_cast<long long>(q) * s.count_qstride);
}

// grid.x = query; outputs are [nq][k] / [nq]
__global__ void __launch_bounds__(MERGE_THREADS) merge_shards_kernel(
    const ShardLists in, int lists, int k, float *__restrict__ out_dist,
    int64_t *__restrict__ out_rowids, int32_t *__restrict__ out_n) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    uint64_t *s = reinterpret_cast<uint64_t *>(merge_smem);
    const int tid = threadIdx.x, q = blockIdx.x;
    const int total = lists * k;
    const int padded = next_pow2(total);
    for (int i = tid; i < padded; i += MERGE_THREADS) {
        uint64_t key = KEY_EMPTY;
        if (i < total) {
            const int l = i / k, p = i - l * k;
            if (p < shard_count(in, l, q)) {
                const float d = shard_dist(in, l, q)[p];
                if (d == d) key = make_key(d, static_cast<uint32_t>(i));
            }
        }
        s[i] = key;
    }
    block_bitonic_sort(s, padded, tid, MERGE_THREADS);
    int found = 0;
    for (int base = 0; base < k; base += MERGE_THREADS) {
        const int i = base + tid;
        const bool valid = i < k && i < padded && s[i] != KEY_EMPTY;
        if (valid) {
            const int src = static_cast<int>(s[i] & 0xFFFFFFFFull);
            const int l = src / k, p = src - l * k;
            out_dist[static_cast<size_t>(q) * k + i] = shard_dist(in, l, q)[p];
            out_rowids[static_cast<size_t>(q) * k + i] = shard_rowids(in, l, q)[p];
        }
        found += __syncthreads_count(valid);
    }
    if (tid == 0) out_n[q] = found;
}

}  // namespace clipdb
