// scan_topk.cuh — K1 + K2: the distance scan fused with top-k selection.
//
// Replaces, for one query, what SQLite + sqlite-vec do per search
// (image_database.py:1564-1574, executed at :1582): score EVERY stored
// float32 row against the query with vec_distance_cosine and keep the k
// smallest (distance, scan sequence) pairs.  The distance vector is never
// written to HBM: each warp keeps a k-entry candidate list in registers, each
// CTA emits one sorted list, and merge.cuh reduces the lists to the answer.
//
// Roofline: HBM.  Algorithmic bytes per query = n * dim * 4 (the row store is
// read exactly once); ~1 flop per byte.  Two variants of the same arithmetic:
//   scan_tma_kernel  dim == 1152 (the reference's only width).  Persistent,
//                    one CTA per SM.  One producer thread streams R-row tiles
//                    (one contiguous R*4608-B span each) global->shared with
//                    cp.async.bulk (TMA engine, SASS UBLKCP) into an mbarrier
//                    ring; groups of R consumer warps take one row each with
//                    conflict-free LDS.128, reduce with warp shuffles and
//                    maintain the candidate lists.  Ring shape = ScanCfg<>.
//   scan_ldg_kernel  any dim (multiple of 4, padded by the loader).  Each warp
//                    grid-strides over rows with 128-bit streaming loads
//                    (ld.global.nc.L1::no_allocate), query staged in shared.
#pragma once

#include "common.cuh"
#include "merge.cuh"

namespace clipdb {

constexpr int METRIC_COSINE = 0;
constexpr int METRIC_L2 = 1;

constexpr int SCAN_DIM = 1152;  // SigLIP 2 SO400M width, `embedding float[1152]` (idb:290-294)
constexpr int SCAN_CHUNKS = SCAN_DIM / 128;  // float4 chunks per lane per row = 9
constexpr int SCAN_ROW_BYTES = SCAN_DIM * 4;  // 4608 = 36 x 128 B
constexpr int SCAN_SMEM_HEADER = 1024;        // barriers + per-stage tile ids

// Shape of the TMA ring.  A tile is R consecutive rows (one contiguous R*4608-byte
// span, one bulk copy); a group of R consumer warps takes one row each; GROUPS
// groups work on different tiles at the same time.
template <int R, int STAGES_, int GROUPS_>
struct ScanCfg {
    static constexpr int TILE_ROWS = R;
    static constexpr int GROUP_WARPS = R;
    static constexpr int STAGES = STAGES_;
    static constexpr int GROUPS = GROUPS_;
    static constexpr int CONSUMER_WARPS = R * GROUPS_;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;  // + producer warp
    static constexpr int STAGE_BYTES = R * SCAN_ROW_BYTES;
    static constexpr int SMEM_BYTES = SCAN_SMEM_HEADER + STAGES_ * STAGE_BYTES;
    static_assert(GROUPS_ <= STAGES_, "each group needs its own stage in flight");
    static_assert(SMEM_BYTES <= 227 * 1024, "ring exceeds shared memory");
    static_assert(STAGES_ * 20 <= SCAN_SMEM_HEADER, "header too small");
    static_assert(CONSUMER_WARPS * 32 * 4 * 8 <= STAGES_ * STAGE_BYTES, "sort scratch must fit the ring");
};
using ScanCfg0 = ScanCfg<8, 6, 2>;    // 36,864-B tiles, 6 stages, 16 consumer warps (default)
using ScanCfg1 = ScanCfg<4, 12, 4>;   // 18,432-B tiles, 12 stages
using ScanCfg2 = ScanCfg<16, 3, 1>;   // 73,728-B tiles, 3 stages
using ScanCfg3 = ScanCfg<8, 4, 2>;    // as 0 with a shallower ring (147 KB)
using ScanCfg4 = ScanCfg<8, 6, 1>;    // as 0 with 8 consumer warps

constexpr int SCAN_ASSIGN_INTERLEAVED = 0;  // tile t -> CTA t mod grid
constexpr int SCAN_ASSIGN_CONTIGUOUS = 1;   // CTA b owns one contiguous span of tiles
constexpr int SCAN_ASSIGN_DYNAMIC = 2;      // chunks of tiles claimed with atomicAdd

constexpr int LDG_WARPS = 8;
constexpr int LDG_THREADS = LDG_WARPS * 32;

// A store may be TIERED (clipdb_reserve_rows): rows [0, split) in HBM at `rows`, rows [split, n) in pinned,
// device-mapped host memory.  `rows_hi` is biased so that rows_hi + pos * ld addresses row pos >= split;
// split is a multiple of 128, so no tile of any kernel straddles it.  All-HBM stores have split = LLONG_MAX.
__device__ __forceinline__ const float *row_address(const float *rows, const float *rows_hi, long long split,
                                                    long long pos, int ld) {
    return (pos < split ? rows : rows_hi) + pos * ld;
}

struct ScanArgs {
    const float *rows;             // [n][ld] float32, scan order
    const float *rows_hi;          // tiered store: biased base of the host-resident rows (see row_address)
    long long split;               // first position that lives in host memory (LLONG_MAX: none)
    const float *query;            // [dim]
    const uint32_t *mask;          // nullable admission bitset
    uint64_t *cand;                // [gridDim.x][cand_stride], each ascending
    ScanSync *sync;                // tile / done / NaN counters, zero between launches (merge.cuh)
    uint64_t *all_keys;            // WRITE_ALL: [n] keys (KEY_EMPTY = not admitted / NaN)
    long long n;
    int dim;
    int ld;           // row stride in floats (multiple of 4)
    int k;            // requested k, <= cand_stride
    int cand_stride;  // 32 * KPL
    int evict_first;  // stream rows with an L2 evict-first policy
    int assign;       // SCAN_ASSIGN_*
    int chunk_tiles;  // tiles claimed per atomicAdd (dynamic assignment)
    int fuse_tail;    // the last CTA to finish merges all lists and decodes the result (TMA kernel)
    DecodeArgs dec;   // used when fuse_tail != 0
    ExchangeArgs xchg;  // world > 1: the tail also exchanges the result with the other shards' GPUs
};

// The exact finish (double sqrt, multiply, divide) is ~70 instructions.  A float32
// estimate 1 - dot*rsqrt(aMag)*rsqrt(bMag) (4 instructions, |error| < 6e-7 for any
// input with |cos| <= 1) decides first whether the row can possibly beat the current
// k-th candidate; only rows within PREFILTER_MARGIN of it pay for the exact finish, so
// results are unchanged.  NaN estimates compare false and take the exact path (counted).
constexpr float PREFILTER_MARGIN = 4e-6f;

// ---- per-warp candidate list ---------------------------------------------------
// 32*KPL slots spread over the lanes' registers, unsorted; `thr` is the largest
// key held (warp-uniform).  Slots >= k are disabled so the list holds exactly
// the k best keys seen so far; a new key enters iff key < thr and replaces the
// current maximum.  Admission is rare (~k/rows-seen), so the common path is one
// compare.
template <int KPL>
struct WarpTopK {
    uint64_t keys[KPL];
    uint64_t thr;
    float thr_d;  // distance above which a row cannot enter (thr's distance + PREFILTER_MARGIN)

    __device__ __forceinline__ void init(int k, int lane) {
#pragma unroll
        for (int j = 0; j < KPL; j++) keys[j] = (j * 32 + lane < k) ? KEY_EMPTY : KEY_DISABLED;
        thr = (k > 0) ? KEY_EMPTY : KEY_DISABLED;
        thr_d = __int_as_float(0x7f800000);  // +inf: everything takes the exact path until the list is full
    }

    // key < thr on entry; key is warp-uniform
    __device__ __forceinline__ void insert(uint64_t key, int lane) {
        int slot = -1;
#pragma unroll
        for (int j = KPL - 1; j >= 0; j--)
            if (keys[j] == thr) slot = j;
        unsigned holders = __ballot_sync(FULL_MASK, slot >= 0);
        int owner = __ffs(holders) - 1;
        if (lane == owner) {
#pragma unroll
            for (int j = 0; j < KPL; j++)
                if (j == slot) keys[j] = key;
        }
        uint64_t m = keys[0];
#pragma unroll
        for (int j = 1; j < KPL; j++) m = umax64(m, keys[j]);
        thr = warp_max_u64(m);
        thr_d = (thr == KEY_EMPTY) ? __int_as_float(0x7f800000)
                                   : orderable_f32(static_cast<uint32_t>(thr >> 32)) + PREFILTER_MARGIN;
    }

    __device__ __forceinline__ void dump(uint64_t *dst, int lane) const {
#pragma unroll
        for (int j = 0; j < KPL; j++)
            dst[j * 32 + lane] = (keys[j] == KEY_DISABLED) ? KEY_EMPTY : keys[j];
    }
};

// The reference arithmetic once the three float32 sums exist (oracle_ref.c):
// roots, product, divide and subtract in double, narrowed to float32.
template <int METRIC>
__device__ __forceinline__ float finish_distance(float s0, float s1, double sqrt_b) {
    if (METRIC == METRIC_COSINE)
        return static_cast<float>(1.0 - static_cast<double>(s0) /
                                            (sqrt(static_cast<double>(s1)) * sqrt_b));
    return static_cast<float>(sqrt(static_cast<double>(s0)));
}

// Packed accumulators: Blackwell issues two float32 FMAs per instruction (PTX fma.rn.f32x2,
// SASS FFMA2) on adjacent register pairs, which is exactly how an LDS.128 / LDG.128 delivers a
// float4.  Each lane is the same IEEE fma as the scalar form, so results are bit-identical; the
// row loop drops from 72 to 36 FMA instructions per lane per row (less issue pressure and power
// in a kernel that runs at the board's power cap).
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

struct RowSums {
    float2 d01, d23;   // dot (cosine) / squared difference (L2) partial sums of elements x,y and z,w
    float2 n01, n23;   // row squared-norm partial sums (cosine)
    __device__ __forceinline__ void clear() {
        d01 = d23 = n01 = n23 = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ float first() const { return (d01.x + d01.y) + (d23.x + d23.y); }
    __device__ __forceinline__ float second() const { return (n01.x + n01.y) + (n23.x + n23.y); }
};

template <int METRIC>
__device__ __forceinline__ void accumulate(const float4 v, const float4 q, RowSums &s) {
    if (METRIC == METRIC_COSINE) {
        const float2 vlo = make_float2(v.x, v.y), vhi = make_float2(v.z, v.w);
        ffma2(s.d01, vlo, make_float2(q.x, q.y));
        ffma2(s.d23, vhi, make_float2(q.z, q.w));
        ffma2(s.n01, vlo, vlo);
        ffma2(s.n23, vhi, vhi);
    } else {
        const float2 dlo = make_float2(v.x - q.x, v.y - q.y), dhi = make_float2(v.z - q.z, v.w - q.w);
        ffma2(s.d01, dlo, dlo);
        ffma2(s.d23, dhi, dhi);
    }
}

// Row sums are complete in every lane: turn them into a key and offer it.
template <int KPL, int METRIC, bool WRITE_ALL>
__device__ __forceinline__ void offer_row(float s0, float s1, double sqrt_b, float rsqrt_b,
                                          long long pos, WarpTopK<KPL> &top, unsigned &nan_rows,
                                          uint64_t *all_keys, int lane) {
    if (!WRITE_ALL && METRIC == METRIC_COSINE) {
        const float estimate = fmaf(-s0 * rsqrtf(s1), rsqrt_b, 1.0f);
        if (estimate > top.thr_d) return;
    }
    float d = finish_distance<METRIC>(s0, s1, sqrt_b);
    if (d != d) {
        nan_rows++;
        if (WRITE_ALL && lane == 0) all_keys[pos] = KEY_EMPTY;
        return;
    }
    uint64_t key = make_key(d, static_cast<uint32_t>(pos));
    if (WRITE_ALL) {
        if (lane == 0) all_keys[pos] = key;
    } else if (key < top.thr) {
        top.insert(key, lane);
    }
}

__device__ __forceinline__ bool row_admitted(const uint32_t *mask, long long pos) {
    return mask == nullptr || ((__ldg(mask + (pos >> 5)) >> (pos & 31)) & 1u);
}

// All warps' lists -> one ascending list of cand_stride keys for this CTA.
// `scratch` holds next_pow2(warps * 32 * KPL) keys.
template <int KPL>
__device__ __forceinline__ void emit_cta_list(uint64_t *scratch, int n_lists,
                                              uint64_t *cand_out, int tid, int nthreads) {
    const int total = n_lists * 32 * KPL;
    const int padded = next_pow2(total);
    for (int i = total + tid; i < padded; i += nthreads) scratch[i] = KEY_EMPTY;
    block_bitonic_sort(scratch, padded, tid, nthreads);
    for (int i = tid; i < 32 * KPL; i += nthreads) cand_out[i] = scratch[i];
}

// ---- TMA-ring variant, dim == 1152 ------------------------------------------------
template <typename CFG, int KPL, int METRIC, bool WRITE_ALL>
__global__ void __launch_bounds__(CFG::THREADS, 1) scan_tma_kernel(const ScanArgs a) {
    constexpr int STAGES = CFG::STAGES;
    constexpr int R = CFG::TILE_ROWS;
    extern __shared__ __align__(128) uint8_t scan_smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(scan_smem);
    uint64_t *empty_bar = full_bar + STAGES;
    volatile int *tile_of = reinterpret_cast<volatile int *>(empty_bar + STAGES);  // -1 = no more tiles
    uint8_t *ring = scan_smem + SCAN_SMEM_HEADER;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int total_tiles = static_cast<int>((a.n + R - 1) / R);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CFG::GROUP_WARPS);
        }
        mbar_fence_init();
        // exchange statistics: the first CTA to get here stamps the launch's start
        if (a.xchg.world > 1 && a.xchg.stats) atomicCAS(&a.sync->t_start, 0ull, global_timer_ns());
    }
    __syncthreads();

    WarpTopK<KPL> top;
    top.init(a.k, lane);
    unsigned nan_rows = 0;

    if (warp == CFG::CONSUMER_WARPS) {
        // ===== producer: one thread feeds the ring =====
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            int s = 0;
            uint32_t phase = 0;
            auto push = [&](int tile) {
                mbar_wait(&empty_bar[s], phase ^ 1u);
                tile_of[s] = tile;
                if (tile >= 0) {
                    const long long row0 = static_cast<long long>(tile) * R;
                    const long long left = a.n - row0;
                    const uint32_t bytes = static_cast<uint32_t>((left < R ? left : R) * SCAN_ROW_BYTES);
                    mbar_arrive_expect_tx(&full_bar[s], bytes);
                    const float *src = row_address(a.rows, a.rows_hi, a.split, row0, SCAN_DIM);
                    if (a.evict_first)
                        bulk_copy_g2s_hint(ring + s * CFG::STAGE_BYTES, src, bytes, &full_bar[s], policy);
                    else
                        bulk_copy_g2s(ring + s * CFG::STAGE_BYTES, src, bytes, &full_bar[s]);
                } else {
                    mbar_arrive(&full_bar[s]);  // end marker: completes the phase without data
                }
                if (++s == STAGES) {
                    s = 0;
                    phase ^= 1u;
                }
            };
            const int grid = static_cast<int>(gridDim.x), b = static_cast<int>(blockIdx.x);
            if (a.assign == SCAN_ASSIGN_DYNAMIC) {
                const unsigned chunk = static_cast<unsigned>(a.chunk_tiles > 0 ? a.chunk_tiles : 1);
                unsigned next = atomicAdd(&a.sync->tile_counter, chunk);
                while (next < static_cast<unsigned>(total_tiles)) {
                    const unsigned base = next;
                    next = atomicAdd(&a.sync->tile_counter, chunk);  // claimed ahead: latency hidden by the pushes
                    for (unsigned j = 0; j < chunk && base + j < static_cast<unsigned>(total_tiles); j++)
                        push(static_cast<int>(base + j));
                }
            } else if (a.assign == SCAN_ASSIGN_CONTIGUOUS) {
                const int per = (total_tiles + grid - 1) / grid;
                const int lo = b * per, hi = (lo + per < total_tiles) ? lo + per : total_tiles;
                for (int t = lo; t < hi; t++) push(t);
            } else {
                for (int t = b; t < total_tiles; t += grid) push(t);
            }
            for (int g = 0; g < CFG::GROUPS; g++) push(-1);  // one end marker per consumer group
        }
    } else {
        // ===== consumers: group g takes ring slots g, g+GROUPS, ...; warp wi takes row wi =====
        const int group = warp / CFG::GROUP_WARPS;
        const int wi = warp % CFG::GROUP_WARPS;

        float4 q[SCAN_CHUNKS];
        const float4 *q4 = reinterpret_cast<const float4 *>(a.query);
        float bsum = 0.f;
#pragma unroll
        for (int j = 0; j < SCAN_CHUNKS; j++) {
            q[j] = __ldg(q4 + lane + 32 * j);
            bsum = fmaf(q[j].x, q[j].x, bsum);
            bsum = fmaf(q[j].y, q[j].y, bsum);
            bsum = fmaf(q[j].z, q[j].z, bsum);
            bsum = fmaf(q[j].w, q[j].w, bsum);
        }
        const double sqrt_b = sqrt(static_cast<double>(warp_sum(bsum)));
        const float rsqrt_b = static_cast<float>(1.0 / sqrt_b);

        int s = group;
        uint32_t phase = 0;
        for (;;) {
            mbar_wait(&full_bar[s], phase);
            const int tile = tile_of[s];
            if (tile < 0) break;
            const long long pos = static_cast<long long>(tile) * R + wi;
            if (pos < a.n && row_admitted(a.mask, pos)) {
                const float4 *src =
                    reinterpret_cast<const float4 *>(ring + s * CFG::STAGE_BYTES + wi * SCAN_ROW_BYTES);
                float4 v[SCAN_CHUNKS];
#pragma unroll
                for (int j = 0; j < SCAN_CHUNKS; j++) v[j] = src[lane + 32 * j];
                RowSums sums;
                sums.clear();
#pragma unroll
                for (int j = 0; j < SCAN_CHUNKS; j++) accumulate<METRIC>(v[j], q[j], sums);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);  // row is in registers: free the slot
                float t0 = warp_sum(sums.first());
                float t1 = (METRIC == METRIC_COSINE) ? warp_sum(sums.second()) : 0.f;
                offer_row<KPL, METRIC, WRITE_ALL>(t0, t1, sqrt_b, rsqrt_b, pos, top, nan_rows, a.all_keys, lane);
            } else {
                if (WRITE_ALL && pos < a.n && lane == 0) a.all_keys[pos] = KEY_EMPTY;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
            s += CFG::GROUPS;
            if (s >= STAGES) {
                s -= STAGES;
                phase ^= 1u;
            }
        }
    }

    if (nan_rows && lane == 0) atomicAdd(&a.sync->nan_rows, static_cast<unsigned long long>(nan_rows));
    if (WRITE_ALL) return;

    // every issued copy has been consumed (each full barrier was waited on), so
    // the ring can be reused as sort scratch
    __syncthreads();
    uint64_t *scratch = reinterpret_cast<uint64_t *>(ring);
    if (warp < CFG::CONSUMER_WARPS) top.dump(scratch + warp * 32 * KPL, lane);
    __syncthreads();
    emit_cta_list<KPL>(scratch, CFG::CONSUMER_WARPS,
                       a.cand + static_cast<size_t>(blockIdx.x) * a.cand_stride, tid, CFG::THREADS);
    if (!a.fuse_tail) return;
    if (!last_cta_done(&a.sync->done_counter, tid)) return;
    merge_decode_reset<32 * KPL>(a.cand, static_cast<int>(gridDim.x), scratch, a.dec, a.xchg, a.sync, tid,
                                 CFG::THREADS);
}

// ---- direct-load variant, any dim -----------------------------------------------
// DIM_T > 0: compile-time dim (query chunks in registers); DIM_T == 0: runtime dim.
template <int DIM_T, int KPL, int METRIC, bool WRITE_ALL>
__global__ void __launch_bounds__(LDG_THREADS) scan_ldg_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) uint8_t scan_smem[];
    uint8_t *smem = scan_smem;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int dim = DIM_T > 0 ? DIM_T : a.dim;
    const int ld = DIM_T > 0 ? DIM_T : a.ld;
    const int chunks = ld >> 2;  // float4 per row (padding floats are zero)

    // sort scratch and the query share the dynamic shared memory
    uint64_t *scratch = reinterpret_cast<uint64_t *>(smem);
    float4 *qs = reinterpret_cast<float4 *>(smem + next_pow2(LDG_WARPS * 32 * KPL) * sizeof(uint64_t));
    for (int c = tid; c < chunks; c += LDG_THREADS) {
        float4 v;
        const int base = c * 4;
        v.x = base + 0 < dim ? __ldg(a.query + base + 0) : 0.f;
        v.y = base + 1 < dim ? __ldg(a.query + base + 1) : 0.f;
        v.z = base + 2 < dim ? __ldg(a.query + base + 2) : 0.f;
        v.w = base + 3 < dim ? __ldg(a.query + base + 3) : 0.f;
        qs[c] = v;
    }
    __syncthreads();

    float bsum = 0.f;
    for (int c = lane; c < chunks; c += 32) {
        float4 v = qs[c];
        bsum = fmaf(v.x, v.x, bsum);
        bsum = fmaf(v.y, v.y, bsum);
        bsum = fmaf(v.z, v.z, bsum);
        bsum = fmaf(v.w, v.w, bsum);
    }
    const double sqrt_b = sqrt(static_cast<double>(warp_sum(bsum)));
    const float rsqrt_b = static_cast<float>(1.0 / sqrt_b);

    WarpTopK<KPL> top;
    top.init(a.k, lane);
    unsigned nan_rows = 0;

    const long long warp_global = static_cast<long long>(blockIdx.x) * LDG_WARPS + warp;
    const long long warp_count = static_cast<long long>(gridDim.x) * LDG_WARPS;

    for (long long pos = warp_global; pos < a.n; pos += warp_count) {
        if (!row_admitted(a.mask, pos)) {
            if (WRITE_ALL && lane == 0) a.all_keys[pos] = KEY_EMPTY;
            continue;
        }
        const float4 *src = reinterpret_cast<const float4 *>(row_address(a.rows, a.rows_hi, a.split, pos, ld));
        RowSums sums;
        sums.clear();
        if (DIM_T > 0) {
            constexpr int CH = DIM_T > 0 ? DIM_T / 128 : 1;
            float4 v[CH];
#pragma unroll
            for (int j = 0; j < CH; j++) v[j] = ldg_stream(src + lane + 32 * j);
#pragma unroll
            for (int j = 0; j < CH; j++) accumulate<METRIC>(v[j], qs[lane + 32 * j], sums);
        } else {
            int c = lane;
            for (; c + 96 < chunks; c += 128) {
                float4 v0 = ldg_stream(src + c), v1 = ldg_stream(src + c + 32);
                float4 v2 = ldg_stream(src + c + 64), v3 = ldg_stream(src + c + 96);
                accumulate<METRIC>(v0, qs[c], sums);
                accumulate<METRIC>(v1, qs[c + 32], sums);
                accumulate<METRIC>(v2, qs[c + 64], sums);
                accumulate<METRIC>(v3, qs[c + 96], sums);
            }
            for (; c < chunks; c += 32) accumulate<METRIC>(ldg_stream(src + c), qs[c], sums);
        }
        float t0 = warp_sum(sums.first());
        float t1 = (METRIC == METRIC_COSINE) ? warp_sum(sums.second()) : 0.f;
        offer_row<KPL, METRIC, WRITE_ALL>(t0, t1, sqrt_b, rsqrt_b, pos, top, nan_rows, a.all_keys, lane);
    }

    if (nan_rows && lane == 0) atomicAdd(&a.sync->nan_rows, static_cast<unsigned long long>(nan_rows));
    if (WRITE_ALL) return;

    top.dump(scratch + warp * 32 * KPL, lane);
    __syncthreads();
    emit_cta_list<KPL>(scratch, LDG_WARPS, a.cand + static_cast<size_t>(blockIdx.x) * a.cand_stride,
                       tid, LDG_THREADS);
}

}  // namespace clipdb
