// binary.cuh — the sign-code fallback search.
//
// When `vec0` is empty the reference falls back to its 1-bit codes
// (image_database.py:1591-1629): every stored code is a 1152-byte array of 0/1
// (`(embedding >= 0).astype(np.uint8)`, :1189-1190), the query is quantised the same
// way (:1593) and the score of a row is `np.dot(query_binary, candidate_binary)` —
// the number of positions where BOTH are 1 (an AND-popcount, not a Hamming distance).
// numpy computes that dot product in uint8, so the reference's score is the popcount
// MODULO 256 (probed: 1152 matching ones -> 128); rows are then sorted by
// score/1152 descending with Python's stable sort (ties keep scan order) and the
// first k are returned (:1627-1629).
//
// Here the codes are bit-packed (36 words = 144 B per row instead of 1152 B) and one
// kernel streams them once: HBM-bound, 144 B per row.  Key = (score_max - score) << 32
// | scan position, so ascending key order is the reference's order and the fused
// top-k machinery of the float scan (WarpTopK, the merge tree) is reused unchanged.
// `score_mask` = 0xFF reproduces the reference's uint8 wrap-around, 0xFFFF gives the
// plain popcount.
#pragma once

#include "common.cuh"
#include "scan_topk.cuh"

namespace clipdb {

constexpr int BIN_WORDS = SCAN_DIM / 32;        // 36 words per code
constexpr int BIN_ROW_BYTES = BIN_WORDS * 4;    // 144 B = 9 x 16 B
constexpr int BIN_TILE_ROWS = 256;              // one tile = 36,864 contiguous bytes (as the fp32 scan's)
constexpr int BIN_STAGES = 6;
constexpr int BIN_GROUPS = 2;                   // consumer groups of 8 warps: one thread per row of a tile
constexpr int BIN_GROUP_WARPS = BIN_TILE_ROWS / 32;
constexpr int BIN_CONSUMER_WARPS = BIN_GROUPS * BIN_GROUP_WARPS;
constexpr int BIN_THREADS = (BIN_CONSUMER_WARPS + 1) * 32;
constexpr int BIN_STAGE_BYTES = BIN_TILE_ROWS * BIN_ROW_BYTES;
constexpr int BIN_SMEM_BYTES = SCAN_SMEM_HEADER + BIN_STAGES * BIN_STAGE_BYTES;
static_assert(BIN_CONSUMER_WARPS * 32 * 4 * 8 <= BIN_STAGES * BIN_STAGE_BYTES, "sort scratch must fit the ring");

struct BinaryArgs {
    const uint32_t *codes;         // [n][36] bit-packed, bit (i & 31) of word (i >> 5) = element i
    const uint32_t *query;         // [36]
    const uint32_t *mask;          // nullable admission bitset
    const uint32_t *seq;           // nullable: tie-break sequence of each row (a permutation of 0..n-1) used
                                   // instead of its position (SQLite returns a filtered statement's rows in
                                   // file_path index order, see clipdb_set_code_mask)
    uint64_t *cand;                // [gridDim.x][cand_stride]
    uint64_t *all_keys;            // WRITE_ALL: [n]
    ScanSync *sync;                // tile / done counters, zero between launches (merge.cuh)
    long long n;
    int k;
    int cand_stride;
    int chunk_tiles;
    uint32_t score_mask;           // 0xFF: the reference's uint8 wrap-around; 0xFFFF: plain popcount
    uint32_t score_max;            // 255 or 1152: key = (score_max - score) << 32 | position
    int fuse_tail;                 // the last CTA to finish merges all lists and decodes the result
    DecodeArgs dec;
};

// bytes (0/1) -> bits.  One thread per output word; `bad` counts bytes other than 0/1 (the
// reference would multiply by the raw byte value; refused here rather than silently differing).
__global__ void pack_codes_kernel(const uint8_t *__restrict__ bytes, long long n_words,
                                  uint32_t *__restrict__ words, unsigned int *__restrict__ bad) {
    const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(bytes + w * 32);
    const uint4 lo = __ldg(src), hi = __ldg(src + 1);
    const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t out = 0, other = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        other |= v[j] & 0xFEFEFEFEu;
        // bytes b0..b3 (little endian) -> bits 4j .. 4j+3
        const uint32_t x = v[j] & 0x01010101u;
        out |= ((x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xFu) << (4 * j);
    }
    words[w] = out;
    if (other) atomicAdd(bad, 1u);
}

// ids_by_seq[seq[pos]] = ids ? ids[pos] : pos — what the decode step looks up when keys carry seq
__global__ void ids_by_sequence_kernel(const uint32_t *__restrict__ seq, const int64_t *__restrict__ ids, long long n,
                                       int64_t *__restrict__ ids_by_seq, unsigned int *__restrict__ bad) {
    const long long pos = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const uint32_t s = seq[pos];
    if (s >= n) {
        atomicAdd(bad, 1u);
        return;
    }
    ids_by_seq[s] = ids ? ids[pos] : pos;
}

template <int KPL, bool WRITE_ALL>
__global__ void __launch_bounds__(BIN_THREADS, 1) binary_scan_kernel(const BinaryArgs a) {
    extern __shared__ __align__(128) uint8_t scan_smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(scan_smem);
    uint64_t *empty_bar = full_bar + BIN_STAGES;
    volatile int *tile_of = reinterpret_cast<volatile int *>(empty_bar + BIN_STAGES);
    uint8_t *ring = scan_smem + SCAN_SMEM_HEADER;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_tiles = static_cast<int>((a.n + BIN_TILE_ROWS - 1) / BIN_TILE_ROWS);

    if (tid == 0) {
        for (int s = 0; s < BIN_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], BIN_GROUP_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    WarpTopK<KPL> top;
    top.init(a.k, lane);

    if (warp == BIN_CONSUMER_WARPS) {
        // ===== producer: one thread streams 36,864-byte tiles through the TMA engine =====
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            int s = 0;
            uint32_t phase = 0;
            auto push = [&](int tile) {
                mbar_wait(&empty_bar[s], phase ^ 1u);
                tile_of[s] = tile;
                if (tile >= 0) {
                    const long long row0 = static_cast<long long>(tile) * BIN_TILE_ROWS;
                    const long long left = a.n - row0;
                    const uint32_t bytes =
                        static_cast<uint32_t>((left < BIN_TILE_ROWS ? left : BIN_TILE_ROWS) * BIN_ROW_BYTES);
                    mbar_arrive_expect_tx(&full_bar[s], bytes);
                    bulk_copy_g2s_hint(ring + s * BIN_STAGE_BYTES, a.codes + row0 * BIN_WORDS, bytes, &full_bar[s],
                                       policy);
                } else {
                    mbar_arrive(&full_bar[s]);
                }
                if (++s == BIN_STAGES) {
                    s = 0;
                    phase ^= 1u;
                }
            };
            const unsigned chunk = static_cast<unsigned>(a.chunk_tiles > 0 ? a.chunk_tiles : 1);
            unsigned next = atomicAdd(&a.sync->tile_counter, chunk);
            while (next < static_cast<unsigned>(total_tiles)) {
                const unsigned base = next;
                next = atomicAdd(&a.sync->tile_counter, chunk);
                for (unsigned j = 0; j < chunk && base + j < static_cast<unsigned>(total_tiles); j++)
                    push(static_cast<int>(base + j));
            }
            for (int g = 0; g < BIN_GROUPS; g++) push(-1);
        }
    } else {
        // ===== consumers: a group of 8 warps takes a tile, one thread per row =====
        const int group = warp / BIN_GROUP_WARPS;
        const int wi = warp % BIN_GROUP_WARPS;
        uint4 q[BIN_WORDS / 4];
#pragma unroll
        for (int j = 0; j < BIN_WORDS / 4; j++) q[j] = __ldg(reinterpret_cast<const uint4 *>(a.query) + j);

        int s = group;
        uint32_t phase = 0;
        for (;;) {
            mbar_wait(&full_bar[s], phase);
            const int tile = tile_of[s];
            if (tile < 0) break;
            const int r = wi * 32 + lane;
            const long long pos = static_cast<long long>(tile) * BIN_TILE_ROWS + r;
            // 144-byte row stride: the 8 lanes of an LDS.128 phase hit banks 0,4,...,28 — conflict-free
            const uint4 *src = reinterpret_cast<const uint4 *>(ring + s * BIN_STAGE_BYTES + r * BIN_ROW_BYTES);
            uint4 v[BIN_WORDS / 4];
            const bool in_range = pos < a.n;
            if (in_range) {
#pragma unroll
                for (int j = 0; j < BIN_WORDS / 4; j++) v[j] = src[j];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);   // rows are in registers: free the slot
            uint64_t key = KEY_EMPTY;
            if (in_range && row_admitted(a.mask, pos)) {
                uint32_t score = 0;
#pragma unroll
                for (int j = 0; j < BIN_WORDS / 4; j++)
                    score += __popc(v[j].x & q[j].x) + __popc(v[j].y & q[j].y) + __popc(v[j].z & q[j].z) +
                             __popc(v[j].w & q[j].w);
                score &= a.score_mask;
                const uint32_t tie = a.seq ? __ldg(a.seq + pos) : static_cast<uint32_t>(pos);
                key = (static_cast<uint64_t>(a.score_max - score) << 32) | tie;
            }
            if (WRITE_ALL) {
                if (in_range) a.all_keys[pos] = key;   // any order: the radix sort follows
            } else {
                unsigned pending = __ballot_sync(FULL_MASK, key < top.thr);
                while (pending) {   // warp-uniform inserts; rare once the list holds k rows
                    const int src_lane = __ffs(pending) - 1;
                    const uint64_t kk = __shfl_sync(FULL_MASK, key, src_lane);
                    if (kk < top.thr) top.insert(kk, lane);
                    pending &= pending - 1;
                }
            }
            s += BIN_GROUPS;
            if (s >= BIN_STAGES) {
                s -= BIN_STAGES;
                phase ^= 1u;
            }
        }
    }
    if (WRITE_ALL) return;

    __syncthreads();
    uint64_t *scratch = reinterpret_cast<uint64_t *>(ring);
    if (warp < BIN_CONSUMER_WARPS) top.dump(scratch + warp * 32 * KPL, lane);
    __syncthreads();
    emit_cta_list<KPL>(scratch, BIN_CONSUMER_WARPS, a.cand + static_cast<size_t>(blockIdx.x) * a.cand_stride, tid,
                       BIN_THREADS);
    if (!a.fuse_tail) return;
    if (!last_cta_done(&a.sync->done_counter, tid)) return;
    ExchangeArgs none{};
    merge_decode_reset<32 * KPL>(a.cand, static_cast<int>(gridDim.x), scratch, a.dec, none, a.sync, tid, BIN_THREADS);
}

}  // namespace clipdb
