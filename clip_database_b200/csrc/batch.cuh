// batch.cuh — K4: batched queries as a dense Q x D contraction on tcgen05 tensor cores,
// a threshold-filter epilogue, and an exact float32 re-rank.
//
// No reference equivalent: the reference answers strictly one query at a time
// (image_database.py:2070-2299).  BASELINE configs[2] (B = 256, k = 100, 10M rows) asks
// for many concurrent sessions' queries to share one pass over the store.  Results must
// equal the single-query path (scan_topk.cuh) exactly, so the tensor cores only
// PRE-SELECT:
//
//   1. bf16 copy of the store D^ [n][1152] (+ fp32 1/||row||), bf16 queries Q^ [256][1152].
//   2. score(q,row) = Q^ . D^ on tcgen05 (UMMA 128x256x16, fp32 accumulators in TMEM),
//      operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a 4-stage
//      mbarrier ring; warp-specialised: 1 TMA thread, 1 MMA thread, 4 epilogue warps
//      reading the accumulators with tcgen05.ld while the next tile's MMAs run (two
//      accumulator buffers = all 512 TMEM columns).
//   3. pass A (every 16th tile): the epilogue reduces u = score/||row|| to one maximum per
//      (query, 32-row group) with a warp transpose-reduce; per query, tau = the k-th largest
//      group maximum  =>  at least k distinct rows have u >= tau.
//   4. pass B (all tiles): the epilogue keeps (q,row) iff u >= tau - 2E||q||, where
//      E bounds |cos^ - cos| for bf16-rounded operands: the measured rounding-error norms
//      of the rows (max over the store) and of the query, + accumulation slack (~2.6e-3).
//      Every row of the true top-k satisfies this (it has cos >= cos_k >= tau/||q|| - E),
//      so the candidate set is a superset of the answer — a guarantee, not a heuristic.
//   5. re-rank: each candidate's distance is recomputed with K1's exact float32/double
//      arithmetic and the same 64-bit keys, so ids and distance bits equal the
//      single-query path.  Queries whose candidate list overflows (or whose norm is
//      zero) are flagged and re-run through the exact scan by the host.
//
// Roofline: 2*B*n*1152 flop per batch on the tensor pipe (5.9 TFLOP at B=256, n=10M) and
// n*2304 B of HBM for the bf16 store; the re-rank reads ~candidates*4608 B.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "merge.cuh"
#include "scan_topk.cuh"

namespace clipdb {

constexpr int BQ_N = 256;        // queries per pass (UMMA N)
constexpr int BQ_M = 128;        // rows per tile (UMMA M, one TMEM lane per row)
constexpr int BQ_BLOCK_K = 64;   // bf16 per k-block = one 128-byte swizzle row
constexpr int BQ_K_BLOCKS = SCAN_DIM / BQ_BLOCK_K;  // 18
constexpr int BQ_UMMA_K = 16;    // bf16 per tcgen05.mma
constexpr int BQ_STAGES = 4;
constexpr int BQ_A_BYTES = BQ_M * BQ_BLOCK_K * 2;   // 16,384
constexpr int BQ_B_BYTES = BQ_N * BQ_BLOCK_K * 2;   // 32,768
constexpr int BQ_STAGE_BYTES = BQ_A_BYTES + BQ_B_BYTES;
constexpr int BQ_THREADS = 192;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int BQ_HEADER = 2048;  // barriers, TMEM base, thresholds
constexpr int BQ_QUEUE_SMEM = 4 * 256 * 12;  // = BQ_QUEUE_BYTES (the epilogue warps' hit queues)
constexpr int BQ_SMEM_BYTES = BQ_HEADER + BQ_STAGES * BQ_STAGE_BYTES + 1024 + BQ_QUEUE_SMEM;  // + alignment slack
constexpr int BQ_TMEM_COLS = 512;
constexpr int BQ_SAMPLE_STRIDE_MAX = 256;  // pass A samples at least 1/256 of the rows (host picks the ratio)
// |cos^ - cos| <= e_d + e_q (1 + e_d) + slack, where e_d = max over rows of ||d^ - d|| / ||d|| and
// e_q = ||q^ - q|| / ||q|| are the MEASURED bf16 rounding-error norms (Cauchy-Schwarz on
// q.(d^-d) + (q^-q).d^) and the slack covers the tensor core's fp32 accumulation of 1152 exact
// bf16 x bf16 products (<= 1152 * 2^-23 relative to ||q^|| ||d^||) and the fp32 path's own 3e-7.
constexpr float BQ_ACCUM_SLACK = 2.0e-4f;

// idesc for kind::f16: D=F32, A=B=BF16, both K-major, N=256, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t BQ_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BQ_N >> 3) << 17) | ((BQ_M >> 4) << 24);

constexpr int BQ_FLAG_OVERFLOW = 1;    // candidate list overflowed its capacity
constexpr int BQ_FLAG_BAD_QUERY = 2;   // zero / non-finite query norm

// ---- PTX wrappers ---------------------------------------------------------------------

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4, LBO=1, SBO=64, version 1, layout SWIZZLE_128B=2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- store / query preparation -----------------------------------------------------------

// The bf16 copy is stored PRE-TILED: block (tile T, k-block kb) = 128 rows x 64 bf16 = one
// contiguous 16 KB span at ((T*18 + kb) * 128 + r) * 64, i.e. exactly one TMA box.  Every
// operand load is then a sequential 16 KB burst and the whole store streams in address order
// (a row-major copy makes each box 128 separate 128-byte pieces 2304 B apart, which held the
// contraction to ~50 % of DRAM efficiency).
__device__ __forceinline__ size_t tiled_offset(long long row, int k) {  // element offset of (row, k)
    const long long tile = row / BQ_M;
    const int r = static_cast<int>(row - tile * BQ_M);
    const int kb = k / BQ_BLOCK_K, kk = k - kb * BQ_BLOCK_K;
    return ((static_cast<size_t>(tile) * BQ_K_BLOCKS + kb) * BQ_M + r) * BQ_BLOCK_K + kk;
}

// one warp per row: bf16 copy of the UNIT-NORMALISED row n = d/||d|| (so the contraction yields
// u = q^ . n^ directly and the epilogue needs no per-row scale), the count of rows whose norm is 0
// or not finite (stored as NaN: never pass a >= test), and the largest rounding-error norm
// ||n^ - n|| over the store (float32 n; its own 2^-23-level error is inside BQ_ACCUM_SLACK).
// Converts rows [row_begin, row_end); `rows` is biased so that rows + r * 1152 addresses row r (the source may
// be the resident store, its host-resident tier, or the caller's buffer of rows being appended).
__global__ void __launch_bounds__(256) build_bf16_store_kernel(const float *__restrict__ rows, long long row_begin,
                                                               long long row_end, __nv_bfloat16 *__restrict__ out,
                                                               unsigned long long *__restrict__ bad_rows,
                                                               unsigned int *__restrict__ max_row_err_bits) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    unsigned bad = 0;
    float worst = 0.f;
    for (long long r = row_begin + warp; r < row_end; r += warps) {
        const float4 *src = reinterpret_cast<const float4 *>(rows + r * SCAN_DIM);
        float4 v[SCAN_CHUNKS];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < SCAN_CHUNKS; j++) {
            v[j] = ldg_stream(src + lane + 32 * j);
            ss = fmaf(v[j].x, v[j].x, ss);
            ss = fmaf(v[j].y, v[j].y, ss);
            ss = fmaf(v[j].z, v[j].z, ss);
            ss = fmaf(v[j].w, v[j].w, ss);
        }
        ss = warp_sum(ss);
        float inv = 1.0f / sqrtf(ss);
        const bool ok = ss > 0.f && isfinite(inv) && isfinite(ss);
        if (!ok) inv = __int_as_float(0x7fc00000);  // NaN row
        float es = 0.f;
#pragma unroll
        for (int j = 0; j < SCAN_CHUNKS; j++) {
            const float x = v[j].x * inv, y = v[j].y * inv, z = v[j].z * inv, w = v[j].w * inv;
            __nv_bfloat162 lo = __floats2bfloat162_rn(x, y), hi = __floats2bfloat162_rn(z, w);
            const float2 rl = __bfloat1622float2(lo), rh = __bfloat1622float2(hi);
            es = fmaf(rl.x - x, rl.x - x, es);
            es = fmaf(rl.y - y, rl.y - y, es);
            es = fmaf(rh.x - z, rh.x - z, es);
            es = fmaf(rh.y - w, rh.y - w, es);
            uint2 packed;
            packed.x = *reinterpret_cast<uint32_t *>(&lo);
            packed.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(out + tiled_offset(r, (lane + 32 * j) * 4)) = packed;
        }
        es = warp_sum(es);
        if (lane == 0) {
            if (!ok) bad++;
            else worst = fmaxf(worst, sqrtf(es) * 1.0001f);  // ||n^ - n||, ||n|| = 1, rounded up
        }
    }
    if (lane == 0) {
        if (bad) atomicAdd(bad_rows, static_cast<unsigned long long>(bad));
        atomicMax(max_row_err_bits, __float_as_uint(worst));  // non-negative floats order as uints
    }
}

// one CTA per query slot (256 slots): bf16 copy (zeros past nq), ||q||
__global__ void __launch_bounds__(128) prep_queries_kernel(const float *__restrict__ q, int nq,
                                                           __nv_bfloat16 *__restrict__ qb,
                                                           float *__restrict__ q_norm,
                                                           float *__restrict__ q_err) {
    __shared__ float red[8];
    const int slot = blockIdx.x, tid = threadIdx.x;
    float ss = 0.f, es = 0.f;
    for (int i = tid; i < SCAN_DIM; i += 128) {
        const float v = slot < nq ? q[static_cast<size_t>(slot) * SCAN_DIM + i] : 0.f;
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        qb[static_cast<size_t>(slot) * SCAN_DIM + i] = b;
        const float e = __bfloat162float(b) - v;
        ss = fmaf(v, v, ss);
        es = fmaf(e, e, es);
    }
    ss = warp_sum(ss);
    es = warp_sum(es);
    if ((tid & 31) == 0) {
        red[tid >> 5] = ss;
        red[4 + (tid >> 5)] = es;
    }
    __syncthreads();
    if (tid == 0) {
        const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
        q_norm[slot] = nrm;
        q_err[slot] = nrm > 0.f ? sqrtf(red[4] + red[5] + red[6] + red[7]) / nrm * 1.0001f : 0.f;
    }
}

// ---- the contraction ---------------------------------------------------------------------------

// Pass A visits one tile out of every `stride`.  A plain multiple of the stride would alias with
// any periodic layout of the data (a power-of-two stride only ever sees rows r with r mod 1024 < 128:
// whole clusters of a round-robin-clustered store were invisible to the sample and their queries got
// noise-level thresholds), so the tile inside each window of `stride` is picked by a hash of the
// window index.  Windows past the end of the store sample nothing (their rows fail the range test).
__device__ __forceinline__ int sampled_tile(int t, int stride) {
    if (stride <= 1) return t;
    const uint32_t h = (static_cast<uint32_t>(t) * 0x9E3779B1u) >> 8;
    return t * stride + static_cast<int>(h % static_cast<uint32_t>(stride));
}

struct BatchGemmArgs {
    const float *thr;           // [256] FILTER: keep iff u >= thr[q]
    float *scores;              // DUMP: [sample_groups][256] group maxima of u
    unsigned int *cand_count;   // FILTER: [256]
    unsigned int *cand_rows;    // FILTER: [256][cand_cap]
    float *cand_u;              // FILTER: [256][cand_cap] the candidate's u (for the second, tighter threshold)
    const uint32_t *mask;       // nullable admission bitset (folder pre-filter, idb:1509-1530)
    long long n;
    long long sample_groups;    // DUMP: number of 32-row groups (= 4 per sampled tile)
    int total_tiles;            // ceil(n / 128)
    int tile_stride;            // 1 (FILTER) or the sampling stride (DUMP)
    int cand_cap;
    unsigned int *tile_counter; // CTA-pair kernel: dynamic tile scheduler, zeroed before the launch
    int static_tiles;           // CTA-pair kernel: 1 = static interleave (pair p takes tiles p, p + pairs, ...)
    unsigned long long *clock;  // nullable (profiling): [0] += SM cycles, [1] += nanoseconds, [2] += 1 per launch,
                                // taken by CTA 0 around its whole run: cycles / ns = the SM clock the launch ran at
};

// ---- filter epilogue ---------------------------------------------------------------------------
// Per 32 accumulator columns of one row: a 32-bit hit mask (one FSETP + one predicated OR per
// element, no branches).  Hits are rare (a few per warp per tile) but each one needs a slot in
// its query's candidate list, i.e. a global atomicAdd whose ~700-cycle round trip would sit in
// the epilogue's critical path: with hits in most warp-chunks (sparse pass-A sampling) that made
// the whole contraction epilogue-bound (6.0 ms instead of 4.5 ms, profiles/).  So hits go into a
// per-warp shared-memory queue (warp prefix sum, no global traffic) that is flushed 32 entries
// at a time — one atomic per lane, all in flight together — when it fills up and at the end.
// Dense bursts (a threshold of -inf keeps everything) bypass the queue.
constexpr int BQ_QUEUE_CAP = 256;     // entries per epilogue warp
constexpr int BQ_QUEUE_BURST = 64;    // a chunk with more hits than this goes straight to global memory
constexpr int BQ_QUEUE_BYTES = 4 * BQ_QUEUE_CAP * 12;   // 4 epilogue warps x (row, query, u)
static_assert(BQ_QUEUE_BYTES == BQ_QUEUE_SMEM, "queue size and shared-memory budget disagree");

struct HitQueue {
    uint32_t *row;   // [BQ_QUEUE_CAP] this warp's slice of shared memory
    uint32_t *qry;
    uint32_t *u;
    int fill;        // warp-uniform
};

__device__ __forceinline__ void push_candidate(const BatchGemmArgs &a, uint32_t q, uint32_t row, uint32_t u_bits) {
    const unsigned pos = atomicAdd(a.cand_count + q, 1u);
    if (pos < static_cast<unsigned>(a.cand_cap)) {
        a.cand_rows[static_cast<size_t>(q) * a.cand_cap + pos] = row;
        a.cand_u[static_cast<size_t>(q) * a.cand_cap + pos] = __uint_as_float(u_bits);
    }
}

__device__ __forceinline__ void flush_queue(HitQueue &hq, const BatchGemmArgs &a, int lane) {
    __syncwarp();
    for (int i = lane; i < hq.fill; i += 32) push_candidate(a, hq.qry[i], hq.row[i], hq.u[i]);
    __syncwarp();
    hq.fill = 0;
}

__device__ __forceinline__ void filter_chunk(const uint32_t (&r)[32], const float *thr, bool row_ok, uint32_t row,
                                             int q0, const BatchGemmArgs &a, HitQueue &hq, int lane) {
    uint32_t hits = 0;
#pragma unroll
    for (int j = 0; j < 32; j++)
        if (__uint_as_float(r[j]) >= thr[j]) hits |= 1u << j;
    if (!row_ok) hits = 0;
    if (!__any_sync(FULL_MASK, hits != 0)) return;
    const int cnt = __popc(hits);
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    if (total > BQ_QUEUE_BURST) {
#pragma unroll   // static indices: r[] must stay in registers
        for (int j = 0; j < 32; j++)
            if ((hits >> j) & 1u) push_candidate(a, static_cast<uint32_t>(q0 + j), row, r[j]);
        return;
    }
    if (hq.fill + total > BQ_QUEUE_CAP) flush_queue(hq, a, lane);
    int at = hq.fill + incl - cnt;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if ((hits >> j) & 1u) {
            hq.row[at] = row;
            hq.qry[at] = static_cast<uint32_t>(q0 + j);
            hq.u[at] = r[j];
            at++;
        }
    }
    hq.fill += total;
}

template <bool DUMP>
__global__ void __launch_bounds__(BQ_THREADS, 1) batch_gemm_kernel(const __grid_constant__ CUtensorMap map_rows,
                                                                   const __grid_constant__ CUtensorMap map_q,
                                                                   const BatchGemmArgs a) {
    extern __shared__ __align__(16) uint8_t bq_smem_raw[];
    // header: barriers | tmem base | thresholds ; tiles start at the next 1024-byte boundary
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(bq_smem_raw);
    uint64_t *empty_bar = full_bar + BQ_STAGES;
    uint64_t *tmem_full = empty_bar + BQ_STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;          // [2]
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *thr_s = reinterpret_cast<float *>(bq_smem_raw + 1024);  // [256]
    const uint32_t raw_addr = smem_u32(bq_smem_raw);
    const uint32_t tiles_addr = (raw_addr + BQ_HEADER + 1023u) & ~1023u;
    uint8_t *tiles = bq_smem_raw + (tiles_addr - raw_addr);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int eff_tiles = (a.total_tiles + a.tile_stride - 1) / a.tile_stride;   // tiles this launch visits
    const int first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < BQ_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);   // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (!DUMP)
        for (int i = tid; i < BQ_N; i += BQ_THREADS) thr_s[i] = a.thr[i];
    if (warp == 1) tmem_alloc(tmem_base_slot, BQ_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int t = first; t < eff_tiles; t += step) {
                const int tile128 = sampled_tile(t, a.tile_stride);
                for (int kb = 0; kb < BQ_K_BLOCKS; kb++) {
                    mbar_wait(&empty_bar[s], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], BQ_STAGE_BYTES);
                    uint8_t *stage = tiles + s * BQ_STAGE_BYTES;
                    tma_load_2d(stage, &map_rows, 0, (tile128 * BQ_K_BLOCKS + kb) * BQ_M, &full_bar[s]);
                    tma_load_2d(stage + BQ_A_BYTES, &map_q, kb * BQ_BLOCK_K, 0, &full_bar[s]);
                    if (++s == BQ_STAGES) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = first; t < eff_tiles; t += step, it++) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BQ_N);
                for (int kb = 0; kb < BQ_K_BLOCKS; kb++) {
                    mbar_wait(&full_bar[s], phase);
                    tc_fence_after();
                    const uint32_t a_addr = tiles_addr + s * BQ_STAGE_BYTES;
                    const uint32_t b_addr = a_addr + BQ_A_BYTES;
#pragma unroll
                    for (int k = 0; k < BQ_BLOCK_K / BQ_UMMA_K; k++) {
                        umma_bf16(tmem_d, umma_desc_sw128(a_addr + k * BQ_UMMA_K * 2),
                                  umma_desc_sw128(b_addr + k * BQ_UMMA_K * 2), BQ_IDESC,
                                  static_cast<uint32_t>((kb | k) != 0));
                    }
                    umma_commit(&empty_bar[s]);   // frees the smem stage when these MMAs retire
                    if (++s == BQ_STAGES) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tmem_full[acc]);     // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====
        const int lane_base = 32 * (warp & 3);
        HitQueue hq;
        hq.row = reinterpret_cast<uint32_t *>(tiles + BQ_STAGES * BQ_STAGE_BYTES) + (warp & 3) * 3 * BQ_QUEUE_CAP;
        hq.qry = hq.row + BQ_QUEUE_CAP;
        hq.u = hq.qry + BQ_QUEUE_CAP;
        hq.fill = 0;
        int it = 0;
        for (int t = first; t < eff_tiles; t += step, it++) {
            const int acc = it & 1;
            const long long row = static_cast<long long>(sampled_tile(t, a.tile_stride)) * BQ_M + lane_base + lane;
            const bool row_ok = row < a.n && row_admitted(a.mask, row);
            mbar_wait(&tmem_full[acc], static_cast<uint32_t>(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) +
                                   static_cast<uint32_t>(acc * BQ_N);
#pragma unroll 1
            for (int c = 0; c < BQ_N / 32; c++) {
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_ld_wait();
                if (DUMP) {
                    // max over this warp's 32 rows for each of the 32 columns, column L ending in lane L
                    float m[32];
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        m[j] = row_ok ? __uint_as_float(r[j]) : __int_as_float(0xff800000);
#pragma unroll
                    for (int half = 16; half >= 1; half >>= 1) {
                        const bool upper = (lane & half) != 0;
#pragma unroll
                        for (int j = 0; j < half; j++) {
                            const float send = upper ? m[j] : m[j + half];
                            const float keep = upper ? m[j + half] : m[j];
                            m[j] = fmaxf(keep, __shfl_xor_sync(FULL_MASK, send, half));  // fmaxf drops NaN
                        }
                    }
                    const long long group = static_cast<long long>(t) * 4 + (warp & 3);
                    if (group < a.sample_groups) a.scores[group * BQ_N + c * 32 + lane] = m[0];
                } else {
                    filter_chunk(r, thr_s + c * 32, row_ok, static_cast<uint32_t>(row), c * 32, a, hq, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (!DUMP) flush_queue(hq, a, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BQ_TMEM_COLS);
}

// ---- the contraction, CTA-pair form (cta_group::2) ------------------------------------------------
// Two CTAs of a cluster (one TPC) work on 256 rows x 256 queries: each loads its own 128
// rows of D^ and HALF of the query block (128 queries), and the leader's single thread issues
// UMMA 256x256x16 that reads both CTAs' shared memory and writes each CTA's TMEM.  Per 128
// rows an SM now ingests 295 KB of rows + 295 KB of queries instead of 295 + 590 KB: the
// 1-CTA kernel is bound by that ingest (ncu: 69 GB through the SMs' L2 ports at 11.9 TB/s,
// tensor pipe 61 % busy), not by HBM or the tensor pipe.
// The kernel is instantiated for NQ = 64, 128 and 256 queries per pass (UMMA N).  At NQ = 256 the
// pass is tensor-bound; with fewer queries the same MMAs shrink with N (UMMA 256 x NQ x 16), the
// query block re-read per tile shrinks too, and the pass becomes a pure HBM stream of the bf16
// store (23 GB at 10M rows): a handful of interactive queries — or ONE — costs ~3.3 ms instead of
// one 6.2 ms float32 scan each, with identical results after the re-rank.  Smaller stages leave
// room for a deeper ring.
constexpr int BP_A_BYTES = BQ_M * BQ_BLOCK_K * 2;          // 16,384: this CTA's 128 rows
constexpr int BP_SMEM_LIMIT = 232448;                      // 227 KB per CTA
template <int NQ>
struct PairCfg {
    static_assert(NQ == 64 || NQ == 128 || NQ == 256, "queries per pass");
    static constexpr int B_BYTES = (NQ / 2) * BQ_BLOCK_K * 2;   // this CTA's NQ/2 queries, one k-block
    // NQ = 64: this CTA's 32 queries are 73,728 B over all 18 k-blocks — they stay RESIDENT in shared memory for the
    // whole launch (loaded once) and the ring carries rows only.  The pass is an HBM stream of the bf16 store; with
    // the query block re-read per tile it pushed 25 % more bytes through L2 than it read from DRAM (28.8 vs 23.0 GB
    // per launch, profiles/r02_batch_gemm_nq64_ncu.md) and twice the TMA operations.
    static constexpr bool Q_RESIDENT = NQ == 64;
    static constexpr int Q_BYTES = Q_RESIDENT ? BQ_K_BLOCKS * B_BYTES : 0;
    static constexpr int STAGE_BYTES = BP_A_BYTES + (Q_RESIDENT ? 0 : B_BYTES);
    static constexpr int FIT = (BP_SMEM_LIMIT - 64 - BQ_HEADER - 1024 - BQ_QUEUE_SMEM - Q_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > 10 ? 10 : FIT;          // 6 / 8 / 8 for NQ = 256 / 128 / 64
    static constexpr int SMEM_BYTES = BQ_HEADER + Q_BYTES + STAGES * STAGE_BYTES + 1024 + BQ_QUEUE_SMEM;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((NQ >> 3) << 17) | ((256u >> 4) << 24);
    static constexpr int TMEM_COLS = 2 * NQ;                    // two accumulator buffers
    static_assert(STAGES * 16 + 32 + 8 <= 448, "barriers must fit below the query barrier");
    static_assert(Q_BYTES % 1024 == 0, "the ring must stay 1024-byte aligned behind the resident queries");
};
constexpr uint32_t BP_PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// TMA load whose completion bytes are credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & BP_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
// same, with an L2 eviction-priority hint (the bf16 rows stream through once; the query block is the
// data worth keeping in L2)
__device__ __forceinline__ void tma_load_2d_pair_hint(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                                      uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & BP_PEER_MASK), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// when the MMAs issued so far retire, arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint64_t *bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    // default semantics (release at CTA scope): what is being ordered is this warp's TMEM reads,
    // already complete (tcgen05.wait::ld + fence); a cluster-scope release would also wait for
    // the filter's global atomics to drain (ERRBAR, visible in the first profile)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// store a 32-bit value at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void st_shared_on_cta(const void *local, uint32_t cta, uint32_t v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(cta));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}
// arrive on the barrier at this offset in CTA `cta`, releasing this thread's earlier (remote) stores
__device__ __forceinline__ void mbar_arrive_release_on_cta(uint64_t *bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

template <bool DUMP, int NQ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BQ_THREADS, 1)
    batch_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_qhalf,
                           const BatchGemmArgs a) {
    using CFG = PairCfg<NQ>;
    constexpr int BP_STAGES = CFG::STAGES, BP_STAGE_BYTES = CFG::STAGE_BYTES;
    extern __shared__ __align__(16) uint8_t bq_smem_raw[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(bq_smem_raw);   // waited on in the leader only
    uint64_t *empty_bar = full_bar + BP_STAGES;
    uint64_t *tmem_full = empty_bar + BP_STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;          // [2], the leader's copy counts both CTAs' epilogues
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    // Dynamic tile scheduler.  The leader's producer thread claims pair tiles from a global counter
    // (CTA pairs drift apart: with a static interleave the fastest SMs idled ~10 % of the launch,
    // ncu smsp__cycles_active vs sm__cycles_elapsed) and publishes each claim in BOTH CTAs: a ring
    // of BP_TQ tile indices guarded by one mbarrier per entry.  Everyone who needs the tile index
    // (the peer's producer, the MMA issuer, all epilogue warps) waits on that entry; -1 ends the
    // launch.  An entry is reused BP_TQ = 4 iterations later, by which time its readers are done:
    // the producer cannot run ahead of iteration it-1's MMAs (the operand ring is shorter than a
    // tile), those waited for the accumulator freed at the END of epilogue it-3.
    constexpr int BP_TQ = 4;
    uint64_t *q_bar = reinterpret_cast<uint64_t *>(bq_smem_raw + 448);   // resident queries landed (leader's copy counts)
    uint64_t *tq_full = reinterpret_cast<uint64_t *>(bq_smem_raw + 512);
    volatile int *tile_ring = reinterpret_cast<volatile int *>(bq_smem_raw + 512 + BP_TQ * 8);
    float *thr_s = reinterpret_cast<float *>(bq_smem_raw + 1024);
    const uint32_t raw_addr = smem_u32(bq_smem_raw);
    const uint32_t q_addr = (raw_addr + BQ_HEADER + 1023u) & ~1023u;     // resident query block (Q_RESIDENT), then the ring
    const uint32_t tiles_addr = q_addr + CFG::Q_BYTES;
    uint8_t *q_tiles = bq_smem_raw + (q_addr - raw_addr);
    uint8_t *tiles = q_tiles + CFG::Q_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    long long clk0 = 0;
    unsigned long long ns0 = 0;
    if (a.clock && blockIdx.x == 0 && tid == 0) {
        clk0 = clock64();
        ns0 = global_timer_ns();
    }
    const int total_pair_tiles = (a.total_tiles + 1) / 2;                       // 256 rows each
    const int eff_tiles = (total_pair_tiles + a.tile_stride - 1) / a.tile_stride;

    if (tid == 0) {
        for (int s = 0; s < BP_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 8);   // 4 epilogue warps in each of the two CTAs
        }
        for (int i = 0; i < BP_TQ; i++) mbar_init(&tq_full[i], 1);
        mbar_init(q_bar, 1);
        mbar_fence_init();
    }
    if (!DUMP)
        for (int i = tid; i < NQ; i += BQ_THREADS) thr_s[i] = a.thr[i];
    if (warp == 1) tmem_alloc_pair(tmem_base_slot, CFG::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();    // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs); the leader's also schedules =====
        if (lane == 0) {
            const uint64_t stream_policy = l2_policy_evict_first();
            unsigned static_next = blockIdx.x >> 1;
            auto claim = [&]() {
                unsigned t;
                if (a.static_tiles) {
                    t = static_next;
                    static_next += gridDim.x >> 1;
                } else {
                    t = atomicAdd(a.tile_counter, 1u);
                }
                return t < static_cast<unsigned>(eff_tiles) ? static_cast<int>(t) : -1;
            };
            if (CFG::Q_RESIDENT) {
                // this CTA's half of the query block, all 18 k-blocks, once; both CTAs' bytes are credited to the
                // leader's barrier, which the MMA issuer waits on before its first instruction (and always waits
                // on, so no CTA can leave while a copy into its shared memory is still in flight)
                if (leader) mbar_arrive_expect_tx(q_bar, 2 * CFG::Q_BYTES);
                for (int kb = 0; kb < BQ_K_BLOCKS; kb++)
                    tma_load_2d_pair(q_tiles + kb * CFG::B_BYTES, &map_qhalf, kb * BQ_BLOCK_K,
                                     static_cast<int>(rank) * (NQ / 2), q_bar);
            }
            int s = 0;
            uint32_t phase = 0;
            int next = leader ? claim() : 0;
            for (int it = 0;; it++) {
                const int slot = it & (BP_TQ - 1);
                int t;
                if (leader) {
                    t = next;
                    tile_ring[slot] = t;
                    st_shared_on_cta(const_cast<int *>(tile_ring + slot), 1, static_cast<uint32_t>(t));
                    mbar_arrive(&tq_full[slot]);
                    mbar_arrive_release_on_cta(&tq_full[slot], 1);
                    if (t >= 0) next = claim();   // claimed ahead: the latency hides behind this tile's loads
                } else {
                    mbar_wait_cluster(&tq_full[slot], static_cast<uint32_t>(it / BP_TQ) & 1u);
                    t = tile_ring[slot];
                }
                if (t < 0) break;
                const int tile128 = sampled_tile(t, a.tile_stride) * 2 + static_cast<int>(rank);
                for (int kb = 0; kb < BQ_K_BLOCKS; kb++) {
                    mbar_wait(&empty_bar[s], phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * BP_STAGE_BYTES);  // both CTAs' bytes
                    uint8_t *stage = tiles + s * BP_STAGE_BYTES;
                    tma_load_2d_pair_hint(stage, &map_rows, 0, (tile128 * BQ_K_BLOCKS + kb) * BQ_M, &full_bar[s],
                                          stream_policy);
                    if (!CFG::Q_RESIDENT)
                        tma_load_2d_pair(stage + BP_A_BYTES, &map_qhalf, kb * BQ_BLOCK_K,
                                         static_cast<int>(rank) * (NQ / 2), &full_bar[s]);
                    if (++s == BP_STAGES) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            if (CFG::Q_RESIDENT) {
                mbar_wait_cluster(q_bar, 0);
                tc_fence_after();
            }
            for (int it = 0;; it++) {
                mbar_wait(&tq_full[it & (BP_TQ - 1)], static_cast<uint32_t>(it / BP_TQ) & 1u);
                if (tile_ring[it & (BP_TQ - 1)] < 0) break;
                const int acc = it & 1;
                mbar_wait_cluster(&tmem_empty[acc], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * NQ);
                for (int kb = 0; kb < BQ_K_BLOCKS; kb++) {
                    mbar_wait_cluster(&full_bar[s], phase);
                    tc_fence_after();
                    const uint32_t a_addr = tiles_addr + s * BP_STAGE_BYTES;
                    const uint32_t b_addr = CFG::Q_RESIDENT ? q_addr + kb * CFG::B_BYTES : a_addr + BP_A_BYTES;
#pragma unroll
                    for (int k = 0; k < BQ_BLOCK_K / BQ_UMMA_K; k++) {
                        umma_bf16_pair(tmem_d, umma_desc_sw128(a_addr + k * BQ_UMMA_K * 2),
                                       umma_desc_sw128(b_addr + k * BQ_UMMA_K * 2), CFG::IDESC,
                                       static_cast<uint32_t>((kb | k) != 0));
                    }
                    umma_commit_pair(&empty_bar[s]);   // frees this stage in both CTAs
                    if (++s == BP_STAGES) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_pair(&tmem_full[acc]);     // accumulators complete in both CTAs
            }
        }
    } else {
        // ===== epilogue (both CTAs): warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====
        const int lane_base = 32 * (warp & 3);
        HitQueue hq;
        hq.row = reinterpret_cast<uint32_t *>(tiles + BP_STAGES * BP_STAGE_BYTES) + (warp & 3) * 3 * BQ_QUEUE_CAP;
        hq.qry = hq.row + BQ_QUEUE_CAP;
        hq.u = hq.qry + BQ_QUEUE_CAP;
        hq.fill = 0;
        for (int it = 0;; it++) {
            mbar_wait_cluster(&tq_full[it & (BP_TQ - 1)], static_cast<uint32_t>(it / BP_TQ) & 1u);
            const int t = tile_ring[it & (BP_TQ - 1)];
            if (t < 0) break;
            const int acc = it & 1;
            const long long t128 = static_cast<long long>(sampled_tile(t, a.tile_stride)) * 2 + rank;   // 128-row tile index
            const long long row = t128 * BQ_M + lane_base + lane;
            const bool row_ok = row < a.n && row_admitted(a.mask, row);
            mbar_wait(&tmem_full[acc], static_cast<uint32_t>(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) +
                                   static_cast<uint32_t>(acc * NQ);
#pragma unroll 1
            for (int c = 0; c < NQ / 32; c++) {
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_ld_wait();
                if (DUMP) {
                    float m[32];
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        m[j] = row_ok ? __uint_as_float(r[j]) : __int_as_float(0xff800000);
#pragma unroll
                    for (int half = 16; half >= 1; half >>= 1) {
                        const bool upper = (lane & half) != 0;
#pragma unroll
                        for (int j = 0; j < half; j++) {
                            const float send = upper ? m[j] : m[j + half];
                            const float keep = upper ? m[j + half] : m[j];
                            m[j] = fmaxf(keep, __shfl_xor_sync(FULL_MASK, send, half));
                        }
                    }
                    const long long group = (static_cast<long long>(t) * 2 + rank) * 4 + (warp & 3);
                    if (group < a.sample_groups) a.scores[group * BQ_N + c * 32 + lane] = m[0];
                } else {
                    filter_chunk(r, thr_s + c * 32, row_ok, static_cast<uint32_t>(row), c * 32, a, hq, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_on_cta(&tmem_empty[acc], 0);   // the leader's barrier
        }
        if (!DUMP) flush_queue(hq, a, lane);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();    // no CTA leaves (or frees TMEM) while its peer may still touch it
    if (warp == 1) tmem_dealloc_pair(tmem_base, CFG::TMEM_COLS);
    if (a.clock && blockIdx.x == 0 && tid == 0) {
        atomicAdd(a.clock, static_cast<unsigned long long>(clock64() - clk0));
        atomicAdd(a.clock + 1, global_timer_ns() - ns0);
        atomicAdd(a.clock + 2, 1ull);
    }
}

// ---- per-query selection in two levels ----------------------------------------------------
// Both the threshold step (k-th best sampled u) and the re-rank (k best exact distances) are
// "top-k of one query's items".  Level 1: grid (query, part) — each CTA keeps warp-level
// candidate lists over its share of the items and writes one sorted list of 32*KPL keys.
// Level 2: one CTA per query sorts the parts' lists and either derives the threshold or
// decodes the final result.  Keys make the outcome independent of how items were split.
constexpr int BQ_SEL_WARPS = 8;
constexpr int BQ_SEL_THREADS = BQ_SEL_WARPS * 32;
constexpr int BQ_THR_PARTS = 2;      // parts per query over the sampled group maxima
constexpr int BQ_RERANK_PARTS = 4;   // parts per query over the candidate rows

template <int KPL>
__device__ __forceinline__ void write_part_list(WarpTopK<KPL> &top, uint64_t *scratch, uint64_t *dst, int tid,
                                                int warp, int lane) {
    top.dump(scratch + warp * 32 * KPL, lane);
    __syncthreads();
    block_bitonic_sort(scratch, BQ_SEL_WARPS * 32 * KPL, tid, BQ_SEL_THREADS);
    for (int i = tid; i < 32 * KPL; i += BQ_SEL_THREADS) dst[i] = scratch[i];
}

// ---- k-th largest of a query's values: radix select -------------------------------------------------
// Thresholds need one order statistic per query (the k-th largest of ~10^4 floats), not a sorted
// list.  Four passes over the values, most significant byte first: a shared-memory histogram of
// the byte among the values that match the digits found so far, then one warp walks the 256 bins
// from the top to find the bin holding the k-th largest.  Exact (all 32 bits of the orderable
// encoding), ~10 us for 256 queries; the register top-k lists used by the scan took 0.2-0.3 ms
// here because with k = 100 most of a warp's few hundred items are admissions.
// A CTA serves QPC queries: thread t works for query q0 + t % QPC, so that values laid out
// query-minor (scores[group][256]) are still read a sector at a time.
constexpr int BQ_SELECT_PASSES = 3;   // 24 of 32 bits: the bucket's lower edge is <= the k-th largest by < 2^-15 relative

template <int QPC>
struct SelectSmem {
    unsigned hist[QPC][256];
    unsigned prefix[QPC];      // digits found so far (high bits of the answer)
    unsigned remaining[QPC];   // rank still to be located inside the current prefix
    unsigned short_of_k[QPC];  // fewer than k valid values
};

// valid floats -> uint32 ascending; invalid (NaN, -inf) values are skipped
__device__ __forceinline__ bool select_key(float v, uint32_t &key) {
    key = f32_orderable(v);
    return v == v && v > __int_as_float(0xff800000);
}

// `load(ql, i)` returns the i-th value of local query ql; `count(ql)` its number of values.
// Returns (to every thread) a value t <= the k-th largest valid value of query ql such that at
// least k values are >= t, or -inf when there are fewer than k valid values.  The callers only
// need such a lower bound, so the last byte is not resolved.
template <int QPC, int THREADS, typename Load, typename Count>
__device__ __forceinline__ float radix_select_kth_largest(SelectSmem<QPC> &sm, int k, int ql_out, Load load, Count count) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ql = tid % QPC, slot = tid / QPC;
    constexpr int SLOTS = THREADS / QPC;
    constexpr int U = 8;   // independent loads in flight per thread
    if (tid < QPC) {
        sm.prefix[tid] = 0;
        sm.remaining[tid] = static_cast<unsigned>(k);
        sm.short_of_k[tid] = 0;
    }
    const unsigned n_items = count(ql);
#pragma unroll 1
    for (int pass = 0; pass < BQ_SELECT_PASSES; pass++) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < QPC * 256; i += THREADS) (&sm.hist[0][0])[i] = 0;
        __syncthreads();
        const unsigned want = sm.prefix[ql];
        for (unsigned base = slot; base < n_items; base += SLOTS * U) {
            float v[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                const unsigned i = base + j * SLOTS;
                v[j] = i < n_items ? load(ql, i) : __int_as_float(0xff800000);
            }
#pragma unroll
            for (int j = 0; j < U; j++) {
                uint32_t key;
                if (select_key(v[j], key) && (pass == 0 || (key >> ((shift + 8) & 31)) == want))
                    atomicAdd(&sm.hist[ql][(key >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (warp < QPC) {
            // lane L owns bins 255-8L .. 248-8L (descending value order across lanes)
            unsigned c[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c[j] = sm.hist[warp][255 - 8 * lane - j];
                mine += c[j];
            }
            unsigned incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned v = __shfl_up_sync(FULL_MASK, incl, off);
                if (lane >= off) incl += v;
            }
            const unsigned rem = sm.remaining[warp];
            const unsigned owners = __ballot_sync(FULL_MASK, incl >= rem);
            if (owners == 0) {
                if (lane == 0) sm.short_of_k[warp] = 1;   // only possible in pass 0
            } else if (lane == __ffs(owners) - 1) {
                unsigned above = incl - mine;   // values in higher bins
                int j = 0;
#pragma unroll
                for (int t = 0; t < 8; t++)
                    if (t == j && above + c[t] < rem) {
                        above += c[t];
                        j++;
                    }
                sm.prefix[warp] = (sm.prefix[warp] << 8) | static_cast<unsigned>(255 - 8 * lane - j);
                sm.remaining[warp] = rem - above;
            }
        }
        __syncthreads();
    }
    if (sm.short_of_k[ql_out]) return __int_as_float(0xff800000);
    return orderable_f32(sm.prefix[ql_out] << (32 - 8 * BQ_SELECT_PASSES));
}

// pass A -> thresholds: tau = k-th largest sampled group maximum of u; thr = tau - 2E||q||
constexpr int BQ_THR_QPC = 4;
constexpr int BQ_THR_THREADS = 1024;
__global__ void __launch_bounds__(BQ_THR_THREADS) batch_threshold_kernel(const float *__restrict__ scores,
                                                                         long long groups,
                                                                         const float *__restrict__ q_norm,
                                                                         const float *__restrict__ q_err,
                                                                         const unsigned int *__restrict__ max_row_err_bits,
                                                                         int nq, int k, float *__restrict__ thr,
                                                                         float *__restrict__ margin,
                                                                         int *__restrict__ flags) {
    __shared__ SelectSmem<BQ_THR_QPC> sm;
    const int q0 = blockIdx.x * BQ_THR_QPC;
    if (q0 >= nq) {   // padding slots only: never produce candidates, nothing to select
        if (threadIdx.x < BQ_THR_QPC) {
            thr[q0 + threadIdx.x] = __int_as_float(0x7f800000);
            margin[q0 + threadIdx.x] = 0.f;
        }
        return;
    }
    const int mine = threadIdx.x < BQ_THR_QPC ? threadIdx.x : 0;
    const float tau = radix_select_kth_largest<BQ_THR_QPC, BQ_THR_THREADS>(
        sm, k, mine, [&](int ql, unsigned i) { return __ldg(scores + static_cast<long long>(i) * BQ_N + q0 + ql); },
        [&](int) { return static_cast<unsigned>(groups); });
    if (threadIdx.x < BQ_THR_QPC) {
        const int q = q0 + threadIdx.x;
        float t = __int_as_float(0x7f800000), mg = 0.f;   // padding slots (q >= nq) never produce candidates
        if (q < nq) {
            const float qn = q_norm[q];
            int f = 0;
            if (!(qn > 0.f) || !isfinite(qn)) {
                f = BQ_FLAG_BAD_QUERY;
            } else {
                const float ed = __uint_as_float(*max_row_err_bits), eq = q_err[q];
                const float bound = ed + eq * (1.0f + ed) + BQ_ACCUM_SLACK;   // E
                mg = 2.0f * bound * qn;
                t = tau - mg;   // tau = -inf (fewer than k samples): keep everything
            }
            flags[q] = f;
        }
        thr[q] = t;
        margin[q] = mg;
    }
}

// ---- second threshold: exact k-th best u over the WHOLE store -------------------------------------
// Pass B kept every row with u >= tau - 2E||q||, tau being the k-th best u of a SAMPLE.  tau2 = the
// k-th largest u among those candidates is at least as large, k distinct rows have u >= tau2, and
// the same argument as for tau (so cos_k >= tau2/||q|| - E, so every true top-k row has
// u >= tau2 - 2E||q||) lets the re-rank skip every candidate below tau2 - 2E||q||.  That cuts the
// rows whose float32 originals are gathered from ~n*k/n_sample per query to a few times k, which
// is what allows pass A to sample sparsely.  The survivors are compacted into a dense list.
constexpr int BQ_SELECT_THREADS = 256;
__global__ void __launch_bounds__(BQ_SELECT_THREADS) batch_refine_kernel(const float *__restrict__ cand_u,
                                                                         const unsigned int *__restrict__ cand_rows,
                                                                         const unsigned int *__restrict__ cand_count,
                                                                         int cand_cap, int k, int refine,
                                                                         const float *__restrict__ margin,
                                                                         unsigned int *__restrict__ surv_rows,
                                                                         unsigned int *__restrict__ surv_count) {
    __shared__ SelectSmem<1> sm;
    __shared__ unsigned n_surv;
    const int q = blockIdx.x, tid = threadIdx.x;
    const unsigned total = cand_count[q];
    const unsigned count = total > static_cast<unsigned>(cand_cap) ? static_cast<unsigned>(cand_cap) : total;
    const float *u = cand_u + static_cast<size_t>(q) * cand_cap;
    if (tid == 0) n_surv = 0;
    float thr2 = __int_as_float(0xff800000);
    if (refine) {   // tau2 = -inf (fewer than k candidates): re-rank them all
        thr2 = radix_select_kth_largest<1, BQ_SELECT_THREADS>(
                   sm, k, 0, [&](int, unsigned i) { return __ldg(u + i); }, [&](int) { return count; }) -
               margin[q];
    } else {
        __syncthreads();
    }
    const unsigned int *rows = cand_rows + static_cast<size_t>(q) * cand_cap;
    unsigned int *dst = surv_rows + static_cast<size_t>(q) * cand_cap;
    for (unsigned i = tid; i < count; i += BQ_SELECT_THREADS)
        if (__ldg(u + i) >= thr2) dst[atomicAdd(&n_surv, 1u)] = rows[i];
    __syncthreads();
    if (tid == 0) surv_count[q] = n_surv;
}

// ---- exact re-rank: K1's arithmetic over each query's candidate rows ------------------------------
struct RerankArgs {
    const float *rows;            // fp32 store
    const float *rows_hi;         // tiered store: host-resident rows, read zero-copy over PCIe (row_address)
    long long split;
    const float *queries;         // [nq][1152] fp32
    const unsigned int *cand_count;   // pass B's count (overflow check)
    const unsigned int *surv_rows;    // [nq][cand_cap] candidates that passed the second threshold
    const unsigned int *surv_count;   // [nq]
    uint64_t *part_keys;          // [nq][BQ_RERANK_PARTS][32*KPL]
    int cand_cap;
    int k;
    int *flags;                   // |= BQ_FLAG_OVERFLOW
    DecodeArgs dec;               // out arrays are [nq][k]; out_n/out_nan [nq]
    const unsigned long long *bad_rows;  // rows of the store with NaN distance for any query
    ExchangeArgs xchg;            // world > 1: the finish kernel also exchanges each query's list with the
                                  // other shards' GPUs and merges (merge.cuh); flags then hold the OR over shards
};

template <int KPL>
__global__ void __launch_bounds__(BQ_SEL_THREADS) batch_rerank_kernel(const RerankArgs a) {
    __shared__ uint64_t scratch[BQ_SEL_WARPS * 32 * KPL];
    const int q = blockIdx.x, part = blockIdx.y, parts = gridDim.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned count = a.surv_count[q];

    float4 qv[SCAN_CHUNKS];
    const float4 *q4 = reinterpret_cast<const float4 *>(a.queries + static_cast<size_t>(q) * SCAN_DIM);
    float bsum = 0.f;
#pragma unroll
    for (int j = 0; j < SCAN_CHUNKS; j++) {
        qv[j] = __ldg(q4 + lane + 32 * j);
        bsum = fmaf(qv[j].x, qv[j].x, bsum);
        bsum = fmaf(qv[j].y, qv[j].y, bsum);
        bsum = fmaf(qv[j].z, qv[j].z, bsum);
        bsum = fmaf(qv[j].w, qv[j].w, bsum);
    }
    const double sqrt_b = sqrt(static_cast<double>(warp_sum(bsum)));
    const float rsqrt_b = static_cast<float>(1.0 / sqrt_b);

    WarpTopK<KPL> top;
    top.init(a.k, lane);
    unsigned nan_rows = 0;
    const unsigned int *list = a.surv_rows + static_cast<size_t>(q) * a.cand_cap;
    // survivors are few (a few times k): deal them out two at a time over all of the query's warps,
    // each warp keeping two rows' loads in flight (the gather is latency-bound, not bandwidth-bound)
    const unsigned pair_stride = static_cast<unsigned>(parts * BQ_SEL_WARPS) * 2u;
    {
        for (unsigned j0 = (part * BQ_SEL_WARPS + warp) * 2u; j0 < count; j0 += pair_stride) {
            const bool two = j0 + 1 < count;
            const long long pos0 = list[j0];
            const long long pos1 = list[two ? j0 + 1 : j0];
            const float4 *src0 = reinterpret_cast<const float4 *>(row_address(a.rows, a.rows_hi, a.split, pos0, SCAN_DIM));
            const float4 *src1 = reinterpret_cast<const float4 *>(row_address(a.rows, a.rows_hi, a.split, pos1, SCAN_DIM));
            float4 v0[SCAN_CHUNKS], v1[SCAN_CHUNKS];
#pragma unroll
            for (int j = 0; j < SCAN_CHUNKS; j++) v0[j] = ldg_stream(src0 + lane + 32 * j);
#pragma unroll
            for (int j = 0; j < SCAN_CHUNKS; j++) v1[j] = ldg_stream(src1 + lane + 32 * j);
            {
                RowSums sums;
                sums.clear();
#pragma unroll
                for (int j = 0; j < SCAN_CHUNKS; j++) accumulate<METRIC_COSINE>(v0[j], qv[j], sums);
                const float t0 = warp_sum(sums.first());
                const float t1 = warp_sum(sums.second());
                offer_row<KPL, METRIC_COSINE, false>(t0, t1, sqrt_b, rsqrt_b, pos0, top, nan_rows, nullptr, lane);
            }
            if (two) {
                RowSums sums;
                sums.clear();
#pragma unroll
                for (int j = 0; j < SCAN_CHUNKS; j++) accumulate<METRIC_COSINE>(v1[j], qv[j], sums);
                const float t0 = warp_sum(sums.first());
                const float t1 = warp_sum(sums.second());
                offer_row<KPL, METRIC_COSINE, false>(t0, t1, sqrt_b, rsqrt_b, pos1, top, nan_rows, nullptr, lane);
            }
        }
    }
    write_part_list<KPL>(top, scratch, a.part_keys + (static_cast<size_t>(q) * parts + part) * (32 * KPL), tid, warp,
                         lane);
}

// level 2 for the re-rank: merge the parts, decode the k best (or, sharded: exchange this shard's k best
// for the query with the peer GPUs over NVLink and decode the merged list)
constexpr int BQ_FINISH_KEYS = 2048;   // >= parts * 128 and >= next_pow2(world * k)
__global__ void __launch_bounds__(256) batch_rerank_finish_kernel(const RerankArgs a, int parts, int stride) {
    __shared__ uint64_t s[BQ_FINISH_KEYS];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int total = parts * stride, padded = next_pow2(total);
    for (int i = tid; i < padded; i += 256)
        s[i] = i < total ? a.part_keys[static_cast<size_t>(q) * total + i] : KEY_EMPTY;
    block_bitonic_sort(s, padded, tid, 256);
    DecodeArgs d = a.dec;
    d.out_rowids += static_cast<size_t>(q) * a.k;
    d.out_dist += static_cast<size_t>(q) * a.k;
    d.out_n += q;
    if (d.out_nan) d.out_nan += q;
    const bool overflow = a.cand_count[q] > static_cast<unsigned>(a.cand_cap);
    if (a.xchg.world > 1) {
        const int G = a.xchg.world, bank = static_cast<int>(a.xchg.epoch & 1u);
        const int local_flags = a.flags[q] | (overflow ? BQ_FLAG_OVERFLOW : 0);
        for (int i = padded + tid; i < BQ_FINISH_KEYS; i += 256) s[i] = KEY_EMPTY;
        __syncthreads();
        exchange_merge_decode<BQ_RERANK_PARTS * 128>(s, d, a.xchg, 2 * G + bank * G * XCHG_BATCH + q, XCHG_BATCH,
                                                      static_cast<long long>(*a.bad_rows), local_flags, a.flags + q, tid,
                                                      256);
        return;
    }
    int found = 0;
    for (int base = 0; base < a.k; base += 256) {
        const int i = base + tid;
        const bool valid = i < a.k && i < padded && s[i] != KEY_EMPTY;
        if (valid) decode_one(d, i, s[i]);
        found += __syncthreads_count(valid);
    }
    if (tid == 0) {
        *d.out_n = found;
        if (d.out_nan) *d.out_nan = static_cast<int64_t>(*a.bad_rows);
        if (overflow) a.flags[q] |= BQ_FLAG_OVERFLOW;
    }
}

}  // namespace clipdb
