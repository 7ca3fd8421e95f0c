// common.cuh — keys, PTX wrappers and block-level sort shared by the kernels.
//
// Result ordering.  SQLite's ORDER BY distance ASC LIMIT k sorter
// (image_database.py:1572-1573) keeps the k smallest (distance, scan sequence)
// pairs.  Every kernel here carries one 64-bit key per row,
//     key = orderable(float32 distance) << 32 | scan position,
// so that unsigned integer order on keys IS that lexicographic order and all
// selection / merging is bit-exact integer work.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace clipdb {

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;  // "no row" (sorts last)
constexpr uint64_t KEY_DISABLED = 0ull;                // list slot beyond k (never the max)
constexpr unsigned FULL_MASK = 0xFFFFFFFFu;

// float32 -> uint32 whose unsigned order equals the float order.  -0.0 is
// folded into +0.0 first (SQLite compares them equal, ties go to sequence).
__device__ __forceinline__ uint32_t f32_orderable(float d) {
    d = d + 0.0f;
    uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float orderable_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t make_key(float d, uint32_t pos) {
    return (static_cast<uint64_t>(f32_orderable(d)) << 32) | pos;
}

__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }

// ---- shared-memory / mbarrier / bulk-copy (TMA engine) PTX -----------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 1-D bulk copy global -> shared through the TMA engine; completion is
// reported as `bytes` transaction bytes on `bar`.  src, dst and bytes must be
// 16-byte multiples.  (SASS: UBLKCP.)
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s_hint(void *dst, const void *src, uint32_t bytes,
                                                   uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// streaming 128-bit global load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// ---- warp reductions ---------------------------------------------------------

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
    return v;
}

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = umax64(v, __shfl_xor_sync(FULL_MASK, v, off));
    return v;
}

// ---- block-wide bitonic sort of n (power of two) keys in shared memory -------
// Ascending.  All `nthreads` threads of the block must call it.
__device__ __forceinline__ void block_bitonic_sort(uint64_t *s, int n, int tid, int nthreads) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < (n >> 1); i += nthreads) {
                int lo = 2 * i - (i & (stride - 1));  // index with bit `stride` clear
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t a = s[lo], b = s[hi];
                if ((a > b) == up) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace clipdb
