// blend.cuh — K3: weighted query blend, negative prompts, L2 normalise.
//
// Restates the numpy float32 arithmetic of the reference, one CTA per query:
//   image_database.py:1387       e = w1*e1 + w2*e2        (two rounded products, one rounded sum)
//   image_database.py:1390-1395  e /= ||e||  if ||e|| > 0 else e = e1
//   image_database.py:555 / 586-587   e = e - w_i*neg_i   (sequential, list order)
//   image_database.py:557-570 / 590-603  normalise; zero norm -> restore e1, or re-blend
// Every elementwise operation uses the *_rn intrinsics so nvcc cannot contract
// a multiply and an add into one FMA (numpy performs them as separate float32
// operations).  Only the squared-norm summation order differs from numpy's
// BLAS dot, hence a float tolerance on this kernel rather than bit equality.
#pragma once

#include "common.cuh"

namespace clipdb {

constexpr int BLEND_THREADS = 256;
constexpr int BLEND_POSITIVE_ZERO_NORM = 1;
constexpr int BLEND_NEGATIVE_ZERO_NORM = 2;

struct BlendArgs {
    const float *e1;     // [batch][dim]
    const float *e2;     // nullable [batch][dim]
    const float *w;      // [batch][2] normalised weights (used when e2 != null)
    const float *negs;   // nullable [batch][n_neg][dim]
    const float *neg_w;  // [batch][n_neg]
    float *out;          // [batch][dim]
    int32_t *flags;      // nullable [batch]
    int n_neg;
    int dim;
};

// sum over the block; every thread gets the total.  `red` holds 32 floats.
__device__ __forceinline__ float blend_block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < BLEND_THREADS / 32) ? red[lane] : 0.f;
    return warp_sum(t);
}

__global__ void __launch_bounds__(BLEND_THREADS) blend_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) uint8_t blend_smem[];
    float *e = reinterpret_cast<float *>(blend_smem);  // [dim] working vector
    __shared__ float red[32];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int dim = a.dim;
    const float *e1 = a.e1 + static_cast<size_t>(b) * dim;
    const float *e2 = a.e2 ? a.e2 + static_cast<size_t>(b) * dim : nullptr;
    const float w1 = e2 ? a.w[2 * b] : 1.f;
    const float w2 = e2 ? a.w[2 * b + 1] : 0.f;
    float *out = a.out + static_cast<size_t>(b) * dim;
    int flags = 0;

    // positive part
    if (e2) {
        float ss = 0.f;
        for (int i = tid; i < dim; i += BLEND_THREADS) {
            const float m = __fadd_rn(__fmul_rn(w1, e1[i]), __fmul_rn(w2, e2[i]));
            e[i] = m;
            ss = fmaf(m, m, ss);
        }
        const float nrm = __fsqrt_rn(blend_block_sum(ss, red));
        if (nrm > 0.f) {
            for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = __fdiv_rn(e[i], nrm);
        } else {
            flags |= BLEND_POSITIVE_ZERO_NORM;
            for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = e1[i];
        }
    } else {
        for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = e1[i];
    }

    // negative prompts (each thread only touches its own elements: no barrier needed)
    if (a.n_neg > 0 && a.negs) {
        const float *negs = a.negs + static_cast<size_t>(b) * a.n_neg * dim;
        const float *nw = a.neg_w + static_cast<size_t>(b) * a.n_neg;
        float ss = 0.f;
        for (int i = tid; i < dim; i += BLEND_THREADS) {
            float x = e[i];
            for (int j = 0; j < a.n_neg; j++)
                x = __fsub_rn(x, __fmul_rn(nw[j], negs[static_cast<size_t>(j) * dim + i]));
            e[i] = x;
            ss = fmaf(x, x, ss);
        }
        const float nrm = __fsqrt_rn(blend_block_sum(ss, red));
        if (nrm > 0.f) {
            for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = __fdiv_rn(e[i], nrm);
        } else {
            flags |= BLEND_NEGATIVE_ZERO_NORM;
            if (!e2) {
                for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = e1[i];
            } else {
                float s2 = 0.f;
                for (int i = tid; i < dim; i += BLEND_THREADS) {
                    const float m = __fadd_rn(__fmul_rn(w1, e1[i]), __fmul_rn(w2, e2[i]));
                    e[i] = m;
                    s2 = fmaf(m, m, s2);
                }
                const float n2 = __fsqrt_rn(blend_block_sum(s2, red));
                if (n2 > 0.f)
                    for (int i = tid; i < dim; i += BLEND_THREADS) e[i] = __fdiv_rn(e[i], n2);
            }
        }
    }

    for (int i = tid; i < dim; i += BLEND_THREADS) out[i] = e[i];
    if (a.flags && tid == 0) a.flags[b] = flags;
}

}  // namespace clipdb
