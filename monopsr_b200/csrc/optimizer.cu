// monopsr_b200/csrc/optimizer.cu -- fused multi-tensor train-op.
//
// Replaces slim.learning.create_train_op(total_loss, optimizer, clip_gradient_norm=1.0)
// (core/trainer.py:76-81) with tf.train.AdamOptimizer (beta 0.9/0.999, eps 1e-8) wrapped in
// tf.contrib.opt.MovingAverageOptimizer(average_decay=0.9999) (builders/optimizer_builder.py:56-82):
//   per VARIABLE  g <- g * clip / max(||g||, clip)        (tf.clip_by_norm, not a global norm)
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; p <- p - lr_t m / (sqrt(v) + eps)
//   ema <- ema - (1-decay)(ema - p)
// over one flat fp32 arena (about 100 M parameters): two launches for the whole model.
#include "common.cuh"
#include "../../include/monopsr_b200_net.h"

namespace mpb {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
opt_sumsq_kernel(const mpb_opt_chunk* __restrict__ chunks, const float* __restrict__ grad, float gscale,
                 float* __restrict__ partial) {
    const mpb_opt_chunk ch = chunks[blockIdx.x];
    const float* g = grad + ch.start;
    float acc = 0.f;
    const int n4 = ch.len >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += kOptThreads) {
        const float4 t = g4[i];
        acc = fmaf(t.x * gscale, t.x * gscale, acc);
        acc = fmaf(t.y * gscale, t.y * gscale, acc);
        acc = fmaf(t.z * gscale, t.z * gscale, acc);
        acc = fmaf(t.w * gscale, t.w * gscale, acc);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < ch.len; i += kOptThreads) {
        const float v = g[i] * gscale;
        acc = fmaf(v, v, acc);
    }
    __shared__ float red[kOptThreads / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kOptThreads / 32; i++) s += red[i];
        partial[blockIdx.x] = s;      // one slot per chunk, combined in a FIXED order by the update kernel
    }
}

// ||g||^2 of the variable that chunk `me` belongs to: the chunks of a variable are contiguous in the table, so a
// warp scans outwards from its own chunk and adds the per-chunk partials lane-strided + shuffle tree.  The order
// depends only on the table, so every CTA of a variable -- and every data-parallel rank -- gets the same bits
// (an atomicAdd accumulation made the clip factor, hence the replicas' parameters, differ in the last bits).
__device__ __forceinline__ float tensor_sumsq(const mpb_opt_chunk* __restrict__ chunks, int nchunks, int me,
                                              const float* __restrict__ partial) {
    const int t = chunks[me].tensor;
    int lo = me, hi = me + 1;
    while (lo > 0 && chunks[lo - 1].tensor == t) --lo;
    while (hi < nchunks && chunks[hi].tensor == t) ++hi;
    float acc = 0.f;
    for (int i = lo + (threadIdx.x & 31); i < hi; i += 32) acc += partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// hyper[0] = lr_t = lr * sqrt(1-b2^t)/(1-b1^t) (device resident so a captured CUDA graph can be replayed)
__global__ void __launch_bounds__(kOptThreads)
opt_adam_ema_kernel(const mpb_opt_chunk* __restrict__ chunks, int nchunks, float* __restrict__ param,
                    const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                    float* __restrict__ ema, const float* __restrict__ partial, const float* __restrict__ hyper,
                    float gscale, float clip, float b1, float b2, float eps, float ema_decay) {
    const mpb_opt_chunk ch = chunks[blockIdx.x];
    __shared__ float s_nrm2;
    if (threadIdx.x < 32) {
        const float t = tensor_sumsq(chunks, nchunks, blockIdx.x, partial);
        if (threadIdx.x == 0) s_nrm2 = t;
    }
    __syncthreads();
    const float nrm = sqrtf(s_nrm2);
    const float cf = gscale * clip / fmaxf(nrm, clip);
    const float lr_t = hyper[0];
    // 128-bit main loop (every tensor and chunk starts on a 16-byte boundary), scalar tail
    const int n4 = ch.len >> 2;
    {
        const float4* g4 = reinterpret_cast<const float4*>(grad + ch.start);
        float4* m4 = reinterpret_cast<float4*>(m + ch.start);
        float4* v4 = reinterpret_cast<float4*>(v + ch.start);
        float4* p4 = reinterpret_cast<float4*>(param + ch.start);
        float4* e4 = reinterpret_cast<float4*>(ema + ch.start);
#pragma unroll 2
        for (int i = threadIdx.x; i < n4; i += kOptThreads) {
            const float4 g = __ldcs(g4 + i);
            float4 mi = m4[i], vi = v4[i], pi = p4[i], ei = e4[i];
#define MPB_ADAM(c)                                                   \
            {                                                         \
                const float gc = g.c * cf;                            \
                mi.c = b1 * mi.c + (1.f - b1) * gc;                   \
                vi.c = b2 * vi.c + (1.f - b2) * gc * gc;              \
                pi.c = pi.c - lr_t * mi.c / (sqrtf(vi.c) + eps);      \
                ei.c = ei.c - (1.f - ema_decay) * (ei.c - pi.c);      \
            }
            MPB_ADAM(x) MPB_ADAM(y) MPB_ADAM(z) MPB_ADAM(w)
#undef MPB_ADAM
            m4[i] = mi; v4[i] = vi; p4[i] = pi; e4[i] = ei;
        }
    }
    for (int i = (n4 << 2) + threadIdx.x; i < ch.len; i += kOptThreads) {
        const long j = ch.start + i;
        const float g = grad[j] * cf;
        const float mi = b1 * m[j] + (1.f - b1) * g;
        const float vi = b2 * v[j] + (1.f - b2) * g * g;
        m[j] = mi;
        v[j] = vi;
        const float p = param[j] - lr_t * mi / (sqrtf(vi) + eps);
        param[j] = p;
        const float e = ema[j];
        ema[j] = e - (1.f - ema_decay) * (e - p);
    }
}

}  // namespace mpb

MPB_API int mpb_opt_step_range(int nchunks, const mpb_opt_chunk* chunks, int tensor0, int ntensors, float* param,
                               const float* grad, float* m, float* v, float* ema, float* norm2, const float* hyper,
                               float grad_scale, float clip_norm, float beta1, float beta2, float eps, float ema_decay,
                               void* stream) {
    using namespace mpb;
    if (nchunks <= 0 || !chunks || !param || !grad || !m || !v || !ema || !norm2 || !hyper || tensor0 < 0) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    (void)ntensors;
    opt_sumsq_kernel<<<nchunks, kOptThreads, 0, s>>>(chunks, grad, grad_scale, norm2);
    MPB_LAUNCH_CHECK();
    opt_adam_ema_kernel<<<nchunks, kOptThreads, 0, s>>>(chunks, nchunks, param, grad, m, v, ema, norm2, hyper, grad_scale,
                                                       clip_norm, beta1, beta2, eps, ema_decay);
    MPB_LAUNCH_CHECK();
    return 0;
}

MPB_API int mpb_opt_step(int nchunks, const mpb_opt_chunk* chunks, int ntensors, float* param, const float* grad,
                         float* m, float* v, float* ema, float* norm2, const float* hyper, float grad_scale,
                         float clip_norm, float beta1, float beta2, float eps, float ema_decay, void* stream) {
    return mpb_opt_step_range(nchunks, chunks, 0, ntensors, param, grad, m, v, ema, norm2, hyper, grad_scale, clip_norm,
                              beta1, beta2, eps, ema_decay, stream);
}
