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
                 float* __restrict__ norm2) {
    const mpb_opt_chunk ch = chunks[blockIdx.x];
    const float* g = grad + ch.start;
    float acc = 0.f;
    for (int i = threadIdx.x; i < ch.len; i += kOptThreads) {
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
        atomicAdd(&norm2[ch.tensor], s);
    }
}

// hyper[0] = lr_t = lr * sqrt(1-b2^t)/(1-b1^t) (device resident so a captured CUDA graph can be replayed)
__global__ void __launch_bounds__(kOptThreads)
opt_adam_ema_kernel(const mpb_opt_chunk* __restrict__ chunks, float* __restrict__ param,
                    const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                    float* __restrict__ ema, const float* __restrict__ norm2, const float* __restrict__ hyper,
                    float gscale, float clip, float b1, float b2, float eps, float ema_decay) {
    const mpb_opt_chunk ch = chunks[blockIdx.x];
    const float nrm = sqrtf(norm2[ch.tensor]);
    const float cf = gscale * clip / fmaxf(nrm, clip);
    const float lr_t = hyper[0];
    for (int i = threadIdx.x; i < ch.len; i += kOptThreads) {
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

MPB_API int mpb_opt_step(int nchunks, const mpb_opt_chunk* chunks, int ntensors, float* param, const float* grad,
                         float* m, float* v, float* ema, float* norm2, const float* hyper, float grad_scale,
                         float clip_norm, float beta1, float beta2, float eps, float ema_decay, void* stream) {
    using namespace mpb;
    if (nchunks <= 0 || !chunks || !param || !grad || !m || !v || !ema || !norm2 || !hyper) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    MPB_CUDA_TRY(cudaMemsetAsync(norm2, 0, sizeof(float) * ntensors, s));
    opt_sumsq_kernel<<<nchunks, kOptThreads, 0, s>>>(chunks, grad, grad_scale, norm2);
    MPB_LAUNCH_CHECK();
    opt_adam_ema_kernel<<<nchunks, kOptThreads, 0, s>>>(chunks, param, grad, m, v, ema, norm2, hyper, grad_scale,
                                                       clip_norm, beta1, beta2, eps, ema_decay);
    MPB_LAUNCH_CHECK();
    return 0;
}
