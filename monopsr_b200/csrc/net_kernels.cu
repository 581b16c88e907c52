// monopsr_b200/csrc/net_kernels.cu -- the bandwidth-bound layers around the tcgen05 GEMM core.
//
// NHWC fp32 everywhere.  Each kernel cites the reference graph op it replaces (the reference
// obtains all of these from TensorFlow 1.8 / TF-slim; semantics restated from TF defaults).
#include "common.cuh"
#include <cuda_fp16.h>
#include "../../include/monopsr_b200_net.h"
#include <math.h>

namespace mpb {

// Operand rounding: every GEMM operand produced here is rounded to tf32 (round-to-nearest; the tensor core itself
// truncates).  The 3xTF32 forward path (tc_gemm_x3_kernel) wants the operands UNROUNDED instead: one switch for all
// producers of this file, 1 by default, flipped by mpb_set_operand_rounding().
__constant__ int c_round_operands = 1;
__device__ __forceinline__ float rtf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return c_round_operands ? __uint_as_float(u) : x;
}
// h3 forward: optional fp16 hi/lo split copy ([hi | lo] per 32-channel block, see csrc/split16.cu) written by the
// producer itself next to its fp32 result.  v = 4 consecutive channels c..c+3 of pixel `row` (pitch ld floats).
__device__ __forceinline__ void store_split16(unsigned char* y16, long row, int ld, int c, float4 v, int* overflow) {
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    unsigned char* d = y16 + ((size_t)row * ld + (c & ~31)) * 4 + (c & 31) * 2;
    *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(d + 64) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    if (overflow && fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) > 65504.f) *overflow = 1;
}


// ---------------------------------------------------------------- weight preparation
// Frozen (inference-mode) BN folded into the conv weights (feature_extractor.py:228-242:
// is_training=False, eps 1e-5, scale=True):  wf = tf32(w * s), s = gamma*rsqrt(var+eps),
// shift = beta - mean*s.  The GEMM then needs only "+ shift" in its epilogue, and the
// data-gradient pass can use wf directly.
__global__ void fold_bn_kernel(int cout, int K, const float* __restrict__ w, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ wf,
                               float* __restrict__ scale, float* __restrict__ shift) {
    const int co = blockIdx.x;
    const float s = gamma[co] * rsqrtf(var[co] + eps);
    if (threadIdx.x == 0) {
        scale[co] = s;
        shift[co] = beta[co] - mean[co] * s;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) wf[(size_t)co * K + k] = rtf32(w[(size_t)co * K + k] * s);
}

// 128-bit accesses, four independent loads in flight per thread (the scalar one-element-per-thread version ran
// at 2.2 TB/s); scalar tail for n % 4 and unaligned bases
__global__ void __launch_bounds__(256)
round_copy_kernel(size_t n, const float* __restrict__ src, float* __restrict__ dst) {
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const size_t n4 = vec ? n >> 2 : 0;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldcs(s4 + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; u++)
            d4[i + u * stride] = make_float4(rtf32(v[u].x), rtf32(v[u].y), rtf32(v[u].z), rtf32(v[u].w));
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldcs(s4 + i);
        d4[i] = make_float4(rtf32(v.x), rtf32(v.y), rtf32(v.z), rtf32(v.w));
    }
    for (size_t j = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) dst[j] = rtf32(src[j]);
}

// d(gamma), d(beta) of a frozen BN from the conv's weight gradient:
//   dbeta[c] = colsum(g)[c];  dgamma[c] = rowdot(w, dw)[c]/gamma[c] - mean[c]*dbeta[c]*rsqrt(var+eps)
__global__ void bn_param_grad_kernel(int cout, int K, const float* __restrict__ w, const float* __restrict__ dw,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ var, float eps, const float* __restrict__ dbeta,
                                     float* __restrict__ dgamma) {
    const int co = blockIdx.x;
    float acc = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) acc += w[(size_t)co * K + k] * dw[(size_t)co * K + k];
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
        dgamma[co] = s / gamma[co] - mean[co] * dbeta[co] * rsqrtf(var[co] + eps);
    }
}


// ---- whole-model variants: one launch over every (layer, output channel) pair instead of
// one launch per layer (208 frozen-BN convs in the two towers)
// One WARP per (layer, output channel) row, 8 rows per CTA, 128-bit loads with four in flight per lane (rows are
// 64 .. 4608 floats, every row starts on a 16-byte boundary because K % 4 == 0 for the tensor-core convs; the
// 7x7x3 stem rows, K = 147, take the scalar path).  The block-per-row scalar versions ran at 2.3-2.6 TB/s.
__global__ void __launch_bounds__(256)
fold_bn_multi_kernel(int total_rows, const mpb_bn_layer* __restrict__ layers, const int* __restrict__ row2layer, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= total_rows) return;
    const mpb_bn_layer L = layers[row2layer[row]];
    const int co = row - L.row0;
    const float s = L.gamma[co] * rsqrtf(L.var[co] + eps);
    if (lane == 0) {
        L.scale[co] = s;
        L.shift[co] = L.beta[co] - L.mean[co] * s;
    }
    const float* w = L.w + (size_t)co * L.K;
    float* wf = L.wf + (size_t)co * L.K;
    if ((L.K & 3) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        float4* f4 = reinterpret_cast<float4*>(wf);
        const int n4 = L.K >> 2;
        int k = lane;
        for (; k + 96 < n4; k += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = __ldcs(w4 + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; u++)
                f4[k + 32 * u] = make_float4(rtf32(v[u].x * s), rtf32(v[u].y * s), rtf32(v[u].z * s), rtf32(v[u].w * s));
        }
        for (; k < n4; k += 32) {
            const float4 v = __ldcs(w4 + k);
            f4[k] = make_float4(rtf32(v.x * s), rtf32(v.y * s), rtf32(v.z * s), rtf32(v.w * s));
        }
    } else {
        for (int k = lane; k < L.K; k += 32) wf[k] = rtf32(w[k] * s);
    }
}
__global__ void __launch_bounds__(256)
bn_param_grad_multi_kernel(int row_begin, int total_rows, const mpb_bn_layer* __restrict__ layers,
                           const int* __restrict__ row2layer, float eps) {
    const int row = row_begin + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= total_rows) return;
    const mpb_bn_layer L = layers[row2layer[row]];
    const int co = row - L.row0;
    const float* w = L.w + (size_t)co * L.K;
    const float* dw = L.dw + (size_t)co * L.K;
    float acc = 0.f;
    if ((L.K & 3) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        const float4* d4 = reinterpret_cast<const float4*>(dw);
        const int n4 = L.K >> 2;
        int k = lane;
        for (; k + 32 < n4; k += 64) {
            const float4 a0 = __ldcs(w4 + k), a1 = __ldcs(w4 + k + 32), b0 = __ldcs(d4 + k), b1 = __ldcs(d4 + k + 32);
            acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc); acc = fmaf(a0.z, b0.z, acc); acc = fmaf(a0.w, b0.w, acc);
            acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc); acc = fmaf(a1.z, b1.z, acc); acc = fmaf(a1.w, b1.w, acc);
        }
        for (; k < n4; k += 32) {
            const float4 a0 = __ldcs(w4 + k), b0 = __ldcs(d4 + k);
            acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc); acc = fmaf(a0.z, b0.z, acc); acc = fmaf(a0.w, b0.w, acc);
        }
    } else {
        for (int k = lane; k < L.K; k += 32) acc = fmaf(w[k], dw[k], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) L.dgamma[co] = acc / L.gamma[co] - L.mean[co] * L.dbeta[co] * rsqrtf(L.var[co] + eps);
}

// ---------------------------------------------------------------- stem: conv 7x7/2 + BN + ReLU
// resnet_utils.conv2d_same(net, 64, 7, stride=2) (nets/resnet_v1.py:234): pad 3|3, VALID.
// Cin=3 -> K=147: kept off the tensor cores.  One thread = FOUR consecutive output pixels of a row x 8 channels: the 24
// weights of a tap (two LDS.128 per input channel from a k-major copy in shared memory) are used for four pixels, the
// 13 x 3 input values of a filter row for all seven taps.  (One pixel x 16 channels per thread issued one LDS per FMA:
// 134 us for the full image at the head of the full-image tower's chain, with nothing else to run beside it.)
constexpr int kStemK = 147;
constexpr int kStemPx = 4, kStemCg = 8;
__global__ void __launch_bounds__(256)
stem_fwd_kernel(int nimg, int Hin, int Win, int Ho, int Wo, const float* __restrict__ x,
                const float* __restrict__ wf, const float* __restrict__ shift, float* __restrict__ y) {
    __shared__ __align__(16) float swt[kStemK * 64];                 // [k][co]
    for (int i = threadIdx.x; i < 64 * kStemK; i += blockDim.x)      // conflict-free stores; the strided reads hit L2
        swt[i] = __ldg(wf + (size_t)(i & 63) * kStemK + (i >> 6));
    __syncthreads();
    const int cg = threadIdx.x & 7;                                  // 8 groups of 8 output channels
    const int qpr = (Wo + kStemPx - 1) / kStemPx;                    // pixel quads per output row
    const long quad = (long)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    if (quad >= (long)nimg * Ho * qpr) return;
    const int n = (int)(quad / ((long)Ho * qpr)), rem = (int)(quad % ((long)Ho * qpr)), oh = rem / qpr, ow0 = (rem % qpr) * kStemPx;
    float acc[kStemPx][kStemCg];
#pragma unroll
    for (int p = 0; p < kStemPx; p++)
#pragma unroll
        for (int c = 0; c < kStemCg; c++) acc[p][c] = 0.f;
    for (int kh = 0; kh < 7; kh++) {
        const int ih = oh * 2 - 3 + kh;
        if (ih < 0 || ih >= Hin) continue;
        const float* xr = x + ((size_t)n * Hin + ih) * Win * 3;
        float px[2 * kStemPx + 5][3];                                // input columns ow0*2-3 .. ow0*2+9
#pragma unroll
        for (int j = 0; j < 2 * kStemPx + 5; j++) {
            const int iw = ow0 * 2 - 3 + j;
            const bool in = iw >= 0 && iw < Win;
            px[j][0] = in ? __ldg(xr + (size_t)iw * 3) : 0.f;
            px[j][1] = in ? __ldg(xr + (size_t)iw * 3 + 1) : 0.f;
            px[j][2] = in ? __ldg(xr + (size_t)iw * 3 + 2) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 7; kw++) {
            float wv[3][kStemCg];
#pragma unroll
            for (int ci = 0; ci < 3; ci++) {
                const float4* wp = reinterpret_cast<const float4*>(swt + ((kh * 7 + kw) * 3 + ci) * 64 + cg * kStemCg);
                const float4 a = wp[0], b = wp[1];
                wv[ci][0] = a.x; wv[ci][1] = a.y; wv[ci][2] = a.z; wv[ci][3] = a.w;
                wv[ci][4] = b.x; wv[ci][5] = b.y; wv[ci][6] = b.z; wv[ci][7] = b.w;
            }
#pragma unroll
            for (int p = 0; p < kStemPx; p++) {
                const int j = kw + 2 * p;
#pragma unroll
                for (int c = 0; c < kStemCg; c++)
                    acc[p][c] = fmaf(px[j][2], wv[2][c], fmaf(px[j][1], wv[1][c], fmaf(px[j][0], wv[0][c], acc[p][c])));
            }
        }
    }
    const float4 s0 = *reinterpret_cast<const float4*>(shift + cg * kStemCg), s1 = *reinterpret_cast<const float4*>(shift + cg * kStemCg + 4);
#pragma unroll
    for (int p = 0; p < kStemPx; p++) {
        if (ow0 + p >= Wo) break;
        float* yp = y + ((((size_t)n * Ho + oh) * Wo) + ow0 + p) * 64 + cg * kStemCg;
        *reinterpret_cast<float4*>(yp) = make_float4(rtf32(fmaxf(acc[p][0] + s0.x, 0.f)), rtf32(fmaxf(acc[p][1] + s0.y, 0.f)),
                                                     rtf32(fmaxf(acc[p][2] + s0.z, 0.f)), rtf32(fmaxf(acc[p][3] + s0.w, 0.f)));
        *reinterpret_cast<float4*>(yp + 4) = make_float4(rtf32(fmaxf(acc[p][4] + s1.x, 0.f)), rtf32(fmaxf(acc[p][5] + s1.y, 0.f)),
                                                         rtf32(fmaxf(acc[p][6] + s1.z, 0.f)), rtf32(fmaxf(acc[p][7] + s1.w, 0.f)));
    }
}

// dW[co][kh][kw][ci] += scale[co] * sum_pix g[pix][co] * x[pix*2-3+k][ci].  A CTA owns a strip of
// output pixels, accumulates the 64x147 products in registers (thread = (co, 9..10 k's)), one
// RED per weight per CTA.
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(int nimg, int Hin, int Win, int Ho, int Wo, const float* __restrict__ x,
                  const float* __restrict__ g, const float* __restrict__ scale, float* __restrict__ dw,
                  int pix_per_cta) {
    // 16 pixels are staged per barrier pair (gradients and 7x7x3 patches), so a thread runs 16 x 37 FMAs between
    // synchronisations instead of 37
    constexpr int PT = 16;
    __shared__ float sg[PT][64];
    __shared__ __align__(16) float sx[PT][4 * 40];       // 4 parts of 37 taps, each padded to 40: a part is 10 LDS.128
    __shared__ int spix[PT][3];
    const long total = (long)nimg * Ho * Wo;
    const long p0 = (long)blockIdx.x * pix_per_cta, pend = min(total, p0 + pix_per_cta);
    const int co = threadIdx.x & 63, part = threadIdx.x >> 6;   // 4 parts over the 147 taps
    float acc[40];
#pragma unroll
    for (int i = 0; i < 40; i++) acc[i] = 0.f;
    for (long pix0 = p0; pix0 < pend; pix0 += PT) {
        __syncthreads();
        if (threadIdx.x < PT) {
            const long pix = pix0 + threadIdx.x;
            const int n = (int)(pix / (Ho * Wo)), rem = (int)(pix % (Ho * Wo));
            spix[threadIdx.x][0] = n; spix[threadIdx.x][1] = rem / Wo; spix[threadIdx.x][2] = rem % Wo;
        }
        for (int i = threadIdx.x; i < PT * 64; i += blockDim.x) {
            const long pix = pix0 + (i >> 6);
            sg[i >> 6][i & 63] = pix < pend ? g[(size_t)pix * 64 + (i & 63)] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < PT * kStemK; i += blockDim.x) {
            const int q = i / kStemK, k = i - q * kStemK;
            const int t = k / 3, ci = k - t * 3;
            const int n = spix[q][0], ih = spix[q][1] * 2 - 3 + t / 7, iw = spix[q][2] * 2 - 3 + t % 7;
            sx[q][(k / 37) * 40 + k % 37] = (pix0 + q < pend && ih >= 0 && ih < Hin && iw >= 0 && iw < Win)
                                                ? x[(((size_t)n * Hin + ih) * Win + iw) * 3 + ci] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int q = 0; q < PT; q++) {
            const float gv = sg[q][co];
            const float4* xp = reinterpret_cast<const float4*>(&sx[q][part * 40]);
#pragma unroll
            for (int i4 = 0; i4 < 10; i4++) {                  // (the 3 padding slots of a part hold garbage: never written back)
                const float4 v = xp[i4];
                acc[4 * i4] = fmaf(gv, v.x, acc[4 * i4]);
                acc[4 * i4 + 1] = fmaf(gv, v.y, acc[4 * i4 + 1]);
                acc[4 * i4 + 2] = fmaf(gv, v.z, acc[4 * i4 + 2]);
                acc[4 * i4 + 3] = fmaf(gv, v.w, acc[4 * i4 + 3]);
            }
        }
    }
    const float s = scale[co];
#pragma unroll
    for (int i = 0; i < 37; i++) {
        const int k = part * 37 + i;
        if (k < kStemK && acc[i] != 0.f) atomicAdd(&dw[(size_t)co * kStemK + k], acc[i] * s);
    }
}

// ---------------------------------------------------------------- max pools
// slim.max_pool2d([3,3], stride=2, padding='SAME') (nets/resnet_v1.py:235): even H,W -> window
// rows 2o..2o+2 clipped at the bottom/right edge.
__global__ void maxpool3s2_fwd_kernel(int nimg, int H, int W, int C4, int Ho, int Wo,
                                      const float4* __restrict__ x, float4* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * Ho * Wo * C4;
    if (i >= total) return;
    const int c = (int)(i % C4);
    long p = i / C4;
    const int ow = (int)(p % Wo); p /= Wo;
    const int oh = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dh = 0; dh < 3; dh++) {
        const int ih = oh * 2 + dh;
        if (ih >= H) break;
        for (int dw = 0; dw < 3; dw++) {
            const int iw = ow * 2 + dw;
            if (iw >= W) break;
            const float4 v = x[(((size_t)n * H + ih) * W + iw) * C4 + c];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    y[i] = m;
}

// gather form (deterministic): dx[in] = (x[in]>0) * sum over the <=4 windows containing `in` whose
// FIRST maximum is `in`.  The (x>0) factor is the ReLU of the stem that produced x.
// thread = (input pixel, 4 channels): the <= 4 windows x 9 taps are 16-byte loads and the "first maximum" test runs on
// the four channels at once (the one-channel version: 57-63 us at the very end of the full-image tower's backward chain)
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(int nimg, int H, int W, int C4, int Ho, int Wo,
                      const float4* __restrict__ x, const float4* __restrict__ dy, float4* __restrict__ dx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * H * W * C4;
    if (i >= total) return;
    const int c = (int)(i % C4);
    long p = i / C4;
    const int iw = (int)(p % W); p /= W;
    const int ih = (int)(p % H);
    const int n = (int)(p / H);
    const float4 xv = x[i];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xv.x > 0.f || xv.y > 0.f || xv.z > 0.f || xv.w > 0.f) {
        for (int oh = max(0, (ih - 1) / 2); oh <= min(Ho - 1, ih / 2); oh++) {
            if (ih < oh * 2 || ih > oh * 2 + 2) continue;
            for (int ow = max(0, (iw - 1) / 2); ow <= min(Wo - 1, iw / 2); ow++) {
                if (iw < ow * 2 || iw > ow * 2 + 2) continue;
                // per channel: is (ih,iw) the FIRST maximum of window (oh,ow)?
                bool fx = true, fy = true, fz = true, fw = true;
                for (int dh = 0; dh < 3; dh++) {
                    const int jh = oh * 2 + dh;
                    if (jh >= H) break;
                    for (int dw = 0; dw < 3; dw++) {
                        const int jw = ow * 2 + dw;
                        if (jw >= W) break;
                        const float4 v = x[(((size_t)n * H + jh) * W + jw) * C4 + c];
                        const bool before = (jh < ih) || (jh == ih && jw < iw);
                        fx = fx && !(v.x > xv.x || (before && v.x == xv.x));
                        fy = fy && !(v.y > xv.y || (before && v.y == xv.y));
                        fz = fz && !(v.z > xv.z || (before && v.z == xv.z));
                        fw = fw && !(v.w > xv.w || (before && v.w == xv.w));
                    }
                }
                const float4 g = dy[(((size_t)n * Ho + oh) * Wo + ow) * C4 + c];
                if (fx) acc.x += g.x;
                if (fy) acc.y += g.y;
                if (fz) acc.z += g.z;
                if (fw) acc.w += g.w;
            }
        }
    }
    dx[i] = make_float4(xv.x > 0.f ? acc.x : 0.f, xv.y > 0.f ? acc.y : 0.f, xv.z > 0.f ? acc.z : 0.f, xv.w > 0.f ? acc.w : 0.f);
}

// slim.max_pool2d([2,2]) (builders/net_builder.py:60,68): VALID, stride 2
__global__ void maxpool2_fwd_kernel(int nimg, int H, int W, int C4, const float* __restrict__ x, int ldx,
                                    float* __restrict__ y, int ldy) {
    const int Ho = H / 2, Wo = W / 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * Ho * Wo * C4;
    if (i >= total) return;
    const int c = (int)(i % C4) * 4;
    long p = i / C4;
    const int ow = (int)(p % Wo); p /= Wo;
    const int oh = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const float* b = x + (((size_t)n * H + oh * 2) * W + ow * 2) * ldx + c;
    const float4 a0 = *reinterpret_cast<const float4*>(b), a1 = *reinterpret_cast<const float4*>(b + ldx);
    const float4 a2 = *reinterpret_cast<const float4*>(b + (size_t)W * ldx),
                 a3 = *reinterpret_cast<const float4*>(b + (size_t)W * ldx + ldx);
    float4 m;
    m.x = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
    m.y = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
    m.z = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
    m.w = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
    *reinterpret_cast<float4*>(y + (((size_t)n * Ho + oh) * Wo + ow) * ldy + c) = m;
}

// dx (same geometry as x) = or += dy routed to the first maximum of each 2x2 window
__global__ void maxpool2_bwd_kernel(int nimg, int H, int W, int C, const float* __restrict__ x, int ldx,
                                    const float* __restrict__ dy, int ldy, float* __restrict__ dx, int lddx,
                                    int accumulate) {
    const int Ho = H / 2, Wo = W / 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * Ho * Wo * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    long p = i / C;
    const int ow = (int)(p % Wo); p /= Wo;
    const int oh = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const size_t b = (((size_t)n * H + oh * 2) * W + ow * 2);
    float v[4] = {x[b * ldx + c], x[(b + 1) * ldx + c], x[(b + W) * ldx + c], x[(b + W + 1) * ldx + c]};
    int best = 0;
    for (int k = 1; k < 4; k++)
        if (v[k] > v[best]) best = k;
    const float g = dy[(((size_t)n * Ho + oh) * Wo + ow) * ldy + c];
    const size_t off[4] = {b, b + 1, b + W, b + W + 1};
    for (int k = 0; k < 4; k++) {
        const float t = (k == best) ? g : 0.f;
        float* d = dx + off[k] * lddx + c;
        *d = accumulate ? (*d + t) : t;
    }
}

// ---------------------------------------------------------------- crop_and_resize + 2x2 max pool
// tf.image.crop_and_resize(full_img_encoder_out, boxes_2d_norm, box_ind=0, (24,24)) followed by
// slim.max_pool2d([2,2]) (builders/net_builder.py:54-60), fused: the (nbox,24,24,C) intermediate
// never exists.  Sample coordinate: in_y = y1 (H-1) + i (y2-y1)(H-1)/(crop-1); outside [0,H-1]
// -> 0 (extrapolation_value); bilinear between floor and ceil.
struct CropCoord { int y0, y1i, x0, x1i; float ly, lx; bool valid; };
__device__ __forceinline__ CropCoord crop_coord(const float* box, int i, int j, int crop, int H, int W) {
    CropCoord c;
    const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
    const float in_y = y1 * (H - 1) + i * ((y2 - y1) * (H - 1) / (crop - 1));
    const float in_x = x1 * (W - 1) + j * ((x2 - x1) * (W - 1) / (crop - 1));
    c.valid = in_y >= 0.f && in_y <= (float)(H - 1) && in_x >= 0.f && in_x <= (float)(W - 1);
    const float fy = floorf(in_y), fx = floorf(in_x);
    c.y0 = min(max((int)fy, 0), H - 1);
    c.y1i = min(max((int)ceilf(in_y), 0), H - 1);
    c.x0 = min(max((int)fx, 0), W - 1);
    c.x1i = min(max((int)ceilf(in_x), 0), W - 1);
    c.ly = in_y - fy;
    c.lx = in_x - fx;
    return c;
}
__device__ __forceinline__ float crop_sample(const float* feat, int W, int C, int ch, const CropCoord& c) {
    if (!c.valid) return 0.f;
    const float tl = feat[((size_t)c.y0 * W + c.x0) * C + ch], tr = feat[((size_t)c.y0 * W + c.x1i) * C + ch];
    const float bl = feat[((size_t)c.y1i * W + c.x0) * C + ch], br = feat[((size_t)c.y1i * W + c.x1i) * C + ch];
    const float top = tl + (tr - tl) * c.lx, bot = bl + (br - bl) * c.lx;
    return top + (bot - top) * c.ly;
}

// thread = (box, pooled pixel, 4 channels): 16-byte loads of the four bilinear corners (the one-channel-per-thread
// versions took 75 / 103 us on the critical path either side of the full-image tower, profiles/r2_notes.md)
__device__ __forceinline__ float4 crop_sample4(const float* feat, int W, int C, int ch, const CropCoord& c) {
    if (!c.valid) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 tl = *reinterpret_cast<const float4*>(feat + ((size_t)c.y0 * W + c.x0) * C + ch);
    const float4 tr = *reinterpret_cast<const float4*>(feat + ((size_t)c.y0 * W + c.x1i) * C + ch);
    const float4 bl = *reinterpret_cast<const float4*>(feat + ((size_t)c.y1i * W + c.x0) * C + ch);
    const float4 br = *reinterpret_cast<const float4*>(feat + ((size_t)c.y1i * W + c.x1i) * C + ch);
    float4 o;
#define MPB_BIL(f) { const float top = tl.f + (tr.f - tl.f) * c.lx, bot = bl.f + (br.f - bl.f) * c.lx; o.f = top + (bot - top) * c.ly; }
    MPB_BIL(x) MPB_BIL(y) MPB_BIL(z) MPB_BIL(w)
#undef MPB_BIL
    return o;
}

__global__ void __launch_bounds__(256)
crop_pool_fwd_kernel(int H, int W, int C, const float* __restrict__ feat, int nbox,
                     const float* __restrict__ boxes, int crop, float* __restrict__ out, int ldo) {
    const int P = crop / 2, C4 = C >> 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nbox * P * P * C4;
    if (i >= total) return;
    const int ch = (int)(i % C4) * 4;
    long p = i / C4;
    const int pw = (int)(p % P); p /= P;
    const int ph = (int)(p % P);
    const int b = (int)(p / P);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const CropCoord c = crop_coord(boxes + b * 4, ph * 2 + (k >> 1), pw * 2 + (k & 1), crop, H, W);
        const float4 v = crop_sample4(feat, W, C, ch, c);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    *reinterpret_cast<float4*>(out + (((size_t)b * P + ph) * P + pw) * ldo + ch) =
        make_float4(rtf32(m.x), rtf32(m.y), rtf32(m.z), rtf32(m.w));
}

__global__ void __launch_bounds__(256)
crop_pool_bwd_kernel(int H, int W, int C, const float* __restrict__ feat, int nbox,
                     const float* __restrict__ boxes, int crop, const float* __restrict__ dout,
                     int ldd, float* __restrict__ dfeat) {
    const int P = crop / 2, C4 = C >> 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nbox * P * P * C4;
    if (i >= total) return;
    const int ch = (int)(i % C4) * 4;
    long p = i / C4;
    const int pw = (int)(p % P); p /= P;
    const int ph = (int)(p % P);
    const int b = (int)(p / P);
    const float4 g4 = *reinterpret_cast<const float4*>(dout + (((size_t)b * P + ph) * P + pw) * ldd + ch);
    if (g4.x == 0.f && g4.y == 0.f && g4.z == 0.f && g4.w == 0.f) return;
    // first maximum of the four samples, per channel (strict >: the forward's fmaxf keeps the first of equals)
    CropCoord cs[4];
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int bx = 0, by = 0, bz = 0, bw = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        cs[k] = crop_coord(boxes + b * 4, ph * 2 + (k >> 1), pw * 2 + (k & 1), crop, H, W);
        const float4 v = crop_sample4(feat, W, C, ch, cs[k]);
        if (v.x > m.x) { m.x = v.x; bx = k; }
        if (v.y > m.y) { m.y = v.y; by = k; }
        if (v.z > m.z) { m.z = v.z; bz = k; }
        if (v.w > m.w) { m.w = v.w; bw = k; }
    }
    // the gradient of each channel goes to the corners of ITS winning sample: per sample k, one vector RED per corner
    // with the channels that did not pick k zeroed (channels mostly agree, so most vectors are skipped)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 g = make_float4(bx == k ? g4.x : 0.f, by == k ? g4.y : 0.f, bz == k ? g4.z : 0.f, bw == k ? g4.w : 0.f);
        if ((g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) || !cs[k].valid) continue;
        const CropCoord& c = cs[k];
        const float w00 = (1.f - c.ly) * (1.f - c.lx), w01 = (1.f - c.ly) * c.lx, w10 = c.ly * (1.f - c.lx), w11 = c.ly * c.lx;
        atomicAdd(reinterpret_cast<float4*>(dfeat + ((size_t)c.y0 * W + c.x0) * C + ch), make_float4(g.x * w00, g.y * w00, g.z * w00, g.w * w00));
        atomicAdd(reinterpret_cast<float4*>(dfeat + ((size_t)c.y0 * W + c.x1i) * C + ch), make_float4(g.x * w01, g.y * w01, g.z * w01, g.w * w01));
        atomicAdd(reinterpret_cast<float4*>(dfeat + ((size_t)c.y1i * W + c.x0) * C + ch), make_float4(g.x * w10, g.y * w10, g.z * w10, g.w * w10));
        atomicAdd(reinterpret_cast<float4*>(dfeat + ((size_t)c.y1i * W + c.x1i) * C + ch), make_float4(g.x * w11, g.y * w11, g.z * w11, g.w * w11));
    }
}

// ---------------------------------------------------------------- bilinear resize, align_corners=True
// tf.image.resize_images(x, (OH,OW), align_corners=True) (builders/net_builder.py:73-75,82-84)
__global__ void resize_ac_fwd_kernel(int nimg, int H, int W, int C4, int OH, int OW,
                                     const float4* __restrict__ x, float4* __restrict__ y,
                                     unsigned char* __restrict__ y16, int* __restrict__ overflow) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * OH * OW * C4;
    if (i >= total) return;
    const int c = (int)(i % C4);
    long p = i / C4;
    const int ow = (int)(p % OW); p /= OW;
    const int oh = (int)(p % OH);
    const int n = (int)(p / OH);
    const float sy = oh * ((float)(H - 1) / (float)(OH - 1)), sx = ow * ((float)(W - 1) / (float)(OW - 1));
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float4 tl = x[(((size_t)n * H + y0) * W + x0) * C4 + c], tr = x[(((size_t)n * H + y0) * W + x1) * C4 + c];
    const float4 bl = x[(((size_t)n * H + y1) * W + x0) * C4 + c], br = x[(((size_t)n * H + y1) * W + x1) * C4 + c];
    float4 o;
#define MPB_LERP(f) { float t = tl.f + (tr.f - tl.f) * lx, b = bl.f + (br.f - bl.f) * lx; o.f = rtf32(t + (b - t) * ly); }
    MPB_LERP(x) MPB_LERP(y) MPB_LERP(z) MPB_LERP(w)
#undef MPB_LERP
    y[i] = o;
    if (y16) store_split16(y16, i / C4, C4 * 4, c * 4, o, overflow);
}

// Backward of the bilinear align-corners resize in GATHER form: one thread per input pixel x 4 channels sums the
// (at most 4 x 4) output pixels that interpolate from it, with the weights recomputed by the forward's own float
// expressions.  No zero-fill, no atomics, deterministic (the scatter version spent 60 + 150 us in memset + RED.ADD
// on the decoder's critical path).
__device__ __forceinline__ int resize_taps(int y, int H, int OH, int* o, float* w) {
    // output indices oh (and weights) with y0(oh) == y or y1(oh) == y
    const float s = (float)(H - 1) / (float)(OH - 1), inv = (float)(OH - 1) / (float)(H - 1);
    const int lo = max(0, (int)floorf((y - 1) * inv) - 1), hi = min(OH - 1, (int)ceilf((y + 1) * inv) + 1);
    int n = 0;
    for (int oh = lo; oh <= hi && n < 6; oh++) {
        const float sy = oh * s;
        const int y0 = (int)floorf(sy), y1 = min(y0 + 1, H - 1);
        const float ly = sy - y0;
        float wt = 0.f;
        if (y0 == y) wt += 1.f - ly;
        if (y1 == y) wt += ly;
        if (y0 == y || y1 == y) { o[n] = oh; w[n] = wt; n++; }
    }
    return n;
}

__global__ void resize_ac_bwd_kernel(int nimg, int H, int W, int C4, int OH, int OW, const float4* __restrict__ dy,
                                     float4* __restrict__ dx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)nimg * H * W * C4;
    if (i >= total) return;
    const int c = (int)(i % C4);
    long p = i / C4;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int n = (int)(p / H);
    int oy[6], ox[6];
    float wy[6], wx[6];
    const int ny = resize_taps(y, H, OH, oy, wy), nx = resize_taps(x, W, OW, ox, wx);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < ny; a++) {
        const float4* row = dy + ((size_t)n * OH + oy[a]) * OW * C4 + c;
        for (int b = 0; b < nx; b++) {
            const float4 g = __ldg(row + (size_t)ox[b] * C4);
            const float wt = wy[a] * wx[b];
            acc.x = fmaf(g.x, wt, acc.x); acc.y = fmaf(g.y, wt, acc.y);
            acc.z = fmaf(g.z, wt, acc.z); acc.w = fmaf(g.w, wt, acc.w);
        }
    }
    dx[i] = acc;
}

// ---------------------------------------------------------------- train-mode batch norm (+beta, ReLU)
// slim.batch_norm(is_training=True) defaults (center, no scale, eps 1e-3, decay 0.999) on the
// decoder convs (builders/net_builder.py:77-89): statistics over all M = nimg*H*W rows.
// stats: CTA = 32 channels x 8 row-groups, fp32 partials, final combine in double.
// thread = 4 channels (one 16-byte load) x a strided set of rows, 4 rows in flight; CTA = 32 channel quads x 8 row
// groups; fp32 partials per thread, combined in double.  (The one-float-per-thread version reached ~1.5 TB/s.)
__global__ void __launch_bounds__(256)
bn_stats_kernel(int M, int C, const float* __restrict__ z, double* __restrict__ psum, double* __restrict__ psq) {
    const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int rg = threadIdx.x >> 5;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (c < C) {
        int r = r0 + rg;
        for (; r + 24 < r1; r += 32) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = *reinterpret_cast<const float4*>(z + (size_t)(r + 8 * u) * C + c);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
                q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y);
                q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
            }
        }
        for (; r < r1; r += 8) {
            const float4 v = *reinterpret_cast<const float4*>(z + (size_t)r * C + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
        }
    }
    __shared__ float4 ss[8][32], sq[8][32];
    ss[rg][threadIdx.x & 31] = s;
    sq[rg][threadIdx.x & 31] = q;
    __syncthreads();
    if (rg == 0 && c < C) {
        double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; k++) {
            const float4 x = ss[k][threadIdx.x], y = sq[k][threadIdx.x];
            a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
            b[0] += y.x; b[1] += y.y; b[2] += y.z; b[3] += y.w;
        }
        for (int j = 0; j < 4; j++) { atomicAdd(&psum[c + j], a[j]); atomicAdd(&psq[c + j], b[j]); }
    }
}
__global__ void bn_finalize_kernel(int M, int C, const double* __restrict__ psum, const double* __restrict__ psq,
                                   float* __restrict__ mean, float* __restrict__ var,
                                   float* __restrict__ mmean, float* __restrict__ mvar, float decay) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = psum[c] / M;
    const double v = fmax(psq[c] / M - m * m, 0.0);
    mean[c] = (float)m;
    var[c] = (float)v;
    if (mmean) {   // UPDATE_OPS: moving = moving*decay + batch*(1-decay)  (batch var as TF: biased)
        mmean[c] = mmean[c] * decay + (float)m * (1.f - decay);
        mvar[c] = mvar[c] * decay + (float)v * (1.f - decay);
    }
}
__global__ void bn_apply_kernel(long total4, int C4, const float4* __restrict__ z, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ beta, float eps,
                                float4* __restrict__ y, unsigned char* __restrict__ y16, int* __restrict__ overflow) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % C4) * 4;
    const float4 v = z[i];
    float4 o;
    o.x = rtf32(fmaxf((v.x - mean[c]) * rsqrtf(var[c] + eps) + beta[c], 0.f));
    o.y = rtf32(fmaxf((v.y - mean[c + 1]) * rsqrtf(var[c + 1] + eps) + beta[c + 1], 0.f));
    o.z = rtf32(fmaxf((v.z - mean[c + 2]) * rsqrtf(var[c + 2] + eps) + beta[c + 2], 0.f));
    o.w = rtf32(fmaxf((v.w - mean[c + 3]) * rsqrtf(var[c + 3] + eps) + beta[c + 3], 0.f));
    y[i] = o;
    if (y16) store_split16(y16, i / C4, C4 * 4, c, o, overflow);
}
// backward: g = dy*(y>0); s1 = sum g; s2 = sum g*xhat; dz = rstd*(g - s1/M - xhat*s2/M); dbeta = s1
// same thread layout as bn_stats_kernel (4 channels per thread, 2 rows in flight: three tensors are read)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(int M, int C, const float* __restrict__ z, const float* __restrict__ y,
                     const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ var,
                     float eps, double* __restrict__ s1, double* __restrict__ s2) {
    const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int rg = threadIdx.x >> 5;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (c < C) {
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), vr = *reinterpret_cast<const float4*>(var + c);
        const float4 rs = make_float4(rsqrtf(vr.x + eps), rsqrtf(vr.y + eps), rsqrtf(vr.z + eps), rsqrtf(vr.w + eps));
        auto acc = [&](const float4& zz, const float4& yy, const float4& dd) {
            const float gx = yy.x > 0.f ? dd.x : 0.f, gy = yy.y > 0.f ? dd.y : 0.f;
            const float gz = yy.z > 0.f ? dd.z : 0.f, gw = yy.w > 0.f ? dd.w : 0.f;
            a.x += gx; a.y += gy; a.z += gz; a.w += gw;
            b.x = fmaf(gx, (zz.x - mu.x) * rs.x, b.x); b.y = fmaf(gy, (zz.y - mu.y) * rs.y, b.y);
            b.z = fmaf(gz, (zz.z - mu.z) * rs.z, b.z); b.w = fmaf(gw, (zz.w - mu.w) * rs.w, b.w);
        };
        int r = r0 + rg;
        for (; r + 8 < r1; r += 16) {
            const size_t o0 = (size_t)r * C + c, o1 = (size_t)(r + 8) * C + c;
            const float4 z0 = *reinterpret_cast<const float4*>(z + o0), z1 = *reinterpret_cast<const float4*>(z + o1);
            const float4 y0 = *reinterpret_cast<const float4*>(y + o0), y1 = *reinterpret_cast<const float4*>(y + o1);
            const float4 d0 = *reinterpret_cast<const float4*>(dy + o0), d1 = *reinterpret_cast<const float4*>(dy + o1);
            acc(z0, y0, d0);
            acc(z1, y1, d1);
        }
        for (; r < r1; r += 8) {
            const size_t o0 = (size_t)r * C + c;
            acc(*reinterpret_cast<const float4*>(z + o0), *reinterpret_cast<const float4*>(y + o0),
                *reinterpret_cast<const float4*>(dy + o0));
        }
    }
    __shared__ float4 sa[8][32], sb[8][32];
    sa[rg][threadIdx.x & 31] = a;
    sb[rg][threadIdx.x & 31] = b;
    __syncthreads();
    if (rg == 0 && c < C) {
        double p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; k++) {
            const float4 x = sa[k][threadIdx.x], w = sb[k][threadIdx.x];
            p[0] += x.x; p[1] += x.y; p[2] += x.z; p[3] += x.w;
            q[0] += w.x; q[1] += w.y; q[2] += w.z; q[3] += w.w;
        }
        for (int j = 0; j < 4; j++) { atomicAdd(&s1[c + j], p[j]); atomicAdd(&s2[c + j], q[j]); }
    }
}
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(long total4, int M, int C, const float* __restrict__ z, const float* __restrict__ y,
                    const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                    const double* __restrict__ s1, const double* __restrict__ s2, float* __restrict__ dz,
                    float* __restrict__ dbeta) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % (C / 4)) * 4;
    const float4 zz = reinterpret_cast<const float4*>(z)[i], yy = reinterpret_cast<const float4*>(y)[i];
    const float4 dd = reinterpret_cast<const float4*>(dy)[i];
    float zc[4] = {zz.x, zz.y, zz.z, zz.w}, yc[4] = {yy.x, yy.y, yy.z, yy.w}, dc[4] = {dd.x, dd.y, dd.z, dd.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float rs = rsqrtf(var[c + j] + eps);
        const float xh = (zc[j] - mean[c + j]) * rs;
        const float g = yc[j] > 0.f ? dc[j] : 0.f;
        const float m1 = (float)(s1[c + j] / M), m2 = (float)(s2[c + j] / M);
        o[j] = rtf32(rs * (g - m1 - xh * m2));
    }
    reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (i * 4 < C)
        for (int j = 0; j < 4; j++) dbeta[i * 4 + j] = (float)s1[i * 4 + j];
}

// ---------------------------------------------------------------- train-mode batch norm, ONE launch per direction
// statistics -> grid barrier -> apply in the same kernel: the apply phase re-reads from L2 what the same CTA streamed a
// moment ago (the 48x48x128 tensors are 38 MB each, L2 is 126 MB), the finalize kernel disappears, and the decoder chain
// -- which runs alone on the GPU -- loses two launches and their gaps per layer and direction.  The grid is 2 CTAs per SM
// (fits beside anything the other streams run), so all CTAs are resident and a counter barrier is safe; the counter
// lives behind the 2C accumulators in `scratch` (2C + 2 doubles, zeroed by the memset that precedes the launch).
__device__ __forceinline__ void bn_grid_barrier(unsigned long long* counter, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        // bounded (~1 s): if the grid were ever not co-resident this must be an error, not a hung GPU
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned long long*>(counter) < nblocks) {
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
bn_train_fwd_fused_kernel(int M, int C, const float* __restrict__ z, const float* __restrict__ beta, float eps,
                          float* __restrict__ y, float* __restrict__ mean, float* __restrict__ var,
                          float* __restrict__ mmean, float* __restrict__ mvar, float decay, double* __restrict__ scratch,
                          unsigned char* __restrict__ y16, int* __restrict__ overflow) {
    double* psum = scratch;
    double* psq = scratch + C;
    const int lane32 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane32) * 4;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (c < C) {
        int r = r0 + rg;
        for (; r + 56 < r1; r += 64) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = *reinterpret_cast<const float4*>(z + (size_t)(r + 8 * u) * C + c);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
                q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y);
                q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
            }
        }
        for (; r < r1; r += 8) {
            const float4 v = *reinterpret_cast<const float4*>(z + (size_t)r * C + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
        }
    }
    __shared__ float4 ss[8][32], sq[8][32];
    ss[rg][lane32] = s;
    sq[rg][lane32] = q;
    __syncthreads();
    if (rg == 0 && c < C) {
        double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; k++) {
            const float4 x = ss[k][lane32], w = sq[k][lane32];
            a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
            b[0] += w.x; b[1] += w.y; b[2] += w.z; b[3] += w.w;
        }
        for (int j = 0; j < 4; j++) { atomicAdd(&psum[c + j], a[j]); atomicAdd(&psq[c + j], b[j]); }
    }
    bn_grid_barrier(reinterpret_cast<unsigned long long*>(scratch + 2 * C), gridDim.x * gridDim.y);
    if (c >= C) return;
    float mu[4], rs[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {        // the arithmetic of bn_finalize_kernel + bn_apply_kernel
        const double m = __ldcg(psum + c + j) / M;
        const double v = fmax(__ldcg(psq + c + j) / M - m * m, 0.0);
        mu[j] = (float)m;
        rs[j] = rsqrtf((float)v + eps);
        bt[j] = beta[c + j];
        if (blockIdx.y == 0 && rg == 0) {
            mean[c + j] = (float)m;
            var[c + j] = (float)v;
            if (mmean) {   // UPDATE_OPS: moving = moving*decay + batch*(1-decay)  (batch var as TF: biased)
                mmean[c + j] = mmean[c + j] * decay + (float)m * (1.f - decay);
                mvar[c + j] = mvar[c + j] * decay + (float)v * (1.f - decay);
            }
        }
    }
    for (int r = r0 + rg; r < r1; r += 8) {
        const float4 v = *reinterpret_cast<const float4*>(z + (size_t)r * C + c);
        float4 o;
        o.x = rtf32(fmaxf((v.x - mu[0]) * rs[0] + bt[0], 0.f));
        o.y = rtf32(fmaxf((v.y - mu[1]) * rs[1] + bt[1], 0.f));
        o.z = rtf32(fmaxf((v.z - mu[2]) * rs[2] + bt[2], 0.f));
        o.w = rtf32(fmaxf((v.w - mu[3]) * rs[3] + bt[3], 0.f));
        *reinterpret_cast<float4*>(y + (size_t)r * C + c) = o;
        if (y16) store_split16(y16, r, C, c, o, overflow);
    }
}

__global__ void __launch_bounds__(256)
bn_train_bwd_fused_kernel(int M, int C, const float* __restrict__ z, const float* __restrict__ y,
                          const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ var,
                          float eps, double* __restrict__ scratch, float* __restrict__ dz, float* __restrict__ dbeta) {
    double* s1 = scratch;
    double* s2 = scratch + C;
    const int lane32 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane32) * 4;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    float4 mu = a, rs = a;
    if (c < C) {
        mu = *reinterpret_cast<const float4*>(mean + c);
        const float4 vr = *reinterpret_cast<const float4*>(var + c);
        rs = make_float4(rsqrtf(vr.x + eps), rsqrtf(vr.y + eps), rsqrtf(vr.z + eps), rsqrtf(vr.w + eps));
        auto acc = [&](const float4& zz, const float4& yy, const float4& dd) {
            const float gx = yy.x > 0.f ? dd.x : 0.f, gy = yy.y > 0.f ? dd.y : 0.f;
            const float gz = yy.z > 0.f ? dd.z : 0.f, gw = yy.w > 0.f ? dd.w : 0.f;
            a.x += gx; a.y += gy; a.z += gz; a.w += gw;
            b.x = fmaf(gx, (zz.x - mu.x) * rs.x, b.x); b.y = fmaf(gy, (zz.y - mu.y) * rs.y, b.y);
            b.z = fmaf(gz, (zz.z - mu.z) * rs.z, b.z); b.w = fmaf(gw, (zz.w - mu.w) * rs.w, b.w);
        };
        int r = r0 + rg;
        for (; r + 16 < r1; r += 24) {
            float4 zv[3], yv[3], dv[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const size_t o = (size_t)(r + 8 * u) * C + c;
                zv[u] = *reinterpret_cast<const float4*>(z + o);
                yv[u] = *reinterpret_cast<const float4*>(y + o);
                dv[u] = *reinterpret_cast<const float4*>(dy + o);
            }
#pragma unroll
            for (int u = 0; u < 3; u++) acc(zv[u], yv[u], dv[u]);
        }
        for (; r < r1; r += 8) {
            const size_t o = (size_t)r * C + c;
            acc(*reinterpret_cast<const float4*>(z + o), *reinterpret_cast<const float4*>(y + o),
                *reinterpret_cast<const float4*>(dy + o));
        }
    }
    __shared__ float4 sa[8][32], sb[8][32];
    sa[rg][lane32] = a;
    sb[rg][lane32] = b;
    __syncthreads();
    if (rg == 0 && c < C) {
        double p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; k++) {
            const float4 x = sa[k][lane32], w = sb[k][lane32];
            p[0] += x.x; p[1] += x.y; p[2] += x.z; p[3] += x.w;
            q[0] += w.x; q[1] += w.y; q[2] += w.z; q[3] += w.w;
        }
        for (int j = 0; j < 4; j++) { atomicAdd(&s1[c + j], p[j]); atomicAdd(&s2[c + j], q[j]); }
    }
    bn_grid_barrier(reinterpret_cast<unsigned long long*>(scratch + 2 * C), gridDim.x * gridDim.y);
    if (c >= C) return;
    float m1[4], m2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {        // the arithmetic of bn_bwd_apply_kernel
        const double t1 = __ldcg(s1 + c + j), t2 = __ldcg(s2 + c + j);
        m1[j] = (float)(t1 / M);
        m2[j] = (float)(t2 / M);
        if (blockIdx.y == 0 && rg == 0) dbeta[c + j] = (float)t1;
    }
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
    for (int r = r0 + rg; r < r1; r += 8) {
        const size_t o = (size_t)r * C + c;
        const float4 zz = *reinterpret_cast<const float4*>(z + o), yy = *reinterpret_cast<const float4*>(y + o);
        const float4 dd = *reinterpret_cast<const float4*>(dy + o);
        const float zc[4] = {zz.x, zz.y, zz.z, zz.w}, yc[4] = {yy.x, yy.y, yy.z, yy.w}, dc[4] = {dd.x, dd.y, dd.z, dd.w};
        float ov[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float xh = (zc[j] - muv[j]) * rsv[j];
            const float g = yc[j] > 0.f ? dc[j] : 0.f;
            ov[j] = rtf32(rsv[j] * (g - m1[j] - xh * m2[j]));
        }
        *reinterpret_cast<float4*>(dz + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
}

// ---------------------------------------------------------------- xyz head: conv3x3 128 -> 3 (+bias)
// add_inst_xyz_maps_local (monopsr_output_builder.py:95-108).  N=3 output channels: bandwidth
// bound on the (32,48,48,128) map features; one warp per output pixel, lanes over the 128
// channels (float4 each), weights [3][9][128] in shared memory.
// A warp owns kXyzPx = 8 consecutive pixels of one image row (W % 8 == 0): the 3 x 10 input pixels it needs are loaded
// once (30 LDG.128 per lane instead of 72), the nine (tap, output) weight vectors of a filter row sit in registers for
// the ten columns, and the 24 partial sums are reduced across the lanes with a halving butterfly (27 SHFL, not 120).
// (One warp per pixel re-read every input pixel nine times: 340 MB through L1/L2 for a 38 MB tensor, 62 us.)
constexpr int kXyzPx = 8;
template <int H_, int BIT>
__device__ __forceinline__ void xyz_halve(const float* v, float* u, int lane) {
    const bool up = (lane & BIT) != 0;
#pragma unroll
    for (int i = 0; i < H_; i++) {
        const float keep = up ? v[H_ + i] : v[i];
        const float send = up ? v[i] : v[H_ + i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
    }
}
__global__ void __launch_bounds__(256)
xyzhead_fwd_kernel(int nimg, int H, int W, const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ y) {
    __shared__ float4 sw[3 * 9 * 32];
    for (int i = threadIdx.x; i < 3 * 9 * 32; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(w)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpr = W / kXyzPx;                                      // warps per image row
    const long wid = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long)nimg * H * wpr) return;
    const int n = (int)(wid / ((long)H * wpr)), rem = (int)(wid % ((long)H * wpr)), h = rem / wpr, w0 = (rem % wpr) * kXyzPx;
    float acc[kXyzPx * 3];
#pragma unroll
    for (int i = 0; i < kXyzPx * 3; i++) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        const int ih = h + ky - 1;
        if (ih < 0 || ih >= H) continue;                             // warp-uniform
        float4 wt[3][3];                                             // [kx][output]
#pragma unroll
        for (int kx = 0; kx < 3; kx++)
#pragma unroll
            for (int o = 0; o < 3; o++) wt[kx][o] = sw[(o * 9 + ky * 3 + kx) * 32 + lane];
        const float4* row = reinterpret_cast<const float4*>(x + (((size_t)n * H + ih) * W) * 128) + lane;
#pragma unroll
        for (int cx = -1; cx <= kXyzPx; cx++) {
            const int iw = w0 + cx;
            if (iw < 0 || iw >= W) continue;                         // warp-uniform
            const float4 v = row[(size_t)iw * 32];
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const int p = cx - kx + 1;                           // output pixel that sees column cx through tap kx
                if (p < 0 || p >= kXyzPx) continue;                  // compile-time after unrolling
#pragma unroll
                for (int o = 0; o < 3; o++) {
                    const float4 q = wt[kx][o];
                    acc[p * 3 + o] += v.x * q.x + v.y * q.y + v.z * q.z + v.w * q.w;
                }
            }
        }
    }
    float u12[12], u6[6], u3[3];
    xyz_halve<12, 16>(acc, u12, lane);
    xyz_halve<6, 8>(u12, u6, lane);
    xyz_halve<3, 4>(u6, u3, lane);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 2);
        u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 1);
    }
    if ((lane & 3) == 0) {                                           // lanes 4j .. 4j+3 hold pixel j
        float* d = y + ((((size_t)n * H + h) * W) + w0 + (lane >> 2)) * 3;
        d[0] = u3[0] + bias[0];
        d[1] = u3[1] + bias[1];
        d[2] = u3[2] + bias[2];
    }
}
// dX[p][ci] = sum_{t,co} dY[p - off(t)][co] * w[co][t][ci]   (thread = 4 channels; a warp owns 8 consecutive pixels of a
// row, so the 27 weight vectors are fetched from shared memory once per 8 pixels and dY comes from 3 x 10 x 3 broadcast loads)
__global__ void __launch_bounds__(256)
xyzhead_dgrad_kernel(int nimg, int H, int W, const float* __restrict__ dy, const float* __restrict__ w,
                     float* __restrict__ dx) {
    __shared__ float4 sw[3 * 9 * 32];
    for (int i = threadIdx.x; i < 3 * 9 * 32; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(w)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpr = W / kXyzPx;
    const long wid = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long)nimg * H * wpr) return;
    const int n = (int)(wid / ((long)H * wpr)), rem = (int)(wid % ((long)H * wpr)), h = rem / wpr, w0 = (rem % wpr) * kXyzPx;
    float4 acc[kXyzPx];
#pragma unroll
    for (int i = 0; i < kXyzPx; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        const int ih = h - (ky - 1);                                 // dY row that reaches this row through filter row ky
        if (ih < 0 || ih >= H) continue;
        float4 wt[3][3];
#pragma unroll
        for (int kx = 0; kx < 3; kx++)
#pragma unroll
            for (int o = 0; o < 3; o++) wt[kx][o] = sw[(o * 9 + ky * 3 + kx) * 32 + lane];
        const float* grow = dy + (((size_t)n * H + ih) * W) * 3;
#pragma unroll
        for (int cx = -1; cx <= kXyzPx; cx++) {
            const int iw = w0 + cx;
            if (iw < 0 || iw >= W) continue;
            const float g0 = __ldg(grow + (size_t)iw * 3), g1 = __ldg(grow + (size_t)iw * 3 + 1), g2 = __ldg(grow + (size_t)iw * 3 + 2);
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const int p = cx + kx - 1;                           // dX pixel p reads dY[p - (kx - 1)] through tap kx
                if (p < 0 || p >= kXyzPx) continue;
                const float4 q0 = wt[kx][0], q1 = wt[kx][1], q2 = wt[kx][2];
                acc[p].x += g0 * q0.x + g1 * q1.x + g2 * q2.x;
                acc[p].y += g0 * q0.y + g1 * q1.y + g2 * q2.y;
                acc[p].z += g0 * q0.z + g1 * q1.z + g2 * q2.z;
                acc[p].w += g0 * q0.w + g1 * q1.w + g2 * q2.w;
            }
        }
    }
    float4* d = reinterpret_cast<float4*>(dx + ((((size_t)n * H + h) * W) + w0) * 128) + lane;
#pragma unroll
    for (int i = 0; i < kXyzPx; i++) d[(size_t)i * 32] = acc[i];
}
// dW[co][t][ci] += sum_p dY[p][co] * x[p + off(t)][ci];  db[co] += sum_p dY[p][co]
// CTA = 128 threads (ci) x a strip of pixels; 27 accumulators per thread.
__global__ void __launch_bounds__(128)
xyzhead_wgrad_kernel(int nimg, int H, int W, const float* __restrict__ x, const float* __restrict__ dy,
                     float* __restrict__ dw, float* __restrict__ db, int pix_per_cta) {
    const int ci = threadIdx.x;
    const long total = (long)nimg * H * W;
    const long p0 = (long)blockIdx.x * pix_per_cta, p1 = min(total, p0 + pix_per_cta);
    float acc[27];
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] = 0.f;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (long pix = p0; pix < p1; pix++) {
        const int n = (int)(pix / (H * W)), rem = (int)(pix % (H * W)), h = rem / W, ww = rem % W;
        const float g0 = dy[pix * 3], g1 = dy[pix * 3 + 1], g2 = dy[pix * 3 + 2];
        b0 += g0; b1 += g1; b2 += g2;
        if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const int ih = h + t / 3 - 1, iw = ww + t % 3 - 1;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            const float v = x[(((size_t)n * H + ih) * W + iw) * 128 + ci];
            acc[t] = fmaf(g0, v, acc[t]);
            acc[9 + t] = fmaf(g1, v, acc[9 + t]);
            acc[18 + t] = fmaf(g2, v, acc[18 + t]);
        }
    }
#pragma unroll
    for (int i = 0; i < 27; i++) atomicAdd(&dw[(size_t)i * 128 + ci], acc[i]);
    if (ci == 0) { atomicAdd(&db[0], b0); atomicAdd(&db[1], b1); atomicAdd(&db[2], b2); }
}

// ---------------------------------------------------------------- small dense heads (N <= 32 outputs)
// slim.fully_connected(features, n, activation_fn=None) for lwh / alpha / cen_y / cen_z
// (monopsr_output_builder.py:283,469,580,633).  w is [n][K].
__global__ void fc_small_fwd_kernel(int B, int K, int N, const float* __restrict__ x, int ldx,
                                    const float* __restrict__ w, const float* __restrict__ bias,
                                    float* __restrict__ y, int ldy) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * N) return;
    const int b = warp / N, n = warp % N;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(x[(size_t)b * ldx + k], w[(size_t)n * K + k], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[(size_t)b * ldy + n] = acc + bias[n];
}
// dx[b][k] (+)= sum_n dy[b][n] w[n][k];  dw[n][k] += sum_b dy[b][n] x[b][k];  db[n] += sum_b dy[b][n]
// grid (K tiles, B + N): row r < B of the grid computes dx[r][:], row B + n computes dw[n][:]; dy (B x N, a few hundred
// floats) is staged in shared memory.  (The first version looped over B x N + N x B dependent global loads in 8 CTAs:
// 16-119 us per call, four calls on the FC stacks' backward chain, which gates the towers' backward pass.)
__global__ void __launch_bounds__(128)
fc_small_bwd_kernel(int B, int K, int N, const float* __restrict__ x, int ldx,
                    const float* __restrict__ w, const float* __restrict__ dy, int ldy,
                    float* __restrict__ dx, int lddx, int accumulate_dx,
                    float* __restrict__ dw, float* __restrict__ db) {
    extern __shared__ float sdy[];                       // [B][N]
    for (int i = threadIdx.x; i < B * N; i += blockDim.x) sdy[i] = dy[(size_t)(i / N) * ldy + i % N];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (r < B) {
        if (k < K) {
            float acc = 0.f;
            for (int n = 0; n < N; n++) acc = fmaf(sdy[r * N + n], w[(size_t)n * K + k], acc);
            float* d = dx + (size_t)r * lddx + k;
            *d = accumulate_dx ? (*d + acc) : acc;
        }
    } else {
        const int n = r - B;
        if (k < K) {
            float acc = 0.f;
            for (int b = 0; b < B; b++) acc = fmaf(sdy[b * N + n], x[(size_t)b * ldx + k], acc);
            dw[(size_t)n * K + k] += acc;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            float acc = 0.f;
            for (int b = 0; b < B; b++) acc += sdy[b * N + n];
            db[n] += acc;
        }
    }
}

// ---------------------------------------------------------------- elementwise helpers
// y = relu(x + bias[c]) (optionally tf32-rounded): finishes split-K FC layers
__global__ void bias_relu_kernel(long total, int C, const float* __restrict__ x, int ldx, const float* __restrict__ bias,
                                 int relu, int round, float* __restrict__ y, int ldy) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long r = i / C;
    float v = x[r * ldx + c] + (bias ? bias[c] : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    y[r * ldy + c] = round ? rtf32(v) : v;
}
// g = (y>0 ? dy : 0) (tf32-rounded), colsum[c] += sum_r g   (ReLU + bias backward)
__global__ void __launch_bounds__(256)
relu_bwd_colsum_kernel(int M, int C, const float* __restrict__ y, int ldy, const float* __restrict__ dy, int lddy,
                       float* __restrict__ g, int ldg, float* __restrict__ colsum) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float s = 0.f;
    if (c < C)
        for (int r = r0 + rg; r < r1; r += 8) {
            const float v = y[(size_t)r * ldy + c] > 0.f ? dy[(size_t)r * lddy + c] : 0.f;
            const float rv = rtf32(v);
            g[(size_t)r * ldg + c] = rv;
            s += rv;
        }
    __shared__ float ss[8][32];
    ss[rg][threadIdx.x & 31] = s;
    __syncthreads();
    if (rg == 0 && c < C && colsum) {
        float a = 0.f;
        for (int k = 0; k < 8; k++) a += ss[k][threadIdx.x];
        atomicAdd(&colsum[c], a);
    }
}
__global__ void __launch_bounds__(256) zero_fill_kernel(long n, float* __restrict__ p) {
    const long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const long stride = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        p4[i] = z; p4[i + stride] = z; p4[i + 2 * stride] = z; p4[i + 3 * stride] = z;
    }
    for (; i < n4; i += stride) p4[i] = z;
    for (long j = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) p[j] = 0.f;
}
__global__ void add_inplace_kernel(long n, float* __restrict__ a, const float* __restrict__ b) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

}  // namespace mpb

// =================================================================== C ABI
using namespace mpb;
#define ST ((cudaStream_t)stream)
static inline unsigned nblk(long total, int t) { return (unsigned)((total + t - 1) / t); }

MPB_API int mpb_fold_bn(int cout, int K, const float* w, const float* gamma, const float* beta, const float* mean,
                        const float* var, float eps, float* wf, float* scale, float* shift, void* stream) {
    fold_bn_kernel<<<cout, 256, 0, ST>>>(cout, K, w, gamma, beta, mean, var, eps, wf, scale, shift);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_round_copy(long n, const float* src, float* dst, void* stream) {
    round_copy_kernel<<<(int)min((long)num_sms() * 16, (long)nblk(n, 1024)), 256, 0, ST>>>((size_t)n, src, dst);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_param_grad(int cout, int K, const float* w, const float* dw, const float* gamma, const float* mean,
                              const float* var, float eps, const float* dbeta, float* dgamma, void* stream) {
    bn_param_grad_kernel<<<cout, 256, 0, ST>>>(cout, K, w, dw, gamma, mean, var, eps, dbeta, dgamma);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_fold_bn_multi(int total_rows, const mpb_bn_layer* layers, const int* row2layer, float eps, void* stream) {
    if (total_rows <= 0 || !layers || !row2layer) return -1;
    fold_bn_multi_kernel<<<ceil_div(total_rows, 8), 256, 0, ST>>>(total_rows, layers, row2layer, eps);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_param_grad_range(int row_begin, int row_end, const mpb_bn_layer* layers, const int* row2layer, float eps,
                                    void* stream) {
    if (row_begin < 0 || row_end < row_begin || !layers || !row2layer) return -1;
    if (row_end == row_begin) return 0;
    bn_param_grad_multi_kernel<<<ceil_div(row_end - row_begin, 8), 256, 0, ST>>>(row_begin, row_end, layers, row2layer, eps);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_param_grad_multi(int total_rows, const mpb_bn_layer* layers, const int* row2layer, float eps, void* stream) {
    if (total_rows <= 0) return -1;
    return mpb_bn_param_grad_range(0, total_rows, layers, row2layer, eps, stream);
}
MPB_API int mpb_stem_fwd(int nimg, int Hin, int Win, const float* x, const float* wf, const float* shift, float* y,
                         void* stream) {
    const int Ho = (Hin + 6 - 7) / 2 + 1, Wo = (Win + 6 - 7) / 2 + 1;
    const int qpr = (Wo + kStemPx - 1) / kStemPx;
    stem_fwd_kernel<<<nblk((long)nimg * Ho * qpr, 32), 256, 0, ST>>>(nimg, Hin, Win, Ho, Wo, x, wf, shift, y);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_stem_wgrad(int nimg, int Hin, int Win, const float* x, const float* g, const float* scale, float* dw,
                           void* stream) {
    const int Ho = (Hin + 6 - 7) / 2 + 1, Wo = (Win + 6 - 7) / 2 + 1;
    const long total = (long)nimg * Ho * Wo;
    const int ppc = 64;
    stem_wgrad_kernel<<<nblk(total, ppc), 256, 0, ST>>>(nimg, Hin, Win, Ho, Wo, x, g, scale, dw, ppc);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_maxpool3s2_fwd(int nimg, int H, int W, int C, const float* x, float* y, void* stream) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    maxpool3s2_fwd_kernel<<<nblk((long)nimg * Ho * Wo * (C / 4), 256), 256, 0, ST>>>(
        nimg, H, W, C / 4, Ho, Wo, (const float4*)x, (float4*)y);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_maxpool3s2_bwd(int nimg, int H, int W, int C, const float* x, const float* dy, float* dx, void* stream) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    if (C % 4) return -1;
    maxpool3s2_bwd_kernel<<<nblk((long)nimg * H * W * (C / 4), 256), 256, 0, ST>>>(
        nimg, H, W, C / 4, Ho, Wo, reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy),
        reinterpret_cast<float4*>(dx));
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_maxpool2_fwd(int nimg, int H, int W, int C, const float* x, int ldx, float* y, int ldy, void* stream) {
    maxpool2_fwd_kernel<<<nblk((long)nimg * (H / 2) * (W / 2) * (C / 4), 256), 256, 0, ST>>>(nimg, H, W, C / 4, x, ldx, y, ldy);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_maxpool2_bwd(int nimg, int H, int W, int C, const float* x, int ldx, const float* dy, int ldy,
                             float* dx, int lddx, int accumulate, void* stream) {
    maxpool2_bwd_kernel<<<nblk((long)nimg * (H / 2) * (W / 2) * C, 256), 256, 0, ST>>>(nimg, H, W, C, x, ldx, dy, ldy, dx,
                                                                                      lddx, accumulate);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_crop_pool_fwd(int H, int W, int C, const float* feat, int nbox, const float* boxes_norm, int crop,
                              float* out, int ldo, void* stream) {
    if (C % 4 || ldo % 4 || (reinterpret_cast<uintptr_t>(feat) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return -1;
    crop_pool_fwd_kernel<<<nblk((long)nbox * (crop / 2) * (crop / 2) * (C / 4), 256), 256, 0, ST>>>(H, W, C, feat, nbox,
                                                                                                  boxes_norm, crop, out, ldo);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_crop_pool_bwd(int H, int W, int C, const float* feat, int nbox, const float* boxes_norm, int crop,
                              const float* dout, int ldd, float* dfeat, void* stream) {
    if (C % 4 || ldd % 4 || (reinterpret_cast<uintptr_t>(feat) & 15u) || (reinterpret_cast<uintptr_t>(dout) & 15u) ||
        (reinterpret_cast<uintptr_t>(dfeat) & 15u))
        return -1;
    MPB_CUDA_TRY(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)H * W * C, ST));
    crop_pool_bwd_kernel<<<nblk((long)nbox * (crop / 2) * (crop / 2) * (C / 4), 256), 256, 0, ST>>>(H, W, C, feat, nbox,
                                                                                                  boxes_norm, crop, dout, ldd, dfeat);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_resize_ac_fwd16(int nimg, int H, int W, int C, const float* x, int OH, int OW, float* y, void* y16,
                                int* overflow, void* stream) {
    if (y16 && C % 32) return -1;
    resize_ac_fwd_kernel<<<nblk((long)nimg * OH * OW * (C / 4), 256), 256, 0, ST>>>(nimg, H, W, C / 4, OH, OW,
                                                                                  (const float4*)x, (float4*)y,
                                                                                  (unsigned char*)y16, overflow);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_resize_ac_fwd(int nimg, int H, int W, int C, const float* x, int OH, int OW, float* y, void* stream) {
    return mpb_resize_ac_fwd16(nimg, H, W, C, x, OH, OW, y, nullptr, nullptr, stream);
}
MPB_API int mpb_resize_ac_bwd(int nimg, int H, int W, int C, const float* dy, int OH, int OW, float* dx, void* stream) {
    if (C % 4 || H < 2 || W < 2 || OH < H || OW < W || OH > 3 * H || OW > 3 * W) return -1;   // <= 6 taps per axis
    resize_ac_bwd_kernel<<<nblk((long)nimg * H * W * (C / 4), 256), 256, 0, ST>>>(
        nimg, H, W, C / 4, OH, OW, reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx));
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_train_fwd16(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean,
                               float* var, float* moving_mean, float* moving_var, float decay, double* scratch,
                               void* y16, int* overflow, void* stream) {
    if (y16 && C % 32) return -1;
    MPB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, ST));
    if (C % 4) return -1;
    dim3 g(ceil_div(C, 128), min(num_sms() * 4, ceil_div(M, 64)));
    bn_stats_kernel<<<g, 256, 0, ST>>>(M, C, z, scratch, scratch + C);
    MPB_LAUNCH_CHECK();
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, ST>>>(M, C, scratch, scratch + C, mean, var, moving_mean, moving_var, decay);
    MPB_LAUNCH_CHECK();
    bn_apply_kernel<<<nblk((long)M * C / 4, 256), 256, 0, ST>>>((long)M * C / 4, C / 4, (const float4*)z, mean, var, beta, eps,
                                                              (float4*)y, (unsigned char*)y16, overflow);
    MPB_LAUNCH_CHECK();
    return 0;
}
// single-launch variants (see bn_train_fwd_fused_kernel): scratch = 2C + 2 doubles
// the launch is refused (-2) unless two CTAs of the kernel fit one SM: the barrier needs the whole grid resident
template <typename K>
static bool bn_fused_fits(K kernel) {
    int per_sm = 0;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) == cudaSuccess && per_sm >= 2;
}
static dim3 bn_fused_grid(int M, int C) {
    const int gx = ceil_div(C, 128);
    int gy = max(1, (2 * num_sms()) / gx);                  // 2 CTAs per SM in total: resident beside anything
    gy = min(gy, ceil_div(M, 8));
    return dim3(gx, gy);
}
MPB_API int mpb_bn_train_fwd_fused(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean,
                                   float* var, float* moving_mean, float* moving_var, float decay, double* scratch,
                                   void* y16, int* overflow, void* stream) {
    if (M <= 0 || C <= 0 || C % 4 || (y16 && C % 32) || !z || !beta || !y || !mean || !var || !scratch) return -1;
    static const bool fits = bn_fused_fits(bn_train_fwd_fused_kernel);
    if (!fits) return -2;
    MPB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * (2 * C + 2), ST));
    bn_train_fwd_fused_kernel<<<bn_fused_grid(M, C), 256, 0, ST>>>(M, C, z, beta, eps, y, mean, var, moving_mean, moving_var,
                                                                  decay, scratch, (unsigned char*)y16, overflow);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_train_bwd_fused(int M, int C, const float* z, const float* mean, const float* var, float eps,
                                   const float* y, const float* dy, float* dz, float* dbeta, double* scratch, void* stream) {
    if (M <= 0 || C <= 0 || C % 4 || !z || !mean || !var || !y || !dy || !dz || !dbeta || !scratch) return -1;
    static const bool fits = bn_fused_fits(bn_train_bwd_fused_kernel);
    if (!fits) return -2;
    MPB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * (2 * C + 2), ST));
    bn_train_bwd_fused_kernel<<<bn_fused_grid(M, C), 256, 0, ST>>>(M, C, z, y, dy, mean, var, eps, scratch, dz, dbeta);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_train_fwd(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean,
                             float* var, float* moving_mean, float* moving_var, float decay, double* scratch,
                             void* stream) {
    return mpb_bn_train_fwd16(M, C, z, beta, eps, y, mean, var, moving_mean, moving_var, decay, scratch, nullptr, nullptr,
                              stream);
}
MPB_API int mpb_set_operand_rounding(int on) {
    const int v = on ? 1 : 0;
    MPB_CUDA_TRY(cudaMemcpyToSymbol(mpb::c_round_operands, &v, sizeof(v)));
    return 0;
}
// slim.batch_norm(is_training=False): the moving statistics instead of the batch's (validation / inference graphs:
// MonoPSRModel is built with is_training = (train_val_test == 'train'), monopsr_model.py:139, net_builder.py:39,79,87)
MPB_API int mpb_bn_infer_fwd16(int M, int C, const float* z, const float* beta, const float* moving_mean,
                               const float* moving_var, float eps, float* y, void* y16, int* overflow, void* stream) {
    if (M <= 0 || C <= 0 || C % 4 || !z || !beta || !moving_mean || !moving_var || !y || (y16 && C % 32)) return -1;
    bn_apply_kernel<<<nblk((long)M * C / 4, 256), 256, 0, ST>>>((long)M * C / 4, C / 4, (const float4*)z, moving_mean,
                                                              moving_var, beta, eps, (float4*)y, (unsigned char*)y16, overflow);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bn_infer_fwd(int M, int C, const float* z, const float* beta, const float* moving_mean,
                             const float* moving_var, float eps, float* y, void* stream) {
    return mpb_bn_infer_fwd16(M, C, z, beta, moving_mean, moving_var, eps, y, nullptr, nullptr, stream);
}
MPB_API int mpb_bn_train_bwd(int M, int C, const float* z, const float* mean, const float* var, float eps, const float* y,
                             const float* dy, float* dz, float* dbeta, double* scratch, void* stream) {
    MPB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, ST));
    if (C % 4) return -1;
    dim3 g(ceil_div(C, 128), min(num_sms() * 4, ceil_div(M, 64)));
    bn_bwd_reduce_kernel<<<g, 256, 0, ST>>>(M, C, z, y, dy, mean, var, eps, scratch, scratch + C);
    MPB_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<nblk((long)M * C / 4, 256), 256, 0, ST>>>((long)M * C / 4, M, C, z, y, dy, mean, var, eps, scratch,
                                                                  scratch + C, dz, dbeta);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_xyzhead_fwd(int nimg, int H, int W, const float* x, const float* w, const float* bias, float* y, void* stream) {
    if (W % kXyzPx) return -1;
    xyzhead_fwd_kernel<<<nblk((long)nimg * H * (W / kXyzPx), 8), 256, 0, ST>>>(nimg, H, W, x, w, bias, y);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_xyzhead_dgrad(int nimg, int H, int W, const float* w, const float* dy, float* dx, void* stream) {
    if (W % kXyzPx) return -1;
    xyzhead_dgrad_kernel<<<nblk((long)nimg * H * (W / kXyzPx), 8), 256, 0, ST>>>(nimg, H, W, dy, w, dx);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_xyzhead_wgrad(int nimg, int H, int W, const float* x, const float* dy, float* dw, float* db, void* stream) {
    const int ppc = 64;
    xyzhead_wgrad_kernel<<<nblk((long)nimg * H * W, ppc), 128, 0, ST>>>(nimg, H, W, x, dy, dw, db, ppc);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_xyzhead_bwd(int nimg, int H, int W, const float* x, const float* w, const float* dy, float* dx, float* dw,
                            float* db, void* stream) {
    const int st = mpb_xyzhead_dgrad(nimg, H, W, w, dy, dx, stream);
    return st ? st : mpb_xyzhead_wgrad(nimg, H, W, x, dy, dw, db, stream);
}
MPB_API int mpb_fc_small_fwd(int B, int K, int N, const float* x, int ldx, const float* w, const float* bias, float* y,
                             int ldy, void* stream) {
    fc_small_fwd_kernel<<<nblk((long)B * N * 32, 256), 256, 0, ST>>>(B, K, N, x, ldx, w, bias, y, ldy);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_fc_small_bwd(int B, int K, int N, const float* x, int ldx, const float* w, const float* dy, int ldy,
                             float* dx, int lddx, int accumulate_dx, float* dw, float* db, void* stream) {
    if (B <= 0 || K <= 0 || N <= 0 || (size_t)B * N * sizeof(float) > 48 * 1024) return -1;
    fc_small_bwd_kernel<<<dim3(ceil_div(K, 128), B + N), 128, sizeof(float) * B * N, ST>>>(B, K, N, x, ldx, w, dy, ldy, dx, lddx,
                                                                                       accumulate_dx, dw, db);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_bias_relu(long rows, int C, const float* x, int ldx, const float* bias, int relu, int round, float* y,
                          int ldy, void* stream) {
    bias_relu_kernel<<<nblk(rows * C, 256), 256, 0, ST>>>(rows * C, C, x, ldx, bias, relu, round, y, ldy);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_relu_bwd_colsum(int M, int C, const float* y, int ldy, const float* dy, int lddy, float* g, int ldg,
                                float* colsum, void* stream) {
    dim3 grid(ceil_div(C, 32), min(256, ceil_div(M, 64)));
    relu_bwd_colsum_kernel<<<grid, 256, 0, ST>>>(M, C, y, ldy, dy, lddy, g, ldg, colsum);
    MPB_LAUNCH_CHECK();
    return 0;
}
// zero-fill that leaves the SMs to others: ONE 256-thread CTA per SM streaming 16-byte stores.  (A framework fill kernel
// with ~100k CTAs takes every CTA slot of the chip for its 55 us, and the step's first kernels queue behind it.)
MPB_API int mpb_zero_fill(long n, float* p, void* stream) {
    if (n < 0 || (n && !p) || (reinterpret_cast<uintptr_t>(p) & 15u)) return -1;
    if (n == 0) return 0;
    zero_fill_kernel<<<num_sms(), 256, 0, ST>>>(n, p);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_add_inplace(long n, float* a, const float* b, void* stream) {
    add_inplace_kernel<<<nblk(n, 256), 256, 0, ST>>>(n, a, b);
    MPB_LAUNCH_CHECK();
    return 0;
}
