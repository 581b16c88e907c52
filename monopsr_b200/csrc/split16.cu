// monopsr_b200/csrc/split16.cu -- fp16 hi/lo "split copies" of fp32 GEMM operands for the h3 forward kernel
// (csrc/tc_gemm.cu, H3 branch of tc_gemm_tma_kernel; include/monopsr_b200_net.h::mpb_tc_gemm_h3).
//
// x = hi + lo + O(2^-22 |x|):  hi = fp16(x) (round to nearest), lo = fp16(x - hi)  (x - hi is exact in fp32).
// A split copy has the byte geometry of the fp32 matrix it shadows: each group of 32 consecutive elements of a row
// (128 bytes = one SWIZZLE_128B shared-memory row of the GEMM) becomes 32 hi halves followed by 32 lo halves
// ([hi | lo], A operand / activations) or the other way round ([lo | hi], B operand / weights), so that
//     sum over the 64 halves of  A_row * B_row  =  A_hi.B_lo + A_lo.B_hi      (one K=64 kind::f16 block)
// and the first half of A against the second half of B is A_hi.B_hi.
// fp16 has 5 exponent bits: activations are split as they are (|x| > 65504 raises the sticky overflow flag; elements
// below 2^-3 keep fewer than 11 bits of lo, an ABSOLUTE error <= 2^-25, harmless next to O(1) activations), weights
// are pre-scaled per output channel by the power of two that puts the row maximum into [2^13, 2^14).
#include "common.cuh"
#include "../../include/monopsr_b200_net.h"
#include <cuda_fp16.h>

namespace mpb {

__device__ __forceinline__ void split16_pair_(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// one thread = 4 consecutive elements (one 16-byte load, two 8-byte stores); bandwidth-bound: 4 B read + 4 B written
// per element
__global__ void __launch_bounds__(256)
split16_kernel(long total4, int C4, const float* __restrict__ src, int lds, unsigned char* __restrict__ dst, int ldd,
               int b_operand, int* __restrict__ overflow) {
    float mx = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const long r = i / C4;
        const int c = (int)(i - r * C4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds + c));
        uint32_t h0, l0, h1, l1;
        split16_pair_(v.x, v.y, h0, l0);
        split16_pair_(v.z, v.w, h1, l1);
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        unsigned char* d = dst + ((size_t)r * ldd + (c & ~31)) * 4 + (c & 31) * 2;
        *reinterpret_cast<uint2*>(d + (b_operand ? 64 : 0)) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(d + (b_operand ? 0 : 64)) = make_uint2(l0, l1);
    }
    if (overflow && mx > 65504.f) *overflow = 1;
}

// One warp per (layer, output channel) row: pass 1 = row maximum of |w * s|, pass 2 = scaled [lo | hi] split copy.
// (The second read of the row comes from L1/L2: rows are 256 B .. 72 KB.)
// CH = float4 chunks per lane kept in registers between the passes (rows of <= 128 * CH floats); 0 = re-read path.
// The rows are handed over in classes by length (rowlist): short rows take a few registers and run at full occupancy,
// and no warp carries 72 registers of buffer for a 256-float row.
template <int CH>
__global__ void __launch_bounds__(256)
split16_weights_multi_kernel(int nrows, const int* __restrict__ rowlist, const mpb_w16_layer* __restrict__ layers,
                             const int* __restrict__ row2layer, float eps) {
    const int ri = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ri >= nrows) return;
    const int row = rowlist ? rowlist[ri] : ri;
    const mpb_w16_layer L = layers[row2layer[row]];
    const int co = row - L.row0;
    const float s = L.gamma ? L.gamma[co] * rsqrtf(L.var[co] + eps) : 1.f;
    const float4* w4 = reinterpret_cast<const float4*>(L.w + (size_t)co * L.K);
    const int n4 = L.K >> 2;
    // rows of up to 2304 floats (every tower conv) stay in registers between the two passes: the weights are read once
    constexpr int kRegChunks = CH > 0 ? CH : 1;
    const bool in_regs = CH > 0 && n4 <= kRegChunks * 32;
    float4 buf[kRegChunks];
    float mx = 0.f;
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < kRegChunks; j++) {
            const int k = lane + 32 * j;
            buf[j] = k < n4 ? __ldg(w4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(buf[j].x), fabsf(buf[j].y))), fmaxf(fabsf(buf[j].z), fabsf(buf[j].w)));
        }
    } else {
        for (int k = lane; k < n4; k += 32) {
            const float4 v = __ldg(w4 + k);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mx *= fabsf(s);
    // power of two sc with mx * sc in [2^13, 2^14); exponent clamped so that sc and 1 / sc stay normal
    float sc = 1.f;
    if (mx > 0.f && mx < 3.0e38f) {
        int e;
        frexpf(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
        sc = ldexpf(1.f, max(-100, min(100, 14 - e)));
    }
    if (lane == 0) {
        L.inv_scale[co] = 1.f / sc;
        if (L.scale) L.scale[co] = s;
        if (L.shift) L.shift[co] = L.beta[co] - L.mean[co] * s;
    }
    const float m = s * sc;
    unsigned char* d = reinterpret_cast<unsigned char*>(L.w16) + (size_t)co * L.K * 4;
    float4* wf4 = L.wf ? reinterpret_cast<float4*>(L.wf + (size_t)co * L.K) : nullptr;
    auto emit = [&](int k, const float4 v) {
        if (wf4) wf4[k] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);      // folded fp32 weights (backward operand)
        uint32_t h0, l0, h1, l1;
        split16_pair_(v.x * m, v.y * m, h0, l0);
        split16_pair_(v.z * m, v.w * m, h1, l1);
        const int c = k * 4;
        unsigned char* q = d + (size_t)(c & ~31) * 4 + (c & 31) * 2;
        *reinterpret_cast<uint2*>(q) = make_uint2(l0, l1);            // [lo | hi]
        *reinterpret_cast<uint2*>(q + 64) = make_uint2(h0, h1);
    };
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < kRegChunks; j++) {
            const int k = lane + 32 * j;
            if (k < n4) emit(k, buf[j]);
        }
    } else {
        for (int k = lane; k < n4; k += 32) emit(k, __ldg(w4 + k));
    }
}

}  // namespace mpb

MPB_API int mpb_split16(long rows, int C, const float* src, int lds, void* dst16, int ldd, int b_operand, int* overflow,
                        void* stream) {
    using namespace mpb;
    if (rows < 0 || C <= 0 || C % 32 || lds % 4 || ldd % 32 || !src || !dst16) return -1;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(dst16) & 127u)) return -1;
    if (rows == 0) return 0;
    const long total4 = rows * (C / 4);
    const long blocks = (total4 + 255) / 256;
    const int grid = (int)(blocks < (long)num_sms() * 16 ? blocks : (long)num_sms() * 16);
    split16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total4, C / 4, src, lds, (unsigned char*)dst16, ldd, b_operand,
                                                          overflow);
    MPB_LAUNCH_CHECK();
    return 0;
}

MPB_API int mpb_split16_weights_rows(int nrows, const int* rowlist, int max_row_floats, const mpb_w16_layer* layers,
                                     const int* row2layer, float eps, void* stream) {
    using namespace mpb;
    if (nrows < 0 || !layers || !row2layer) return -1;
    if (nrows == 0) return 0;
    const dim3 g(ceil_div(nrows, 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (max_row_floats <= 256) split16_weights_multi_kernel<2><<<g, 256, 0, st>>>(nrows, rowlist, layers, row2layer, eps);
    else if (max_row_floats <= 1024) split16_weights_multi_kernel<8><<<g, 256, 0, st>>>(nrows, rowlist, layers, row2layer, eps);
    else if (max_row_floats <= 2304) split16_weights_multi_kernel<18><<<g, 256, 0, st>>>(nrows, rowlist, layers, row2layer, eps);
    else split16_weights_multi_kernel<0><<<g, 256, 0, st>>>(nrows, rowlist, layers, row2layer, eps);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_split16_weights_multi(int total_rows, const mpb_w16_layer* layers, const int* row2layer, float eps,
                                      void* stream) {
    if (total_rows <= 0) return -1;
    return mpb_split16_weights_rows(total_rows, nullptr, 1 << 30, layers, row2layer, eps, stream);
}
