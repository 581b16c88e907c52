// monopsr_b200/csrc/approxmatch.cu -- approximate-EMD ops for sm_100a.
//
// Replaces approxmatch / matchcost / matchcostgrad{1,2} and their launchers
// (reference: src/tf_ops/approxmatch/tf_approxmatch_g.cu:1-295).  Same mathematics as
// the reference GPU kernel (10 annealing levels level=-4^j, j=7..-2; three n x m sweeps
// per level; match stored (b,m,n) with [l,k] at l*n+k), different machine mapping:
//
//   * one THREAD-BLOCK CLUSTER per batch element instead of one CTA (the reference keeps
//     at most 32 of 148 SMs busy).  Each CTA of the cluster owns a slice of the dataset
//     points (k) and a slice of the query points (l); per-level ratios are all-gathered
//     through distributed shared memory, two cluster barriers per level;
//   * both clouds live in shared memory as float4 {x,y,z,weight} for the whole kernel;
//     the per-point state (remainL/remainR/ratioL/ratioR) never touches HBM -- the
//     caller's `temp` scratch is unused;
//   * `match` is written exactly ONCE: the per-level ratioL/ratioR are kept on chip and
//     the final pass evaluates  match[l,k] = sum_j exp(level_j d_kl) ratioL_j[k] ratioR_j[l]
//     (the reference zero-fills match and read-modify-writes it in HBM ten times);
//   * the level-0 pass (exp == 1) has no third sweep (nothing consumes remainL after it).
//
// The op is MUFU(ex2)/FP32 bound with a 10-deep serial dependency chain; HBM only
// matters for the single match write.  See DESIGN.md "approxmatch roofline".
#include "common.cuh"
#include "../../include/monopsr_b200_tfops.h"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace mpb {

constexpr int kAmThreads = 512;
#ifndef MPB_AM_UNROLL
#define MPB_AM_UNROLL 8
#endif
constexpr int kAmUnroll = MPB_AM_UNROLL;     // candidate pairs in flight per lane in a sweep
#ifndef MPB_AM_UNROLL_FINAL
#define MPB_AM_UNROLL_FINAL 1
#endif
constexpr int kAmUnrollFinal = MPB_AM_UNROLL_FINAL;      // query rows in flight per lane in the final pass
constexpr int kAmLevels = 10;   // j = 7 .. -2  (tf_approxmatch_g.cu:21)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float am_d2(const float4& a, const float4& b) {
    float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

// Blackwell packed fp32 (two IEEE operations per instruction): the sweeps are FP32-issue bound
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// A cloud in shared memory: points stored in PAIRS, structure-of-arrays inside a pair:
//   {x0,x1,y0,y1} {z0,z1,w0,w1}   (8 floats per 2 points) so a pair loads as two LDS.128 whose halves
// are already the packed operands.  Accessors for single points:
__device__ __forceinline__ int am_ix(int k, int c) { return (k >> 1) * 8 + c * 2 + (k & 1); }   // c: 0 x,1 y,2 z,3 w
__device__ __forceinline__ float4 am_point(const float* cl, int k) {
    return make_float4(cl[am_ix(k, 0)], cl[am_ix(k, 1)], cl[am_ix(k, 2)], cl[am_ix(k, 3)]);
}

struct AmSmem {
    // carve-up of dynamic shared memory; all sizes in floats
    int n_pad, m_pad, own_n, own_m;
    __host__ __device__ size_t p1() const { return 0; }                                   // float4[n_pad]
    __host__ __device__ size_t p2() const { return p1() + 4 * (size_t)n_pad; }            // float4[m_pad]
    __host__ __device__ size_t rr() const { return p2() + 4 * (size_t)m_pad; }            // remainR full [m_pad]
    __host__ __device__ size_t hl() const { return rr() + m_pad; }                        // ratioL history [levels][n_pad]
    __host__ __device__ size_t hr() const { return hl() + (size_t)kAmLevels * n_pad; }    // ratioR history own [own_m][12]
    __host__ __device__ size_t rl() const { return hr() + 12 * (size_t)own_m; }           // remainL own [own_n]
    __host__ __device__ size_t total() const { return rl() + own_n; }
};

// Pick S (lanes sharing one owner point) so that owners*S fills the CTA in whole rounds.
__host__ __device__ inline int am_pick_split(int owners) {
    int best_s = 1;
    float best_w = 1e30f;
#ifndef MPB_AM_MAXS
#define MPB_AM_MAXS 2    // wider splits make 8 lanes read 32 B-strided pairs (bank conflicts): measured slower
#endif
    for (int s = 1; s <= MPB_AM_MAXS; s <<= 1) {
        float units = (float)owners * s / kAmThreads;
        float w = ceilf(units) / units;
        if (w < best_w - 1e-3f) {
            best_w = w;
            best_s = s;
        }
    }
    return best_s;
}

// sum over the `cnt` points of `other` (pair-SoA, weights in the w slots), pairs strided over S lanes,
// of exp2(c*d2)*w -- two candidates per packed instruction
__device__ __forceinline__ float am_sweep(const float4 me, const float* __restrict__ other, int cnt,
                                          int s, int S, float c) {
    const float4* o4 = reinterpret_cast<const float4*>(other);
    const uint64_t mx = pk(me.x, me.x), my = pk(me.y, me.y), mz = pk(me.z, me.z), c2 = pk(c, c);
    uint64_t acc = pk(0.f, 0.f);
    const int npairs = (cnt + 1) >> 1;
#pragma unroll(kAmUnroll)
    for (int j = s; j < npairs; j += S) {
        const float4 a = o4[2 * j], b = o4[2 * j + 1];      // {x0,x1,y0,y1} {z0,z1,w0,w1}
        const uint64_t dx = sub2(pk(a.x, a.y), mx), dy = sub2(pk(a.z, a.w), my), dz = sub2(pk(b.x, b.y), mz);
        const uint64_t d2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float a0, a1;
        upk(mul2(c2, d2), a0, a1);
        acc = fma2(pk(ex2_approx(a0), ex2_approx(a1)), pk(b.z, b.w), acc);
    }
    float lo, hi;
    upk(acc, lo, hi);
    return lo + hi;
}

__device__ __forceinline__ float group_sum(float v, int S) {
    for (int o = S >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kAmThreads, 1)
approxmatch_cluster_kernel(int n, int m, const float* __restrict__ xyz1,
                           const float* __restrict__ xyz2, float* __restrict__ match) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = cluster.num_blocks();
    const int rank = cluster.block_rank();
    const int bi = blockIdx.x / C;
    const int tid = threadIdx.x;

    AmSmem L;
    L.n_pad = (n + 3) & ~3;
    L.m_pad = (m + 3) & ~3;
    L.own_n = ceil_div(n, C);
    L.own_m = ceil_div(m, C);
    float* P1 = smem + L.p1();     // pair-SoA clouds (see am_ix)
    float* P2 = smem + L.p2();
    float* RR = smem + L.rr();
    float* HL = smem + L.hl();
    float* HR = smem + L.hr();
    float* RL = smem + L.rl();

    const int k0 = min(n, rank * L.own_n), k1 = min(n, k0 + L.own_n);
    const int l0 = min(m, rank * L.own_m), l1 = min(m, l0 + L.own_m);

    float multiL, multiR;   // integer division as in tf_approxmatch_g.cu:4-10 (quirk Q5)
    if (n >= m) { multiL = 1.f; multiR = (float)(n / m); }
    else        { multiL = (float)(m / n); multiR = 1.f; }

    const float* g1 = xyz1 + (size_t)bi * n * 3;
    const float* g2 = xyz2 + (size_t)bi * m * 3;
    for (int t = tid; t < 4 * L.n_pad; t += kAmThreads) P1[t] = 0.f;     // pad points: weight 0
    for (int t = tid; t < 4 * L.m_pad; t += kAmThreads) P2[t] = 0.f;
    __syncthreads();
    for (int t = tid; t < n * 3; t += kAmThreads) P1[am_ix(t / 3, t % 3)] = g1[t];
    for (int t = tid; t < m * 3; t += kAmThreads) P2[am_ix(t / 3, t % 3)] = g2[t];
    for (int t = tid; t < m; t += kAmThreads) { RR[t] = multiR; P2[am_ix(t, 3)] = multiR; }
    for (int t = tid; t < k1 - k0; t += kAmThreads) RL[t] = multiL;
    __syncthreads();
    cluster.sync();   // every CTA's smem is initialised before any remote write lands

    const int SL = am_pick_split(L.own_n);   // lanes per owned dataset point
    const int SR = am_pick_split(L.own_m);   // lanes per owned query point

    for (int lev = 0; lev < kAmLevels; lev++) {
        const int j = 7 - lev;
        const float level = (j == -2) ? 0.f : -exp2f(2.f * (float)j);   // -4^j
        const float c = level * kLog2e;
        // ---- sweep 1: ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d) remainR[l]) ----
        for (int u = tid; u < ((k1 - k0) * SL + 31) / 32 * 32; u += kAmThreads) {
            int o = u / SL, s = u % SL;
            bool act = o < k1 - k0;
            float sum = 0.f;
            if (act) sum = am_sweep(am_point(P1, k0 + o), P2, m, s, SL, c);
            sum = group_sum(sum, SL);
            if (act && s == 0) {
                float ratio = RL[o] / (1e-9f + sum);
                for (int r = 0; r < C; r++) {   // all-gather through DSMEM
                    float* rs = cluster.map_shared_rank(smem, r);
                    rs[L.p1() + am_ix(k0 + o, 3)] = ratio;
                    rs[L.hl() + (size_t)lev * L.n_pad + k0 + o] = ratio;
                }
            }
        }
        cluster.sync();
        // ---- sweep 2: per query l ----
        for (int u = tid; u < ((l1 - l0) * SR + 31) / 32 * 32; u += kAmThreads) {
            int o = u / SR, s = u % SR;
            bool act = o < l1 - l0;
            float sum = 0.f;
            if (act) sum = am_sweep(am_point(P2, l0 + o), P1, n, s, SR, c);
            sum = group_sum(sum, SR);
            if (act && s == 0) {
                float rem = RR[l0 + o];
                float sumr = sum * rem;
                float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
                float ratio = consumption * rem;
                float nrem = fmaxf(0.0f, rem - sumr);
                HR[o * 12 + lev] = ratio;
                for (int r = 0; r < C; r++) {
                    float* rs = cluster.map_shared_rank(smem, r);
                    rs[L.p2() + am_ix(l0 + o, 3)] = ratio;
                    rs[L.rr() + l0 + o] = nrem;
                }
            }
        }
        cluster.sync();
        if (lev == kAmLevels - 1) break;   // nothing consumes remainL after the last level
        // ---- sweep 3: remainL[k] -= ratioL[k] * sum_l exp(level d) ratioR[l] ----
        for (int u = tid; u < ((k1 - k0) * SL + 31) / 32 * 32; u += kAmThreads) {
            int o = u / SL, s = u % SL;
            bool act = o < k1 - k0;
            float sum = 0.f;
            if (act) sum = am_sweep(am_point(P1, k0 + o), P2, m, s, SL, c);
            sum = group_sum(sum, SL);
            if (act && s == 0) RL[o] = fmaxf(0.0f, RL[o] - sum * P1[am_ix(k0 + o, 3)]);
        }
        __syncthreads();
        for (int t = tid; t < m; t += kAmThreads) P2[am_ix(t, 3)] = RR[t];   // weights for next sweep 1
        __syncthreads();
    }

    // ---- final pass: write match[l, :] for own query rows, once, coalesced over k ----
    float* mout = match + (size_t)bi * n * m;
    float cl[kAmLevels];
#pragma unroll
    for (int q = 0; q < kAmLevels; q++) {
        int j = 7 - q;
        cl[q] = (j == -2) ? 0.f : -exp2f(2.f * (float)j) * kLog2e;
    }
    for (int kb = 0; kb < n; kb += kAmThreads) {
        const int k = kb + tid;
        const bool act = k < n;
        float4 me = act ? am_point(P1, k) : make_float4(0.f, 0.f, 0.f, 0.f);
        float rl[kAmLevels];
#pragma unroll
        for (int q = 0; q < kAmLevels; q++) rl[q] = act ? HL[(size_t)q * L.n_pad + k] : 0.f;
#pragma unroll(kAmUnrollFinal)
        for (int l = l0; l < l1; l++) {
            const float4 o = am_point(P2, l);
            const float4* hr4 = reinterpret_cast<const float4*>(HR + (size_t)(l - l0) * 12);
            float4 ra = hr4[0], rb = hr4[1], rc = hr4[2];
            float rr[12] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w, rc.x, rc.y, rc.z, rc.w};
            float d = am_d2(me, o);
            float acc = rl[kAmLevels - 1] * rr[kAmLevels - 1];   // level 0: exp == 1
#pragma unroll
            for (int q = 0; q < kAmLevels - 1; q++)
                acc = fmaf(ex2_approx(cl[q] * d), rl[q] * rr[q], acc);
            if (act) __stcs(&mout[(size_t)l * n + k], acc);   // streaming: written once, not re-read here
        }
    }
    cluster.sync();   // keep smem alive until all remote accesses in the cluster are done
}

// ---- generic fallback (clouds too large for the on-chip path): one CTA per batch
// element, state in global scratch, match read-modify-written per level.  Correctness
// path only; same sweeps as above.
__global__ void __launch_bounds__(kAmThreads)
approxmatch_fallback_kernel(int b, int n, int m, const float* __restrict__ xyz1,
                            const float* __restrict__ xyz2, float* __restrict__ match,
                            float* __restrict__ scratch) {
    const int tid = threadIdx.x;
    float multiL, multiR;
    if (n >= m) { multiL = 1.f; multiR = (float)(n / m); }
    else        { multiL = (float)(m / n); multiR = 1.f; }
    for (int bi = blockIdx.x; bi < b; bi += gridDim.x) {
        float* remainL = scratch + (size_t)bi * (n + m) * 2;
        float* remainR = remainL + n;
        float* ratioL = remainR + m;
        float* ratioR = ratioL + n;
        const float* p1 = xyz1 + (size_t)bi * n * 3;
        const float* p2 = xyz2 + (size_t)bi * m * 3;
        float* mt = match + (size_t)bi * n * m;
        for (size_t t = tid; t < (size_t)n * m; t += kAmThreads) mt[t] = 0.f;
        for (int t = tid; t < n; t += kAmThreads) remainL[t] = multiL;
        for (int t = tid; t < m; t += kAmThreads) remainR[t] = multiR;
        __syncthreads();
        for (int lev = 0; lev < kAmLevels; lev++) {
            const int j = 7 - lev;
            const float c = ((j == -2) ? 0.f : -exp2f(2.f * (float)j)) * kLog2e;
            for (int k = tid; k < n; k += kAmThreads) {
                float4 me = make_float4(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2], 0.f);
                float sum = 1e-9f;
                for (int l = 0; l < m; l++) {
                    float4 o = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
                    sum = fmaf(ex2_approx(c * am_d2(me, o)), remainR[l], sum);
                }
                ratioL[k] = remainL[k] / sum;
            }
            __syncthreads();
            for (int l = tid; l < m; l += kAmThreads) {
                float4 me = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
                float sum = 0.f;
                for (int k = 0; k < n; k++) {
                    float4 o = make_float4(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2], 0.f);
                    sum = fmaf(ex2_approx(c * am_d2(me, o)), ratioL[k], sum);
                }
                float rem = remainR[l];
                float sumr = sum * rem;
                float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * rem;
                remainR[l] = fmaxf(0.0f, rem - sumr);
            }
            __syncthreads();
            for (int k = tid; k < n; k += kAmThreads) {
                float4 me = make_float4(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2], 0.f);
                float rl = ratioL[k];
                float sum = 0.f;
                for (int l = 0; l < m; l++) {
                    float4 o = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
                    float w = ex2_approx(c * am_d2(me, o)) * rl * ratioR[l];
                    mt[(size_t)l * n + k] += w;
                    sum += w;
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - sum);
            }
            __syncthreads();
        }
    }
}

// ---- matchcost: cost[b] = sum_{l,k} sqrt(d2) * match[l,k]  (tf_approxmatch_g.cu:183-225)
// Grid (row tiles, b); a CTA streams kMcRows rows of match once (coalesced over k),
// block-reduces, and writes one partial; a second tiny kernel sums the partials in a
// fixed order (deterministic, no atomics).
constexpr int kMcThreads = 256;
constexpr int kMcRows = 16;

__global__ void __launch_bounds__(kMcThreads)
matchcost_partial_kernel(int n, int m, const float* __restrict__ xyz1,
                         const float* __restrict__ xyz2, const float* __restrict__ match,
                         float* __restrict__ partial) {
    __shared__ float4 q[kMcRows];
    __shared__ float red[kMcThreads / 32];
    const int bi = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int l0 = tile * kMcRows, rows = min(kMcRows, m - l0);
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + (size_t)bi * m * 3;
    const float* mt = match + (size_t)bi * n * m + (size_t)l0 * n;
    if (tid < kMcRows) {
        int l = l0 + min(tid, rows - 1);
        q[tid] = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
    }
    __syncthreads();
    float acc = 0.f;
    for (int k = tid; k < n; k += kMcThreads) {
        float4 me = make_float4(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2], 0.f);
        float w[kMcRows];
#pragma unroll
        for (int r = 0; r < kMcRows; r++) w[r] = r < rows ? __ldcs(&mt[(size_t)r * n + k]) : 0.f;
#pragma unroll
        for (int r = 0; r < kMcRows; r++) acc = fmaf(sqrtf(am_d2(me, q[r])), w[r], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < kMcThreads / 32; w++) s += red[w];
        partial[(size_t)bi * gridDim.x + tile] = s;
    }
}

__global__ void matchcost_final_kernel(int tiles, const float* __restrict__ partial,
                                       float* __restrict__ out) {
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (a no-op unless launched as a programmatic dependent)
    const int bi = blockIdx.x;
    float acc = 0.f;
    for (int t = threadIdx.x; t < tiles; t += 32) acc += __ldcg(&partial[(size_t)bi * tiles + t]);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[bi] = acc;
}

// ---- matchcostgrad (tf_approxmatch_g.cu:229-291) ----
// grad1[k] = sum_l match[l,k] (p1_k - p2_l) rsqrt(max(d2,1e-20)): thread owns k, a CTA
//   is 64 k-columns x 4 row groups, rows strided over the groups, smem-reduced (fixed order).
// grad2[l] = sum_k match[l,k] (p2_l - p1_k) rsqrt(...): one warp per row, float4 loads.
constexpr int kG1Cols = 64, kG1Groups = 4;

__global__ void __launch_bounds__(kG1Cols * kG1Groups)
matchcostgrad1_kernel(int n, int m, const float* __restrict__ xyz1,
                      const float* __restrict__ xyz2, const float* __restrict__ match,
                      float* __restrict__ grad1) {
    extern __shared__ __align__(16) float sm[];
    float4* q = reinterpret_cast<float4*>(sm);                 // [chunk]
    const int bi = blockIdx.y, tid = threadIdx.x;
    const int col = tid % kG1Cols, grp = tid / kG1Cols;
    const int k = blockIdx.x * kG1Cols + col;
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + (size_t)bi * m * 3;
    const float* mt = match + (size_t)bi * n * m;
    const bool act = k < n;
    float x = act ? p1[k * 3] : 0.f, y = act ? p1[k * 3 + 1] : 0.f, z = act ? p1[k * 3 + 2] : 0.f;
    float ax = 0.f, ay = 0.f, az = 0.f;
    constexpr int kChunk = 1024;
    for (int c0 = 0; c0 < m; c0 += kChunk) {
        int cnt = min(kChunk, m - c0);
        __syncthreads();
        for (int t = tid; t < cnt; t += kG1Cols * kG1Groups)
            q[t] = make_float4(p2[(c0 + t) * 3], p2[(c0 + t) * 3 + 1], p2[(c0 + t) * 3 + 2], 0.f);
        __syncthreads();
        if (act) {
#pragma unroll 4
            for (int l = grp; l < cnt; l += kG1Groups) {
                float4 o = q[l];
                float dx = x - o.x, dy = y - o.y, dz = z - o.z;
                float w = __ldcs(&mt[(size_t)(c0 + l) * n + k]) *
                          rsqrtf(fmaxf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)), 1e-20f));
                ax = fmaf(dx, w, ax);
                ay = fmaf(dy, w, ay);
                az = fmaf(dz, w, az);
            }
        }
    }
    __syncthreads();
    float* red = sm;   // reuse: [groups][cols][3]
    red[(grp * kG1Cols + col) * 3 + 0] = ax;
    red[(grp * kG1Cols + col) * 3 + 1] = ay;
    red[(grp * kG1Cols + col) * 3 + 2] = az;
    __syncthreads();
    if (grp == 0 && act) {
        for (int c = 0; c < 3; c++) {
            float s = 0.f;
            for (int g = 0; g < kG1Groups; g++) s += red[(g * kG1Cols + col) * 3 + c];
            grad1[((size_t)bi * n + k) * 3 + c] = s;
        }
    }
}

constexpr int kG2Warps = 8;

__global__ void __launch_bounds__(kG2Warps * 32)
matchcostgrad2_kernel(int n, int m, const float* __restrict__ xyz1,
                      const float* __restrict__ xyz2, const float* __restrict__ match,
                      float* __restrict__ grad2) {
    const int bi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.x * kG2Warps + warp;
    if (l >= m) return;
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + ((size_t)bi * m + l) * 3;
    const float* row = match + (size_t)bi * n * m + (size_t)l * n;
    const float x = p2[0], y = p2[1], z = p2[2];
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll 4
    for (int k = lane; k < n; k += 32) {
        float dx = x - p1[k * 3], dy = y - p1[k * 3 + 1], dz = z - p1[k * 3 + 2];
        float w = __ldcs(&row[k]) * rsqrtf(fmaxf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)), 1e-20f));
        ax = fmaf(dx, w, ax);
        ay = fmaf(dy, w, ay);
        az = fmaf(dz, w, az);
    }
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    if (lane == 0) {
        float* g = grad2 + ((size_t)bi * m + l) * 3;
        g[0] = ax;
        g[1] = ay;
        g[2] = az;
    }
}


// =====================================================================================
// Streaming versions of matchcost / matchcostgrad for n % 4 == 0 (the HBM-bound pair: `match` is b*m*n*4 bytes,
// 134 MB at 32 x 1024^2, everything else is KBs).  `match` is read exactly ONCE with 16-byte streaming loads,
// eight rows (eight independent LDG.128) in flight per thread:
//   * a CTA owns kMsRows... rows x one slab of 1024 columns; thread t owns columns 4t..4t+3 of the slab, so the
//     dataset points and the column accumulators (grad1) live in registers for the whole tile;
//   * matchcostgrad is FUSED (the reference, tf_approxmatch_g.cu:229-295, and the first version here read `match`
//     twice): w = match * rsqrt(max(d2, 1e-20)) is evaluated once per element and feeds both
//     grad1[k] += (p1_k - p2_l) w  (column sums: registers) and grad2[l] -= (p1_k - p2_l) w (row sums);
//   * row sums of 8 rows x 3 coordinates are reduced across the warp with a halving butterfly (27 SHFL for 24
//     values instead of 120), across warps through shared memory in a fixed order;
//   * per-tile column partials / per-slab row partials go to a stream-ordered scratch and are summed by a small
//     second kernel in a FIXED order: deterministic, no atomics (the reference's grad kernels are deterministic
//     too; only its nn_distance gradient uses atomics).
// FP32 work is issued as packed f32x2 (two columns per instruction): ~9 instructions per element, which keeps
// the kernel under the HBM time (20.5 us at 6555 GB/s for 134 MB).
// =====================================================================================
constexpr int kMsThreads = 256;
constexpr int kMsSlab = kMsThreads * 4;      // columns per CTA
constexpr int kMsBatch = 8;                  // rows in flight per thread

__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One row of this thread's 4 columns (zero beyond the tile).  The kernels keep a RING of kMsBatch such loads per thread:
// slot j is consumed and immediately refilled with the row kMsBatch further down, so eight LDG.128 per thread are in
// flight at every moment of the CTA's life -- including its prologue, because the ring is primed before the query
// points are staged.  (The first version loaded eight rows, waited, computed, loaded the next eight: 3.8 TB/s.)
// The load itself is UNCONDITIONAL (row clamped into the tile, inactive threads point at column 0): ptxas sinks
// predicated loads down to their first use, which would serialise the ring; out-of-range values are zeroed at use.
__device__ __forceinline__ float4 ms_ld(const float* mt, int r, int rows, int n) {
    return ldcs4(mt + (size_t)min(r, rows - 1) * n);
}
__device__ __forceinline__ float4 ms_use(float4 v, bool valid) {
    return valid ? v : make_float4(0.f, 0.f, 0.f, 0.f);
}

// the four dataset points of this thread as packed pairs: X01 = {x0,x1}, X23 = {x2,x3}, ...
struct MsCols { uint64_t x01, x23, y01, y23, z01, z23; };
__device__ __forceinline__ MsCols ms_load_cols(const float* p1, int k, bool act) {
    MsCols c;
    if (act) {      // 12 consecutive floats, 16-byte aligned (k % 4 == 0)
        const float4 a = __ldg(reinterpret_cast<const float4*>(p1 + (size_t)k * 3));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p1 + (size_t)k * 3 + 4));
        const float4 d = __ldg(reinterpret_cast<const float4*>(p1 + (size_t)k * 3 + 8));
        c.x01 = pk(a.x, a.w); c.y01 = pk(a.y, b.x); c.z01 = pk(a.z, b.y);
        c.x23 = pk(b.z, d.y); c.y23 = pk(b.w, d.z); c.z23 = pk(d.x, d.w);
    } else {
        c.x01 = c.x23 = c.y01 = c.y23 = c.z01 = c.z23 = pk(0.f, 0.f);
    }
    return c;
}

#ifndef MPB_MS_COST_MINB
#define MPB_MS_COST_MINB 4          // 64 registers, no spills: a third more loads in flight per SM
#endif
__global__ void __launch_bounds__(kMsThreads, MPB_MS_COST_MINB)
matchcost_stream_kernel(int n, int m, int rows_per_tile, const float* __restrict__ xyz1,
                        const float* __restrict__ xyz2, const float* __restrict__ match,
                        float* __restrict__ partial, int* __restrict__ counters, float* __restrict__ out) {
    extern __shared__ __align__(16) float4 ms_q[];           // [rows_per_tile] query points of the tile
    __shared__ float red[kMsThreads / 32];
    const int tile = blockIdx.x, slab = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x;
    const int l0 = tile * rows_per_tile, rows = min(rows_per_tile, m - l0);
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + (size_t)bi * m * 3;
    const int k = slab * kMsSlab + tid * 4;
    const bool act = k < n;
    const float* mt = match + ((size_t)bi * m + l0) * n + (act ? k : 0);
    float4 w[kMsBatch];
#pragma unroll
    for (int j = 0; j < kMsBatch; j++) w[j] = ms_ld(mt, j, rows, n);
    for (int t = tid; t < rows_per_tile; t += kMsThreads) {
        const int l = l0 + min(t, rows - 1);
        ms_q[t] = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
    }
    const MsCols c = ms_load_cols(p1, k, act);
    __syncthreads();
    uint64_t acc = pk(0.f, 0.f);
    for (int r0 = 0; r0 < rows; r0 += kMsBatch) {
#pragma unroll
        for (int j = 0; j < kMsBatch; j++) {
            const float4 cur = ms_use(w[j], act && r0 + j < rows);
            w[j] = ms_ld(mt, r0 + kMsBatch + j, rows, n);
            asm volatile("" ::: "memory");        // keep the refill HERE: eight rows ahead of its use
            const float4 q = ms_q[min(r0 + j, rows_per_tile - 1)];
            const uint64_t qx = pk(q.x, q.x), qy = pk(q.y, q.y), qz = pk(q.z, q.z);
            uint64_t dx = sub2(c.x01, qx), dy = sub2(c.y01, qy), dz = sub2(c.z01, qz);
            float a0, a1, a2, a3;
            upk(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), a0, a1);
            dx = sub2(c.x23, qx); dy = sub2(c.y23, qy); dz = sub2(c.z23, qz);
            upk(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), a2, a3);
            acc = fma2(pk(sqrt_approx(a0), sqrt_approx(a1)), pk(cur.x, cur.y), acc);
            acc = fma2(pk(sqrt_approx(a2), sqrt_approx(a3)), pk(cur.z, cur.w), acc);
        }
    }
    float lo, hi;
    upk(acc, lo, hi);
    float v = lo + hi;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    // the LAST CTA of this batch element to arrive sums all partials in a fixed order (deterministic, one launch)
    const int per_elem = gridDim.x * gridDim.y;
    __shared__ int s_last;
    if (tid == 0) {
        float s_ = 0.f;
        for (int w_ = 0; w_ < kMsThreads / 32; w_++) s_ += red[w_];
        partial[(size_t)bi * per_elem + slab * gridDim.x + tile] = s_;
        __threadfence();
        s_last = atomicAdd(&counters[bi], 1) == per_elem - 1;
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        float a = 0.f;
        for (int t = tid; t < per_elem; t += 32) a += __ldcg(&partial[(size_t)bi * per_elem + t]);
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (tid == 0) out[bi] = a;
    }
}

// fixed-order sum of T partial planes of n4 float4 each (the tail of the last CTA of a batch element): four planes
// and two outputs per thread in flight, so the tail is a few L2 round trips instead of T of them
__device__ __forceinline__ void ms_sum_planes(const float4* __restrict__ q, float4* __restrict__ o, int n4, int T, int tid,
                                              int nthr = kMsThreads) {
    for (int i0 = tid; i0 < n4; i0 += 2 * nthr) {
        const int i1 = i0 + nthr;
        const bool two = i1 < n4;
        const int j1 = two ? i1 : i0;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
        int t = 0;
        for (; t + 4 <= T; t += 4) {
            float4 v[4], u[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[j] = __ldcg(q + (size_t)(t + j) * n4 + i0);
                u[j] = __ldcg(q + (size_t)(t + j) * n4 + j1);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w;
                c.x += u[j].x; c.y += u[j].y; c.z += u[j].z; c.w += u[j].w;
            }
        }
        for (; t < T; t++) {
            const float4 v = __ldcg(q + (size_t)t * n4 + i0), u = __ldcg(q + (size_t)t * n4 + j1);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            c.x += u.x; c.y += u.y; c.z += u.z; c.w += u.w;
        }
        o[i0] = a;
        if (two) o[i1] = c;
    }
}

// keep-or-send halving step of the butterfly: after it, u[i] (i < H) holds the sum over the lane pair
// {lane, lane ^ BIT} of value (lane & BIT ? H + i : i)
template <int H, int BIT>
__device__ __forceinline__ void ms_halve(const float* v, float* u, int lane) {
    const bool up = (lane & BIT) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
        const float keep = up ? v[H + i] : v[i];
        const float send = up ? v[i] : v[H + i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
    }
}

__global__ void __launch_bounds__(kMsThreads, 3)
matchcostgrad_stream_kernel(int n, int m, int rows_per_tile, const float* __restrict__ xyz1,
                            const float* __restrict__ xyz2, const float* __restrict__ match,
                            float* __restrict__ part1,     // [b][tiles][n][3]   column partials (grad1)
                            float* __restrict__ part2,     // [b][slabs][m][3]   row partials (grad2, sign included)
                            int* __restrict__ counters, float* __restrict__ grad1, float* __restrict__ grad2) {
    extern __shared__ __align__(16) float4 ms_q[];           // [rows_per_tile] query points, then [warps][rows][3] row sums
    const int tile = blockIdx.x, slab = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int l0 = tile * rows_per_tile, rows = min(rows_per_tile, m - l0);
    float* rsum = reinterpret_cast<float*>(ms_q + rows_per_tile);      // [8 warps][rows_per_tile][3]
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + (size_t)bi * m * 3;
    const int k = slab * kMsSlab + tid * 4;
    const bool act = k < n;
    const float* mt = match + ((size_t)bi * m + l0) * n + (act ? k : 0);
    float4 w[kMsBatch];
#pragma unroll
    for (int j = 0; j < kMsBatch; j++) w[j] = ms_ld(mt, j, rows, n);
    for (int t = tid; t < rows_per_tile; t += kMsThreads) {
        const int l = l0 + min(t, rows - 1);
        ms_q[t] = make_float4(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], 0.f);
    }
    const MsCols c = ms_load_cols(p1, k, act);
    __syncthreads();
    uint64_t gx01 = pk(0.f, 0.f), gx23 = gx01, gy01 = gx01, gy23 = gx01, gz01 = gx01, gz23 = gx01;
    for (int r0 = 0; r0 < rows; r0 += kMsBatch) {
        float v[kMsBatch * 3];
#pragma unroll
        for (int r = 0; r < kMsBatch; r++) {
            const float4 cur = ms_use(w[r], act && r0 + r < rows);
            w[r] = ms_ld(mt, r0 + kMsBatch + r, rows, n);
            asm volatile("" ::: "memory");        // keep the refill HERE: eight rows ahead of its use
            const float4 q = ms_q[min(r0 + r, rows_per_tile - 1)];
            const uint64_t qx = pk(q.x, q.x), qy = pk(q.y, q.y), qz = pk(q.z, q.z);
            const uint64_t dx0 = sub2(c.x01, qx), dy0 = sub2(c.y01, qy), dz0 = sub2(c.z01, qz);
            const uint64_t dx1 = sub2(c.x23, qx), dy1 = sub2(c.y23, qy), dz1 = sub2(c.z23, qz);
            float a0, a1, a2, a3;
            upk(fma2(dz0, dz0, fma2(dx0, dx0, mul2(dy0, dy0))), a0, a1);
            upk(fma2(dz1, dz1, fma2(dx1, dx1, mul2(dy1, dy1))), a2, a3);
            const uint64_t w01 = mul2(pk(cur.x, cur.y), pk(rsqrt_approx(fmaxf(a0, 1e-20f)), rsqrt_approx(fmaxf(a1, 1e-20f))));
            const uint64_t w23 = mul2(pk(cur.z, cur.w), pk(rsqrt_approx(fmaxf(a2, 1e-20f)), rsqrt_approx(fmaxf(a3, 1e-20f))));
            gx01 = fma2(dx0, w01, gx01); gy01 = fma2(dy0, w01, gy01); gz01 = fma2(dz0, w01, gz01);
            gx23 = fma2(dx1, w23, gx23); gy23 = fma2(dy1, w23, gy23); gz23 = fma2(dz1, w23, gz23);
            float s0, s1;
            upk(fma2(dx1, w23, mul2(dx0, w01)), s0, s1); v[r * 3 + 0] = s0 + s1;
            upk(fma2(dy1, w23, mul2(dy0, w01)), s0, s1); v[r * 3 + 1] = s0 + s1;
            upk(fma2(dz1, w23, mul2(dz0, w01)), s0, s1); v[r * 3 + 2] = s0 + s1;
        }
        // 24 row sums across the 32 lanes: 12 + 6 + 3 keep-or-send steps, then 2 plain steps on 3 values
        float u12[12], u6[6], u3[3];
        ms_halve<12, 16>(v, u12, lane);
        ms_halve<6, 8>(u12, u6, lane);
        ms_halve<3, 4>(u6, u3, lane);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 2);
            u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 1);
        }
        // lanes 4j..4j+3 now hold row j: index = (lane & 16 ? 12 : 0) + (lane & 8 ? 6 : 0) + (lane & 4 ? 3 : 0) = 3 * (lane >> 2)
        if ((lane & 3) == 0) {
            const int r = r0 + (lane >> 2);
            if (r < rows_per_tile) {
                float* d = rsum + ((size_t)warp * rows_per_tile + r) * 3;
                d[0] = u3[0]; d[1] = u3[1]; d[2] = u3[2];
            }
        }
    }
    if (act) {          // column partials of this tile: 12 consecutive floats per thread
        float x0, x1, x2, x3, y0, y1, y2, y3, z0, z1, z2, z3;
        upk(gx01, x0, x1); upk(gx23, x2, x3); upk(gy01, y0, y1); upk(gy23, y2, y3); upk(gz01, z0, z1); upk(gz23, z2, z3);
        float4* d = reinterpret_cast<float4*>(part1 + (((size_t)bi * gridDim.x + tile) * n + k) * 3);
        d[0] = make_float4(x0, y0, z0, x1);
        d[1] = make_float4(y1, z1, x2, y2);
        d[2] = make_float4(z2, x3, y3, z3);
    }
    __syncthreads();
    for (int t = tid; t < rows * 3; t += kMsThreads) {       // fixed order over the 8 warps
        float s_ = 0.f;
#pragma unroll
        for (int w_ = 0; w_ < kMsThreads / 32; w_++) s_ += rsum[(size_t)w_ * rows_per_tile * 3 + t];
        part2[(((size_t)bi * gridDim.y + slab) * m + l0) * 3 + t] = -s_;          // grad2 uses (p2 - p1)
    }
    // the LAST CTA of this batch element to arrive sums the per-tile column partials (and, with several slabs, the
    // per-slab row partials) in a fixed order: deterministic, no second launch
    const int per_elem = gridDim.x * gridDim.y;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&counters[bi], 1) == per_elem - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int T = gridDim.x, S = gridDim.y;
    ms_sum_planes(reinterpret_cast<const float4*>(part1 + (size_t)bi * T * n * 3),
                  reinterpret_cast<float4*>(grad1 + (size_t)bi * n * 3), n * 3 / 4, T, tid);      // n % 4 == 0
    if (S > 1) {
        const float* q2 = part2 + (size_t)bi * S * m * 3;
        for (int i = tid; i < m * 3; i += kMsThreads) {
            float a = 0.f;
            for (int t = 0; t < S; t++) a += __ldcg(q2 + (size_t)t * m * 3 + i);
            grad2[(size_t)bi * m * 3 + i] = a;
        }
    }
}

// ---------------------------------------------------------------- match_cost_grad with the bulk-copy engine feeding it
// The register ring above cannot be both deep and cheap: eight LDG.128 per thread are 32 registers that ptxas refills
// only at the end of a batch, so the loads of a warp arrive in bursts and every row pays address, clamp and select
// instructions.  Here `match` is moved by cp.async.bulk (1-D TMA, no tensor map: a tile row is `slabw` contiguous floats)
// into a ring of kMtStages shared-memory stages of kMtRows rows each, completion on mbarriers.  Loads in flight no longer
// occupy registers, the warps take turns issuing four copies per stage and re-arm a stage two blocks after it was
// consumed (so the issuer never waits for a slow warp), and the inner loop is arithmetic only:
//  * every thread owns CG groups of four columns of a slab of `slabw` = 4 * CG * blockDim.x columns (all threads active:
//    slabw divides n); with CG = 2 the per-row overheads -- query point, row-sum butterfly, barrier checks -- are shared
//    by eight elements instead of four;
//  * tiles hold whole 8-row batches (m % 8 == 0) and are walked in sub-tiles of <= kMtSubRows rows: the per-warp row sums
//    of a sub-tile are combined (fixed order) and written out at its end, so a tile can be hundreds of rows tall and
//    the number of column-partial planes the last CTA has to add stays small;
//  * query points are staged pre-duplicated ({x,x,y,y} {z,z,-,-}: the packed operands come straight out of LDS.128 +
//    LDS.64), the next sub-tile's points wait in registers;
//  * the zero-distance guard is folded into the distance as an additive 1e-37 (for d = 0 every difference is 0 as well
//    and the product is 0, as with the reference's max(d, 1e-20)).
#ifndef MPB_MT_ROWS
#define MPB_MT_ROWS 4
#endif
#ifndef MPB_MT_STAGES
#define MPB_MT_STAGES 3
#endif
#ifndef MPB_MT_LAG
#define MPB_MT_LAG 2
#endif
constexpr int kMtRows = MPB_MT_ROWS;        // rows per stage
constexpr int kMtStages = MPB_MT_STAGES;
constexpr int kMtSubRows = 80;              // rows per sub-tile (shared memory of the per-warp row sums)
constexpr int kMtLag = MPB_MT_LAG;          // a stage is re-armed kMtLag blocks after its block was consumed:
                                            // kMtStages - kMtLag blocks are in flight while one is being consumed
static_assert(kMtLag >= 1 && kMtLag < kMtStages && kMsBatch % kMtRows == 0, "stage ring geometry");

__device__ __forceinline__ uint32_t ms_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ms_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;      // bounded spin: a protocol bug must trap, not hang the GPU
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void ms_issue_block(uint32_t full_bar, uint32_t dst, const float* src, size_t row_stride,
                                               uint32_t row_bytes, uint64_t policy) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar), "r"(row_bytes * kMtRows) : "memory");
    if (row_stride * 4 == row_bytes) {       // the slab is the whole row: the block is one contiguous piece
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
            ::"r"(dst), "l"(src), "r"(row_bytes * kMtRows), "r"(full_bar), "l"(policy)
            : "memory");
        return;
    }
#pragma unroll
    for (int r = 0; r < kMtRows; r++)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
            ::"r"(dst + r * row_bytes), "l"(src + (size_t)r * row_stride), "r"(row_bytes), "r"(full_bar), "l"(policy)
            : "memory");
}

template <int CG> struct MtCfg;
template <> struct MtCfg<1> { static constexpr int kThreads = 256, kMinBlocks = 3; };
template <> struct MtCfg<2> { static constexpr int kThreads = 128, kMinBlocks = 4; };

template <int CG>
__global__ void __launch_bounds__(MtCfg<CG>::kThreads, MtCfg<CG>::kMinBlocks)
matchcostgrad_tma_kernel(int n, int m, int rows_per_tile, int sub_rows, const float* __restrict__ xyz1,
                         const float* __restrict__ xyz2, const float* __restrict__ match,
                         float* __restrict__ part1,     // [b][tiles][n][3]   column partials (grad1)
                         float* __restrict__ part2) {   // [b][slabs][m][3]   row partials (grad2, sign included)
    extern __shared__ __align__(128) unsigned char mt_smem[];
    const int tile = blockIdx.x, slab = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
    const int slabw = nthr * 4 * CG;
    const uint32_t row_bytes = (uint32_t)slabw * 4u, stage_bytes = row_bytes * kMtRows;
    const int l0 = tile * rows_per_tile, rows = min(rows_per_tile, m - l0);        // a multiple of 8, >= 8
    float4* qd = reinterpret_cast<float4*>(mt_smem + (size_t)kMtStages * stage_bytes);       // [sub_rows][2]
    float* rsum = reinterpret_cast<float*>(qd + 2 * sub_rows);                              // [warps][sub_rows][3]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rsum + (size_t)nwarps * sub_rows * 3);      // (sub_rows % 8 == 0: aligned)
    const uint32_t full0 = ms_smem_u32(bars), empty0 = full0 + 8 * kMtStages, stage0 = ms_smem_u32(mt_smem);
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + ((size_t)bi * m + l0) * 3;                               // the tile's query points
    const size_t ns = (size_t)n;
    const float* mt = match + ((size_t)bi * m + l0) * ns + (size_t)slab * slabw;     // the tile's row 0, this slab
    const int nblk = rows / kMtRows;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (tid == 0) {
        for (int s_ = 0; s_ < kMtStages; s_++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s_), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s_), "r"(nwarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int j = 0; j < kMtStages && j < nblk; j++)
            ms_issue_block(full0 + 8 * j, stage0 + j * stage_bytes, mt + (size_t)j * kMtRows * ns, ns, row_bytes, policy);
    }
    // query points: this sub-tile's into shared memory, the next one's into registers (sub_rows <= blockDim.x)
    float nq0 = 0.f, nq1 = 0.f, nq2 = 0.f;
    if (tid < min(sub_rows, rows)) {
        const float x = p2[tid * 3], y = p2[tid * 3 + 1], z = p2[tid * 3 + 2];
        qd[2 * tid] = make_float4(x, x, y, y);
        qd[2 * tid + 1] = make_float4(z, z, 0.f, 0.f);
    }
    if (sub_rows + tid < rows && tid < sub_rows) {
        nq0 = p2[(sub_rows + tid) * 3]; nq1 = p2[(sub_rows + tid) * 3 + 1]; nq2 = p2[(sub_rows + tid) * 3 + 2];
    }
    MsCols c[CG];
#pragma unroll
    for (int g = 0; g < CG; g++) c[g] = ms_load_cols(p1, slab * slabw + (tid + g * nthr) * 4, true);
    __syncthreads();
    const uint64_t eps2 = pk(1e-37f, 1e-37f), zero2 = pk(0.f, 0.f);
    uint64_t gx01[CG], gx23[CG], gy01[CG], gy23[CG], gz01[CG], gz23[CG];
#pragma unroll
    for (int g = 0; g < CG; g++) gx01[g] = gx23[g] = gy01[g] = gy23[g] = gz01[g] = gz23[g] = zero2;
    int stage = 0, j = 0, turn = 0;
    uint32_t phase = 0;
    for (int sub0 = 0; sub0 < rows; sub0 += sub_rows) {
        const int srows = min(sub_rows, rows - sub0);
        for (int rb = 0; rb < srows; rb += kMsBatch) {
            float v[kMsBatch * 3];
#pragma unroll
            for (int h = 0; h < kMsBatch / kMtRows; h++, j++) {
                {
                    // block j - lag was consumed by every warp a while ago: re-arm its stage with block j - lag + stages
                    // (operands computed in uniform control flow; the warps take turns, one lane issues)
                    const int jo = j - kMtLag, so = (stage + kMtStages - kMtLag) % kMtStages;
                    const uint32_t ebar = empty0 + 8 * so, fbar = full0 + 8 * so, dst = stage0 + so * stage_bytes;
                    const uint32_t eph = (so > stage ? phase ^ 1u : phase);        // phase of block jo's use of its stage
                    const float* src = mt + (size_t)(jo + kMtStages) * kMtRows * ns;
                    if (jo >= 0 && jo + kMtStages < nblk && warp == turn && lane == 0) {
                        ms_mbar_wait(ebar, eph);
                        ms_issue_block(fbar, dst, src, ns, row_bytes, policy);
                    }
                    if (++turn == nwarps) turn = 0;
                }
                ms_mbar_wait(full0 + 8 * stage, phase);
                const float4* st = reinterpret_cast<const float4*>(mt_smem + (size_t)stage * stage_bytes) + tid;
                const float4* q = qd + 2 * (rb + h * kMtRows);
#pragma unroll
                for (int r = 0; r < kMtRows; r++) {
                    const float4 qa = q[2 * r];
                    const float2 qb = *reinterpret_cast<const float2*>(q + 2 * r + 1);
                    const uint64_t qx = pk(qa.x, qa.y), qy = pk(qa.z, qa.w), qz = pk(qb.x, qb.y);
                    uint64_t sx = zero2, sy = zero2, sz = zero2;
#pragma unroll
                    for (int g = 0; g < CG; g++) {
                        const float4 cur = st[(size_t)r * nthr * CG + g * nthr];
                        const uint64_t dx0 = sub2(c[g].x01, qx), dy0 = sub2(c[g].y01, qy), dz0 = sub2(c[g].z01, qz);
                        const uint64_t dx1 = sub2(c[g].x23, qx), dy1 = sub2(c[g].y23, qy), dz1 = sub2(c[g].z23, qz);
                        float a0, a1, a2, a3;
                        upk(fma2(dz0, dz0, fma2(dx0, dx0, fma2(dy0, dy0, eps2))), a0, a1);
                        upk(fma2(dz1, dz1, fma2(dx1, dx1, fma2(dy1, dy1, eps2))), a2, a3);
                        const uint64_t w01 = mul2(pk(cur.x, cur.y), pk(rsqrt_approx(a0), rsqrt_approx(a1)));
                        const uint64_t w23 = mul2(pk(cur.z, cur.w), pk(rsqrt_approx(a2), rsqrt_approx(a3)));
                        gx01[g] = fma2(dx0, w01, gx01[g]); gy01[g] = fma2(dy0, w01, gy01[g]); gz01[g] = fma2(dz0, w01, gz01[g]);
                        gx23[g] = fma2(dx1, w23, gx23[g]); gy23[g] = fma2(dy1, w23, gy23[g]); gz23[g] = fma2(dz1, w23, gz23[g]);
                        if (g == 0) {
                            sx = fma2(dx1, w23, mul2(dx0, w01)); sy = fma2(dy1, w23, mul2(dy0, w01)); sz = fma2(dz1, w23, mul2(dz0, w01));
                        } else {
                            sx = fma2(dx1, w23, fma2(dx0, w01, sx)); sy = fma2(dy1, w23, fma2(dy0, w01, sy));
                            sz = fma2(dz1, w23, fma2(dz0, w01, sz));
                        }
                    }
                    float s0, s1;
                    const int vi = (h * kMtRows + r) * 3;
                    upk(sx, s0, s1); v[vi + 0] = s0 + s1;
                    upk(sy, s0, s1); v[vi + 1] = s0 + s1;
                    upk(sz, s0, s1); v[vi + 2] = s0 + s1;
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * stage) : "memory");
                if (++stage == kMtStages) { stage = 0; phase ^= 1u; }
            }
            // 24 row sums across the 32 lanes: 12 + 6 + 3 keep-or-send steps, then 2 plain steps on 3 values
            float u12[12], u6[6], u3[3];
            ms_halve<12, 16>(v, u12, lane);
            ms_halve<6, 8>(u12, u6, lane);
            ms_halve<3, 4>(u6, u3, lane);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 2);
                u3[i] += __shfl_xor_sync(0xffffffffu, u3[i], 1);
            }
            if ((lane & 3) == 0) {       // lanes 4j..4j+3 hold row j of the batch
                float* d = rsum + ((size_t)warp * sub_rows + rb + (lane >> 2)) * 3;
                d[0] = u3[0]; d[1] = u3[1]; d[2] = u3[2];
            }
        }
        // end of the sub-tile: row sums out (fixed order over the warps), next query points in
        __syncthreads();
        for (int t = tid; t < srows * 3; t += nthr) {
            float s_ = 0.f;
            for (int w_ = 0; w_ < nwarps; w_++) s_ += rsum[(size_t)w_ * sub_rows * 3 + t];
            part2[(((size_t)bi * gridDim.y + slab) * m + l0 + sub0) * 3 + t] = -s_;          // grad2 uses (p2 - p1)
        }
        if (sub0 + sub_rows < rows) {
            if (tid < sub_rows) {
                qd[2 * tid] = make_float4(nq0, nq0, nq1, nq1);
                qd[2 * tid + 1] = make_float4(nq2, nq2, 0.f, 0.f);
                const int t2 = sub0 + 2 * sub_rows + tid;
                if (t2 < rows) { nq0 = p2[t2 * 3]; nq1 = p2[t2 * 3 + 1]; nq2 = p2[t2 * 3 + 2]; }
            }
            __syncthreads();
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the plane-summing launch may take its seats
#pragma unroll
    for (int g = 0; g < CG; g++) {       // column partials of this tile: 12 consecutive floats per column group
        float x0, x1, x2, x3, y0, y1, y2, y3, z0, z1, z2, z3;
        upk(gx01[g], x0, x1); upk(gx23[g], x2, x3); upk(gy01[g], y0, y1); upk(gy23[g], y2, y3);
        upk(gz01[g], z0, z1); upk(gz23[g], z2, z3);
        const int k = slab * slabw + (tid + g * nthr) * 4;
        float4* d = reinterpret_cast<float4*>(part1 + (((size_t)bi * gridDim.x + tile) * n + k) * 3);
        d[0] = make_float4(x0, y0, z0, x1);
        d[1] = make_float4(y1, z1, x2, y2);
        d[2] = make_float4(z2, x3, y3, z3);
    }
}

// match_cost on the same feed: no row or column sums to keep, one partial per CTA, summed per batch element by
// matchcost_final_kernel as a programmatic dependent launch (no counters to clear, no last-CTA tail)
template <int CG>
__global__ void __launch_bounds__(MtCfg<CG>::kThreads, MtCfg<CG>::kMinBlocks)
matchcost_tma_kernel(int n, int m, int rows_per_tile, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                     const float* __restrict__ match, float* __restrict__ partial) {      // [b][slabs * tiles]
    extern __shared__ __align__(128) unsigned char mt_smem[];
    const int tile = blockIdx.x, slab = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
    const int slabw = nthr * 4 * CG;
    const uint32_t row_bytes = (uint32_t)slabw * 4u, stage_bytes = row_bytes * kMtRows;
    const int l0 = tile * rows_per_tile, rows = min(rows_per_tile, m - l0);        // a multiple of 8, >= 8
    float4* qd = reinterpret_cast<float4*>(mt_smem + (size_t)kMtStages * stage_bytes);       // [rows_per_tile][2]
    float* red = reinterpret_cast<float*>(qd + 2 * rows_per_tile);                          // [8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8);
    const uint32_t full0 = ms_smem_u32(bars), empty0 = full0 + 8 * kMtStages, stage0 = ms_smem_u32(mt_smem);
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + ((size_t)bi * m + l0) * 3;
    const size_t ns = (size_t)n;
    const float* mt = match + ((size_t)bi * m + l0) * ns + (size_t)slab * slabw;
    const int nblk = rows / kMtRows;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (tid == 0) {
        for (int s_ = 0; s_ < kMtStages; s_++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s_), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s_), "r"(nwarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int j = 0; j < kMtStages && j < nblk; j++)
            ms_issue_block(full0 + 8 * j, stage0 + j * stage_bytes, mt + (size_t)j * kMtRows * ns, ns, row_bytes, policy);
    }
    for (int t = tid; t < rows; t += nthr) {
        const float x = p2[t * 3], y = p2[t * 3 + 1], z = p2[t * 3 + 2];
        qd[2 * t] = make_float4(x, x, y, y);
        qd[2 * t + 1] = make_float4(z, z, 0.f, 0.f);
    }
    MsCols c[CG];
#pragma unroll
    for (int g = 0; g < CG; g++) c[g] = ms_load_cols(p1, slab * slabw + (tid + g * nthr) * 4, true);
    __syncthreads();
    uint64_t acc = pk(0.f, 0.f);
    int stage = 0, turn = 0;
    uint32_t phase = 0;
    for (int j = 0; j < nblk; j++) {
        {
            const int jo = j - kMtLag, so = (stage + kMtStages - kMtLag) % kMtStages;
            const uint32_t ebar = empty0 + 8 * so, fbar = full0 + 8 * so, dst = stage0 + so * stage_bytes;
            const uint32_t eph = (so > stage ? phase ^ 1u : phase);
            const float* src = mt + (size_t)(jo + kMtStages) * kMtRows * ns;
            if (jo >= 0 && jo + kMtStages < nblk && warp == turn && lane == 0) {
                ms_mbar_wait(ebar, eph);
                ms_issue_block(fbar, dst, src, ns, row_bytes, policy);
            }
            if (++turn == nwarps) turn = 0;
        }
        ms_mbar_wait(full0 + 8 * stage, phase);
        const float4* st = reinterpret_cast<const float4*>(mt_smem + (size_t)stage * stage_bytes) + tid;
        const float4* q = qd + 2 * (j * kMtRows);
#pragma unroll
        for (int r = 0; r < kMtRows; r++) {
            const float4 qa = q[2 * r];
            const float2 qb = *reinterpret_cast<const float2*>(q + 2 * r + 1);
            const uint64_t qx = pk(qa.x, qa.y), qy = pk(qa.z, qa.w), qz = pk(qb.x, qb.y);
#pragma unroll
            for (int g = 0; g < CG; g++) {
                const float4 cur = st[(size_t)r * nthr * CG + g * nthr];
                uint64_t dx = sub2(c[g].x01, qx), dy = sub2(c[g].y01, qy), dz = sub2(c[g].z01, qz);
                float a0, a1, a2, a3;
                upk(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), a0, a1);
                dx = sub2(c[g].x23, qx); dy = sub2(c[g].y23, qy); dz = sub2(c[g].z23, qz);
                upk(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), a2, a3);
                acc = fma2(pk(sqrt_approx(a0), sqrt_approx(a1)), pk(cur.x, cur.y), acc);
                acc = fma2(pk(sqrt_approx(a2), sqrt_approx(a3)), pk(cur.z, cur.w), acc);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * stage) : "memory");
        if (++stage == kMtStages) { stage = 0; phase ^= 1u; }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float lo, hi;
    upk(acc, lo, hi);
    float v = lo + hi;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
        float s_ = 0.f;
        for (int w_ = 0; w_ < nwarps; w_++) s_ += red[w_];
        partial[(size_t)bi * gridDim.x * gridDim.y + slab * gridDim.x + tile] = s_;
    }
}

// the planes are added by a second, chip-wide launch (programmatic dependent launch: it is resident and waiting when
// the last tile retires); a last-CTA-per-batch-element tail would leave 32 small CTAs chasing L2 latency alone
// (measured: 4 us of 46 at 32 x 1024^2, 16 of 181 at 32 x 2304^2)
__global__ void __launch_bounds__(256)
ms_sum_planes_kernel(const float4* __restrict__ part1, float4* __restrict__ grad1, int n4a, int T,
                     const float4* __restrict__ part2, float4* __restrict__ grad2, int n4b, int S) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int bi = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float4* q;
    float4* o;
    int n4, P;
    const int na = T > 1 ? n4a : 0;
    if (i < na) { q = part1 + (size_t)bi * T * n4a + i; o = grad1 + (size_t)bi * n4a + i; n4 = n4a; P = T; }
    else {
        i -= na;
        if (S <= 1 || i >= n4b) return;
        q = part2 + (size_t)bi * S * n4b + i; o = grad2 + (size_t)bi * n4b + i; n4 = n4b; P = S;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int t = 0;
    for (; t + 8 <= P; t += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ldcg(q + (size_t)(t + j) * n4);
#pragma unroll
        for (int j = 0; j < 8; j++) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
    }
    if (t + 4 <= P) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = __ldcg(q + (size_t)(t + j) * n4);
#pragma unroll
        for (int j = 0; j < 4; j++) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
        t += 4;
    }
    for (; t < P; t++) {
        const float4 v = __ldcg(q + (size_t)t * n4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *o = a;
}

// rows per tile: a multiple of 8, chosen so that the whole launch is ONE wave of long-lived CTAs where possible
// (3 CTAs of 256 threads per SM): T = slots / (b * slabs) tiles per batch element; capped at 128 rows (shared memory
// of the row sums), below which several waves run -- the primed ring keeps the loads flowing across CTA boundaries
inline int ms_rows_per_tile(int b, int m, int slabs, int ctas_per_sm, int cap = 128) {
    const char* e = getenv("MPB_MS_ROWS");
    if (e && atoi(e) >= 8) return min(cap, atoi(e) / 8 * 8);
    const long slots = (long)ctas_per_sm * num_sms();
    const int T = (int)max(1L, slots / ((long)b * slabs));
    int rt = (ceil_div(m, T) + 7) / 8 * 8;
    return max(8, min(cap, rt));
}

template <int CG>
cudaError_t launch_matchcostgrad_tma(int b, int n, int m, int slabw, int rt, int tiles, const float* xyz1, const float* xyz2,
                                     const float* match, float* part1, float* part2, float* grad1, float* grad2,
                                     cudaStream_t s) {
    const int nthr = slabw / (4 * CG), sub = min(kMtSubRows, nthr / 8 * 8), slabs = n / slabw;
    auto smem_of = [](int slabw_, int sub_, int warps) {
        return (size_t)kMtStages * kMtRows * slabw_ * 4 + 32 * (size_t)sub_ + 12 * (size_t)warps * sub_ + 16 * kMtStages;
    };
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(matchcostgrad_tma_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_of(kMsSlab, kMtSubRows, MtCfg<CG>::kThreads / 32));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    matchcostgrad_tma_kernel<CG><<<dim3(tiles, slabs, b), nthr, smem_of(slabw, sub, nthr / 32), s>>>(
        n, m, rt, sub, xyz1, xyz2, match, part1, part2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || (tiles == 1 && slabs == 1)) return e;
    count_launch();
    const int n4a = n * 3 / 4, n4b = m * 3 / 4, work = (tiles > 1 ? n4a : 0) + (slabs > 1 ? n4b : 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(work, 256), b);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ms_sum_planes_kernel, reinterpret_cast<const float4*>(part1),
                              reinterpret_cast<float4*>(grad1), n4a, tiles, reinterpret_cast<const float4*>(part2),
                              reinterpret_cast<float4*>(grad2), n4b, slabs);
}

}  // namespace mpb

MPB_API int mpb_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2,
                            float* match, float* temp, void* stream) {
    using namespace mpb;
    (void)temp;
    if (b < 0 || n < 0 || m < 0) return -1;
    if (b == 0 || n == 0 || m == 0) return 0;
    if (!xyz1 || !xyz2 || !match) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const int sms = num_sms();
    // largest power-of-two cluster (<= 8, portable) that still fits b clusters on the chip
    int C = 1;
    while (C < 8 && (long)b * (C * 2) <= sms) C *= 2;
    // do not slice a cloud thinner than 64 points per CTA
    while (C > 1 && (ceil_div(n, C) < 64 || ceil_div(m, C) < 64)) C /= 2;
    static int max_smem = -1;
    if (max_smem < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    for (; C >= 1; C /= 2) {
        AmSmem L;
        L.n_pad = (n + 3) & ~3;
        L.m_pad = (m + 3) & ~3;
        L.own_n = ceil_div(n, C);
        L.own_m = ceil_div(m, C);
        size_t bytes = L.total() * sizeof(float);
        if (bytes > (size_t)max_smem) {
            if (C == 1) break;
            continue;
        }
        MPB_CUDA_TRY(cudaFuncSetAttribute(approxmatch_cluster_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)b * C);
        cfg.blockDim = dim3(kAmThreads);
        cfg.dynamicSmemBytes = bytes;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        MPB_CUDA_TRY(cudaLaunchKernelEx(&cfg, approxmatch_cluster_kernel, n, m, xyz1, xyz2, match));
        count_launch();
        return 0;
    }
    // fallback: stream-ordered scratch, one CTA per batch element
    float* scratch = nullptr;
    MPB_CUDA_TRY(scratch_alloc((void**)&scratch, sizeof(float) * (size_t)b * (n + m) * 2, s));
    approxmatch_fallback_kernel<<<min(b, 4 * sms), kAmThreads, 0, s>>>(b, n, m, xyz1, xyz2, match, scratch);
    count_launch();
    cudaError_t e = cudaGetLastError();
    scratch_free(scratch, s);
    return cuda_status(e);
}

MPB_API int mpb_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2,
                          const float* match, float* out, void* stream) {
    using namespace mpb;
    if (b < 0 || n < 0 || m < 0) return -1;
    if (b == 0) return 0;
    if (!out) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) return cuda_status(cudaMemsetAsync(out, 0, sizeof(float) * b, s));
    if (!xyz1 || !xyz2 || !match) return -1;
    if (b > 65535) return -1;
    static const int cg_env = getenv("MPB_MC_CG") ? atoi(getenv("MPB_MC_CG")) : -1;      // 0: register-ring kernel
    if (cg_env != 0 && n % 128 == 0 && m % 8 == 0 && (reinterpret_cast<uintptr_t>(match) & 15u) == 0 &&
        (reinterpret_cast<uintptr_t>(xyz1) & 15u) == 0) {
        // bulk-copy-fed kernel (matchcost_tma_kernel), per-CTA partials summed by a programmatic dependent launch
        const int cg = (n % 256 == 0 && cg_env != 1) ? 2 : 1;
        int slabw = 0;
        for (int w_ = kMsSlab; w_ >= 128 * cg; w_ -= 128 * cg)
            if (n % w_ == 0) { slabw = w_; break; }
        const int slabs = n / slabw;
        const char* e_ = getenv("MPB_MS_ROWS");
        int rt = e_ && atoi(e_) >= 8 ? min(kMtSubRows, atoi(e_) / 8 * 8) : kMtSubRows;
        while (!e_ && rt > 8 && (long)b * slabs * ceil_div(m, rt) < 2L * num_sms()) rt -= 8;
        const int tiles = ceil_div(m, rt), nthr = slabw / (4 * cg);
        float* partial = nullptr;
        MPB_CUDA_TRY(scratch_alloc((void**)&partial, sizeof(float) * (size_t)b * slabs * tiles, s));
        const size_t smem = (size_t)kMtStages * kMtRows * slabw * 4 + 32 * (size_t)rt + 32 + 16 * kMtStages;
        static bool attr_set = false;
        if (!attr_set) {
            const int mx = kMtStages * kMtRows * kMsSlab * 4 + 32 * kMtSubRows + 32 + 16 * kMtStages;
            MPB_CUDA_TRY(cudaFuncSetAttribute(matchcost_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            MPB_CUDA_TRY(cudaFuncSetAttribute(matchcost_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            attr_set = true;
        }
        if (cg == 2) matchcost_tma_kernel<2><<<dim3(tiles, slabs, b), nthr, smem, s>>>(n, m, rt, xyz1, xyz2, match, partial);
        else matchcost_tma_kernel<1><<<dim3(tiles, slabs, b), nthr, smem, s>>>(n, m, rt, xyz1, xyz2, match, partial);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(b);
            cfg.blockDim = dim3(32);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, matchcost_final_kernel, slabs * tiles, (const float*)partial, out);
            count_launch();
        }
        scratch_free(partial, s);
        return cuda_status(e);
    }
    if (n % 4 == 0 && (reinterpret_cast<uintptr_t>(match) & 15u) == 0 && (reinterpret_cast<uintptr_t>(xyz1) & 15u) == 0) {
        const int slabs = ceil_div(n, kMsSlab), rt = ms_rows_per_tile(b, m, slabs, MPB_MS_COST_MINB), tiles = ceil_div(m, rt);
        float* partial = nullptr;
        MPB_CUDA_TRY(scratch_alloc((void**)&partial, sizeof(float) * ((size_t)b * slabs * tiles + b), s));
        int* counters = reinterpret_cast<int*>(partial + (size_t)b * slabs * tiles);
        MPB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * b, s));
        matchcost_stream_kernel<<<dim3(tiles, slabs, b), kMsThreads, sizeof(float4) * rt, s>>>(n, m, rt, xyz1, xyz2, match,
                                                                                              partial, counters, out);
        count_launch();
        cudaError_t e = cudaGetLastError();
        scratch_free(partial, s);
        return cuda_status(e);
    }
    const int tiles = ceil_div(m, kMcRows);
    float* partial = nullptr;
    MPB_CUDA_TRY(scratch_alloc((void**)&partial, sizeof(float) * (size_t)b * tiles, s));
    matchcost_partial_kernel<<<dim3(tiles, b), kMcThreads, 0, s>>>(n, m, xyz1, xyz2, match, partial);
    count_launch();
    matchcost_final_kernel<<<b, 32, 0, s>>>(tiles, partial, out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    scratch_free(partial, s);
    return cuda_status(e);
}

MPB_API int mpb_matchcostgrad(int b, int n, int m, const float* xyz1, const float* xyz2,
                              const float* match, float* grad1, float* grad2, void* stream) {
    using namespace mpb;
    if (b < 0 || n < 0 || m < 0) return -1;
    if (b == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        if (n && grad1) MPB_CUDA_TRY(cudaMemsetAsync(grad1, 0, sizeof(float) * (size_t)b * n * 3, s));
        if (m && grad2) MPB_CUDA_TRY(cudaMemsetAsync(grad2, 0, sizeof(float) * (size_t)b * m * 3, s));
        return 0;
    }
    if (!xyz1 || !xyz2 || !match || !grad1 || !grad2) return -1;
    if (b > 65535) return -1;
    if (n % 4 == 0 && (reinterpret_cast<uintptr_t>(match) & 15u) == 0 && (reinterpret_cast<uintptr_t>(xyz1) & 15u) == 0) {
        // fused single pass over `match`: fed by bulk copies where the geometry allows (matchcostgrad_tma_kernel: a slab
        // width that divides n, all lanes of every warp busy, whole 8-row batches), else the register-ring kernel
        static const int cg_env = getenv("MPB_MG_CG") ? atoi(getenv("MPB_MG_CG")) : -1;      // 0: ring kernel, 1 / 2: groups
        int cg = 0, slabw = 0;
        if (cg_env != 0 && m % 8 == 0) {
            if (n % 256 == 0 && cg_env != 1) cg = 2;
            else if (n % 128 == 0 && cg_env != 2) cg = 1;
            if (cg)
                for (int w_ = kMsSlab; w_ >= 128 * cg; w_ -= 128 * cg)
                    if (n % w_ == 0) { slabw = w_; break; }
        }
        const int slabs = cg ? n / slabw : ceil_div(n, kMsSlab);
        // bulk-copy kernel: tiles of one sub-tile (80 rows) measured best at both sizes -- 39.3 us at 32 x 1024^2 (13 tiles
        // per element, one wave; 64 rows: 41.3, 96: 45.4) and 137.6 us at 32 x 2304^2 (29 tiles x 3 slabs, several waves;
        // 336-row tiles in one wave: 162.1) -- shorter only when that would leave SMs without a CTA
        int rt = ms_rows_per_tile(b, m, slabs, 3);
        if (cg) {
            const char* e_ = getenv("MPB_MS_ROWS");
            rt = e_ && atoi(e_) >= 8 ? atoi(e_) / 8 * 8 : kMtSubRows;
            while (!e_ && rt > 8 && (long)b * slabs * ceil_div(m, rt) < 2L * num_sms()) rt -= 8;
        }
        const int tiles = ceil_div(m, rt);
        const bool direct1 = cg && tiles == 1;         // one tile: the column sums ARE grad1
        float *part1 = nullptr, *part2 = nullptr;
        const size_t n1 = direct1 ? 0 : (size_t)b * tiles * n * 3, n2 = slabs > 1 ? (size_t)b * slabs * m * 3 : 0;
        float* scratch = nullptr;
        MPB_CUDA_TRY(scratch_alloc((void**)&scratch, sizeof(float) * (n1 + n2 + b), s));
        part1 = direct1 ? grad1 : scratch;
        part2 = slabs > 1 ? scratch + n1 : grad2;       // one slab: the row sums ARE grad2
        cudaError_t e;
        if (cg == 2) {
            e = launch_matchcostgrad_tma<2>(b, n, m, slabw, rt, tiles, xyz1, xyz2, match, part1, part2, grad1, grad2, s);
        } else if (cg == 1) {
            e = launch_matchcostgrad_tma<1>(b, n, m, slabw, rt, tiles, xyz1, xyz2, match, part1, part2, grad1, grad2, s);
        } else {
            int* counters = reinterpret_cast<int*>(scratch + n1 + n2);
            MPB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * b, s));
            const size_t smem = sizeof(float4) * rt + sizeof(float) * (kMsThreads / 32) * rt * 3;
            matchcostgrad_stream_kernel<<<dim3(tiles, slabs, b), kMsThreads, smem, s>>>(n, m, rt, xyz1, xyz2, match, part1,
                                                                                       part2, counters, grad1, grad2);
            e = cudaGetLastError();
        }
        count_launch();
        scratch_free(scratch, s);
        return cuda_status(e);
    }
    size_t sm1 = sizeof(float) * 4 * 1024;   // float4[1024] >= [4][64][3] floats
    matchcostgrad1_kernel<<<dim3(ceil_div(n, kG1Cols), b), kG1Cols * kG1Groups, sm1, s>>>(
        n, m, xyz1, xyz2, match, grad1);
    MPB_LAUNCH_CHECK();
    matchcostgrad2_kernel<<<dim3(ceil_div(m, kG2Warps), b), kG2Warps * 32, 0, s>>>(
        n, m, xyz1, xyz2, match, grad2);
    MPB_LAUNCH_CHECK();
    return 0;
}
