// monopsr_b200/csrc/tc_gemm.cu -- see tc_gemm.cuh for the design.
#include "tc_gemm.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace mpb {

#ifdef MPB_TC_TRACE
// debug build only: per-CTA clock64 timeline of the TMA kernel ([cta][64] slots), see tools/gemm_trace.py
__device__ long long* g_tc_trace = nullptr;
#define TC_TR(slot) do { if (trc) trc[(slot)] = clock64(); } while (0)
#define TC_TR_DECL long long* trc = (g_tc_trace && cta_lin < 1024) ? g_tc_trace + cta_lin * 64 : nullptr
#define TC_TR_ARG , long long* trc
#define TC_TR_PASS , trc
#else
#define TC_TR(slot) do { } while (0)
#define TC_TR_DECL
#define TC_TR_ARG
#define TC_TR_PASS
#endif

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a protocol bug must trap, not hang the GPU box
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;   // src-size 0 => 16 zero bytes are written
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// fp16 hi/lo split of two fp32 values: hi = fp16(x) (round to nearest), lo = fp16(x - hi); x - hi is exact in fp32
__device__ __forceinline__ void split16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B.
//  K-major : rows of 128 B (32 tf32 of K), 8-row groups 1024 B apart (SBO); LBO unused (=1)
//  MN-major: K-rows of 128 B (32 tf32 of M/N), 4-K-row groups SBO apart, 32-element MN groups LBO apart
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// tf32 MN-major operands only exist as SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows x
// 128 B, the 32-byte chunk index XORed with (K-row % 4)  [cute Layout_MN_SW128_32B_Atom]
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) |
           ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of 16-byte chunk c (0..7) inside K-row r of an MN-major group
__device__ __forceinline__ uint32_t mn_swz(int c, int r) {
    return (uint32_t)((((((c & 7) >> 1) ^ (r & 3)) << 1) | (c & 1)) << 4);
}

__device__ __forceinline__ int tap_delta(int tap, int kh, int kw, int dil, int W) {
    int th = tap / kw - kh / 2, tw = tap % kw - kw / 2;
    return (th * W + tw) * dil;
}

// ------------------------------------------------------------------ fused epilogue
// Called by 4 or 8 warps; `quad` (= warp index % 4) selects the 32 TMEM lanes (= accumulator rows) a
// warp may read, `half`/`nhalf` split the 32-column chunks of the tile between the warps of one quad.
// tcgen05.ld hands every thread one ROW (32 consecutive columns); global memory wants whole 128-byte
// lines.  Each 32x32 chunk is therefore transposed through a 4 KB XOR-swizzled shared-memory tile
// (the pipeline stages are free once the accumulator is complete) with 128-bit accesses on both
// sides (conflict-free: float4 slot j of row r lives at slot j ^ (r & 7)).  Afterwards a lane owns
// 4 consecutive columns of 8 rows, so every global access (residual / mask reads, the store, the
// second tf32-rounded store, split-K RED.ADD.v4) is a 16-byte vector access, 4 full lines per warp
// instruction, the per-column BN shift lives in registers and the column sums (d beta) cost two
// shuffle stages.  All epilogue options are branch-free selects inside the row loop: a lone warp is
// bound by branch / issue latency, not by arithmetic (the first version of this loop, with a
// warp-uniform branch per option and row, took ~7000 clk per chunk; see profiles/r1_notes.md).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Cluster split-K (ks > 1): the ks CTAs of a thread-block cluster hold partial accumulators of the SAME
// output tile (disjoint K ranges).  32-column chunk c is owned by cluster rank c % ks.  Phase 1
// (tc_epilogue_dump): every CTA copies the chunks it does not own from TMEM into its own shared memory,
// already in the swizzled staging layout.  After a cluster barrier the owner (tc_epilogue) adds the peers'
// partials straight out of their shared memory (DSMEM, ld.shared::cluster.v4) in the transposed domain and
// runs the normal fused epilogue -- the partial sums never touch HBM/L2, the output needs no zero-fill and
// no atomics, and the epilogue work of a tile is spread over the ks CTAs.
__device__ __forceinline__ float4 ld_dsmem4(uint32_t local_saddr, uint32_t peer) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(peer));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}

// Staging in the (idle) pipeline stages: one 4 KB transposition scratch tile (32 rows x 128 B) per epilogue warp,
// then, for cluster split-K, one 16 KB region (4 quads) per chunk this CTA does NOT own, packed densely:
// dump_idx = position of chunk c among the non-owned chunks of cluster rank d.
__device__ __forceinline__ int dump_idx(int c, int d, int ks) { return c - (c > d ? (c - d + ks - 1) / ks : 0); }
__device__ __forceinline__ float* dump_region(float* stg_base, int nepi, int idx, int quad) {
    return stg_base + (nepi + idx * 4 + quad) * 1024;
}

template <int BN, int OP>
__device__ __forceinline__ void tc_epilogue_dump(const TcGemmParams& p, uint32_t tmem_full_bar, uint32_t tmem_acc,
                                                 int quad, int half, int nhalf, int lane, int m0,
                                                 float* stg_base, int nepi, int ks, int rank TC_TR_ARG) {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (quad == 0 && half == 0 && lane == 0) TC_TR(52);
    if (ks == 1) return;
    const int row0 = m0 + quad * 32;
    const int nrows = (OP == TC_WGRAD) ? p.Cout : p.M;
    float rs = 1.f;
    if (p.rowscale && row0 + lane < nrows) rs = __ldg(p.rowscale + row0 + lane);
    const uint32_t trow = tmem_acc + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c = half; c < BN / 32; c += nhalf) {
        if (c % ks == rank) continue;
        float4* stg4 = reinterpret_cast<float4*>(dump_region(stg_base, nepi, dump_idx(c, rank, ks), quad));
        float x[32];
        tmem_ld32(trow + c * 32, x);
#pragma unroll
        for (int j = 0; j < 8; j++)
            stg4[lane * 8 + (j ^ (lane & 7))] =
                make_float4(x[4 * j] * rs, x[4 * j + 1] * rs, x[4 * j + 2] * rs, x[4 * j + 3] * rs);
    }
}

struct EpiCtx {
    int quad, row0, nvalid, ldo, ks, rank;
    uint32_t trow;
    float rs;
    bool has_res, has_mask, do_round, has_outr, is_atomic, has_out16;
    float* stg_base;
    float* scratch;            // this warp's 4 KB transposition tile
    int nepi;
    uint32_t cl_x, cl_nx;      // cluster split-K: DSMEM rank of K slice z of this tile = cl_x + cl_nx * z
};

// The per-element epilogue inputs of chunk cc (residual / mask: 8 independent 16-byte loads each).  They do not
// depend on the accumulator, so the first chunk's are issued BEFORE the wait for the main loop and every later
// chunk's as soon as the previous chunk has consumed its copies: their latency never sits on the critical path.
__device__ __forceinline__ void epi_load_res(const TcGemmParams& p, const EpiCtx& c, int cc, int lane, int n0, float4 (&rv)[8]) {
    if (!c.has_res) return;
    const float* rp = p.res + (size_t)(c.row0 + (lane >> 3)) * p.ldr + n0 + cc * 32 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; i++)
        rv[i] = (4 * i + (lane >> 3) < c.nvalid) ? ldg4(rp + (size_t)(4 * i) * p.ldr) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void epi_load_mask(const TcGemmParams& p, const EpiCtx& c, int cc, int lane, int n0, float4 (&mv)[8]) {
    if (!c.has_mask) return;
    const float* mp = p.mask + (size_t)(c.row0 + (lane >> 3)) * p.ldm + n0 + cc * 32 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; i++)
        mv[i] = (4 * i + (lane >> 3) < c.nvalid) ? ldg4(mp + (size_t)(4 * i) * p.ldm) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// per-column scale / shift of chunk cc (this lane's 4 columns): like the residual they do not depend on the accumulator,
// so the first chunk's are fetched before the wait for the main loop and the next chunk's while this one is finished
// (they used to be loaded at the top of every chunk and waited for ~500 clk each time: an L2 round trip per chunk)
struct EpiCols { float4 sc, sh; };
__device__ __forceinline__ void epi_load_cols(const TcGemmParams& p, int cc, int lane, int n0, EpiCols& cv) {
    const int col = n0 + cc * 32 + (lane & 7) * 4;
    cv.sc = p.scale ? ldg4(p.scale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
    cv.sh = p.shift ? ldg4(p.shift + col) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// fused epilogue of one 32-column chunk: accumulator from TMEM (+ the cluster peers' partials).  rv / mv / cv hold
// this chunk's residual / mask / column values on entry and the ones of chunk cc_next (if < BN/32) on exit.
template <int BN, int OP>
__device__ __forceinline__ void epi_chunk(const TcGemmParams& p, const EpiCtx& c, int cc, int cc_next, int lane, int n0,
                                          float4 (&rv)[8], float4 (&mv)[8], EpiCols& cv) {
    const int g = lane >> 3, q = lane & 7;         // after the transpose: rows 4i+g, columns 4q..4q+3
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c0 = cc * 32;
    float4* stg4 = reinterpret_cast<float4*>(c.scratch);
    (void)stg4;
        const int col = n0 + c0 + q * 4;
        const float4 sc = cv.sc, sh = cv.sh;
        if (cc_next < BN / 32) epi_load_cols(p, cc_next, lane, n0, cv);
        (void)one4;
        float x[32];
        {
            tmem_ld32(c.trow + c0, x);
#pragma unroll
            for (int j = 0; j < 8; j++)
                stg4[lane * 8 + (j ^ (lane & 7))] =
                    make_float4(x[4 * j] * c.rs, x[4 * j + 1] * c.rs, x[4 * j + 2] * c.rs, x[4 * j + 3] * c.rs);
            __syncwarp();
            // after the transpose x[4i..4i+3] = columns col..col+3 of row 4i+g
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int rl = 4 * i + g;
                const float4 t = stg4[rl * 8 + (q ^ (rl & 7))];
                x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
            }
            __syncwarp();         // the next chunk overwrites the scratch tile
            if (c.ks > 1) {       // cluster split-K: add the peers' partials out of their shared memory
#pragma unroll 1
                for (int pr = 1; pr < c.ks; pr++) {
                    const int pz = (c.rank + pr) % c.ks;
                    const uint32_t peer = c.cl_x + c.cl_nx * (uint32_t)pz;
                    const uint32_t sreg = smem_u32(dump_region(c.stg_base, c.nepi, dump_idx(cc, pz, c.ks), c.quad));
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int rl = 4 * i + g;
                        const float4 t = ld_dsmem4(sreg + (uint32_t)(rl * 8 + (q ^ (rl & 7))) * 16u, peer);
                        x[4 * i] += t.x; x[4 * i + 1] += t.y; x[4 * i + 2] += t.z; x[4 * i + 3] += t.w;
                    }
                }
            }
        }
        // every option below is ONE warp-uniform branch around a straight run of 32 independent operations
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[4 * i] = fmaf(x[4 * i], sc.x, sh.x);
            x[4 * i + 1] = fmaf(x[4 * i + 1], sc.y, sh.y);
            x[4 * i + 2] = fmaf(x[4 * i + 2], sc.z, sh.z);
            x[4 * i + 3] = fmaf(x[4 * i + 3], sc.w, sh.w);
        }
        if (c.has_res) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[4 * i] += rv[i].x; x[4 * i + 1] += rv[i].y; x[4 * i + 2] += rv[i].z; x[4 * i + 3] += rv[i].w;
            }
            if (cc_next < BN / 32) epi_load_res(p, c, cc_next, lane, n0, rv);
        }
        if (p.relu) {
#pragma unroll
            for (int e = 0; e < 32; e++) x[e] = fmaxf(x[e], 0.f);
        }
        if (c.has_mask) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[4 * i] = mv[i].x > 0.f ? x[4 * i] : 0.f;
                x[4 * i + 1] = mv[i].y > 0.f ? x[4 * i + 1] : 0.f;
                x[4 * i + 2] = mv[i].z > 0.f ? x[4 * i + 2] : 0.f;
                x[4 * i + 3] = mv[i].w > 0.f ? x[4 * i + 3] : 0.f;
            }
            if (cc_next < BN / 32) epi_load_mask(p, c, cc_next, lane, n0, mv);
        }
        if (p.scale2) {
            const float4 s2 = ldg4(p.scale2 + col);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[4 * i] *= s2.x; x[4 * i + 1] *= s2.y; x[4 * i + 2] *= s2.z; x[4 * i + 3] *= s2.w;
            }
        }
        if (c.has_out16) {
            // split copy of the (unrounded) result for the next h3 GEMM: 4 hi halves at the column's place in the
            // [hi | lo] block of its 32-column group, 4 lo halves 64 bytes further
            unsigned char* o16 = reinterpret_cast<unsigned char*>(p.out16) + ((size_t)(c.row0 + g) * p.ldo16 + (col & ~31)) * 4 +
                                 (col & 31) * 2;
            const size_t step16 = (size_t)16 * p.ldo16;
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t h0, l0, h1, l1;
                split16_pair(x[4 * i], x[4 * i + 1], h0, l0);
                split16_pair(x[4 * i + 2], x[4 * i + 3], h1, l1);
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x[4 * i]), fabsf(x[4 * i + 1]))), fmaxf(fabsf(x[4 * i + 2]), fabsf(x[4 * i + 3])));
                if (4 * i + g < c.nvalid) {
                    *reinterpret_cast<uint2*>(o16 + i * step16) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(o16 + i * step16 + 64) = make_uint2(l0, l1);
                }
            }
            if (mx > 65504.f && p.overflow) *p.overflow = 1;
        }
        if (c.do_round) {
#pragma unroll
            for (int e = 0; e < 32; e++) x[e] = round_tf32(x[e]);
        }
        float* op = p.out + (size_t)(c.row0 + g) * c.ldo + col;
        const size_t ostep = (size_t)4 * c.ldo;
        if (c.nvalid >= 32) {
            if (c.is_atomic) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    atomicAdd(reinterpret_cast<float4*>(op + i * ostep), make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]));
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    *reinterpret_cast<float4*>(op + i * ostep) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (4 * i + g < c.nvalid) {
                    const float4 o = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                    if (c.is_atomic) atomicAdd(reinterpret_cast<float4*>(op + i * ostep), o);
                    else *reinterpret_cast<float4*>(op + i * ostep) = o;
                } else {
                    x[4 * i] = x[4 * i + 1] = x[4 * i + 2] = x[4 * i + 3] = 0.f;   // keep tail rows out of colsum
                }
            }
        }
        if (p.colsum) {
            float4 cs = zero4;
#pragma unroll
            for (int i = 0; i < 8; i++) { cs.x += x[4 * i]; cs.y += x[4 * i + 1]; cs.z += x[4 * i + 2]; cs.w += x[4 * i + 3]; }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
                cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
                cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            }
            if (g == 0 && c.nvalid > 0) atomicAdd(reinterpret_cast<float4*>(p.colsum + col), cs);
        }
        if (c.has_outr) {
            float* orp = p.out_r + (size_t)(c.row0 + g) * p.ldor + col;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (4 * i + g < c.nvalid)
                    *reinterpret_cast<float4*>(orp + (size_t)(4 * i) * p.ldor) =
                        make_float4(round_tf32(x[4 * i]), round_tf32(x[4 * i + 1]), round_tf32(x[4 * i + 2]), round_tf32(x[4 * i + 3]));
        }
}

// ks == 1: this CTA owns the whole accumulator; ks > 1 (cluster split-K, see tc_epilogue_dump): it finishes the
// chunks it owns.  (A third variant -- slices meeting in a global fp32 workspace, last arriver finishes the tile --
// was measured 1.5-2x slower than either: L2 RED.ADD throughput plus a fence per slice; profiles/r1_notes.md.)
template <int OP>
__device__ __forceinline__ void epi_setup(EpiCtx& c, const TcGemmParams& p, uint32_t tmem_acc, int quad, int half, int lane,
                                          int m0, float* stg_base, int ks, int rank, int nepi, uint32_t cl_x, uint32_t cl_nx) {
    c.cl_x = cl_x; c.cl_nx = cl_nx;
    c.quad = quad;
    c.row0 = m0 + quad * 32;                       // first accumulator row of this warp
    const int nrows = (OP == TC_WGRAD) ? p.Cout : p.M;
    c.nvalid = nrows - c.row0;                     // warp-uniform; >= 32 except in the last row tile
    c.trow = tmem_acc + ((uint32_t)(quad * 32) << 16);
    c.ldo = (OP == TC_WGRAD) ? p.ldw : p.ldo;
    c.rs = 1.f;
    if (p.rowscale && c.row0 + lane < nrows) c.rs = __ldg(p.rowscale + c.row0 + lane);
    c.has_res = p.res != nullptr; c.has_mask = p.mask != nullptr; c.do_round = p.round_tf32 != 0;
    c.has_outr = p.out_r != nullptr; c.is_atomic = p.atomic != 0; c.has_out16 = p.out16 != nullptr;
    c.stg_base = stg_base;
    c.scratch = stg_base + (quad + 4 * half) * 1024;
    c.nepi = nepi;
    c.ks = ks;
    c.rank = rank;
}

// owned chunks: rank, rank + ks, ...; the j-th of them goes to the warp with half == j % nhalf
template <int BN, int OP>
__device__ __forceinline__ void tc_epilogue(const TcGemmParams& p, const EpiCtx& c, int half, int nhalf, int lane, int n0,
                                            float4 (&rv)[8], float4 (&mv)[8], EpiCols& cv TC_TR_ARG) {
    const int step = nhalf * c.ks;
#pragma unroll 1
    for (int cc = c.rank + half * c.ks; cc < BN / 32; cc += step) {
        epi_chunk<BN, OP>(p, c, cc, cc + step, lane, n0, rv, mv, cv);
        if (c.quad == 0 && lane == 0 && half == 0 && cc == c.rank) TC_TR(53);
    }
    if (c.quad == 0 && half == 0 && lane == 0) TC_TR(54);
    tc_fence_before();
}

// ------------------------------------------------------------------ the kernel
template <int BN, int OP>
__global__ void __launch_bounds__(kTcThreads)
tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024 B alignment
    constexpr uint32_t kStage = kTcABytes + BN * 128;
    constexpr int kTcStages = tc_stages<BN>();
    const uint32_t bar_base = base + kTcStages * kStage;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * kTcStages);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kTcStages) + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = p.kh * p.kw;

    // ---- K range of this CTA
    int nkb;
    if (OP == TC_FWD) nkb = taps * (p.Cin / kTcBK);
    else if (OP == TC_DGRAD) nkb = taps * (p.Cout / kTcBK);
    else nkb = (p.M + kTcBK - 1) / kTcBK;
    const int per = (nkb + p.ksplit - 1) / p.ksplit;
    const int kb0 = blockIdx.z * per;
    const int kb1 = min(nkb, kb0 + per);
    const int nk = kb1 - kb0;
    if (nk <= 0) return;   // uniform per CTA (only possible with split-K)

    if (tid == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(full_bar(s), 128);   // one deferred arrive per producer thread
            mbar_init(empty_bar(s), 1);    // one tcgen05.commit
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_ptr;

    const int m0 = blockIdx.x * kTcBM;   // FWD/DGRAD: first pixel; WGRAD: first output channel
    const int n0 = blockIdx.y * BN;      // FWD: co0; DGRAD: ci0; WGRAD: column in (tap,ci)

    if (warp < 4) {
        // =========================== PRODUCERS ===========================
        int stage = 0;
        uint32_t phase = 0;
        if (OP == TC_FWD || OP == TC_DGRAD) {
            // A rows are fixed per thread: chunk = tid&7, rows (tid>>3)+16j
            const int chunk = tid & 7;
            const int rbase = tid >> 3;
            const int Ck = (OP == TC_FWD) ? p.Cin : p.Cout;   // channels of the gathered tensor
            const int cblocks = Ck / kTcBK;
            unsigned rowmask[8];
            bool rowok[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int m = m0 + rbase + 16 * j;
                rowok[j] = m < p.M;
                rowmask[j] = (p.tapmask && rowok[j]) ? p.tapmask[m] : 0xFFFFu;
            }
            for (int i = 0; i < nk; i++) {
                const int kb = kb0 + i;
                const int tap = kb / cblocks, cb = kb - tap * cblocks;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sA = base + stage * kStage, sB = sA + kTcABytes;
                int delta = (taps > 1) ? tap_delta(tap, p.kh, p.kw, p.dil, p.W) : 0;
                int mtap = tap;
                if (OP == TC_DGRAD) { delta = -delta; mtap = taps - 1 - tap; }
                const float* xsrc = p.X + (size_t)cb * kTcBK + chunk * 4;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int r = rbase + 16 * j;
                    const bool ok = rowok[j] && ((rowmask[j] >> mtap) & 1u);
                    const long pix = ok ? (long)(m0 + r + delta) : 0;
                    cp_async16(sA + r * 128 + ((chunk ^ (r & 7)) << 4), xsrc + pix * p.ldx, ok);
                }
                if (OP == TC_FWD) {
                    // B: BN weight rows, K-major
                    const float* wsrc = p.Wt + (size_t)kb * kTcBK + chunk * 4;
#pragma unroll
                    for (int j = 0; j < BN / 16; j++) {
                        const int r = rbase + 16 * j;
                        cp_async16(sB + r * 128 + ((chunk ^ (r & 7)) << 4),
                                   wsrc + (size_t)(n0 + r) * p.ldw, true);
                    }
                } else {
                    // B: 32 K-rows (co) x BN contiguous ci, MN-major
                    constexpr int cpr = BN / 4;
                    const float* wsrc = p.Wt + (size_t)tap * p.Cin + n0;
#pragma unroll
                    for (int j = 0; j < cpr / 4; j++) {
                        const int idx = tid + 128 * j;
                        const int r = idx / cpr, c = idx % cpr;
                        cp_async16(sB + (c >> 3) * 4096 + r * 128 + mn_swz(c, r),
                                   wsrc + (size_t)(cb * kTcBK + r) * p.ldw + c * 4, true);
                    }
                }
                cp_async_arrive_noinc(full_bar(stage));
                if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
            }
        } else {
            // WGRAD: K = pixels.  A: dY rows (MN-major, 128 co wide); B: gathered X (MN-major, BN ci wide)
            const int tap = n0 / p.Cin, ci0 = n0 - tap * p.Cin;
            const int delta = (taps > 1) ? tap_delta(tap, p.kh, p.kw, p.dil, p.W) : 0;
            constexpr int cpr = BN / 4;
            for (int i = 0; i < nk; i++) {
                const int k0 = (kb0 + i) * kTcBK;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sA = base + stage * kStage, sB = sA + kTcABytes;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int idx = tid + 128 * j;
                    const int r = idx >> 5, c = idx & 31;
                    const int m = k0 + r, co = m0 + c * 4;
                    const bool ok = m < p.M && co < p.Cout;
                    cp_async16(sA + (c >> 3) * 4096 + r * 128 + mn_swz(c, r),
                               p.Y + (ok ? (size_t)m * p.ldy + co : 0), ok);
                }
#pragma unroll
                for (int j = 0; j < cpr / 4; j++) {
                    const int idx = tid + 128 * j;
                    const int r = idx / cpr, c = idx % cpr;
                    const int m = k0 + r;
                    bool ok = m < p.M;
                    if (ok && p.tapmask) ok = (p.tapmask[m] >> tap) & 1u;
                    cp_async16(sB + (c >> 3) * 4096 + r * 128 + mn_swz(c, r),
                               p.X + (ok ? (size_t)(m + delta) * p.ldx + ci0 + c * 4 : 0), ok);
                }
                cp_async_arrive_noinc(full_bar(stage));
                if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
            }
        }

#ifdef MPB_TC_TRACE
        long long* trc = nullptr;
#endif
        float* stg_base = reinterpret_cast<float*>(smem_raw + (base - raw));
        EpiCtx ec;
        float4 rv[8], mv[8];
        epi_setup<OP>(ec, p, tmem_acc, warp, 0, lane, m0, stg_base, 1, 0, 4, 0, 1);
        EpiCols cv;
        epi_load_res(p, ec, 0, lane, n0, rv);
        epi_load_mask(p, ec, 0, lane, n0, mv);
        epi_load_cols(p, 0, lane, n0, cv);
        tc_epilogue_dump<BN, OP>(p, tmem_full_bar, tmem_acc, warp, 0, 1, lane, m0, stg_base, 4, 1, 0 TC_TR_PASS);
        tc_epilogue<BN, OP>(p, ec, 0, 1, lane, n0, rv, mv, cv TC_TR_PASS);
    } else {
        // =========================== MMA ISSUER (warp 4) ===========================
        constexpr bool a_mn = (OP == TC_WGRAD);
        constexpr bool b_mn = (OP != TC_FWD);
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) |
                                   ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(kTcBM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < nk; i++) {
            mbar_wait(full_bar(stage), phase);
            // cp.async wrote through the generic proxy; the MMA reads through the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sA = base + stage * kStage, sB = sA + kTcABytes;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; k++) {
                    const uint64_t ad = a_mn ? desc_mnmajor(sA + k * 1024) : desc_kmajor(sA + k * 32);
                    const uint64_t bd = b_mn ? desc_mnmajor(sB + k * 1024) : desc_kmajor(sB + k * 32);
                    umma_tf32(tmem_acc, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar(stage));               // frees the smem stage when the MMAs retire
                if (i == nk - 1) umma_commit(tmem_full_bar);  // accumulator complete
            }
            __syncwarp();
            if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN)
                     : "memory");
    }
}


// =====================================================================================
// TMA-fed variant: every operand tile is brought in by the Tensor Memory Accelerator
// (cp.async.bulk.tensor), so the whole producer side is ONE thread and no LSU issue slots
// are spent on operand traffic:
//   K-major tiles   : tiled 2-D maps (box 32 floats x rows, SWIZZLE_128B)            -- 1x1 / FC / weights
//   gathered pixels : im2col 4-D maps over the NHWC tensor (C,W,H,N); the filter tap is the
//                     per-instruction {offW,offH}; padding taps are zero-filled by the TMA
//   MN-major tiles  : 32x32 boxes with SWIZZLE_128B_ATOM_32B, one per 32-column group
// Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue.
// =====================================================================================
// epilogue warps / minimum CTAs per SM per tile width (rationale in tc_gemm.cuh)
#ifndef MPB_EW64
#define MPB_EW64 4
#endif
#ifndef MPB_EW128
#define MPB_EW128 4
#endif
#ifndef MPB_EW256
#define MPB_EW256 8
#endif
#ifndef MPB_MINB128
#define MPB_MINB128 2
#endif
#ifndef MPB_MINB256
#define MPB_MINB256 1
#endif
template <int BN> constexpr int tma_epi_warps() { return BN == 64 ? MPB_EW64 : BN == 128 ? MPB_EW128 : MPB_EW256; }
template <int BN> constexpr int tma_threads() { return 64 + 32 * tma_epi_warps<BN>(); }
template <int BN> constexpr int tma_min_blocks() { return BN == 64 ? 2 : BN == 128 ? MPB_MINB128 : MPB_MINB256; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                                int h, int n, int offw, int offh) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"((unsigned short)offw), "h"((unsigned short)offh)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                                   int h, int n, int offw, int offh, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"((unsigned short)offw), "h"((unsigned short)offh), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, unsigned short mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// CN = CTAs per cluster along the N-tile axis.  The CN CTAs of a cluster compute different column
// tiles of the SAME 128 rows, so the A operand is identical for all of them: each CTA fetches
// 1/CN of the A tile and the TMA multicasts it into every CTA's shared memory -- L2 -> SM traffic
// for A drops by CN (the GEMMs of this network are bound by exactly that traffic).
// CSK: cluster split-K -- the cluster spans gridDim.z (= p.ksplit CTAs, disjoint K ranges of one tile) and the
// partial accumulators are reduced through distributed shared memory in the epilogue (see tc_epilogue_dump).
// H3: fp16 hi/lo split forward (mpb_tc_gemm_h3): the maps address the split copies X16 / W16 -- same byte geometry as
// the fp32 operands, so producer, stages and barriers are untouched; only the MMA issue differs (six kind::f16
// instructions per k-block instead of four kind::tf32).
template <int BN, int OP, int CN, bool CSK, bool H3 = false>
__global__ void __launch_bounds__(tma_threads<BN>(), tma_min_blocks<BN>())
tc_gemm_tma_kernel(const __grid_constant__ TcGemmParams p, const __grid_constant__ CUtensorMap mapA,
                   const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    constexpr uint32_t kStage = kTcABytes + BN * 128;
    constexpr int kTcStages = tc_stages<BN>();
    const uint32_t bar_base = base + kTcStages * kStage;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * kTcStages);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kTcStages) + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    (void)cta_lin;
    TC_TR_DECL;
    if (tid == 0) TC_TR(0);
    const int taps = p.kh * p.kw;
    int nkb;
    if (OP == TC_FWD) nkb = taps * (p.Cin / kTcBK);
    else if (OP == TC_DGRAD) nkb = taps * (p.Cout / kTcBK);
    else nkb = (p.M + kTcBK - 1) / kTcBK;
    const int per = (nkb + p.ksplit - 1) / p.ksplit;
    const int kb0 = blockIdx.z * per;
    const int nk = min(nkb, kb0 + per) - kb0;
    if (nk <= 0) return;

    const int m0 = blockIdx.x * kTcBM;
    const int n0 = blockIdx.y * BN;
    const int crank = (CN > 1) ? (int)cluster_ctarank() : 0;
    constexpr unsigned short kMask = (unsigned short)((1u << CN) - 1u);
    constexpr int kSlice = kTcBM / CN;      // A rows (pixels) fetched by this CTA

    if (tid == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(full_bar(s), 1);     // the producer's arrive.expect_tx (+ TMA byte count)
            mbar_init(empty_bar(s), CN);   // one tcgen05.commit from every CTA of the cluster
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // =========================== TMA PRODUCER (warp 0) ===========================
    // The whole warp walks the (warp-uniform) loop and one elected lane issues: loop state stays in uniform
    // registers and ptxas needs no per-lane wrapper around the UTMALDG / SYNCS instructions.  No integer division
    // inside the loop: (tap, channel block) and the pixel coordinates are carried.
    // The first kTcStages k-blocks need no free-slot wait and no other warp: they are issued BEFORE the CTA-wide
    // setup barrier, so the operands of the first MMAs are in flight while TMEM is being allocated (the first TMA
    // used to leave ~2200 clk after CTA entry, now ~600).
    int stage = 0;
    uint32_t phase = 0;
    constexpr uint32_t kBytes = kTcABytes + BN * 128;
    const int dil_ = p.dil;
    int pr_tap = 0, pr_cb = 0, pr_th = 0, pr_tw = 0, pr_bw = 0, pr_bh = 0, pr_img0 = 0, pr_ms = 0, pr_cblocks = 1;   // FWD/DGRAD
    int pr_k0 = 0, pr_img = 0, pr_ph = 0, pr_pw = 0, pr_ci0 = 0;                                                    // WGRAD
    // Two issuing warps: one producer iteration (free-slot wait, expect_tx, 2..9 TMA instructions with their
    // uniform-register set-up) takes ~420 clk of ONE thread's latency, more than the MMAs of a 64- or 128-wide
    // k-block (128 / 256 clk).  Warp 0 therefore issues the even k-blocks and the first epilogue warp (idle during
    // the main loop) the odd ones; both walk the same coordinate sequence and skip the other's iterations.
    constexpr bool kTwoProducers = (CN == 1);
    if (warp == 0 || (kTwoProducers && warp == 2)) {
        const int hw = p.H * p.W;
        if (OP == TC_FWD || OP == TC_DGRAD) {
            const int Ck = (OP == TC_FWD) ? p.Cin : p.Cout;
            pr_cblocks = Ck / kTcBK;
            pr_ms = m0 + crank * kSlice;     // first pixel of this CTA's slice of the A tile: base of the im2col walk
            pr_img0 = pr_ms / hw;
            const int rem0 = pr_ms - pr_img0 * hw, ph0 = rem0 / p.W, pw0 = rem0 - ph0 * p.W;
            pr_bw = pw0 - dil_ * (p.kw / 2); pr_bh = ph0 - dil_ * (p.kh / 2);
            pr_tap = kb0 / pr_cblocks; pr_cb = kb0 - pr_tap * pr_cblocks;
            pr_th = pr_tap / p.kw; pr_tw = pr_tap - pr_th * p.kw;
        } else {
            pr_tap = n0 / p.Cin; pr_ci0 = n0 - pr_tap * p.Cin;
            pr_th = pr_tap / p.kw; pr_tw = pr_tap - pr_th * p.kw;
            pr_k0 = kb0 * kTcBK;
            pr_img = pr_k0 / hw; pr_ph = (pr_k0 - pr_img * hw) / p.W; pr_pw = pr_k0 - pr_img * hw - pr_ph * p.W;
        }
    }
    auto produce = [&](int i, bool wait_free, bool mine) {
        if (mine && wait_free) mbar_wait(empty_bar(stage), phase ^ 1u);
        if (mine && elect_one()) {
            if (i < 16) TC_TR(4 + i);
            const uint32_t sA = base + stage * kStage, sB = sA + kTcABytes;
            mbar_expect_tx(full_bar(stage), kBytes);
            if (OP == TC_FWD || OP == TC_DGRAD) {
                const uint32_t sAs = sA + crank * kSlice * 128;
                if (taps == 1) {
                    if (CN > 1) tma_load_2d_mc(sAs, &mapA, full_bar(stage), pr_cb * kTcBK, pr_ms, kMask);
                    else tma_load_2d(sA, &mapA, full_bar(stage), pr_cb * kTcBK, m0);
                } else {
                    // dX[p] needs dY[p - off]: mirrored tap
                    const int ow = (OP == TC_DGRAD ? p.kw - 1 - pr_tw : pr_tw) * dil_;
                    const int oh = (OP == TC_DGRAD ? p.kh - 1 - pr_th : pr_th) * dil_;
                    if (CN > 1) tma_load_im2col_mc(sAs, &mapA, full_bar(stage), pr_cb * kTcBK, pr_bw, pr_bh, pr_img0, ow, oh, kMask);
                    else tma_load_im2col(sA, &mapA, full_bar(stage), pr_cb * kTcBK, pr_bw, pr_bh, pr_img0, ow, oh);
                }
                if (OP == TC_FWD) {
                    tma_load_2d(sB, &mapB, full_bar(stage), (pr_tap * pr_cblocks + pr_cb) * kTcBK, n0);
                } else {
#pragma unroll
                    for (int g = 0; g < BN / 32; g++)
                        tma_load_2d(sB + g * 4096, &mapB, full_bar(stage), pr_tap * p.Cin + n0 + 32 * g, pr_cb * kTcBK);
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    if (CN > 1) {       // the 4 column groups of the dY tile are split over the cluster
                        if (g / (4 / CN) == crank)
                            tma_load_2d_mc(sA + g * 4096, &mapA, full_bar(stage), m0 + 32 * g, pr_k0, kMask);
                    } else {
                        tma_load_2d(sA + g * 4096, &mapA, full_bar(stage), m0 + 32 * g, pr_k0);
                    }
                }
                if (taps == 1) {
#pragma unroll
                    for (int g = 0; g < BN / 32; g++)
                        tma_load_2d(sB + g * 4096, &mapB, full_bar(stage), pr_ci0 + 32 * g, pr_k0);
                } else {
#pragma unroll
                    for (int g = 0; g < BN / 32; g++)
                        tma_load_im2col(sB + g * 4096, &mapB, full_bar(stage), pr_ci0 + 32 * g, pr_pw - dil_ * (p.kw / 2),
                                        pr_ph - dil_ * (p.kh / 2), pr_img, pr_tw * dil_, pr_th * dil_);
                }
            }
        }
        __syncwarp();
        if (OP == TC_FWD || OP == TC_DGRAD) {
            if (++pr_cb == pr_cblocks) {
                pr_cb = 0; ++pr_tap;
                if (++pr_tw == p.kw) { pr_tw = 0; ++pr_th; }
            }
        } else {
            pr_k0 += kTcBK;
            pr_pw += kTcBK;
            while (pr_pw >= p.W) {
                pr_pw -= p.W;
                if (++pr_ph == p.H) { pr_ph = 0; ++pr_img; }
            }
        }
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
    };
    // PDL: everything above touched only parameters and shared memory; from here on global memory written by the
    // stream predecessor is read (no-op when the launch carries no programmatic dependency)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int npre = (CN == 1) ? min(nk, kTcStages) : 0;   // multicast needs the peers' barriers first
    if (warp == 0) {
        __syncwarp();                 // lane 0's barrier initialisation is visible to the elected lane
        for (int i = 0; i < npre; i++) produce(i, false, true);
    }

    tc_fence_before();
    __syncthreads();
    if (CN > 1) cluster_sync_all();      // every CTA's barriers exist before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_ptr;
#ifdef MPB_TC_TRACE
    if (tid == 0) {
        TC_TR(1);
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (trc) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            trc[2] = smid; trc[56] = nk; trc[3] = (long long)gt;
        }
    }
#endif

    if (warp == 0) {
        for (int i = npre; i < nk; i++) produce(i, true, !kTwoProducers || (i & 1) == 0);
    } else if (warp == 1) {
        // =========================== MMA ISSUER ===========================
        constexpr bool a_mn = (OP == TC_WGRAD);
        constexpr bool b_mn = (OP != TC_FWD);
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) |
                                   ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(kTcBM >> 4) << 24);
        // descriptors of stage 0 / k-step 0; stages and k-steps only move the 14-bit address field
        // (16-byte units; every tile lives below 256 KB, so the add never carries out of the field)
        const uint64_t ad0 = a_mn ? desc_mnmajor(base) : desc_kmajor(base);
        const uint64_t bd0 = b_mn ? desc_mnmajor(base + kTcABytes) : desc_kmajor(base + kTcABytes);
        constexpr uint32_t ka = (a_mn ? 1024u : 32u) >> 4, kb_ = (b_mn ? 1024u : 32u) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < nk; i++) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                if (i < 16) TC_TR(20 + i);
                const uint64_t ad = ad0 + (uint64_t)((stage * kStage) >> 4);
                const uint64_t bd = bd0 + (uint64_t)((stage * kStage) >> 4);
                if (H3) {
                    // a 128-byte row of A is [hi(32) | lo(32)] fp16, of B [lo(32) | hi(32)]: the four K=16 steps over
                    // the whole row give A_hi*B_lo + A_lo*B_hi, then A's first half against B's second half is
                    // A_hi*B_hi.  Small terms first.  (descriptor address units are 16 bytes: +2 = one K=16 step)
                    constexpr uint32_t idesc_h = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        umma_f16(tmem_acc, ad + k * 2, bd + k * 2, idesc_h, (i > 0 || k > 0) ? 1u : 0u);
                    umma_f16(tmem_acc, ad, bd + 4, idesc_h, 1u);
                    umma_f16(tmem_acc, ad + 2, bd + 6, idesc_h, 1u);
                } else {
#pragma unroll
                    for (int k = 0; k < kTcBK / 8; k++)
                        umma_tf32(tmem_acc, ad + k * ka, bd + k * kb_, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                if (CN > 1) umma_commit_mc(empty_bar(stage), kMask);   // frees this stage in every CTA
                else umma_commit(empty_bar(stage));
                if (i == nk - 1) {
                    umma_commit(tmem_full_bar);
                    // PDL: the successor may launch now -- its setup overlaps this CTA's epilogue.  (Releasing it
                    // at CTA start was measured slower: early successors sit on SM slots other streams could use.)
                    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                }
                if (i < 16) TC_TR(36 + i);
            }
            __syncwarp();
            if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
        tc_fence_before();
    }
    constexpr int EW = tma_epi_warps<BN>();
    float* stg_base = reinterpret_cast<float*>(smem_raw + (base - raw));
    const int ks = CSK ? p.ksplit : 1, krank = CSK ? (int)blockIdx.z : 0;
    // a cluster may hold the K slices of SEVERAL row tiles (clusterDim = (cx, 1, ks)): 3 slices x 2 tiles fill
    // whole TPCs, a bare cluster of 3 wastes every fourth SM.  DSMEM rank of slice z of my tile: my_x + cx * z
    uint32_t cl_x = 0, cl_nx = 1;
    if (CSK) {
        asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(cl_x));
        asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cl_nx));
    }
    EpiCtx ec;
    float4 rv[8], mv[8];
    EpiCols cv;
    if (warp >= 2) {
        const int half = (warp - 2) >> 2;
        epi_setup<OP>(ec, p, tmem_acc, warp & 3, half, lane, m0, stg_base, ks, krank, EW, cl_x, cl_nx);
        // residual / mask / column values of this warp's first chunk: in flight while the main loop runs
        if (krank + half * ks < BN / 32) {
            epi_load_res(p, ec, krank + half * ks, lane, n0, rv);
            epi_load_mask(p, ec, krank + half * ks, lane, n0, mv);
            epi_load_cols(p, krank + half * ks, lane, n0, cv);
        }
        if (kTwoProducers && warp == 2) {      // second issuing warp: the odd k-blocks (see above)
            for (int i = 0; i < npre; i++) produce(i, false, false);
            for (int i = npre; i < nk; i++) produce(i, true, (i & 1) == 1);
        }
        tc_epilogue_dump<BN, OP>(p, tmem_full_bar, tmem_acc, warp & 3, half, EW / 4, lane, m0, stg_base, EW, ks, krank
                                 TC_TR_PASS);
    }
    if (CSK) cluster_sync_all();         // every peer's partial chunks are in its shared memory
    if (warp >= 2) tc_epilogue<BN, OP>(p, ec, (warp - 2) >> 2, EW / 4, lane, n0, rv, mv, cv TC_TR_PASS);
    __syncthreads();
    if (tid == 0) TC_TR(55);
    if (CN > 1 || CSK) cluster_sync_all();   // no CTA leaves while peers may still multicast into it / read its smem
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN) : "memory");
#ifdef MPB_TC_TRACE
        if (lane == 0 && trc) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            trc[57] = (long long)gt;
        }
#endif
    }
}

// =====================================================================================
// 3xTF32 forward variant ("x3", opt-in: mpb_tc_gemm_x3 / MPB_PRECISION=x3 in the engine).
//
// kind::tf32 reads the top 19 bits of each fp32 operand (TRUNCATION, measured: profiles/r1_notes.md), so one
// pass carries ~3e-4 relative error per GEMM, which the decoder's train-mode batch norm amplifies past the 1e-3
// parity bar (tools/precision_study.py).  Here the operands stay UNROUNDED fp32 in HBM and every k-block is
// split on chip:  x = hi + lo,  hi = x & 0xFFFFE000 (exactly what the tensor core reads when handed x itself,
// so it needs no copy),  lo = x - hi (exact in fp32, <= 13 significant bits).  Three MMAs per k-step,
//     acc += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo          (the dropped lo*lo term is ~2^-22 relative)
// give fp32-level accuracy (3.6e-7 on a K=2304 contraction, numpy emulation in tests/test_x3_split.py) for the
// HBM / L2 traffic of the single-pass kernel: the TMA brings each tile in once and the four epilogue warps,
// idle during the main loop, write the lo tiles (same swizzled layout, so the hi descriptors + a constant
// offset address them).
//   stage = [A | B | A_lo | B_lo];  barriers: full (TMA landed) -> split (lo written, 128 arrivals) -> MMA ->
//   empty (tcgen05.commit).  Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = splitters, then
//   the same fused epilogue as the single-pass kernel.  FWD only, no multicast / cluster split-K.
// NOT YET RUN ON A B200 (written after the round-1 GPU budget was spent); nothing calls it by default.
// =====================================================================================
#ifndef MPB_X3_STAGES
#define MPB_X3_STAGES 3          // sweepable: MPB_NVCC_EXTRA="-DMPB_X3_STAGES=2 -DMPB_X3_MINB=2" puts two 64-wide CTAs on an SM
#endif
#ifndef MPB_X3_MINB
#define MPB_X3_MINB 1
#endif
constexpr int kX3Stages = MPB_X3_STAGES;
template <int BN> constexpr uint32_t x3_half_bytes() { return kTcABytes + BN * 128; }      // [A | B]
template <int BN> constexpr int x3_smem_bytes() { return kX3Stages * 2 * (int)x3_half_bytes<BN>() + 1024 + 256; }
constexpr int kX3Threads = 192;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float tf32_lo(float x) {      // x - (what kind::tf32 reads of x)
    return __fsub_rn(x, __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

template <int BN>
__global__ void __launch_bounds__(kX3Threads, MPB_X3_MINB)
tc_gemm_x3_kernel(const __grid_constant__ TcGemmParams p, const __grid_constant__ CUtensorMap mapA,
                  const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    constexpr uint32_t kHalf = x3_half_bytes<BN>();
    constexpr uint32_t kStage = 2 * kHalf;
    static_assert(kHalf % 1024 == 0 && kTcABytes % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
    static_assert((kHalf / 16) % 128 == 0, "the splitter threads share the float4s of a stage evenly");
    static_assert(x3_smem_bytes<BN>() <= 227 * 1024, "shared memory per CTA");
    static_assert(((kX3Stages * kStage + 2048) >> 4) < (1u << 14), "descriptor address field: no carry");
    const uint32_t bar_base = base + kX3Stages * kStage;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto split_bar = [&](int s) { return bar_base + 8u * (kX3Stages + s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kX3Stages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (3 * kX3Stages);
    const uint32_t tmem_slot = tmem_full_bar + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef MPB_TC_TRACE
    long long* trc = nullptr;
#endif
    const int taps = p.kh * p.kw;
    const int cblocks = p.Cin / kTcBK;
    const int nkb = taps * cblocks;
    const int per = (nkb + p.ksplit - 1) / p.ksplit;
    const int kb0 = blockIdx.z * per;
    const int nk = min(nkb, kb0 + per) - kb0;
    if (nk <= 0) return;
    const int m0 = blockIdx.x * kTcBM;
    const int n0 = blockIdx.y * BN;

    if (tid == 0) {
        for (int s = 0; s < kX3Stages; s++) {
            mbar_init(full_bar(s), 1);       // the producer's arrive.expect_tx (+ TMA byte count)
            mbar_init(split_bar(s), 128);    // every splitter thread, after its own cross-proxy fence
            mbar_init(empty_bar(s), 1);      // one tcgen05.commit
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- TMA producer state (warp 0; warp-uniform, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    int pr_tap = 0, pr_cb = 0, pr_th = 0, pr_tw = 0, pr_bw = 0, pr_bh = 0, pr_img0 = 0;
    if (warp == 0) {
        const int hw = p.H * p.W;
        pr_img0 = m0 / hw;
        const int rem0 = m0 - pr_img0 * hw, ph0 = rem0 / p.W, pw0 = rem0 - ph0 * p.W;
        pr_bw = pw0 - p.dil * (p.kw / 2); pr_bh = ph0 - p.dil * (p.kh / 2);
        pr_tap = kb0 / cblocks; pr_cb = kb0 - pr_tap * cblocks;
        pr_th = pr_tap / p.kw; pr_tw = pr_tap - pr_th * p.kw;
    }
    auto produce = [&](bool wait_free) {
        if (wait_free) mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
            const uint32_t sA = base + stage * kStage, sB = sA + kTcABytes;
            mbar_expect_tx(full_bar(stage), kHalf);
            if (taps == 1) tma_load_2d(sA, &mapA, full_bar(stage), pr_cb * kTcBK, m0);
            else tma_load_im2col(sA, &mapA, full_bar(stage), pr_cb * kTcBK, pr_bw, pr_bh, pr_img0, pr_tw * p.dil, pr_th * p.dil);
            tma_load_2d(sB, &mapB, full_bar(stage), (pr_tap * cblocks + pr_cb) * kTcBK, n0);
        }
        __syncwarp();
        if (++pr_cb == cblocks) {
            pr_cb = 0; ++pr_tap;
            if (++pr_tw == p.kw) { pr_tw = 0; ++pr_th; }
        }
        if (++stage == kX3Stages) { stage = 0; phase ^= 1u; }
    };
    asm volatile("griddepcontrol.wait;" ::: "memory");      // PDL: global memory of the stream predecessor from here on
    const int npre = min(nk, kX3Stages);
    if (warp == 0) {
        __syncwarp();
        for (int i = 0; i < npre; i++) produce(false);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_ptr;

    float* stg_base = reinterpret_cast<float*>(smem_raw + (base - raw));
    EpiCtx ec;
    float4 rv[8], mv[8];
    EpiCols cv;
    if (warp == 0) {
        for (int i = npre; i < nk; i++) produce(true);
    } else if (warp == 1) {
        // =========================== MMA ISSUER ===========================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(kTcBM >> 4) << 24);
        const uint64_t ad0 = desc_kmajor(base), bd0 = desc_kmajor(base + kTcABytes);
        constexpr uint64_t lo_off = kHalf >> 4;
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nk; i++) {
            mbar_wait(split_bar(st), ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lo tiles came through the generic proxy
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = ad0 + (uint64_t)((st * kStage) >> 4);
                const uint64_t bd = bd0 + (uint64_t)((st * kStage) >> 4);
#pragma unroll
                for (int k = 0; k < kTcBK / 8; k++) {
                    umma_tf32(tmem_acc, ad + lo_off + k * 2, bd + k * 2, idesc, (i > 0 || k > 0) ? 1u : 0u);   // A_lo * B_hi
                    umma_tf32(tmem_acc, ad + k * 2, bd + lo_off + k * 2, idesc, 1u);                            // A_hi * B_lo
                    umma_tf32(tmem_acc, ad + k * 2, bd + k * 2, idesc, 1u);                                     // A_hi * B_hi
                }
                umma_commit(empty_bar(st));
                if (i == nk - 1) {
                    umma_commit(tmem_full_bar);
                    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                }
            }
            __syncwarp();
            if (++st == kX3Stages) { st = 0; ph ^= 1u; }
        }
        tc_fence_before();
    } else {
        // =========================== SPLITTERS, then EPILOGUE (warps 2..5) ===========================
        epi_setup<TC_FWD>(ec, p, tmem_acc, warp & 3, 0, lane, m0, stg_base, 1, 0, 4, 0, 1);
        epi_load_res(p, ec, 0, lane, n0, rv);
        epi_load_mask(p, ec, 0, lane, n0, mv);
        epi_load_cols(p, 0, lane, n0, cv);
        const int t = tid - 64;
        constexpr int n4 = (int)(kHalf / 16);          // float4s of [A | B]; a multiple of 128
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nk; i++) {
            mbar_wait(full_bar(st), ph);
            const uint32_t hi = base + st * kStage + (uint32_t)t * 16u, lo = hi + kHalf;
#pragma unroll 4
            for (int j = 0; j < n4 / 128; j++) {
                const float4 v = lds4(hi + (uint32_t)j * 2048u);
                sts4(lo + (uint32_t)j * 2048u, make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            mbar_arrive(split_bar(st));
            if (++st == kX3Stages) { st = 0; ph ^= 1u; }
        }
        tc_epilogue_dump<BN, TC_FWD>(p, tmem_full_bar, tmem_acc, warp & 3, 0, 1, lane, m0, stg_base, 4, 1, 0 TC_TR_PASS);
        tc_epilogue<BN, TC_FWD>(p, ec, 0, 1, lane, n0, rv, mv, cv TC_TR_PASS);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN) : "memory");
    }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static bool tma_api_ready() {
    if (g_encode_tiled && g_encode_im2col) return true;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
    g_encode_tiled = (EncodeTiledFn)f;
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
    g_encode_im2col = (EncodeIm2colFn)f;
    cudaDriverGetVersion(&g_driver_version);
    return true;
}

// rows x inner fp32 matrix with a row pitch of `pitch` floats; box = 32 floats x box_rows
static bool make_map_2d(CUtensorMap* m, const float* ptr, long inner, long rows, long pitch, int box_rows, bool mn_major) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// NHWC tensor (C channels used, pixel pitch `pitch` floats) walked in im2col order for a kh x kw filter
// with atrous rate r and SAME padding; box = 32 channels x `pixels` consecutive output pixels
static bool make_map_im2col(CUtensorMap* m, const float* ptr, int C, int W, int H, int N, long pitch, int kh, int kw,
                            int r, int pixels, bool mn_major) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * 4 * W, (cuuint64_t)pitch * 4 * W * H};
    // base pixel range [-pad, dim + upper): pad = r*(k/2); upper = pad - (k-1)*r   (WHD order: {W, H})
    int lower[2] = {-r * (kw / 2), -r * (kh / 2)};
    int upper[2] = {r * (kw / 2) - (kw - 1) * r, r * (kh / 2) - (kh - 1) * r};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult res = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, lower, upper, 32,
                                   (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) return false;
    // driver workaround also applied by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors
    if (g_driver_version <= 13010 && (size_t)pitch * 4 * W * H * N < 131072)
        reinterpret_cast<uint64_t*>(m)[1] &= ~(1llu << 21);
    return true;
}

// Programmatic dependent launch (default on, MPB_PDL=0 disables): every GEMM waits (griddepcontrol.wait) for its stream predecessor
// before touching global memory and releases its successor (griddepcontrol.launch_dependents) when its
// accumulator is complete, so the successor's launch latency and setup overlap this grid's epilogue.
static int g_tc_pdl = -1;
static bool tc_gemm_pdl() {
    if (g_tc_pdl < 0) {
        const char* e = getenv("MPB_PDL");
        g_tc_pdl = (e && atoi(e) == 0) ? 0 : 1;
    }
    return g_tc_pdl == 1;
}

// row tiles per cluster of a cluster split-K launch: make the cluster an even number of CTAs (whole TPCs)
static int csk_cluster_x(unsigned gx, unsigned ks) {
    if (ks % 2 == 0 || gx % 2 != 0 || ks * 2 > 8) return 1;
    return 2;
}

template <int BN, int OP, int CN, bool CSK = false, bool H3 = false>
static int launch_tma_cn(const TcGemmParams& p, dim3 grid, cudaStream_t s) {
    alignas(64) CUtensorMap mapA, mapB;
    const int taps = p.kh * p.kw;
    const int nimg = p.M / (p.H * p.W);
    bool ok = true;
    if (OP == TC_FWD || OP == TC_DGRAD) {
        const int Ck = (OP == TC_FWD) ? p.Cin : p.Cout;
        // h3: the split copies have the byte geometry of the fp32 operands (32 floats = 64 halves = one 128-byte
        // block), so the fp32 maps address them unchanged
        const float* Xp = H3 ? reinterpret_cast<const float*>(p.X16) : p.X;
        const float* Wp = H3 ? reinterpret_cast<const float*>(p.W16) : p.Wt;
        if (taps == 1) ok &= make_map_2d(&mapA, Xp, Ck, p.M, p.ldx, kTcBM / CN, false);
        else ok &= make_map_im2col(&mapA, Xp, Ck, p.W, p.H, nimg, p.ldx, p.kh, p.kw, p.dil, kTcBM / CN, false);
        if (OP == TC_FWD) ok &= make_map_2d(&mapB, Wp, (long)taps * p.Cin, p.Cout, p.ldw, BN, false);
        else ok &= make_map_2d(&mapB, p.Wt, (long)taps * p.Cin, p.Cout, p.ldw, 32, true);
    } else {
        ok &= make_map_2d(&mapA, p.Y, p.Cout, p.M, p.ldy, 32, true);
        if (taps == 1) ok &= make_map_2d(&mapB, p.X, p.Cin, p.M, p.ldx, 32, true);
        else ok &= make_map_im2col(&mapB, p.X, p.Cin, p.W, p.H, nimg, p.ldx, p.kh, p.kw, p.dil, 32, true);
    }
    if (!ok) return -2;
    constexpr int smem = tc_smem_bytes<BN>();
    static bool attr_set = false;
    if (!attr_set) {
        MPB_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_tma_kernel<BN, OP, CN, CSK, H3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const bool pdl = tc_gemm_pdl();
    if (CN == 1 && !CSK && !pdl) {
        tc_gemm_tma_kernel<BN, OP, CN, CSK, H3><<<grid, tma_threads<BN>(), smem, s>>>(p, mapA, mapB);
        MPB_LAUNCH_CHECK();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(tma_threads<BN>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CN > 1 || CSK) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CSK ? csk_cluster_x(grid.x, grid.z) : 1;
        at[na].val.clusterDim.y = CN;
        at[na].val.clusterDim.z = CSK ? grid.z : 1;
        na++;
    }
    if (pdl) {      // programmatic dependent launch: this grid may start while its stream predecessor drains
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    MPB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_gemm_tma_kernel<BN, OP, CN, CSK, H3>, p, mapA, mapB));
    count_launch();
    return 0;
}

static int g_tc_cluster = -1;   // max CTAs per cluster along N (1 disables multicast)
template <int BN, int OP>
static int launch_tma(const TcGemmParams& p, dim3 grid, cudaStream_t s) {
    if (p.ksplit > 1 && !p.atomic) return launch_tma_cn<BN, OP, 1, true>(p, grid, s);   // cluster split-K
    if (g_tc_cluster < 0) {
        // Measured on B200 (profiles/r1_notes.md): multicasting A over a cluster of column tiles does
        // NOT speed these GEMMs up (21.1 vs 19.6 ms of GEMM time per step) -- they are bound by the
        // bytes each SM must RECEIVE per MMA, which multicast does not change -- so it is off by default.
        const char* e = getenv("MPB_TC_CLUSTER");
        g_tc_cluster = e ? atoi(e) : 1;
    }
    if (g_tc_cluster >= 4 && grid.y % 4 == 0) return launch_tma_cn<BN, OP, 4>(p, grid, s);
    if (g_tc_cluster >= 2 && grid.y % 2 == 0) return launch_tma_cn<BN, OP, 2>(p, grid, s);
    return launch_tma_cn<BN, OP, 1>(p, grid, s);
}

template <int BN>
static int launch_x3(const TcGemmParams& p, dim3 grid, cudaStream_t s) {
    alignas(64) CUtensorMap mapA, mapB;
    const int taps = p.kh * p.kw;
    const int nimg = p.M / (p.H * p.W);
    bool ok = true;
    if (taps == 1) ok &= make_map_2d(&mapA, p.X, p.Cin, p.M, p.ldx, kTcBM, false);
    else ok &= make_map_im2col(&mapA, p.X, p.Cin, p.W, p.H, nimg, p.ldx, p.kh, p.kw, p.dil, kTcBM, false);
    ok &= make_map_2d(&mapB, p.Wt, (long)taps * p.Cin, p.Cout, p.ldw, BN, false);
    if (!ok) return -2;
    constexpr int smem = x3_smem_bytes<BN>();
    static bool attr_set = false;
    if (!attr_set) {
        MPB_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kX3Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (tc_gemm_pdl()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    MPB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_gemm_x3_kernel<BN>, p, mapA, mapB));
    count_launch();
    return 0;
}

// 3xTF32 forward GEMM (see tc_gemm_x3_kernel): FWD only, BN 64 or 128, whole images in the pixel grid, split-K only
// in its atomic (RED.ADD into a zeroed output) form.  Operands are read as stored -- hand it UNROUNDED fp32.
int tc_gemm_x3_launch(const TcGemmParams& p, int BN, cudaStream_t s) {
    if (p.op != TC_FWD || p.M <= 0 || p.Cin <= 0 || p.Cout <= 0 || p.ksplit < 1) return -1;
    if (BN != 64 && BN != 128) return -1;
    if (p.Cin % kTcBK || p.Cout % BN || p.ldo % 4 || p.ldx % 4 || p.ldw % 4) return -1;
    if (p.ksplit > 1 && !p.atomic) return -1;
    if (p.M % (p.H * p.W) != 0 || !tma_api_ready()) return -1;
    if (p.tapmask) { /* the im2col TMA zero-fills padding taps itself; the mask is only used by the cp.async path */ }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p.out) || !al16(p.out_r) || !al16(p.res) || !al16(p.mask) || !al16(p.scale) || !al16(p.shift) ||
        !al16(p.scale2) || !al16(p.colsum))
        return -1;
    if ((p.res && p.ldr % 4) || (p.mask && p.ldm % 4) || (p.out_r && p.ldor % 4)) return -1;
    const int taps = p.kh * p.kw;
    if (p.ksplit > 1 && (p.ksplit - 1) * ceil_div(taps * (p.Cin / kTcBK), p.ksplit) >= taps * (p.Cin / kTcBK)) return -1;
    dim3 grid(ceil_div(p.M, kTcBM), p.Cout / BN, p.ksplit);
    if (grid.y > 65535 || grid.z > 65535) return -1;
    return BN == 64 ? launch_x3<64>(p, grid, s) : launch_x3<128>(p, grid, s);
}

// fp16-split forward GEMM (see the H3 branch of tc_gemm_tma_kernel): FWD only, whole images in the pixel grid (TMA
// path), split-K in both forms of the single-pass kernel.
int tc_gemm_h3_launch(const TcGemmParams& p, int BN, cudaStream_t s) {
    if (p.op != TC_FWD || p.M <= 0 || p.Cin <= 0 || p.Cout <= 0 || p.ksplit < 1) return -1;
    if (BN != 64 && BN != 128 && BN != 256) return -1;
    if (!p.X16 || !p.W16 || !p.out) return -1;
    if (p.Cin % kTcBK || p.Cout % BN || p.ldo % 4 || p.ldx % 4 || p.ldw % 4) return -1;
    if (p.M % (p.H * p.W) != 0 || !tma_api_ready()) return -1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p.out) || !al16(p.out_r) || !al16(p.res) || !al16(p.mask) || !al16(p.scale) || !al16(p.shift) ||
        !al16(p.scale2) || !al16(p.colsum) || !al16(p.out16) || !al16(p.X16) || !al16(p.W16))
        return -1;
    if ((p.res && p.ldr % 4) || (p.mask && p.ldm % 4) || (p.out_r && p.ldor % 4) || (p.out16 && p.ldo16 % 32)) return -1;
    const int taps = p.kh * p.kw;
    const int nkb_ = taps * (p.Cin / kTcBK);
    if (p.ksplit > 1 && (p.ksplit - 1) * ceil_div(nkb_, p.ksplit) >= nkb_) return -1;
    if (p.ksplit > 1 && !p.atomic) return -1;      // cluster split-K: not instantiated for h3
    dim3 grid(ceil_div(p.M, kTcBM), p.Cout / BN, p.ksplit);
    if (grid.y > 65535 || grid.z > 65535) return -1;
    if (BN == 64) return launch_tma_cn<64, TC_FWD, 1, false, true>(p, grid, s);
    if (BN == 128) return launch_tma_cn<128, TC_FWD, 1, false, true>(p, grid, s);
    return launch_tma_cn<256, TC_FWD, 1, false, true>(p, grid, s);
}

static int g_tc_mode = -1;   // 0 = cp.async producers, 1 = TMA producers
int tc_gemm_mode() {
    if (g_tc_mode < 0) {
        const char* e = getenv("MPB_TC_PRODUCER");
        g_tc_mode = (e && e[0] == 'c') ? 0 : 1;    // MPB_TC_PRODUCER=cpasync selects the LSU-fed variant
        if (g_tc_mode == 1 && !tma_api_ready()) g_tc_mode = 0;
    }
    return g_tc_mode;
}
void tc_gemm_set_cluster(int c) { g_tc_cluster = c < 1 ? 1 : c; }
// how many clusters of (cx, 1, ks) CTAs of the FWD kernel with tile width BN can be resident at once
int tc_gemm_max_clusters(int BN, int cx, int ks) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cx * 64, 1, ks);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cx; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = ks;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaErrorInvalidValue;
#define MPB_Q(bn)                                                                                              \
    if (BN == bn) {                                                                                            \
        cfg.blockDim = dim3(tma_threads<bn>()); cfg.dynamicSmemBytes = tc_smem_bytes<bn>();                    \
        cudaFuncSetAttribute(tc_gemm_tma_kernel<bn, TC_FWD, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             tc_smem_bytes<bn>());                                                             \
        e = cudaOccupancyMaxActiveClusters(&n, tc_gemm_tma_kernel<bn, TC_FWD, 1, true>, &cfg);                 \
    }
    MPB_Q(64) MPB_Q(128) MPB_Q(256)
#undef MPB_Q
    return e == cudaSuccess ? n : -(int)e;
}

#ifdef MPB_TC_TRACE
int tc_gemm_set_trace(void* buf) {
    long long* b = (long long*)buf;
    return (int)cudaMemcpyToSymbol(g_tc_trace, &b, sizeof(b));
}
#endif
void tc_gemm_set_mode(int m) { g_tc_mode = m; if (m == 1 && !tma_api_ready()) g_tc_mode = 0; }

template <int BN, int OP>
static int launch_one(const TcGemmParams& p, dim3 grid, cudaStream_t s) {
    constexpr int smem = tc_smem_bytes<BN>();
    static bool attr_set = false;
    if (!attr_set) {
        MPB_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel<BN, OP>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    tc_gemm_kernel<BN, OP><<<grid, kTcThreads, smem, s>>>(p);
    MPB_LAUNCH_CHECK();
    return 0;
}

int tc_gemm_launch(const TcGemmParams& p, int BN, cudaStream_t s) {
    const int taps = p.kh * p.kw;
    if (p.M <= 0 || p.Cin <= 0 || p.Cout <= 0) return -1;
    if (p.ksplit < 1) return -1;
    dim3 grid;
    if (p.op == TC_FWD) {
        if (p.Cin % kTcBK || p.Cout % BN || p.ldo % 4 || p.ldx % 4 || p.ldw % 4) return -1;
        grid = dim3(ceil_div(p.M, kTcBM), p.Cout / BN, p.ksplit);
    } else if (p.op == TC_DGRAD) {
        if (p.Cout % kTcBK || p.Cin % BN || p.ldo % 4 || p.ldx % 4 || p.ldw % 4) return -1;
        grid = dim3(ceil_div(p.M, kTcBM), p.Cin / BN, p.ksplit);
    } else if (p.op == TC_WGRAD) {
        if (p.Cin % BN || p.Cout % 4 || p.ldy % 4 || p.ldx % 4 || p.ldw % 4) return -1;
        grid = dim3(ceil_div(p.Cout, kTcBM), taps * p.Cin / BN, p.ksplit);
    } else {
        return -1;
    }
    if (grid.y > 65535 || grid.z > 65535) return -1;
    // the epilogue moves 16-byte vectors (4 consecutive columns): every per-column / per-element operand
    // must be 16-byte aligned with a pitch that is a multiple of 4 floats
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p.out) || !al16(p.out_r) || !al16(p.res) || !al16(p.mask) || !al16(p.scale) || !al16(p.shift) ||
        !al16(p.scale2) || !al16(p.colsum))
        return -1;
    if ((p.res && p.ldr % 4) || (p.mask && p.ldm % 4) || (p.out_r && p.ldor % 4)) return -1;
    // split-K: with atomic=1 the slices RED.ADD into a zeroed output (any ksplit); with atomic=0 the slices of a
    // tile form one thread-block cluster and reduce through DSMEM (ksplit <= 8, every slice must be non-empty:
    // a CTA that left early would hang its cluster)
    if (p.ksplit > 1 && !p.atomic) {
        const int nkb_ = p.op == TC_FWD ? taps * (p.Cin / kTcBK) : p.op == TC_DGRAD ? taps * (p.Cout / kTcBK)
                                                                                    : ceil_div(p.M, kTcBK);
        const int per_ = ceil_div(nkb_, p.ksplit);
        if ((p.ksplit - 1) * per_ >= nkb_) return -1;
        if (p.ksplit > 8) return -1;
        {                        // the chunks a CTA does not own are parked in its idle pipeline stages
            const int n = BN / 32, ew = BN == 64 ? MPB_EW64 : BN == 128 ? MPB_EW128 : MPB_EW256;
            const int st = BN == 64 ? MPB_STAGES64 : BN == 128 ? MPB_STAGES128 : MPB_STAGES256;
            if ((n - n / p.ksplit) * 16384 + ew * 4096 > st * (kTcABytes + BN * 128)) return -1;
        }
        if (!(tc_gemm_mode() == 1 && (p.M % (p.H * p.W) == 0))) return -1;
    }
    // the TMA path needs whole images in the pixel grid (im2col walk) and 16-byte aligned pitches
    const bool tma = tc_gemm_mode() == 1 && (p.M % (p.H * p.W) == 0);
#define MPB_TC_CASE(bn)                                                       \
    if (BN == bn) {                                                           \
        if (tma) {                                                            \
            if (p.op == TC_FWD) return launch_tma<bn, TC_FWD>(p, grid, s);    \
            if (p.op == TC_DGRAD) return launch_tma<bn, TC_DGRAD>(p, grid, s);\
            return launch_tma<bn, TC_WGRAD>(p, grid, s);                      \
        }                                                                     \
        if (p.op == TC_FWD) return launch_one<bn, TC_FWD>(p, grid, s);        \
        if (p.op == TC_DGRAD) return launch_one<bn, TC_DGRAD>(p, grid, s);    \
        return launch_one<bn, TC_WGRAD>(p, grid, s);                          \
    }
    MPB_TC_CASE(64)
    MPB_TC_CASE(128)
    MPB_TC_CASE(256)
#undef MPB_TC_CASE
    return -1;
}

}  // namespace mpb
