// monopsr_b200/csrc/tc_gemm.cuh -- tcgen05 (UMMA) implicit-GEMM core for sm_100a.
//
// One kernel family serves every dense contraction of the MonoPSR network
// (SURVEY.md Appendix A): stride-1 SAME convolutions (1x1, 3x3 atrous) on NHWC fp32
// activations, fully-connected layers (a 1x1 "conv" on a 1x1 grid), and their data-
// and weight-gradients.  fp32 operands are fed to `tcgen05.mma.kind::tf32` straight
// from shared memory (128B-swizzled canonical layouts), fp32 accumulators live in
// TMEM and are read back with `tcgen05.ld` for the fused epilogue.
//
//   FWD   out[p, co] = sum_{tap,ci} X[p + off(tap), ci] * Wt[co][tap][ci]
//         A = gathered pixels (K-major), B = weight rows (K-major)
//   DGRAD dX[p, ci]  = sum_{tap,co} dY[p - off(tap), co] * Wt[co][tap][ci]
//         A = gathered pixels of dY (K-major), B = weight rows read as MN-major
//   WGRAD dW[co][tap][ci] += sum_p dY[p, co] * X[p + off(tap), ci]
//         A = dY rows (MN-major), B = gathered pixels of X (MN-major); split-K over
//         pixels with vector RED.ADD into the (pre-zeroed) gradient buffer
//
// so no operand is ever transposed or im2col-materialised in HBM.  Operands are staged
// by 4 producer warps with 16-byte cp.async (zero-fill for padding taps / tails),
// completion is tracked on mbarriers (cp.async.mbarrier.arrive.noinc), one elected
// thread of warp 4 issues the MMAs and recycles stages with tcgen05.commit.
#pragma once
#include "common.cuh"
#include "../../include/monopsr_b200_net.h"

namespace mpb {

enum TcOp { TC_FWD = MPB_TC_FWD, TC_DGRAD = MPB_TC_DGRAD, TC_WGRAD = MPB_TC_WGRAD };
using TcGemmParams = mpb_tc_gemm_params;

constexpr int kTcBM = 128;
constexpr int kTcBK = 32;          // floats per k-block = one 128B swizzle row
// Pipeline depth / residency (measured on the whole training step, profiles/r1_notes.md): a CTA spends a
// third to a half of its life in prologue and epilogue with the tensor pipe idle, so the step is fastest when two
// CTAs (of the same or of different, concurrently running launches) share an SM and fill each other's gaps:
//   BN = 64 : 4 x 24 KB stages, 4 epilogue warps, 2 CTAs/SM
//   BN = 128: 3 x 32 KB stages, 4 epilogue warps, 2 CTAs/SM   (4 stages / 8 warps / 1 CTA per SM is 6 % slower per step)
//   BN = 256: 4 x 48 KB stages, 8 epilogue warps, 1 CTA/SM    (main loop runs at the tcgen05 tf32 rate: 512 clk/k-block)
#ifndef MPB_STAGES64
#define MPB_STAGES64 4
#endif
#ifndef MPB_STAGES128
#define MPB_STAGES128 3
#endif
#ifndef MPB_STAGES256
#define MPB_STAGES256 4
#endif
template <int BN> constexpr int tc_stages() { return BN == 64 ? MPB_STAGES64 : BN == 128 ? MPB_STAGES128 : MPB_STAGES256; }
constexpr int kTcThreads = 160;    // 4 producer/epilogue warps + 1 MMA warp
constexpr int kTcABytes = kTcBM * 128;

template <int BN>
constexpr int tc_smem_bytes() {
    return tc_stages<BN>() * (kTcABytes + BN * 128) + 1024 /*align slack*/ + 256 /*barriers*/;
}

int tc_gemm_launch(const TcGemmParams& p, int BN, cudaStream_t s);
int tc_gemm_x3_launch(const TcGemmParams& p, int BN, cudaStream_t s);   // 3xTF32 forward variant
int tc_gemm_h3_launch(const TcGemmParams& p, int BN, cudaStream_t s);   // fp16 hi/lo split forward variant
int tc_gemm_mode();            // 0 = cp.async producers, 1 = TMA producers (default)
void tc_gemm_set_mode(int m);
void tc_gemm_set_cluster(int c);  // max CTAs per cluster for A-tile multicast (1 = off, default)

}  // namespace mpb
