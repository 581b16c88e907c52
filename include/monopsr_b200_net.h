/*
 * include/monopsr_b200_net.h -- C ABI of the B200-native MonoPSR network kernels.
 *
 * The reference gets these computations from TensorFlow 1.8 (cuDNN/cuBLAS) through graph
 * ops; there is no FFI boundary in the reference for them, so each entry point cites the
 * reference graph code whose arithmetic it replaces.  All pointers are DEVICE pointers;
 * activations are NHWC fp32, weights are [Cout][kh][kw][Cin] fp32 ("OHWI", K-major for the
 * tensor-core forward pass).  Status: 0 ok, >0 cudaError_t, -1 invalid argument.
 */
#ifndef MONOPSR_B200_NET_H_
#define MONOPSR_B200_NET_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tcgen05 implicit-GEMM core (conv 1x1 / 3x3 atrous, stride 1, SAME; FC layers) ----
 * Replaces slim.conv2d / conv2d_same / slim.fully_connected and their gradients
 * (src/object_detection/nets/resnet_v1.py:116-127, nets/resnet_utils.py:111-113,
 *  src/monopsr/builders/net_builder.py:66,79-89, core/models/monopsr/monopsr_output_builder.py:166,186). */
enum { MPB_TC_FWD = 0, MPB_TC_DGRAD = 1, MPB_TC_WGRAD = 2 };

typedef struct mpb_tc_gemm_params {
    int op;              /* MPB_TC_* */
    int H, W;            /* pixel grid of one image (M = nimg*H*W); FC: H=W=1 */
    int kh, kw, dil;     /* filter taps (1x1 or 3x3) and atrous rate */
    int M;               /* pixels (GEMM rows) */
    int Cin, Cout;
    const float* X;      /* FWD: input; DGRAD: dY (gathered operand); WGRAD: input */
    int ldx;             /* floats per pixel of X */
    const float* Y;      /* WGRAD only: dY */
    int ldy;
    const float* Wt;     /* weights [Cout][ldw], K ordering (tap, ci) */
    int ldw;
    float* out;          /* FWD [M][ldo] (Cout cols); DGRAD [M][ldo] (Cin cols); WGRAD dW [Cout][ldw] */
    int ldo;
    const unsigned short* tapmask;   /* [M], bit t <=> tap t in bounds at that pixel; NULL for 1x1 */
    /* fused epilogue: v=acc; v*=scale[c]; v+=shift[c]; v+=res[r][c]; relu; v = mask[r][c]>0 ? v : 0;
     *                 v*=scale2[c]; round-to-tf32; colsum[c]+=v; store | atomic add */
    const float* scale;
    const float* shift;
    const float* res;
    int ldr;
    const float* mask;
    int ldm;
    const float* scale2;
    float* colsum;
    int relu;
    int round_tf32;
    int atomic;
    int ksplit;          /* >=1: split the K loop over gridDim.z (needs atomic=1 and a zeroed out) */
} mpb_tc_gemm_params;

/* BN: tile width in output columns (64, 128 or 256). */
int mpb_tc_gemm(const mpb_tc_gemm_params* p, int BN, void* stream);

/* tapmask[m] for an (nimg,H,W) pixel grid and a kh x kw filter with atrous rate dil. */
int mpb_build_tapmask(int nimg, int H, int W, int kh, int kw, int dil, unsigned short* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONOPSR_B200_NET_H_ */
