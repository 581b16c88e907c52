// monopsr_b200/csrc/net_abi.cu -- C-ABI entry points of the network kernels (include/monopsr_b200_net.h).
#include "tc_gemm.cuh"

namespace mpb {
__global__ void tapmask_kernel(int total, int H, int W, int kh, int kw, int dil, unsigned short* out) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= total) return;
    int rem = m % (H * W), h = rem / W, w = rem % W;
    unsigned bits = 0;
    for (int t = 0; t < kh * kw; t++) {
        int hh = h + (t / kw - kh / 2) * dil, ww = w + (t % kw - kw / 2) * dil;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) bits |= 1u << t;
    }
    out[m] = (unsigned short)bits;
}
}  // namespace mpb

MPB_API int mpb_tc_gemm(const mpb_tc_gemm_params* p, int BN, void* stream) {
    if (!p) return -1;
    return mpb::tc_gemm_launch(*p, BN, (cudaStream_t)stream);
}

MPB_API int mpb_tc_gemm_x3(const mpb_tc_gemm_params* p, int BN, void* stream) {
    if (!p) return -1;
    return mpb::tc_gemm_x3_launch(*p, BN, (cudaStream_t)stream);
}

MPB_API int mpb_tc_gemm_h3(const mpb_tc_gemm_params* p, int BN, void* stream) {
    if (!p) return -1;
    return mpb::tc_gemm_h3_launch(*p, BN, (cudaStream_t)stream);
}

MPB_API int mpb_build_tapmask(int nimg, int H, int W, int kh, int kw, int dil, unsigned short* out,
                              void* stream) {
    if (nimg <= 0 || H <= 0 || W <= 0 || kh * kw > 16 || !out) return -1;
    int total = nimg * H * W;
    mpb::tapmask_kernel<<<mpb::ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(total, H, W, kh, kw, dil, out);
    MPB_LAUNCH_CHECK();
    return 0;
}

MPB_API int mpb_tc_set_producer(int mode) {
    mpb::tc_gemm_set_mode(mode);
    return mpb::tc_gemm_mode();
}

MPB_API int mpb_tc_set_cluster(int max_cluster) {
    mpb::tc_gemm_set_cluster(max_cluster);
    return max_cluster;
}

namespace mpb { int tc_gemm_max_clusters(int BN, int cx, int ks); }
MPB_API int mpb_tc_max_clusters(int BN, int cluster_x, int ksplit) {
    return mpb::tc_gemm_max_clusters(BN, cluster_x, ksplit);
}

#ifdef MPB_TC_TRACE
namespace mpb { int tc_gemm_set_trace(void* buf); }
MPB_API int mpb_tc_set_trace(void* buf) { return mpb::tc_gemm_set_trace(buf); }
#endif
