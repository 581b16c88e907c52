// monopsr_b200/csrc/heads.cu -- closed-form box heads, geometric projections and the losses with
// their hand-derived gradients (the small per-box / per-pixel part of the graph).
//
// Replaces (reference, TF graph code):
//   monopsr_output_builder.py:126-194,200-274  feature concatenation for the two FC stacks
//   :407-438 get_prop_cen_z/get_prop_cen_y ; instance_utils.py:907-953
//   :551-571,573-623 cen_x / cen_y / cen_z / centroids
//   monopsr_model.py:416-461 ; output_builder.py:663-746 ; instance_utils.py:567-681,738-788 ;
//   calib_utils.py:263-280   (local->global, projection error, global depth)
//   monopsr_model.py:554-958 ; losses_custom.py:93-132 ; object_detection/core/losses.py:118-157,283-317
// Everything here is O(num_boxes) or O(num_boxes*48*48): negligible FLOPs, fused to keep the
// launch count low.  Gradient derivations are in DESIGN.md ("heads").
#include "common.cuh"
#include "../../include/monopsr_b200_net.h"
#include <math.h>

namespace mpb {

constexpr int kMap = 48;
constexpr int kPix = kMap * kMap;
constexpr float kImgH = 320.f, kImgW = 1216.f;   // model_config.image_input_shape (yaml:51)
constexpr float kMaxDepth = 45.f;                // depth_range[1] (yaml:31)
constexpr float kCenYNorm = 1.666754f;           // output_builder.py:243
constexpr float kCarYOffset = 0.0648f;           // instance_utils.py:934

__device__ __forceinline__ float huber(float x) { float a = fabsf(x); return a <= 1.f ? 0.5f * x * x : a - 0.5f; }
__device__ __forceinline__ float dhuber(float x) { return fminf(fmaxf(x, -1.f), 1.f); }

__device__ float block_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
    return s;
}

// ---- once per sample: static concat columns and valid-mask statistics
__global__ void heads_static_kernel(mpb_heads_io io) {
    const int b = blockIdx.x;
    __shared__ float red[8];
    float cnt = 0.f;
    for (int p = threadIdx.x; p < kPix; p += blockDim.x) cnt += io.valid_mask[(size_t)b * kPix + p] != 0.f ? 1.f : 0.f;
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) {
        io.maskstats[b] = cnt;
        atomicAdd(&io.maskstats[io.nbox], cnt);
        const float* bx = io.boxes_2d + b * 4;
        const float cu = io.cam_p[2], cv = io.cam_p[6];
        float st[7];
        st[0] = (bx[0] - cv) / (kImgH / 2);
        st[1] = (bx[1] - cu) / (kImgW / 2);
        st[2] = (bx[2] - cv) / (kImgH / 2);
        st[3] = (bx[3] - cu) / (kImgW / 2);
        st[4] = (bx[2] - bx[0]) / kImgH;
        st[5] = io.est_view_angs[b];
        st[6] = io.class_indices[b] == 0 ? 1.f : 0.f;   // tf.one_hot(idx, num_classes=1) (quirk Q7)
        float* f1 = io.feat1 + (size_t)b * io.ld1 + 1024;
        float* f2 = io.feat2 + (size_t)b * io.ld2 + 1024;
        for (int i = 0; i < 7; i++) { f1[i] = st[i]; f2[i] = st[i]; }
        const float div[12] = {1000.f, 1.f, 1000.f, 100.f, 1.f, 1000.f, 1000.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        for (int i = 0; i < 12; i++) f1[7 + i] = io.cam_p[i] / div[i];
        for (int i = 1024 + 19; i < io.ld1; i++) io.feat1[(size_t)b * io.ld1 + i] = 0.f;   // K padding
        for (int i = 1024 + 36; i < io.ld2; i++) io.feat2[(size_t)b * io.ld2 + i] = 0.f;
    }
}

// ---- after the proposal heads: lwh, centroid proposals, dynamic tail of the regression concat
__global__ void heads_mid_kernel(mpb_heads_io io) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.nbox) return;
    const float* bx = io.boxes_2d + b * 4;
    const float f = io.cam_p[0], cv = io.cam_p[6];
    float lwh[3];
    for (int c = 0; c < 3; c++) {
        lwh[c] = io.mean_lwh[b * 3 + c] + io.lwh_offs[b * 3 + c];
        io.lwh[b * 3 + c] = lwh[c];
    }
    const float box_h = bx[2] - bx[0];
    const float pz = f * lwh[2] / box_h + io.prop_cen_z_offset[b];
    const float py = ((bx[2] + bx[0]) / 2.f - cv) * (pz / f) - kCarYOffset;
    io.prop_cen_z[b] = pz;
    io.prop_cen_y[b] = py;
    float* f2 = io.feat2 + (size_t)b * io.ld2 + 1024 + 7;
    for (int c = 0; c < 3; c++) f2[c] = io.lwh_offs[b * 3 + c];
    for (int c = 0; c < 24; c++) f2[3 + c] = io.alpha[b * 24 + c];
    f2[27] = py / kCenYNorm;
    f2[28] = pz / kMaxDepth;
}

// ---- after the regression heads: centroids; (train) pixel + box losses and their gradients.
// One CTA per box.  losses[]: 0 xyz_local, 1 lwh_offs, 2 alpha_bins, 3 alpha_regs, 4 cen_z_offs,
// 5 cen_y_offs, 6 proj_err, 7 depth_global, 8 total.
__global__ void __launch_bounds__(256) heads_final_kernel(mpb_heads_io io, int train) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int N = io.nbox;
    __shared__ float red[8];
    const float* P = io.cam_p;
    const float f = P[0], cu = P[2];
    const float* bx = io.boxes_2d + b * 4;
    const float ev = io.est_view_angs[b];
    const float pz = io.prop_cen_z[b], py = io.prop_cen_y[b];
    const float oy = io.cen_y_offs[b], oz = io.cen_z_offs[b];
    const float cen_y = py + oy, cen_z = pz + oz;
    const float x_off = -P[3] / P[0];
    if (tid == 0) {
        const float cen_x = cen_z * tanf(ev) + x_off;
        io.cen_y[b] = cen_y;
        io.cen_z[b] = cen_z;
        io.cen_x[b] = cen_x;
        io.centroids[b * 3] = cen_x;
        io.centroids[b * 3 + 1] = cen_y;
        io.centroids[b * 3 + 2] = cen_z;
    }
    if (!train) return;

    const float nv_b = fmaxf(io.maskstats[b], 1.f);     // tf.where(num_valid < 1, 1, num_valid)
    const float nv_all = io.maskstats[N];
    // SUM_BY_NONZERO over (N,48,48,3); a point-set loss for this output (xyz_loss_mode 1) is added by the caller
    const float inv_xyz = (nv_all > 0.f && io.xyz_loss_mode == 0) ? io.xyz_loss_weight / N / (3.f * nv_all) : 0.f;
    const float inv_dep = nv_all > 0.f ? 10.f / N / nv_all : 0.f;

    // geometry constants of this box
    const float gv = io.gt_view_angs[b];
    const float cg = cosf(gv), sg = sinf(gv), tg = tanf(gv);
    const float Xb = cen_z * tg + x_off, Yb = cen_y, Zb = cen_z;
    const float v1 = bx[0], u1 = bx[1], v2 = bx[2], u2 = bx[3];
    const float hu = (u2 - u1) / kMap / 2.f, hv = (v2 - v1) / kMap / 2.f;
    const float bw = u2 - u1, bh = v2 - v1;
    // global depth: depth = z_local + cen_z*(1 + k_row), k_row linear in row (quirk Q6: along rows)
    const float sp = (u2 - u1) / kMap / 2.f;
    const float va_l = atan2f((u1 + sp - cu) / f, 1.f), va_r = atan2f((u2 - sp - cu) / f, 1.f);
    const float kl = -tanf(va_l - ev) * tanf(ev), kr = -tanf(va_r - ev) * tanf(ev);

    const float* xl = io.xyz_local + (size_t)b * kPix * 3;
    const float* gl = io.gt_xyz_local + (size_t)b * kPix * 3;
    const float* gg = io.gt_xyz_global + (size_t)b * kPix * 3;
    const float* vm = io.valid_mask + (size_t)b * kPix;
    float* dxl = io.d_xyz_local + (size_t)b * kPix * 3;
    float* dg = io.depth_global + (size_t)b * kPix;

    // ---- pass 1: losses that are plain sums, projection-error sum
    float l_xyz = 0.f, l_dep = 0.f, pe_sum = 0.f;
    for (int p = tid; p < kPix; p += blockDim.x) {
        const int row = p / kMap, col = p % kMap;
        const float v = vm[p];
        const float lx = xl[p * 3], ly = xl[p * 3 + 1], lz = xl[p * 3 + 2];
        l_xyz += (huber(lx - gl[p * 3]) + huber(ly - gl[p * 3 + 1]) + huber(lz - gl[p * 3 + 2])) * v;
        const float t = row / 47.f;
        const float depth = lz + cen_z + cen_z * (kl + (kr - kl) * t);
        dg[p] = depth;
        l_dep += huber(depth - gg[p * 3 + 2]) * v;
        const float gx = cg * lx + sg * lz + Xb, gy = ly + Yb, gz = -sg * lx + cg * lz + Zb;
        const float pu = P[0] * gx + P[1] * gy + P[2] * gz + P[3];
        const float pv = P[4] * gx + P[5] * gy + P[6] * gz + P[7];
        const float pw = P[8] * gx + P[9] * gy + P[10] * gz + P[11];
        const float eu_exp = (u1 + hu) + ((u2 - hu) - (u1 + hu)) * (col / 47.f);
        const float ev_exp = (v1 + hv) + ((v2 - hv) - (v1 + hv)) * (row / 47.f);
        const float eu = (eu_exp - pu / pw) / bw * v, evv = (ev_exp - pv / pw) / bh * v;
        pe_sum += fminf(fmaxf(eu, -2.f), 2.f) + fminf(fmaxf(evv, -2.f), 2.f);
    }
    l_xyz = block_sum(l_xyz, red);
    l_dep = block_sum(l_dep, red);
    pe_sum = block_sum(pe_sum, red);
    const float pe = pe_sum / nv_b;
    const float dpe = 0.1f / N * dhuber(pe) / nv_b;     // d total / d (clipped error sum entry)

    // ---- pass 2: gradients
    float d_ceny = 0.f, d_cenz = 0.f;
    for (int p = tid; p < kPix; p += blockDim.x) {
        const int row = p / kMap, col = p % kMap;
        const float v = vm[p];
        const float lx = xl[p * 3], ly = xl[p * 3 + 1], lz = xl[p * 3 + 2];
        float dlx = inv_xyz * dhuber(lx - gl[p * 3]) * v;
        float dly = inv_xyz * dhuber(ly - gl[p * 3 + 1]) * v;
        float dlz = inv_xyz * dhuber(lz - gl[p * 3 + 2]) * v;
        const float t = row / 47.f;
        const float kk = 1.f + kl + (kr - kl) * t;
        const float ddep = inv_dep * dhuber(dg[p] - gg[p * 3 + 2]) * v;
        dlz += ddep;
        d_cenz += ddep * kk;
        const float gx = cg * lx + sg * lz + Xb, gy = ly + Yb, gz = -sg * lx + cg * lz + Zb;
        const float pu = P[0] * gx + P[1] * gy + P[2] * gz + P[3];
        const float pv = P[4] * gx + P[5] * gy + P[6] * gz + P[7];
        const float pw = P[8] * gx + P[9] * gy + P[10] * gz + P[11];
        const float eu_exp = (u1 + hu) + ((u2 - hu) - (u1 + hu)) * (col / 47.f);
        const float ev_exp = (v1 + hv) + ((v2 - hv) - (v1 + hv)) * (row / 47.f);
        const float eu = (eu_exp - pu / pw) / bw * v, evv = (ev_exp - pv / pw) / bh * v;
        // tf.clip_by_value passes the gradient inside [-2,2] (inclusive)
        const float deu = (eu >= -2.f && eu <= 2.f) ? dpe : 0.f;
        const float dev = (evv >= -2.f && evv <= 2.f) ? dpe : 0.f;
        const float dproj_u = -deu * v / bw, dproj_v = -dev * v / bh;
        const float dpu = dproj_u / pw, dpv = dproj_v / pw;
        const float dpw = -(dproj_u * pu + dproj_v * pv) / (pw * pw);
        const float dgx = P[0] * dpu + P[4] * dpv + P[8] * dpw;
        const float dgy = P[1] * dpu + P[5] * dpv + P[9] * dpw;
        const float dgz = P[2] * dpu + P[6] * dpv + P[10] * dpw;
        dlx += cg * dgx - sg * dgz;
        dly += dgy;
        dlz += sg * dgx + cg * dgz;
        d_cenz += dgx * tg + dgz;
        d_ceny += dgy;
        dxl[p * 3] = dlx;
        dxl[p * 3 + 1] = dly;
        dxl[p * 3 + 2] = dlz;
    }
    d_ceny = block_sum(d_ceny, red);
    d_cenz = block_sum(d_cenz, red);

    if (tid == 0) {
        io.proj_err_norm[b] = pe;
        // ---- box losses (monopsr_model.py:554-958, weights yaml:102-118)
        const float* b3 = io.boxes_3d + b * 7;
        float L[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        L[0] = l_xyz * inv_xyz;
        L[7] = l_dep * inv_dep;
        L[6] = 0.1f / N * huber(pe);
        // lwh: gt offsets are gt_lwh - PRED_lwh (output_builder.py:655-660) => residual 2*offs-(gt-mean) (quirk Q9)
        for (int c = 0; c < 3; c++) {
            const float o = io.lwh_offs[b * 3 + c];
            const float r = 2.f * o - (b3[3 + c] - io.mean_lwh[b * 3 + c]);
            L[1] += huber(r) / N;
            io.d_lwh_offs[b * 3 + c] = 2.f * dhuber(r) / N;
        }
        // alpha bins: softmax CE, label smoothing 0.001 (on 1-eps, off eps/12); TF's op returns
        // softmax-labels as the logits gradient (valid-distribution assumption), weight 0.3
        {
            const float* lg = io.alpha + b * 24;
            float mx = lg[0];
            for (int k = 1; k < 12; k++) mx = fmaxf(mx, lg[k]);
            float se = 0.f;
            for (int k = 0; k < 12; k++) se += expf(lg[k] - mx);
            const float lse = mx + logf(se);
            const int gtb = io.gt_alpha_bins[b];
            float ce = 0.f;
            for (int k = 0; k < 12; k++) {
                const float tgt = (k == gtb) ? 1.f - 0.001f : 0.001f / 12.f;
                const float logp = lg[k] - lse;
                ce -= tgt * logp;
                io.d_alpha[b * 24 + k] = 0.3f / N * (expf(logp) - tgt);
            }
            L[2] = 0.3f / N * ce;
        }
        for (int k = 0; k < 12; k++) {
            const float r = io.alpha[b * 24 + 12 + k] - io.gt_alpha_regs[b * 12 + k];
            const float w = io.gt_alpha_valid_bins[b * 12 + k];
            L[3] += huber(r) * w / N;
            io.d_alpha[b * 24 + 12 + k] = dhuber(r) * w / N;
        }
        const float gt_cen_y = b3[1] - b3[5] / 2.f;      // centroid_type 'middle' (monopsr_model.py:266-270)
        const float gt_cen_z = b3[2];
        const float rz = oz - (gt_cen_z - pz), ry = oy - (gt_cen_y - py);
        L[4] = 0.1f / N * huber(rz);
        L[5] = 0.1f / N * huber(ry);
        const float dz_l = 0.1f / N * dhuber(rz), dy_l = 0.1f / N * dhuber(ry);
        // cen_y = prop_y + oy and cen_z = prop_z + oz feed the pixel losses
        io.d_cen_y_offs[b] = dy_l + d_ceny;
        io.d_cen_z_offs[b] = dz_l + d_cenz;
        io.d_prop_y[b] = dy_l + d_ceny;      // (+ the concat-tail term, added in heads_bwd_mid)
        io.d_prop_z[b] = dz_l + d_cenz;
        float tot = 0.f;
        for (int i = 0; i < 8; i++) {
            atomicAdd(&io.losses[i], L[i]);
            tot += L[i];
        }
        atomicAdd(&io.losses[8], tot);
    }
}

// ---- after the regression fc0 data-gradient: fold the concat-tail gradient into the heads
__global__ void heads_bwd_mid_kernel(mpb_heads_io io) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.nbox) return;
    const float* t = io.d_feat2 + (size_t)b * io.ldd2 + 1024 + 7;
    const float* bx = io.boxes_2d + b * 4;
    const float f = io.cam_p[0], cv = io.cam_p[6];
    float dpy = io.d_prop_y[b] + t[27] / kCenYNorm;
    float dpz = io.d_prop_z[b] + t[28] / kMaxDepth;
    dpz += dpy * ((bx[2] + bx[0]) / 2.f - cv) / f;        // prop_y = box_cv * prop_z / f - c
    for (int c = 0; c < 3; c++) io.d_lwh_offs[b * 3 + c] += t[c];
    io.d_lwh_offs[b * 3 + 2] += dpz * f / (bx[2] - bx[0]); // prop_z = f*lwh_h/box_h + offset
    for (int c = 0; c < 24; c++) io.d_alpha[b * 24 + c] += t[3 + c];
}

// ---- glue of the point-set training losses (see include/monopsr_b200_net.h)
__global__ void pointset_mask_kernel(long npts, const float* __restrict__ pred, const float* __restrict__ gt,
                                     const float* __restrict__ mask, float* __restrict__ p, float* __restrict__ t) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npts * 3; i += (long)gridDim.x * blockDim.x) {
        const float m = mask[i / 3];
        p[i] = pred[i] * m;
        t[i] = gt[i] * m;
    }
}
// ONE CTA: fixed summation order (deterministic loss value), then two plain adds by one thread
__global__ void __launch_bounds__(1024)
pointset_loss_add_kernel(long na, const float* __restrict__ a, long nb, const float* __restrict__ b, float scale,
                         float* __restrict__ losses, int slot, int total_slot, float* __restrict__ fill, long nfill,
                         float fill_value) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long i = threadIdx.x; i < na; i += blockDim.x) acc += (double)a[i];
    if (b) for (long i = threadIdx.x; i < nb; i += blockDim.x) acc += (double)b[i];
    if (fill) for (long i = threadIdx.x; i < nfill; i += blockDim.x) fill[i] = fill_value;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        const float v = (float)(s * (double)scale);
        losses[slot] += v;
        losses[total_slot] += v;
    }
}
__global__ void pointset_grad_add_kernel(long npts, const float* __restrict__ grad, const float* __restrict__ mask,
                                         float scale, float* __restrict__ d_pred) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npts * 3; i += (long)gridDim.x * blockDim.x)
        d_pred[i] += scale * grad[i] * mask[i / 3];
}

}  // namespace mpb

using namespace mpb;
MPB_API int mpb_pointset_mask(long npts, const float* pred, const float* gt, const float* mask, float* p, float* t,
                              void* stream) {
    if (npts < 0 || !pred || !gt || !mask || !p || !t) return -1;
    if (npts == 0) return 0;
    const long blocks = (npts * 3 + 255) / 256;
    pointset_mask_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, (cudaStream_t)stream>>>(npts, pred, gt, mask, p, t);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_pointset_loss_add(long na, const float* a, long nb, const float* b, float scale, float* losses, int slot,
                                  int total_slot, float* fill, long nfill, float fill_value, void* stream) {
    if (na < 0 || nb < 0 || !a || !losses || slot < 0 || total_slot < 0) return -1;
    pointset_loss_add_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(na, a, b ? nb : 0, b, scale, losses, slot, total_slot, fill,
                                                                  nfill, fill_value);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_pointset_grad_add(long npts, const float* grad, const float* mask, float scale, float* d_pred, void* stream) {
    if (npts < 0 || !grad || !mask || !d_pred) return -1;
    if (npts == 0) return 0;
    const long blocks = (npts * 3 + 255) / 256;
    pointset_grad_add_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, (cudaStream_t)stream>>>(npts, grad, mask, scale,
                                                                                                       d_pred);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_heads_static(const mpb_heads_io* io, void* stream) {
    if (!io || io->nbox <= 0) return -1;
    MPB_CUDA_TRY(cudaMemsetAsync(io->maskstats + io->nbox, 0, sizeof(float), (cudaStream_t)stream));
    heads_static_kernel<<<io->nbox, 256, 0, (cudaStream_t)stream>>>(*io);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_heads_mid(const mpb_heads_io* io, void* stream) {
    if (!io || io->nbox <= 0) return -1;
    heads_mid_kernel<<<ceil_div(io->nbox, 64), 64, 0, (cudaStream_t)stream>>>(*io);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_heads_final(const mpb_heads_io* io, int train, void* stream) {
    if (!io || io->nbox <= 0) return -1;
    if (train) MPB_CUDA_TRY(cudaMemsetAsync(io->losses, 0, sizeof(float) * 9, (cudaStream_t)stream));
    heads_final_kernel<<<io->nbox, 256, 0, (cudaStream_t)stream>>>(*io, train);
    MPB_LAUNCH_CHECK();
    return 0;
}
MPB_API int mpb_heads_bwd_mid(const mpb_heads_io* io, void* stream) {
    if (!io || io->nbox <= 0) return -1;
    heads_bwd_mid_kernel<<<ceil_div(io->nbox, 64), 64, 0, (cudaStream_t)stream>>>(*io);
    MPB_LAUNCH_CHECK();
    return 0;
}
