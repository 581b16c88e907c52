// monopsr_b200/csrc/targets.cu -- ground-truth target synthesis (the step immediately before
// the network path).  One launch replaces the 2 x num_boxes per-box TF sub-graphs of MonoPSRModel.build
// (core/models/monopsr/monopsr_model.py:165-203):
//   tf_instance_xyz_crop_from_depth_map  (datasets/kitti/instance_utils.py:395-481)
//   tf_depth_patch_to_pc_map             (datasets/kitti/depth_map_utils.py:161-236)
//   tf_get_tr_mat                        (core/transform_utils.py:36-66)
// thread = (box, roi row, roi column): mask + crop the depth map by the ROUNDED box, nearest-neighbour resize with
// align_corners (TF 1.8 kernel: in = min(roundf(out * (in-1)/(out-1)), in-1)), back-project through the pixel CENTRES
// of the UNROUNDED box, valid = |depth| >= 0.1, then the view-normalised (local) and the camera-frame (global) map.
// Float operation order follows the reference expression by expression (fp32, no FMA contraction where the reference
// has separate TF ops) so that the CPU restatement (oracle/targets.py) is matched bit for bit.
#include "common.cuh"
#include "../../include/monopsr_b200_net.h"

namespace mpb {

__global__ void __launch_bounds__(256)
gt_xyz_from_depth_kernel(int nbox, int H, int W, int roi, const float* __restrict__ depth,
                         const unsigned char* __restrict__ masks, const float* __restrict__ boxes_2d,
                         const float* __restrict__ boxes_3d, int ld3, const float* __restrict__ view_angs,
                         const float* __restrict__ cam_p, int centroid_middle, int rotate_view,
                         float* __restrict__ xyz_local, float* __restrict__ xyz_global, float* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbox * roi * roi) return;
    const int c = i % roi, r = (i / roi) % roi, b = i / (roi * roi);
    const float y1 = boxes_2d[b * 4], x1 = boxes_2d[b * 4 + 1], y2 = boxes_2d[b * 4 + 2], x2 = boxes_2d[b * 4 + 3];
    // tf.to_int32(tf.round(box)): round half to even
    const int r0 = __float2int_rn(y1), c0 = __float2int_rn(x1), r1 = __float2int_rn(y2), c1 = __float2int_rn(x2);
    const int ch = r1 - r0, cw = c1 - c0;
    float d = 0.f;
    if (ch > 0 && cw > 0) {
        const float sh = roi > 1 ? __fdiv_rn((float)(ch - 1), (float)(roi - 1)) : 0.f;
        const float sw = roi > 1 ? __fdiv_rn((float)(cw - 1), (float)(roi - 1)) : 0.f;
        const int sr = min((int)roundf(__fmul_rn((float)r, sh)), ch - 1), sc = min((int)roundf(__fmul_rn((float)c, sw)), cw - 1);
        const int yy = r0 + sr, xx = c0 + sc;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
            d = __fmul_rn(depth[(size_t)yy * W + xx], masks[((size_t)b * H + yy) * W + xx] ? 1.f : 0.f);
    }
    // pixel-centre grid of the unrounded box: tf.linspace(x1 + hw, x2 - hw, n)[c] = start + c * ((stop - start)/(n-1))
    const float pw = __fdiv_rn(__fsub_rn(x2, x1), (float)roi), ph = __fdiv_rn(__fsub_rn(y2, y1), (float)roi);
    const float hw = __fdiv_rn(pw, 2.f), hh = __fdiv_rn(ph, 2.f);
    const float xs = __fadd_rn(x1, hw), xe = __fsub_rn(x2, hw), ys = __fadd_rn(y1, hh), ye = __fsub_rn(y2, hh);
    const float xstep = roi > 1 ? __fdiv_rn(__fsub_rn(xe, xs), (float)(roi - 1)) : 0.f;
    const float ystep = roi > 1 ? __fdiv_rn(__fsub_rn(ye, ys), (float)(roi - 1)) : 0.f;
    const float gx = __fadd_rn(xs, __fmul_rn((float)c, xstep)), gy = __fadd_rn(ys, __fmul_rn((float)r, ystep));
    const float f = cam_p[0], cu = cam_p[2], cv = cam_p[6];
    const float ratio = __fdiv_rn(d, f);
    const float px = __fmul_rn(__fsub_rn(gx, cu), ratio), py = __fmul_rn(__fsub_rn(gy, cv), ratio), pz = d;
    const float v = fabsf(d) >= 0.1f ? 1.f : 0.f;
    // view normalisation: tr = rot_y(-va) @ translate(-centroid), applied as a 4x4 matmul row by row
    const float x_off = __fdiv_rn(-cam_p[3], cam_p[0]);
    float cx = __fsub_rn(boxes_3d[b * ld3], x_off), cy = boxes_3d[b * ld3 + 1], cz = boxes_3d[b * ld3 + 2];
    if (centroid_middle) cy = __fsub_rn(cy, __fdiv_rn(boxes_3d[b * ld3 + 5], 2.f));
    float lx, ly, lz;
    if (rotate_view) {
        const float a = -view_angs[b];
        const float co = cosf(a), si = sinf(a);
        // (rot @ t_mat) rows: [co, 0, si, co*(-cx) + si*(-cz)], [0, 1, 0, -cy], [-si, 0, co, -si*(-cx) + co*(-cz)]
        const float t0 = __fadd_rn(__fmul_rn(co, -cx), __fmul_rn(si, -cz));
        const float t2 = __fadd_rn(__fmul_rn(-si, -cx), __fmul_rn(co, -cz));
        lx = __fadd_rn(__fadd_rn(__fmul_rn(co, px), __fmul_rn(si, pz)), t0);
        ly = __fadd_rn(py, -cy);
        lz = __fadd_rn(__fadd_rn(__fmul_rn(-si, px), __fmul_rn(co, pz)), t2);
    } else {
        lx = __fsub_rn(px, cx); ly = __fsub_rn(py, cy); lz = __fsub_rn(pz, cz);
    }
    float* ol = xyz_local + (size_t)i * 3;
    float* og = xyz_global + (size_t)i * 3;
    ol[0] = lx * v; ol[1] = ly * v; ol[2] = lz * v;
    og[0] = px * v; og[1] = py * v; og[2] = pz * v;
    valid[i] = v;
}

}  // namespace mpb

MPB_API int mpb_gt_xyz_from_depth(int nbox, int H, int W, int roi, const float* depth, const unsigned char* masks,
                                  const float* boxes_2d, const float* boxes_3d, int ld3, const float* view_angs,
                                  const float* cam_p, int centroid_middle, int rotate_view, float* xyz_local,
                                  float* xyz_global, float* valid, void* stream) {
    if (nbox <= 0 || H <= 0 || W <= 0 || roi <= 0 || ld3 < 6 || !depth || !masks || !boxes_2d || !boxes_3d ||
        !view_angs || !cam_p || !xyz_local || !xyz_global || !valid)
        return -1;
    const int total = nbox * roi * roi;
    mpb::gt_xyz_from_depth_kernel<<<mpb::ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        nbox, H, W, roi, depth, masks, boxes_2d, boxes_3d, ld3, view_angs, cam_p, centroid_middle, rotate_view,
        xyz_local, xyz_global, valid);
    MPB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Network inputs from the raw camera image (second half):
//   ImgPreprocessor.preprocess_input (core/img_preprocessor.py:12-35): float(img) - channel means, then
//       tf.image.resize_images(bilinear, align_corners=False: TF1 legacy mapping in = out * in_size/out_size)
//   rgb crops: tf.image.crop_and_resize(img_preprocessed, boxes_2d_norm, 0, img_roi_size)   (monopsr_model.py:222-226)
//   full image: tf.image.resize_bilinear(img_preprocessed, resized_full_img_shape, align_corners=True)   (:228-233)
// Three launches (the 320 x 1216 preprocessed image is materialised once: 4.7 MB, both consumers read it).
namespace mpb {

__device__ __forceinline__ float lerp2(float tl, float tr, float bl, float br, float lx, float ly) {
    const float t = tl + (tr - tl) * lx, b = bl + (br - bl) * lx;
    return t + (b - t) * ly;
}

template <typename TIn>
__global__ void __launch_bounds__(256)
image_preprocess_kernel(int H, int W, int OH, int OW, const TIn* __restrict__ img, float m0, float m1, float m2,
                        float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OH * OW) return;
    const int ow = i % OW, oh = i / OW;
    const float sy = oh * ((float)H / (float)OH), sx = ow * ((float)W / (float)OW);     // legacy (no half-pixel) mapping
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float mean[3] = {m0, m1, m2};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float tl = (float)img[((size_t)y0 * W + x0) * 3 + c] - mean[c], tr = (float)img[((size_t)y0 * W + x1) * 3 + c] - mean[c];
        const float bl = (float)img[((size_t)y1 * W + x0) * 3 + c] - mean[c], br = (float)img[((size_t)y1 * W + x1) * 3 + c] - mean[c];
        out[(size_t)i * 3 + c] = lerp2(tl, tr, bl, br, lx, ly);
    }
}

// tf.image.crop_and_resize, bilinear, extrapolation value 0, one source image (box_ind = 0), C = 3
__global__ void __launch_bounds__(256)
image_crops_kernel(int H, int W, const float* __restrict__ img, int nbox, const float* __restrict__ boxes_norm, int crop,
                   float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbox * crop * crop) return;
    const int cx = i % crop, cy = (i / crop) % crop, b = i / (crop * crop);
    const float y1 = boxes_norm[b * 4], x1 = boxes_norm[b * 4 + 1], y2 = boxes_norm[b * 4 + 2], x2 = boxes_norm[b * 4 + 3];
    // TF's expressions, op by op and without FMA contraction: whether the last row / column of a box that ends exactly at
    // 1.0 is inside the image (in <= size-1) depends on the last bit of this sum
    const float hs = crop > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(y2, y1), (float)(H - 1)), (float)(crop - 1)) : 0.f;
    const float ws = crop > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(x2, x1), (float)(W - 1)), (float)(crop - 1)) : 0.f;
    const float in_y = crop > 1 ? __fadd_rn(__fmul_rn(y1, (float)(H - 1)), __fmul_rn((float)cy, hs))
                                : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(y1, y2)), (float)(H - 1));
    const float in_x = crop > 1 ? __fadd_rn(__fmul_rn(x1, (float)(W - 1)), __fmul_rn((float)cx, ws))
                                : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x1, x2)), (float)(W - 1));
    float* o = out + (size_t)i * 3;
    if (in_y < 0 || in_y > H - 1 || in_x < 0 || in_x > W - 1) { o[0] = o[1] = o[2] = 0.f; return; }
    const int t = (int)floorf(in_y), bo = (int)ceilf(in_y), l = (int)floorf(in_x), r = (int)ceilf(in_x);
    const float ly = in_y - t, lx = in_x - l;
#pragma unroll
    for (int c = 0; c < 3; c++)
        o[c] = lerp2(img[((size_t)t * W + l) * 3 + c], img[((size_t)t * W + r) * 3 + c], img[((size_t)bo * W + l) * 3 + c],
                     img[((size_t)bo * W + r) * 3 + c], lx, ly);
}

// tf.image.resize_bilinear(align_corners=True), C = 3
__global__ void __launch_bounds__(256)
image_resize_ac_kernel(int H, int W, const float* __restrict__ img, int OH, int OW, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OH * OW) return;
    const int ow = i % OW, oh = i / OW;
    const float sy = OH > 1 ? oh * ((float)(H - 1) / (float)(OH - 1)) : 0.f, sx = OW > 1 ? ow * ((float)(W - 1) / (float)(OW - 1)) : 0.f;
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
#pragma unroll
    for (int c = 0; c < 3; c++)
        out[(size_t)i * 3 + c] = lerp2(img[((size_t)y0 * W + x0) * 3 + c], img[((size_t)y0 * W + x1) * 3 + c],
                                       img[((size_t)y1 * W + x0) * 3 + c], img[((size_t)y1 * W + x1) * 3 + c], lx, ly);
}

}  // namespace mpb

MPB_API int mpb_image_inputs(int H, int W, const void* img, int img_is_u8, const float* channel_means, int PH, int PW,
                             float* preprocessed, int nbox, const float* boxes_norm, int crop, float* rgb_crops, int FH,
                             int FW, float* full_img, void* stream) {
    if (H < 2 || W < 2 || PH < 1 || PW < 1 || !img || !channel_means || !preprocessed) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const float m0 = channel_means[0], m1 = channel_means[1], m2 = channel_means[2];   // HOST pointer: 3 constants
    if (img_is_u8)
        mpb::image_preprocess_kernel<unsigned char><<<mpb::ceil_div(PH * PW, 256), 256, 0, s>>>(
            H, W, PH, PW, (const unsigned char*)img, m0, m1, m2, preprocessed);
    else
        mpb::image_preprocess_kernel<float><<<mpb::ceil_div(PH * PW, 256), 256, 0, s>>>(H, W, PH, PW, (const float*)img, m0,
                                                                                       m1, m2, preprocessed);
    MPB_LAUNCH_CHECK();
    if (rgb_crops) {
        if (nbox <= 0 || crop <= 0 || !boxes_norm) return -1;
        mpb::image_crops_kernel<<<mpb::ceil_div(nbox * crop * crop, 256), 256, 0, s>>>(PH, PW, preprocessed, nbox, boxes_norm,
                                                                                     crop, rgb_crops);
        MPB_LAUNCH_CHECK();
    }
    if (full_img) {
        if (FH <= 0 || FW <= 0) return -1;
        mpb::image_resize_ac_kernel<<<mpb::ceil_div(FH * FW, 256), 256, 0, s>>>(PH, PW, preprocessed, FH, FW, full_img);
        MPB_LAUNCH_CHECK();
    }
    return 0;
}
