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
    /* fused epilogue: v=acc; v*=rowscale[r]; v*=scale[c]; v+=shift[c]; v+=res[r][c]; relu; v = mask[r][c]>0 ? v : 0;
     *                 v*=scale2[c]; round-to-tf32; colsum[c]+=v; store | atomic add */
    const float* scale;
    const float* shift;
    const float* res;
    int ldr;
    const float* mask;
    int ldm;
    const float* scale2;
    const float* rowscale;  /* per output ROW factor (WGRAD: folded-BN scale of each output channel) */
    float* colsum;
    int relu;
    int round_tf32;
    int atomic;
    int ksplit;          /* >=1: split the K loop over gridDim.z.  atomic=1: slices RED.ADD into a zeroed `out` (no
                          * other epilogue option makes sense then).  atomic=0: the slices of a tile form one
                          * thread-block cluster (ksplit <= 8), reduce through distributed shared memory and the full
                          * fused epilogue still runs once per tile */
    float* out_r;        /* optional second output [rows][ldor]: the tf32-rounded value (GEMM operand of the next
                          * layer) while `out` keeps the unrounded one (fp32 residual stream) */
    int ldor;
    /* ---- fp16 hi/lo split forward ("h3", mpb_tc_gemm_h3; ignored by the other entry points) ----
     * A "split copy" of an fp32 matrix has the SAME byte geometry (4 bytes per element, same pitch): every group of
     * 32 consecutive elements of a row (128 bytes) holds 32 fp16 `hi` = fp16(x) followed by 32 fp16 `lo` =
     * fp16(x - hi) -- for the weights the other way round, `lo` first -- so that one 128-byte shared-memory row is
     * at once the K=64 operand of the two cross products and, half of it, the K=32 operand of hi*hi. */
    const void* X16;     /* split copy of X ([hi | lo] blocks, pitch ldx floats) */
    const void* W16;     /* split copy of Wt ([lo | hi] blocks, pitch ldw floats), each output channel optionally
                          * pre-scaled by a power of two (undone by `scale`) */
    void* out16;         /* optional third output: the stored value ALSO as a split copy ([hi | lo], pitch ldo16 floats) */
    int ldo16;
    int* overflow;       /* optional sticky flag, set to 1 when a value beyond the fp16 range (|v| > 65504) was split */
} mpb_tc_gemm_params;

/* BN: tile width in output columns (64, 128 or 256). */
int mpb_tc_gemm(const mpb_tc_gemm_params* p, int BN, void* stream);

/* 3xTF32 forward variant of mpb_tc_gemm (op must be MPB_TC_FWD, BN 64 or 128, split-K only with atomic=1): the
 * operands are read as stored -- UNROUNDED fp32 -- split on chip into tf32 hi + lo parts and contracted with three
 * MMAs per k-step, for fp32-level accuracy (the reference's convolutions / matmuls are fp32: net_builder.py:44-89,
 * monopsr_output_builder.py:166-283) at the memory traffic of the single-pass kernel.  Same epilogue options. */
int mpb_tc_gemm_x3(const mpb_tc_gemm_params* p, int BN, void* stream);
/* fp16-split forward variant ("h3"): op must be MPB_TC_FWD, operands are the split copies X16 / W16 (see the struct).
 *   acc = A_hi*B_lo + A_lo*B_hi  (one K=64 kind::f16 block)  +  A_hi*B_hi  (K=32)   -- fp32 accumulation in TMEM;
 * the dropped lo*lo term is 2^-22 relative.  Accuracy is that of the 3xTF32 kernel (fp32-level, the reference's
 * convolutions / matmuls are fp32: net_builder.py:44-89, monopsr_output_builder.py:166-283) for 1.5x instead of
 * 3x the tensor work of a single tf32 pass, with the shared-memory / TMA traffic of the single pass.
 * BN 64, 128 or 256; split-K as in mpb_tc_gemm.  Same epilogue options plus `out16`. */
int mpb_tc_gemm_h3(const mpb_tc_gemm_params* p, int BN, void* stream);
/* split copy of a row-major fp32 matrix: rows x C (C % 32 == 0), pitches in floats; b_operand = 1 writes [lo | hi]
 * blocks (weights), 0 writes [hi | lo] (activations).  overflow: optional sticky flag (see the struct). */
int mpb_split16(long rows, int C, const float* src, int lds, void* dst16, int ldd, int b_operand, int* overflow,
                void* stream);
/* whole-model weight preparation for the h3 kernel: one warp per (layer, output channel) row computes
 * v = w * s (s = gamma * rsqrt(var + eps) for a frozen-BN conv, 1 otherwise), scales the row by the power of two
 * that puts its largest |v| into [2^13, 2^14) -- so that the lo halves stay in fp16's normal range -- writes the
 * [lo | hi] split copy and 1 / scale to inv_scale[row]. */
typedef struct mpb_w16_layer {
    const float* w; const float* gamma; const float* var;     /* gamma / var NULL: no BN fold */
    void* w16; float* inv_scale;
    int cout; int K; int row0; int pad_;
    /* optional: the same pass also does the work of mpb_fold_bn_multi for this layer -- wf = w * s (fp32, the
     * backward pass's operand), scale[co] = s, shift[co] = beta[co] - mean[co] * s -- so the weights are read once */
    float* wf; float* scale; float* shift; const float* beta; const float* mean;
} mpb_w16_layer;
int mpb_split16_weights_multi(int total_rows, const mpb_w16_layer* layers, const int* row2layer, float eps, void* stream);
/* the same over a LIST of global rows (DEVICE array; NULL = rows 0..nrows-1) none longer than max_row_floats: rows up to
 * 256 / 1024 / 2304 floats are kept in registers between the two passes (the weights are read once, and short rows do
 * not pay the register footprint of long ones) */
int mpb_split16_weights_rows(int nrows, const int* rowlist, int max_row_floats, const mpb_w16_layer* layers,
                             const int* row2layer, float eps, void* stream);

/* 1 (default): the elementwise / weight-preparation kernels round every GEMM operand they produce to tf32;
 * 0: they leave it as computed (what mpb_tc_gemm_x3 wants).  Synchronous; set it outside graph capture. */
int mpb_set_operand_rounding(int on);

/* operand staging of the GEMM core: 1 = TMA (cp.async.bulk.tensor, default), 0 = cp.async (LSU) producers.
 * Returns the mode in effect (TMA falls back to 0 when the driver lacks cuTensorMapEncode*). */
int mpb_tc_set_producer(int mode);
/* max CTAs per thread-block cluster sharing (TMA-multicasting) one A tile: 1 (default, off), 2 or 4. */
int mpb_tc_set_cluster(int max_cluster);

/* diagnostics: how many (cluster_x, 1, ksplit)-CTA clusters of the cluster split-K kernel (tile width BN) can be
 * resident at once (cudaOccupancyMaxActiveClusters); < 0: -cudaError_t */
int mpb_tc_max_clusters(int BN, int cluster_x, int ksplit);

/* tapmask[m] for an (nimg,H,W) pixel grid and a kh x kw filter with atrous rate dil. */
int mpb_build_tapmask(int nimg, int H, int W, int kh, int kw, int dil, unsigned short* out, void* stream);

/* ---- weight preparation -------------------------------------------------------------------
 * Frozen (inference-mode) batch norm of the two ResNet towers (feature_extractor.py:228-242:
 * is_training=False, eps=1e-5, scale=True) folded into the conv weights:
 *   wf = tf32(w * s), s = gamma*rsqrt(var+eps), shift = beta - mean*s.   w is [cout][K]. */
int mpb_fold_bn(int cout, int K, const float* w, const float* gamma, const float* beta, const float* mean,
                const float* var, float eps, float* wf, float* scale, float* shift, void* stream);
int mpb_round_copy(long n, const float* src, float* dst, void* stream);   /* dst = tf32(src) */
/* d(gamma) of a frozen BN from the conv weight gradient: rowdot(w,dw)/gamma - mean*dbeta*rsqrt(var+eps) */
int mpb_bn_param_grad(int cout, int K, const float* w, const float* dw, const float* gamma, const float* mean,
                      const float* var, float eps, const float* dbeta, float* dgamma, void* stream);

/* whole-model variants of the two calls above: one launch over all (layer, output channel) rows.
 * layers / row2layer are DEVICE arrays; row2layer[r] = layer index of global row r, layers[i].row0 = first row. */
typedef struct mpb_bn_layer {
    const float* w; const float* gamma; const float* beta; const float* mean; const float* var;
    float* wf; float* scale; float* shift;
    const float* dw; const float* dbeta; float* dgamma;
    int cout; int K; int row0; int pad_;
} mpb_bn_layer;
int mpb_fold_bn_multi(int total_rows, const mpb_bn_layer* layers, const int* row2layer, float eps, void* stream);
/* p[0..n) = 0 with one CTA per SM (the gradient arena is zeroed beside the forward pass without taking its CTA slots) */
int mpb_zero_fill(long n, float* p, void* stream);
int mpb_bn_param_grad_multi(int total_rows, const mpb_bn_layer* layers, const int* row2layer, float eps, void* stream);
/* the same over rows [row_begin, row_end) of the table: d(gamma) of a group of layers as soon as their weight gradients
 * are final (the data-parallel step all-reduces the tower gradients bucket by bucket under the backward pass) */
int mpb_bn_param_grad_range(int row_begin, int row_end, const mpb_bn_layer* layers, const int* row2layer, float eps,
                            void* stream);

/* ---- stem: conv2d_same(7x7, stride 2) + frozen BN + ReLU  (nets/resnet_v1.py:234) ---- */
int mpb_stem_fwd(int nimg, int Hin, int Win, const float* x, const float* wf, const float* shift, float* y, void* stream);
int mpb_stem_wgrad(int nimg, int Hin, int Win, const float* x, const float* g, const float* scale, float* dw, void* stream);

/* ---- pools: slim.max_pool2d([3,3],2,'SAME') (resnet_v1.py:235); slim.max_pool2d([2,2]) (net_builder.py:60,68) */
int mpb_maxpool3s2_fwd(int nimg, int H, int W, int C, const float* x, float* y, void* stream);
int mpb_maxpool3s2_bwd(int nimg, int H, int W, int C, const float* x, const float* dy, float* dx, void* stream); /* dx also masked by x>0 */
int mpb_maxpool2_fwd(int nimg, int H, int W, int C, const float* x, int ldx, float* y, int ldy, void* stream);
int mpb_maxpool2_bwd(int nimg, int H, int W, int C, const float* x, int ldx, const float* dy, int ldy,
                     float* dx, int lddx, int accumulate, void* stream);

/* ---- tf.image.crop_and_resize(feat, boxes, 0, (crop,crop)) + max_pool2d([2,2]) fused (net_builder.py:54-60) */
int mpb_crop_pool_fwd(int H, int W, int C, const float* feat, int nbox, const float* boxes_norm, int crop,
                      float* out, int ldo, void* stream);
int mpb_crop_pool_bwd(int H, int W, int C, const float* feat, int nbox, const float* boxes_norm, int crop,
                      const float* dout, int ldd, float* dfeat, void* stream);   /* zeroes dfeat itself */

/* ---- tf.image.resize_images(align_corners=True) (net_builder.py:73-75,82-84) ---- */
int mpb_resize_ac_fwd(int nimg, int H, int W, int C, const float* x, int OH, int OW, float* y, void* stream);
/* ...16: the same kernels, writing ALSO the fp16 hi/lo split copy of y (y16, same geometry, [hi | lo] blocks; C % 32 == 0)
 * that the h3 forward GEMM consumes -- saves a separate mpb_split16 pass over the decoder activations. */
int mpb_resize_ac_fwd16(int nimg, int H, int W, int C, const float* x, int OH, int OW, float* y, void* y16, int* overflow,
                        void* stream);
int mpb_bn_train_fwd16(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean, float* var,
                       float* moving_mean, float* moving_var, float decay, double* scratch, void* y16, int* overflow,
                       void* stream);
int mpb_bn_infer_fwd16(int M, int C, const float* z, const float* beta, const float* moving_mean, const float* moving_var,
                       float eps, float* y, void* y16, int* overflow, void* stream);
/* train-mode batch norm in ONE launch per direction: statistics, a grid barrier, then the apply phase re-reading from L2
 * what the same CTA just streamed (2 CTAs per SM, all resident).  Same arithmetic and outputs as mpb_bn_train_fwd16 /
 * mpb_bn_train_bwd; scratch must hold 2C + 2 doubles (the barrier counter sits behind the accumulators). */
int mpb_bn_train_fwd_fused(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean, float* var,
                           float* moving_mean, float* moving_var, float decay, double* scratch, void* y16, int* overflow,
                           void* stream);
int mpb_bn_train_bwd_fused(int M, int C, const float* z, const float* mean, const float* var, float eps, const float* y,
                           const float* dy, float* dz, float* dbeta, double* scratch, void* stream);
int mpb_resize_ac_bwd(int nimg, int H, int W, int C, const float* dy, int OH, int OW, float* dx, void* stream); /* overwrites dx (gather form, no atomics) */

/* ---- slim.batch_norm(is_training=True) + ReLU of the map decoder (net_builder.py:77-89) ----
 * scratch: 2*C doubles.  moving_mean/var may be NULL (no UPDATE_OPS). */
int mpb_bn_train_fwd(int M, int C, const float* z, const float* beta, float eps, float* y, float* mean, float* var,
                     float* moving_mean, float* moving_var, float decay, double* scratch, void* stream);
/* inference-mode counterpart (is_training=False: moving statistics, no update) */
int mpb_bn_infer_fwd(int M, int C, const float* z, const float* beta, const float* moving_mean, const float* moving_var,
                     float eps, float* y, void* stream);
int mpb_bn_train_bwd(int M, int C, const float* z, const float* mean, const float* var, float eps, const float* y,
                     const float* dy, float* dz, float* dbeta, double* scratch, void* stream);

/* ---- xyz head: conv3x3 128->3 + bias (monopsr_output_builder.py:95-108); w is [3][3][3][128] ---- */
int mpb_xyzhead_fwd(int nimg, int H, int W, const float* x, const float* w, const float* bias, float* y, void* stream);
int mpb_xyzhead_bwd(int nimg, int H, int W, const float* x, const float* w, const float* dy, float* dx, float* dw,
                    float* db, void* stream);
/* the two halves of mpb_xyzhead_bwd, so that the weight gradient can run beside the data-gradient chain */
int mpb_xyzhead_dgrad(int nimg, int H, int W, const float* w, const float* dy, float* dx, void* stream);
int mpb_xyzhead_wgrad(int nimg, int H, int W, const float* x, const float* dy, float* dw, float* db, void* stream);   /* dw, db accumulate */

/* ---- small dense heads, N<=32 outputs (monopsr_output_builder.py:283,469,580,633); w is [N][K] ---- */
int mpb_fc_small_fwd(int B, int K, int N, const float* x, int ldx, const float* w, const float* bias, float* y, int ldy,
                     void* stream);
int mpb_fc_small_bwd(int B, int K, int N, const float* x, int ldx, const float* w, const float* dy, int ldy, float* dx,
                     int lddx, int accumulate_dx, float* dw, float* db, void* stream);

/* ---- elementwise helpers ---- */
int mpb_bias_relu(long rows, int C, const float* x, int ldx, const float* bias, int relu, int round, float* y, int ldy, void* stream);
int mpb_relu_bwd_colsum(int M, int C, const float* y, int ldy, const float* dy, int lddy, float* g, int ldg,
                        float* colsum, void* stream);
int mpb_add_inplace(long n, float* a, const float* b, void* stream);

/* ---- box heads, geometric projections, losses and their gradients (csrc/heads.cu) ----------
 * Replaces monopsr_output_builder.py:126-274,407-488,551-746, monopsr_model.py:416-461,554-958
 * and the tf_* geometry helpers (instance_utils.py:567-681,738-788,907-953; calib_utils.py:263-280). */
typedef struct mpb_heads_io {
    int nbox;
    /* per-box inputs (placeholders of monopsr_model.py:70-126) */
    const float* boxes_2d;            /* [N][4] y1,x1,y2,x2 px */
    const float* cam_p;               /* [3][4] */
    const int* class_indices;         /* [N] */
    const float* mean_lwh;            /* [N][3] */
    const float* prop_cen_z_offset;   /* [N] */
    const float* est_view_angs;       /* [N] */
    /* ground truth */
    const float* boxes_3d;            /* [N][7] x,y,z,l,w,h,ry */
    const int* gt_alpha_bins;         /* [N] */
    const float* gt_alpha_regs;       /* [N][12] */
    const float* gt_alpha_valid_bins; /* [N][12] */
    const float* gt_view_angs;        /* [N] */
    const float* gt_xyz_local;        /* [N][48][48][3] */
    const float* gt_xyz_global;       /* [N][48][48][3] (z used) */
    const float* valid_mask;          /* [N][48][48] */
    /* network outputs (read) */
    const float* lwh_offs;            /* [N][3] */
    const float* alpha;               /* [N][24] bins | regs */
    const float* cen_y_offs;          /* [N] */
    const float* cen_z_offs;          /* [N] */
    const float* xyz_local;           /* [N][48][48][3] */
    /* derived outputs (written) */
    float* lwh; float* prop_cen_z; float* prop_cen_y; float* cen_x; float* cen_y; float* cen_z; float* centroids;
    float* proj_err_norm;             /* [N] */
    float* depth_global;              /* [N][48][48] */
    float* feat1; int ld1;            /* proposal concat buffer [N][ld1]: cols 0..1023 img_fc, tail written here */
    float* feat2; int ld2;            /* regression concat buffer */
    /* losses and gradients */
    float* losses;                    /* [9] xyz, lwh, alpha_bins, alpha_regs, cen_z, cen_y, proj_err, depth, total */
    float* d_lwh_offs; float* d_alpha; float* d_cen_y_offs; float* d_cen_z_offs; float* d_xyz_local;
    float* d_prop_y; float* d_prop_z; /* scratch [N] */
    const float* d_feat2; int ldd2;   /* gradient of the regression concat buffer (from the fc0 data-gradient) */
    float* maskstats;                 /* [N+1]: valid pixels per box, total */
    /* loss of the local xyz map (yaml loss_config.inst_xyz_map_local = [type, weight]; loss_builder.py:19-84):
     * xyz_loss_mode 0 = smooth_l1_nonzero, evaluated (value and gradient) inside mpb_heads_final;
     * 1 = a point-set loss (chamfer_dist / emd): mpb_heads_final leaves that term out, the caller adds it with
     * mpb_pointset_* and the point-set ops of monopsr_b200_tfops.h. */
    int xyz_loss_mode;
    float xyz_loss_weight;
} mpb_heads_io;
int mpb_heads_static(const mpb_heads_io* io, void* stream);        /* once per sample */
int mpb_heads_mid(const mpb_heads_io* io, void* stream);           /* after lwh / alpha heads */
int mpb_heads_final(const mpb_heads_io* io, int train, void* stream); /* after cen_y / cen_z heads */
int mpb_heads_bwd_mid(const mpb_heads_io* io, void* stream);       /* after the regression fc0 data-gradient */

/* ---- glue of the point-set training losses (ChamferDistance / EarthMoversDistance as loss_config entries;
 * core/losses_custom.py:135-198, builders/loss_builder.py:60-84, monopsr_model.py:580-586) ----
 * clouds: p = pred * mask, t = gt * mask for npts points (mask is per point: masked points become (0,0,0) on both
 * sides, quirk Q8).  */
int mpb_pointset_mask(long npts, const float* pred, const float* gt, const float* mask, float* p, float* t, void* stream);
/* losses[slot] += scale * (sum(a[0..na)) + sum(b[0..nb))) and the same into losses[total_slot]; b may be NULL.
 * fill (may be NULL): fill[0..nfill) = fill_value (the constant upstream gradient of the distances). */
int mpb_pointset_loss_add(long na, const float* a, long nb, const float* b, float scale, float* losses, int slot,
                          int total_slot, float* fill, long nfill, float fill_value, void* stream);
/* d_pred[i][c] += scale * grad[i][c] * mask[i]   (d p / d pred = mask) */
int mpb_pointset_grad_add(long npts, const float* grad, const float* mask, float scale, float* d_pred, void* stream);

/* ---- fused train-op: per-variable clip_by_norm + Adam + EMA (csrc/optimizer.cu) ----
 * Replaces slim.learning.create_train_op(..., clip_gradient_norm=1.0) (core/trainer.py:76-81) with
 * AdamOptimizer + MovingAverageOptimizer (builders/optimizer_builder.py:56-82). */
/* chunks: the variables cut into pieces of the flat arena; the chunks of one variable must be CONTIGUOUS in the
 * table.  norm2: scratch, one float per chunk (per-chunk sums of squares, combined per variable in a fixed order
 * so that every data-parallel replica computes the same clip factor bit for bit). */
typedef struct mpb_opt_chunk { long start; int len; int tensor; } mpb_opt_chunk;
int mpb_opt_step(int nchunks, const mpb_opt_chunk* chunks, int ntensors, float* param, const float* grad,
                 float* m, float* v, float* ema, float* norm2, const float* hyper, float grad_scale,
                 float clip_norm, float beta1, float beta2, float eps, float ema_decay, void* stream);
/* the same over a sub-range of the chunk table: `chunks` points at the first chunk of the range, whose tensors are
 * tensor0 .. tensor0+ntensors-1; norm2 needs one float per chunk of the range.  The update is per variable, so
 * a group of variables can be stepped as soon as its gradients are final, under the rest of the backward pass. */
int mpb_opt_step_range(int nchunks, const mpb_opt_chunk* chunks, int tensor0, int ntensors, float* param,
                       const float* grad, float* m, float* v, float* ema, float* norm2, const float* hyper,
                       float grad_scale, float clip_norm, float beta1, float beta2, float eps, float ema_decay,
                       void* stream);

/* ---- ground-truth target synthesis (csrc/targets.cu) ----
 * Replaces the 2 x num_boxes instance_utils.tf_instance_xyz_crop_from_depth_map sub-graphs of MonoPSRModel.build
 * (core/models/monopsr/monopsr_model.py:165-203; datasets/kitti/instance_utils.py:395-481,
 * datasets/kitti/depth_map_utils.py:161-236, core/transform_utils.py:36-66) with one launch.
 * depth [H][W] fp32, masks [nbox][H][W] bytes (0 / non-0), boxes_2d [nbox][4] = y1,x1,y2,x2 px, boxes_3d [nbox][ld3]
 * (x,y,z,l,w,h,..), view_angs [nbox], cam_p [3][4]; outputs xyz_local / xyz_global [nbox][roi][roi][3] and
 * valid [nbox][roi][roi] (1.0 where |depth| >= 0.1).  All DEVICE pointers. */
int mpb_gt_xyz_from_depth(int nbox, int H, int W, int roi, const float* depth, const unsigned char* masks,
                          const float* boxes_2d, const float* boxes_3d, int ld3, const float* view_angs,
                          const float* cam_p, int centroid_middle, int rotate_view, float* xyz_local,
                          float* xyz_global, float* valid, void* stream);

/* network inputs from the raw camera image: ImgPreprocessor.preprocess_input (core/img_preprocessor.py:12-35: minus
 * channel means, legacy bilinear resize to PH x PW), the RGB instance crops (tf.image.crop_and_resize with the
 * normalised boxes, monopsr_model.py:222-226) and the down-sized full image (resize_bilinear align_corners=True,
 * :228-233).  img [H][W][3] uint8 (img_is_u8) or fp32, DEVICE; channel_means: 3 floats on the HOST; preprocessed
 * [PH][PW][3] is an output too; rgb_crops [nbox][crop][crop][3] and full_img [FH][FW][3] may be NULL. */
int mpb_image_inputs(int H, int W, const void* img, int img_is_u8, const float* channel_means, int PH, int PW,
                     float* preprocessed, int nbox, const float* boxes_norm, int crop, float* rgb_crops, int FH, int FW,
                     float* full_img, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONOPSR_B200_NET_H_ */
