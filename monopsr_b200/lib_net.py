"""ctypes mirror of include/monopsr_b200_net.h (structs + argument types)."""
import ctypes

c_i = ctypes.c_int
c_l = ctypes.c_long
c_f = ctypes.c_float
c_p = ctypes.c_void_p

TC_FWD, TC_DGRAD, TC_WGRAD = 0, 1, 2


class TcGemmParams(ctypes.Structure):
    _fields_ = [
        ("op", c_i), ("H", c_i), ("W", c_i), ("kh", c_i), ("kw", c_i), ("dil", c_i), ("M", c_i),
        ("Cin", c_i), ("Cout", c_i),
        ("X", c_p), ("ldx", c_i),
        ("Y", c_p), ("ldy", c_i),
        ("Wt", c_p), ("ldw", c_i),
        ("out", c_p), ("ldo", c_i),
        ("tapmask", c_p),
        ("scale", c_p), ("shift", c_p),
        ("res", c_p), ("ldr", c_i),
        ("mask", c_p), ("ldm", c_i),
        ("scale2", c_p), ("rowscale", c_p), ("colsum", c_p),
        ("relu", c_i), ("round_tf32", c_i), ("atomic", c_i), ("ksplit", c_i),
        ("out_r", c_p), ("ldor", c_i),
        ("X16", c_p), ("W16", c_p), ("out16", c_p), ("ldo16", c_i), ("overflow", c_p),
    ]


class HeadsIO(ctypes.Structure):
    _fields_ = [
        ("nbox", c_i),
        ("boxes_2d", c_p), ("cam_p", c_p), ("class_indices", c_p), ("mean_lwh", c_p),
        ("prop_cen_z_offset", c_p), ("est_view_angs", c_p),
        ("boxes_3d", c_p), ("gt_alpha_bins", c_p), ("gt_alpha_regs", c_p), ("gt_alpha_valid_bins", c_p),
        ("gt_view_angs", c_p), ("gt_xyz_local", c_p), ("gt_xyz_global", c_p), ("valid_mask", c_p),
        ("lwh_offs", c_p), ("alpha", c_p), ("cen_y_offs", c_p), ("cen_z_offs", c_p), ("xyz_local", c_p),
        ("lwh", c_p), ("prop_cen_z", c_p), ("prop_cen_y", c_p), ("cen_x", c_p), ("cen_y", c_p), ("cen_z", c_p),
        ("centroids", c_p), ("proj_err_norm", c_p), ("depth_global", c_p),
        ("feat1", c_p), ("ld1", c_i), ("feat2", c_p), ("ld2", c_i),
        ("losses", c_p),
        ("d_lwh_offs", c_p), ("d_alpha", c_p), ("d_cen_y_offs", c_p), ("d_cen_z_offs", c_p), ("d_xyz_local", c_p),
        ("d_prop_y", c_p), ("d_prop_z", c_p),
        ("d_feat2", c_p), ("ldd2", c_i),
        ("maskstats", c_p),
        ("xyz_loss_mode", c_i), ("xyz_loss_weight", c_f),
    ]


class BnLayer(ctypes.Structure):
    _fields_ = [("w", c_p), ("gamma", c_p), ("beta", c_p), ("mean", c_p), ("var", c_p), ("wf", c_p), ("scale", c_p),
                ("shift", c_p), ("dw", c_p), ("dbeta", c_p), ("dgamma", c_p), ("cout", c_i), ("K", c_i), ("row0", c_i),
                ("pad_", c_i)]


class W16Layer(ctypes.Structure):
    _fields_ = [("w", c_p), ("gamma", c_p), ("var", c_p), ("w16", c_p), ("inv_scale", c_p), ("cout", c_i), ("K", c_i),
                ("row0", c_i), ("pad_", c_i), ("wf", c_p), ("scale", c_p), ("shift", c_p), ("beta", c_p), ("mean", c_p)]


class OptChunk(ctypes.Structure):
    _fields_ = [("start", c_l), ("len", c_i), ("tensor", c_i)]


_SIGS = {
    "mpb_tc_gemm": [ctypes.POINTER(TcGemmParams), c_i, c_p],
    "mpb_tc_set_producer": [c_i],
    "mpb_tc_set_cluster": [c_i],
    "mpb_tc_max_clusters": [c_i, c_i, c_i],
    "mpb_build_tapmask": [c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "mpb_fold_bn": [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p],
    "mpb_fold_bn_multi": [c_i, c_p, c_p, c_f, c_p],
    "mpb_bn_param_grad_multi": [c_i, c_p, c_p, c_f, c_p],
    "mpb_bn_param_grad_range": [c_i, c_i, c_p, c_p, c_f, c_p],
    "mpb_round_copy": [c_l, c_p, c_p, c_p],
    "mpb_bn_param_grad": [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p],
    "mpb_stem_fwd": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "mpb_stem_wgrad": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "mpb_maxpool3s2_fwd": [c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "mpb_maxpool3s2_bwd": [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "mpb_maxpool2_fwd": [c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p],
    "mpb_maxpool2_bwd": [c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_p],
    "mpb_crop_pool_fwd": [c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p],
    "mpb_crop_pool_bwd": [c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p],
    "mpb_resize_ac_fwd": [c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p],
    "mpb_resize_ac_bwd": [c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p],
    "mpb_bn_train_fwd": [c_i, c_i, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p],
    "mpb_resize_ac_fwd16": [c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p],
    "mpb_bn_train_fwd16": [c_i, c_i, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p],
    "mpb_bn_infer_fwd16": [c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p],
    "mpb_bn_train_fwd_fused": [c_i, c_i, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p],
    "mpb_bn_train_bwd_fused": [c_i, c_i, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "mpb_tc_gemm_x3": [c_p, c_i, c_p],
    "mpb_tc_gemm_h3": [c_p, c_i, c_p],
    "mpb_split16": [c_l, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p],
    "mpb_split16_weights_multi": [c_i, c_p, c_p, c_f, c_p],
    "mpb_split16_weights_rows": [c_i, c_p, c_i, c_p, c_p, c_f, c_p],
    "mpb_set_operand_rounding": [c_i],
    "mpb_bn_infer_fwd": [c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_p, c_p],
    "mpb_bn_train_bwd": [c_i, c_i, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "mpb_xyzhead_fwd": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "mpb_xyzhead_dgrad": [c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "mpb_xyzhead_wgrad": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "mpb_xyzhead_bwd": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "mpb_fc_small_fwd": [c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_p],
    "mpb_fc_small_bwd": [c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p],
    "mpb_bias_relu": [c_l, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_i, c_p],
    "mpb_relu_bwd_colsum": [c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p],
    "mpb_add_inplace": [c_l, c_p, c_p, c_p],
    "mpb_zero_fill": [c_l, c_p, c_p],
    "mpb_heads_static": [ctypes.POINTER(HeadsIO), c_p],
    "mpb_heads_mid": [ctypes.POINTER(HeadsIO), c_p],
    "mpb_heads_final": [ctypes.POINTER(HeadsIO), c_i, c_p],
    "mpb_heads_bwd_mid": [ctypes.POINTER(HeadsIO), c_p],
    "mpb_pointset_mask": [c_l, c_p, c_p, c_p, c_p, c_p, c_p],
    "mpb_pointset_loss_add": [c_l, c_p, c_l, c_p, c_f, c_p, c_i, c_i, c_p, c_l, c_f, c_p],
    "mpb_pointset_grad_add": [c_l, c_p, c_p, c_f, c_p, c_p],
    "mpb_gt_xyz_from_depth": [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p],
    "mpb_image_inputs": [c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p],
    "mpb_opt_step_range": [c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_f, c_f, c_p],
    "mpb_opt_step": [c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_f, c_f, c_p],
}


def declare(lib):
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = c_i
