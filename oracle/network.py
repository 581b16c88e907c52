"""oracle/network.py -- TEST INFRASTRUCTURE ONLY.

fp64 (or fp32, for the CPU timing baseline) RESTATEMENT in plain torch of the MonoPSR
per-instance network of ``monopsr_model_000`` -- NOT TensorFlow, which is an un-vendored,
un-installable dependency of the reference (requirements.txt:13).  Parity against the real
TF1 graph is UNPINNED for the arithmetic of TensorFlow's own kernels (convolution, batch
norm, resize, crop_and_resize: cross-checked against torchvision / torch.nn.functional in
tests/test_host_mirrors.py).  Everything that lives in the reference's Python IS pinned to
that code, executed rather than read -- end to end by running its whole MonoPSRModel
__init__ / build / loss on arrays and comparing all outputs and loss terms
(tests/test_graph_golden.py), and piecewise: the architecture to the record of its graph builders
(tests/test_arch_golden.py), the geometry to its TF functions and numpy twins run on arrays
(tests/test_geometry_tf_golden.py, tests/test_oracle_geometry.py), the loss to its
MonoPSRModel.loss (tests/test_loss_golden.py).  This file follows, line by line:

  nets/resnet_v1.py:78-139,142-254,310-330   bottleneck, resnet_v1, resnet_v1_101
  nets/resnet_utils.py:59-122,125-219,222-272 subsample, conv2d_same, stack_blocks_dense, arg_scope
  core/feature_extractors/faster_rcnn_resnet_v1_feature_extractor.py:197-245
  builders/net_builder.py:30-96               resnet101_4x_squash
  core/models/monopsr/monopsr_output_builder.py:95-108,126-302,407-488,551-746
  core/models/monopsr/monopsr_model.py:138-492 (build), 554-958 (loss)
  datasets/kitti/instance_utils.py:567-681,738-788,907-953 ; calib_utils.py:263-280
  core/transform_utils.py:69-108 ; obj_utils.py:1016-1034
  core/losses_custom.py:93-132 ; object_detection/core/losses.py:118-157,283-317

with TF/TF-slim default semantics restated from knowledge of TF 1.8 (SAME padding,
conv2d_same, crop_and_resize, resize_bilinear(align_corners=True), slim.batch_norm
defaults).  Parameters are a dict name -> numpy array in TF variable layout (conv HWIO,
fully-connected [in,out]); they are DATA handed in by the caller.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

NUM_BOXES = 32
BN_EPS_RESNET = 1e-5      # feature_extractor.py:229
BN_EPS_DECODER = 1e-3     # slim.batch_norm default
MAX_DEPTH = 45.0          # dataset_config.obj_filter_config.depth_range[1] (yaml:31)


# ----------------------------------------------------------------------------- tf32 emulation
# The product feeds fp32 operands to tcgen05 kind::tf32 and rounds them to tf32 (cvt.rna: 10
# mantissa bits, ties away from zero) where they are produced.  With EMULATE_TF32 the oracle
# applies the same rounding at the same points (straight-through gradient), so the remaining
# product-vs-oracle difference is fp32-vs-fp64 accumulation only; with it off the oracle is the
# plain fp64 restatement and the difference measures the tf32 effect itself.
EMULATE_TF32 = False
# Emulation variant: keep the residual stream in full precision (the product would store an
# unrounded copy of every unit output next to the tf32-rounded one that feeds the GEMMs).
RESIDUAL_FULL = False


class _RoundTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        i = x.detach().to(torch.float32).contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def Q(x):
    return _RoundTF32.apply(x) if EMULATE_TF32 else x


# ----------------------------------------------------------------------------- layers
def conv_hwio(x, w, stride=1, rate=1, pad="SAME"):
    """x NHWC, w HWIO (TF).  SAME for stride 1 == symmetric pad rate*(k-1)/2 (k odd)."""
    k = w.shape[0]
    wt = w.permute(3, 2, 0, 1)
    xn = x.permute(0, 3, 1, 2)
    if pad == "SAME":
        assert stride == 1
        y = F.conv2d(xn, wt, stride=1, padding=rate * (k - 1) // 2, dilation=rate)
    else:
        y = F.conv2d(xn, wt, stride=stride, padding=0, dilation=rate)
    return y.permute(0, 2, 3, 1)


def conv2d_same(x, w, stride, rate=1):
    """resnet_utils.py:77-122"""
    if stride == 1:
        return conv_hwio(x, w, 1, rate, "SAME")
    k = w.shape[0]
    k_eff = k + (k - 1) * (rate - 1)
    pad_total = k_eff - 1
    pad_beg = pad_total // 2
    pad_end = pad_total - pad_beg
    x = F.pad(x, (0, 0, pad_beg, pad_end, pad_beg, pad_end))
    return conv_hwio(x, w, stride, rate, "VALID")


def frozen_bn(x, P, scope):
    g, b = P[scope + "/gamma"], P[scope + "/beta"]
    m, v = P[scope + "/moving_mean"], P[scope + "/moving_variance"]
    return (x - m) / torch.sqrt(v + BN_EPS_RESNET) * g + b


def conv_bn(x, P, scope, stride=1, rate=1):
    """conv (no bias) + inference-mode BN.  Emulation mode mirrors the product: BN scale folded
    into the weights, folded weights rounded to tf32, then '+ shift'."""
    w = P[scope + "/weights"]
    bn = scope + "/BatchNorm"
    if not EMULATE_TF32:
        return frozen_bn(conv2d_same(x, w, stride, rate), P, bn)
    sc = P[bn + "/gamma"] / torch.sqrt(P[bn + "/moving_variance"] + BN_EPS_RESNET)
    wf = Q(w * sc)           # HWIO: scale broadcasts over the output-channel axis
    return conv2d_same(x, wf, stride, rate) + (P[bn + "/beta"] - P[bn + "/moving_mean"] * sc)


def max_pool_same_3x3_s2(x):
    """slim.max_pool2d([3,3], stride=2, padding='SAME'): even input => one pad row/col at the end."""
    xn = x.permute(0, 3, 1, 2)
    H, W = xn.shape[2], xn.shape[3]
    ph = max((math.ceil(H / 2) - 1) * 2 + 3 - H, 0)
    pw = max((math.ceil(W / 2) - 1) * 2 + 3 - W, 0)
    xn = F.pad(xn, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
    return F.max_pool2d(xn, 3, 2).permute(0, 2, 3, 1)


def max_pool_2x2(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


def bottleneck(x, P, scope, depth, depth_bottleneck, rate):
    """resnet_v1.py:78-139 with stride forced to 1 (output_stride reached, resnet_utils.py:194-200)."""
    s = scope + "/bottleneck_v1"
    xin = Q(x) if RESIDUAL_FULL else x          # GEMM operand (x already rounded unless RESIDUAL_FULL)
    if x.shape[-1] == depth:
        shortcut = x
    else:
        shortcut = conv_bn(xin, P, s + "/shortcut")
    r = Q(torch.relu(conv_bn(xin, P, s + "/conv1")))
    r = Q(torch.relu(conv_bn(r, P, s + "/conv2", 1, rate)))
    r = conv_bn(r, P, s + "/conv3")
    out = torch.relu(shortcut + r)
    return out if RESIDUAL_FULL else Q(out)


BLOCKS = [("block1", 64, 3), ("block2", 128, 4), ("block3", 256, 23)]   # block4 is never consumed


def resnet101_block3(x, P, scope):
    """feature_extractor.py:197-245 with output_stride=4: stem /2, pool /2, every unit stride 1,
    atrous rate 1/2/4 in block1/2/3 (resnet_utils.py:181-200)."""
    s = scope + "/resnet_v1_101"
    x = Q(torch.relu(conv_bn(x, P, s + "/conv1", 2)))
    x = max_pool_same_3x3_s2(x)
    rate = 1
    for name, base, units in BLOCKS:
        for u in range(1, units + 1):
            x = bottleneck(x, P, "%s/%s/unit_%d" % (s, name, u), base * 4, base, rate)
        rate *= 2
    return x


def crop_and_resize(img, boxes_norm, crop_h, crop_w):
    """tf.image.crop_and_resize (bilinear, extrapolation 0) of ONE image (box_ind all 0). img (1,H,W,C)."""
    _, H, W, C = img.shape
    dt = img.dtype
    y1, x1, y2, x2 = [boxes_norm[:, i] for i in range(4)]
    iy = torch.arange(crop_h, dtype=dt, device=img.device)
    ix = torch.arange(crop_w, dtype=dt, device=img.device)
    in_y = y1[:, None] * (H - 1) + iy[None, :] * ((y2 - y1) * (H - 1) / (crop_h - 1))[:, None]   # (N,ch)
    in_x = x1[:, None] * (W - 1) + ix[None, :] * ((x2 - x1) * (W - 1) / (crop_w - 1))[:, None]   # (N,cw)
    vy = (in_y >= 0) & (in_y <= H - 1)
    vx = (in_x >= 0) & (in_x <= W - 1)
    top = torch.floor(in_y).clamp(0, H - 1)
    bot = torch.ceil(in_y).clamp(0, H - 1)
    lef = torch.floor(in_x).clamp(0, W - 1)
    rig = torch.ceil(in_x).clamp(0, W - 1)
    ly = (in_y - torch.floor(in_y))[:, :, None, None]
    lx = (in_x - torch.floor(in_x))[:, None, :, None]
    im = img[0]

    def g(yy, xx):
        return im[yy.long()[:, :, None], xx.long()[:, None, :]]   # (N,ch,cw,C)

    t = g(top, lef) + (g(top, rig) - g(top, lef)) * lx
    b = g(bot, lef) + (g(bot, rig) - g(bot, lef)) * lx
    out = t + (b - t) * ly
    valid = (vy[:, :, None] & vx[:, None, :])[..., None]
    return torch.where(valid, out, torch.zeros_like(out))


def resize_bilinear_ac(x, oh, ow):
    """tf.image.resize_images(..., align_corners=True)"""
    return F.interpolate(x.permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear",
                         align_corners=True).permute(0, 2, 3, 1)


def train_bn_relu(x, P, scope):
    """slim.batch_norm defaults, is_training=True: batch statistics (biased variance), beta only."""
    mean = x.mean((0, 1, 2))
    var = x.var((0, 1, 2), unbiased=False)
    return Q(torch.relu((x - mean) / torch.sqrt(var + BN_EPS_DECODER) + P[scope + "/beta"])), mean, var


def infer_bn_relu(x, P, scope):
    """slim.batch_norm defaults, is_training=False: moving statistics (validation / inference graphs)."""
    y = (x - P[scope + "/moving_mean"]) / torch.sqrt(P[scope + "/moving_variance"] + BN_EPS_DECODER) + P[scope + "/beta"]
    return Q(torch.relu(y)), P[scope + "/moving_mean"], P[scope + "/moving_variance"]


def fc(x, P, scope, relu=True):
    """slim.fully_connected; the 1024-wide ReLU layers run on the tensor cores in the product
    (weights and outputs tf32-rounded), the small linear heads run in plain fp32."""
    if relu:
        return Q(torch.relu(x @ Q(P[scope + "/weights"]) + P[scope + "/biases"]))
    return x @ P[scope + "/weights"] + P[scope + "/biases"]


# ----------------------------------------------------------------------------- losses
def huber(x, delta=1.0):
    a = x.abs()
    q = torch.clamp(a, max=delta)
    return 0.5 * q * q + delta * (a - q)


def smooth_l1_nonzero(pred, gt, weights):
    """tf.losses.huber_loss(reduction=SUM_BY_NONZERO_WEIGHTS): sum(l*w)/count(w!=0 broadcast to l)."""
    w = weights.expand_as(pred)
    num = (huber(pred - gt) * w).sum()
    cnt = (w != 0).to(pred.dtype).sum()
    return torch.where(cnt > 0, num / torch.clamp(cnt, min=1.0), torch.zeros_like(num))


# ----------------------------------------------------------------------------- the model
def to_torch(P, dtype=torch.float64, device="cpu"):
    return {k: torch.as_tensor(np.asarray(v), device=device).to(dtype if np.asarray(v).dtype.kind == "f" else torch.int64)
            for k, v in P.items()}


def prop_cen_y_from_box(boxes_2d, cam_p, prop_cen_z):
    """tf_est_y_from_box_2d_and_depth (instance_utils.py:907-953; numpy twin :841-904), 'middle' centroids of cars with
    the KITTI trend offset: the box centre row back-projected at depth z, minus 0.0648"""
    N = boxes_2d.shape[0]
    box_cv = ((boxes_2d[:, 2] + boxes_2d[:, 0]) / 2.0 - cam_p[1, 2]).reshape(N, 1)
    return box_cv * (prop_cen_z / cam_p[0, 0]) - 0.0648


def box_heads(P, S, pooled):
    """The FC stacks and box heads on the pooled (N,6,6,512) features (monopsr_output_builder.py:126-302, 407-488,
    551-623): layer names, widths and the order of the concatenated inputs are pinned to the record of the reference's
    own builder code (tests/test_arch_golden.py)."""
    dt, dev = pooled.dtype, pooled.device
    N = pooled.shape[0]
    boxes_2d, cam_p = S["boxes_2d"], S["cam_p"]
    # proposal fc :126-194
    flat = pooled.reshape(N, -1)
    est_view = S["est_view_angs"].reshape(N, 1)
    centre_u, centre_v = cam_p[0, 2], cam_p[1, 2]
    box_ij = boxes_2d - torch.stack([centre_v, centre_u, centre_v, centre_u])
    box_h = (boxes_2d[:, 2] - boxes_2d[:, 0]).reshape(N, 1)
    img_h, img_w = 320.0, 1216.0                      # model_config.image_input_shape
    box_h_norm = box_h / img_h
    box_ij_norm = box_ij / torch.tensor([img_h / 2, img_w / 2, img_h / 2, img_w / 2], dtype=dt, device=dev)
    # tf.one_hot(squeeze(class_indices), num_classes=1): 1.0 only for index 0 (quirk Q7)
    one_hot = (S["class_indices"].reshape(N, 1) == 0).to(dt)
    cam_norm = cam_p.reshape(1, 12) / torch.tensor(
        [1000.0, 1.0, 1000.0, 100.0, 1.0, 1000.0, 1000.0, 1.0, 1.0, 1.0, 1.0, 1.0], dtype=dt, device=dev)
    p = "output/proposal_fc/proposal_fc"
    img_fc = fc(flat, P, p + "/img_fc")
    feat = torch.cat([img_fc, box_ij_norm, box_h_norm, est_view, one_hot, cam_norm.expand(N, 12)], dim=1)
    feat = fc(fc(feat, P, p + "/fc0"), P, p + "/fc1")
    lwh_offs = fc(feat, P, "output/lwh/lwh", relu=False)
    lwh = S["mean_lwh"] + lwh_offs
    alpha = fc(feat, P, "output/alpha", relu=False)
    alpha_bins, alpha_regs = alpha[:, :12], alpha[:, 12:24]
    out = {"lwh": lwh, "lwh_offs": lwh_offs, "alpha_bins": alpha_bins, "alpha_regs": alpha_regs,
           "view_ang": est_view, "_est_view": est_view}

    # centroid proposals :407-438 ; instance_utils.py:907-953
    f = cam_p[0, 0]
    prop_cen_z = (f * lwh[:, 2] / (boxes_2d[:, 2] - boxes_2d[:, 0]) + S["prop_cen_z_offset"]).reshape(N, 1)
    prop_cen_y = prop_cen_y_from_box(boxes_2d, cam_p, prop_cen_z)
    out["prop_cen_z"] = prop_cen_z

    # regression fc :200-274
    p = "output/regression_fc/regression_fc"
    img_fc2 = fc(flat, P, p + "/img_fc")
    feat2 = torch.cat([img_fc2, box_ij_norm, box_h_norm, est_view, one_hot, lwh_offs, alpha_bins, alpha_regs,
                       prop_cen_y / 1.666754, prop_cen_z / MAX_DEPTH], dim=1)
    feat2 = fc(fc(feat2, P, p + "/fc0"), P, p + "/fc1")
    cen_y_offs = fc(feat2, P, "output/cen_y/cen_y", relu=False)
    cen_z_offs = fc(feat2, P, "output/cen_z_offs/cen_z", relu=False)
    cen_y = prop_cen_y + cen_y_offs
    cen_z = prop_cen_z + cen_z_offs
    x_offset = -cam_p[0, 3] / cam_p[0, 0]
    cen_x = cen_z * torch.tan(est_view) + x_offset
    out.update({"cen_y": cen_y, "cen_y_offs": cen_y_offs, "cen_z": cen_z, "cen_z_offs": cen_z_offs,
                "cen_x": cen_x, "centroids": torch.cat([cen_x, cen_y, cen_z], dim=1)})

    return out


def train_projections(xyz_local, valid, boxes_2d, cam_p, est_view, gt_view, cen_y, cen_z):
    """The train / val-only geometry of the graph: local map -> camera frame -> image, expected pixel-centre grid,
    normalised projection error, global depth map.  Numpy twins in the reference (pinned in
    tests/test_oracle_geometry.py): inst_points_local_to_global (instance_utils.py:552-564), project_pc_to_image
    (calib_utils.py:245-260), get_exp_proj_uv_map(use_pixel_centres=True) (instance_utils.py:684-735)."""
    N = xyz_local.shape[0]
    dt, dev = xyz_local.dtype, xyz_local.device
    f, centre_u = cam_p[0, 0], cam_p[0, 2]
    x_offset = -cam_p[0, 3] / cam_p[0, 0]
    # xyz projection (monopsr_model.py:416-448; output_builder :663-746)
    proj_cen = torch.cat([cen_z * torch.tan(gt_view) + x_offset, cen_y, cen_z], dim=1)
    c, s = torch.cos(gt_view)[:, :, None], torch.sin(gt_view)[:, :, None]      # (N,1,1)
    lx, ly, lz = xyz_local[..., 0], xyz_local[..., 1], xyz_local[..., 2]      # (N,48,48)
    gx = c * lx + s * lz + proj_cen[:, 0, None, None]
    gy = ly + proj_cen[:, 1, None, None]
    gz = -s * lx + c * lz + proj_cen[:, 2, None, None]
    pu = cam_p[0, 0] * gx + cam_p[0, 1] * gy + cam_p[0, 2] * gz + cam_p[0, 3]
    pv = cam_p[1, 0] * gx + cam_p[1, 1] * gy + cam_p[1, 2] * gz + cam_p[1, 3]
    pw = cam_p[2, 0] * gx + cam_p[2, 1] * gy + cam_p[2, 2] * gz + cam_p[2, 3]
    proj_u, proj_v = pu / pw, pv / pw
    # expected uv at pixel centres (instance_utils.py:738-788): linspace(start+half, stop-half, 48)
    lin = torch.arange(48, dtype=dt, device=dev) / 47.0
    v1, u1, v2, u2 = [boxes_2d[:, i] for i in range(4)]
    hu, hv = (u2 - u1) / 48 / 2.0, (v2 - v1) / 48 / 2.0
    grid_u = (u1 + hu)[:, None] + ((u2 - hu) - (u1 + hu))[:, None] * lin[None, :]
    grid_v = (v1 + hv)[:, None] + ((v2 - hv) - (v1 + hv))[:, None] * lin[None, :]
    exp_u = grid_u[:, None, :].expand(N, 48, 48)
    exp_v = grid_v[:, :, None].expand(N, 48, 48)
    bw, bh = (u2 - u1)[:, None, None], (v2 - v1)[:, None, None]
    vm = valid[..., 0]
    eu = torch.clamp((exp_u - proj_u) / bw * vm, -2.0, 2.0)
    ev = torch.clamp((exp_v - proj_v) / bh * vm, -2.0, 2.0)
    nvalid = vm.sum((1, 2))
    nvalid = torch.where(nvalid < 1.0, torch.ones_like(nvalid), nvalid)
    proj_err_norm = (eu.sum((1, 2)) + ev.sum((1, 2))) / nvalid

    # global depth map (instance_utils.py:607-681, rotate_view=True, quirk Q6)
    x1b, x2b = boxes_2d[:, 1], boxes_2d[:, 3]
    sp = (x2b - x1b) / 48 / 2.0
    va_l = torch.atan2((x1b + sp - centre_u) / f, torch.ones_like(x1b)).reshape(N, 1)
    va_r = torch.atan2((x2b - sp - centre_u) / f, torch.ones_like(x1b)).reshape(N, 1)
    inst_xz = cen_z / torch.cos(est_view)
    l_o = inst_xz / torch.cos(va_l - est_view)
    r_o = inst_xz / torch.cos(va_r - est_view)
    off_l = (l_o * torch.sin(va_l - est_view) * torch.sin(est_view)).reshape(N)
    off_r = (r_o * torch.sin(va_r - est_view) * torch.sin(est_view)).reshape(N)
    off = (-off_l)[:, None] + ((-off_r) - (-off_l))[:, None] * lin[None, :]       # (N,48), along ROWS
    depth_global = xyz_local[..., 2:3] + cen_z.reshape(N, 1, 1, 1) + off.reshape(N, 48, 1, 1)
    return {"global_xyz": torch.stack([gx, gy, gz], dim=-1), "proj_uv": torch.stack([proj_u, proj_v], dim=-1),
            "exp_uv": torch.stack([exp_u, exp_v], dim=-1), "proj_err_norm": proj_err_norm,
            "inst_depth_map_global": depth_global}


def forward(P, S, train=True, projections=None):
    """P: dict of torch tensors (see to_torch); S: dict of torch tensors (the synthetic feed_dict).
    Returns (output_dict, aux) -- output_dict keys follow core/constants.py KEY_*.
    train: the reference's is_training (train_val_test == 'train': batch statistics in the decoder's batch norm);
    projections (default = train): the train-or-val part of the graph (projection error, global depth map) -- the
    'val' graph is forward(train=False, projections=True)."""
    if projections is None:
        projections = train
    dt = S["rgb_crops"].dtype
    dev = S["rgb_crops"].device
    N = S["rgb_crops"].shape[0]
    boxes_2d = S["boxes_2d"]
    cam_p = S["cam_p"]
    out = {}

    # net_builder.py:30-96
    crop_feat = resnet101_block3(S["rgb_crops"], P, "FirstStageFeatureExtractor_crop")
    full_feat = resnet101_block3(S["full_img"], P, "FirstStageFeatureExtractor_full")
    large = crop_and_resize(full_feat, S["boxes_2d_norm"], 24, 24)
    full_crop = Q(max_pool_2x2(large))
    concat = torch.cat([crop_feat, full_crop], dim=3)
    squashed = Q(torch.relu(conv_hwio(concat, Q(P["squash/1x1_conv/weights"])) + P["squash/1x1_conv/biases"]))
    pooled = max_pool_2x2(squashed)
    x = Q(resize_bilinear_ac(squashed, 24, 24))
    bn_stats = {}
    bn_relu = train_bn_relu if train else infer_bn_relu      # is_training = (train_val_test == 'train'), monopsr_model.py:139
    for i in (1, 2):
        sc = "map_decoder/conv2/conv2_%d" % i
        x, m, v = bn_relu(conv_hwio(x, Q(P[sc + "/weights"])), P, sc + "/BatchNorm")
        bn_stats[sc] = (m, v)
    x = Q(resize_bilinear_ac(x, 48, 48))
    for i in (1, 2):
        sc = "map_decoder/conv3/conv3_%d" % i
        x, m, v = bn_relu(conv_hwio(x, Q(P[sc + "/weights"])), P, sc + "/BatchNorm")
        bn_stats[sc] = (m, v)
    map_features = x

    # output builder :95-108
    sc = "output/inst_xyz_map_local/inst_xyz_map_local"
    xyz_local = conv_hwio(map_features, P[sc + "/weights"]) + P[sc + "/biases"]
    out["inst_xyz_map_local"] = xyz_local
    valid = S["gt_valid_mask_maps"]
    out["valid_mask_maps"] = valid

    heads = box_heads(P, S, pooled)
    out.update(heads)
    est_view, cen_y, cen_z = out.pop("_est_view"), out["cen_y"], out["cen_z"]

    aux = {"map_features": map_features, "features_pooled": pooled, "features_squashed": squashed,
           "crop_feat": crop_feat, "full_feat": full_feat, "bn_stats": bn_stats, "concat": concat}
    if not projections:
        return out, aux

    g = train_projections(xyz_local, valid, boxes_2d, cam_p, est_view, S["gt_view_angs"].reshape(N, 1), cen_y, cen_z)
    out["proj_err_norm"] = g["proj_err_norm"]
    out["inst_depth_map_global"] = g["inst_depth_map_global"]
    return out, aux


def regression_targets(out, S):
    """gt_dict entries of the offset heads as the output builder creates them (monopsr_output_builder.py:441-488,
    573-609, 625-661; gt centroid: monopsr_model.py:262-277, centroid_type 'middle').  Pinned to those methods
    executed on arrays: tests/test_loss_golden.py::test_regression_targets."""
    b3 = S["boxes_3d"]
    gt_cen_y = b3[:, 1:2] - b3[:, 5:6] / 2.0          # centroid_type 'middle' (yaml:20)
    gt_cen_z = b3[:, 2:3]
    prop_cen_y = out["cen_y"] - out["cen_y_offs"]
    return {
        # gt_lwh - pred_lwh: a graph tensor that depends on the prediction; TF differentiates through it
        # (reference quirk Q9, kept as is)
        "lwh_offs": b3[:, 3:6] - out["lwh"],
        "cen_z_offs": gt_cen_z - out["prop_cen_z"],
        "cen_y_offs": gt_cen_y - prop_cen_y,
    }


def chamfer_loss(pred, gt, valid):
    """losses_custom.ChamferDistance (losses_custom.py:169-198): both maps times the mask, (B, h*w, 3) clouds, squared
    nearest-neighbour distances in both directions, sum / B"""
    B = pred.shape[0]
    p, t = (pred * valid).reshape(B, -1, 3), (gt * valid).reshape(B, -1, 3)
    tot = 0.0
    for b in range(B):
        d = ((p[b][:, None, :] - t[b][None, :, :]) ** 2).sum(-1)
        tot = tot + d.min(1).values.sum() + d.min(0).values.sum()
    return tot / B


def emd_loss(pred, gt, valid, match_fn):
    """losses_custom.EarthMoversDistance (losses_custom.py:135-166): match = approx_match(p, t) is a constant
    (ops.NoGradient), cost_b = sum_{l,k} match[l,k] * ||t_l - p_k||, sum / B.  match_fn(p, t) -> (B, m, n) supplies the
    matching (the CPU oracle of the op, or the op itself, which is checked against that oracle separately)"""
    B = pred.shape[0]
    p, t = (pred * valid).reshape(B, -1, 3), (gt * valid).reshape(B, -1, 3)
    match = match_fn(p.detach(), t.detach())
    tot = 0.0
    for b in range(B):
        d2 = ((t[b][:, None, :] - p[b][None, :, :]) ** 2).sum(-1)
        # the gradient of the op clamps the squared distance at 1e-20 (tf_approxmatch_g.cu:243,281): sqrt(0) has none
        tot = tot + (match[b].to(d2.dtype) * torch.sqrt(d2.clamp_min(1e-20))).sum()
    return tot / B


def loss(out, S, xyz_loss=("smooth_l1_nonzero", 100.0), match_fn=None):
    """monopsr_model.py:554-958 with the weights of monopsr_model_000.yaml:102-118.  xyz_loss: the yaml entry
    loss_config.inst_xyz_map_local = [type, weight] (loss_builder.py:19-84)."""
    N = NUM_BOXES
    valid = S["gt_valid_mask_maps"]
    L = {}
    kind, weight = xyz_loss
    if kind == "smooth_l1_nonzero":
        xl = smooth_l1_nonzero(out["inst_xyz_map_local"], S["gt_inst_xyz_maps_local"], valid)
    elif kind == "chamfer_dist":
        xl = chamfer_loss(out["inst_xyz_map_local"], S["gt_inst_xyz_maps_local"], valid)
    elif kind == "emd":
        xl = emd_loss(out["inst_xyz_map_local"], S["gt_inst_xyz_maps_local"], valid, match_fn)
    else:
        raise ValueError("Invalid loss type", kind)
    L["inst_xyz_map_local"] = weight * xl / N
    T = regression_targets(out, S)
    L["lwh_offs"] = 1.0 * huber(out["lwh_offs"] - T["lwh_offs"]).sum() / N
    eps = 0.001
    onehot = torch.full((N, 12), eps / 12, dtype=out["alpha_bins"].dtype, device=out["alpha_bins"].device)
    onehot[torch.arange(N, device=onehot.device), S["gt_alpha_bins"].long()] = 1.0 - eps
    logp = torch.log_softmax(out["alpha_bins"], dim=1)
    L["alpha_bins"] = 0.3 * (-(onehot * logp).sum(1)).sum() / N
    L["alpha_regs"] = 1.0 * (huber(out["alpha_regs"] - S["gt_alpha_regs"]) * S["gt_alpha_valid_bins"]).sum() / N
    L["cen_z_offs"] = 0.1 * huber(out["cen_z_offs"] - T["cen_z_offs"]).sum() / N
    L["cen_y_offs"] = 0.1 * huber(out["cen_y_offs"] - T["cen_y_offs"]).sum() / N
    L["proj_err"] = 0.1 * huber(out["proj_err_norm"]).sum() / N        # SUM_BY_NONZERO over 32 ones
    L["inst_depth_map_global"] = 10.0 * smooth_l1_nonzero(out["inst_depth_map_global"],
                                                          S["gt_inst_xyz_maps_global"][..., 2:3], valid) / N
    total = sum(L.values())
    return L, total
