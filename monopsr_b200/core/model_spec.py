"""Parameter table and synthetic feed of ``monopsr_model_000`` (names follow the TF variable scopes).

Reference: the variables created by builders/net_builder.py:30-96 (two
``FirstStageFeatureExtractor_{crop,full}/resnet_v1_101`` towers, ``squash``, ``map_decoder``),
core/models/monopsr/monopsr_output_builder.py (``output/...`` heads) and the placeholders of
core/models/monopsr/monopsr_model.py:70-126.  Public parameter layout
is TensorFlow's (conv HWIO, fully-connected [in,out]) so checkpoints map 1:1; the engine
converts to its own OHWI / padded-K device layout.
"""
import math

import numpy as np

NUM_BOXES = 32
CROP = 48
FULL_H, FULL_W = 160, 608
IMG_H, IMG_W = 320, 1216
BLOCKS = [("block1", 64, 3, 1), ("block2", 128, 4, 2), ("block3", 256, 23, 4)]   # name, base depth, units, atrous rate
ENCODERS = ("FirstStageFeatureExtractor_crop", "FirstStageFeatureExtractor_full")
KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]], np.float32)   # tests/datasets/Kitti/object/training/calib/000000.txt P2


def conv_layers(enc):
    """(scope, kh, cin, cout, rate) of every conv+frozen-BN of one encoder, in execution order."""
    s = enc + "/resnet_v1_101"
    out = [(s + "/conv1", 7, 3, 64, 1)]
    cin = 64
    for name, base, units, rate in BLOCKS:
        for u in range(1, units + 1):
            p = "%s/%s/unit_%d/bottleneck_v1" % (s, name, u)
            if cin != base * 4:
                out.append((p + "/shortcut", 1, cin, base * 4, 1))
            out.append((p + "/conv1", 1, cin, base, 1))
            out.append((p + "/conv2", 3, base, base, rate))
            out.append((p + "/conv3", 1, base, base * 4, 1))
            cin = base * 4
    return out


def param_table():
    """list of (name, shape, kind); kind in weights|gamma|beta|moving_mean|moving_variance|biases."""
    T = []
    for enc in ENCODERS:
        for scope, k, cin, cout, _ in conv_layers(enc):
            T.append((scope + "/weights", (k, k, cin, cout), "weights"))
            for n in ("gamma", "beta", "moving_mean", "moving_variance"):
                T.append((scope + "/BatchNorm/" + n, (cout,), n))
    T.append(("squash/1x1_conv/weights", (1, 1, 2048, 512), "weights"))
    T.append(("squash/1x1_conv/biases", (512,), "biases"))
    for blk, cin, cout in (("conv2", 512, 256), ("conv3", 256, 128)):
        for i in (1, 2):
            sc = "map_decoder/%s/%s_%d" % (blk, blk, i)
            T.append((sc + "/weights", (3, 3, cin if i == 1 else cout, cout), "weights"))
            T.append((sc + "/BatchNorm/beta", (cout,), "beta"))
            T.append((sc + "/BatchNorm/moving_mean", (cout,), "moving_mean"))
            T.append((sc + "/BatchNorm/moving_variance", (cout,), "moving_variance"))
    sc = "output/inst_xyz_map_local/inst_xyz_map_local"
    T.append((sc + "/weights", (3, 3, 128, 3), "weights"))
    T.append((sc + "/biases", (3,), "biases"))
    for p, cat in (("output/proposal_fc/proposal_fc", 1043), ("output/regression_fc/regression_fc", 1060)):
        for name, kin in (("img_fc", 18432), ("fc0", cat), ("fc1", 1024)):
            T.append(("%s/%s/weights" % (p, name), (kin, 1024), "weights"))
            T.append(("%s/%s/biases" % (p, name), (1024,), "biases"))
    for sc, n in (("output/lwh/lwh", 3), ("output/alpha", 24), ("output/cen_y/cen_y", 1),
                  ("output/cen_z_offs/cen_z", 1)):
        T.append((sc + "/weights", (1024, n), "weights"))
        T.append((sc + "/biases", (n,), "biases"))
    return T


TRAINABLE_KINDS = ("weights", "gamma", "beta", "biases")


def init_params(seed=0, randomize_bn=False, tame=True):
    """Seeded random initialisation: variance-scaling (fan-in, factor 2) for the ResNet convs
    (resnet_utils.py:260), Xavier-uniform for everything else (slim default), zero biases, BN
    gamma=1 beta=0 mean=0 var=1 -- or randomised BN statistics for parity tests.

    tame=True keeps activations O(1) the way a trained checkpoint does (a randomly initialised
    inference-mode-BN ResNet-101 doubles its variance in each of the 33 units): the stem is
    scaled for +-50 pixel inputs and the last BN gamma of every unit is ~0.25."""
    rng = np.random.RandomState(seed)
    P = {}
    for name, shape, kind in param_table():
        if kind == "weights":
            if len(shape) == 4:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            else:
                fan_in, fan_out = shape
            if name.startswith("FirstStage"):
                w = rng.standard_normal(shape) * math.sqrt(2.0 / fan_in) * 0.8796   # truncated-normal correction
                w = np.clip(w, -2 * math.sqrt(2.0 / fan_in), 2 * math.sqrt(2.0 / fan_in))
            else:
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                w = rng.uniform(-lim, lim, shape)
            if tame and name.endswith("resnet_v1_101/conv1/weights"):
                w = w * 0.02
            P[name] = w.astype(np.float32)
        elif kind in ("gamma", "moving_variance"):
            v = rng.uniform(0.6, 1.4, shape) if randomize_bn else np.ones(shape)
            if tame and kind == "gamma" and "/conv3/BatchNorm" in name:
                v = v * 0.25
            P[name] = v.astype(np.float32)
        elif kind in ("beta", "moving_mean"):
            P[name] = (rng.standard_normal(shape) * 0.1 if randomize_bn else np.zeros(shape)).astype(np.float32)
        else:
            P[name] = (rng.standard_normal(shape) * 0.05 if randomize_bn else np.zeros(shape)).astype(np.float32)
    return P


def synthetic_sample(seed=0, num_boxes=NUM_BOXES):
    """A valid feed without KITTI: one image worth of crops + targets.
    All float arrays are float32; shapes follow monopsr_model.py:70-126."""
    rng = np.random.RandomState(seed + 1000)
    N = num_boxes
    S = {}
    S["rgb_crops"] = (rng.standard_normal((N, CROP, CROP, 3)) * 50).astype(np.float32)     # mean-subtracted pixel scale
    S["full_img"] = (rng.standard_normal((1, FULL_H, FULL_W, 3)) * 50).astype(np.float32)
    # boxes inside a 375 x 1242 KITTI image, [y1,x1,y2,x2] px
    h = rng.uniform(30, 150, N)
    w = rng.uniform(40, 250, N)
    y1 = rng.uniform(100, 370 - h)
    x1 = rng.uniform(5, 1237 - w)
    boxes = np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)
    S["boxes_2d"] = boxes
    S["boxes_2d_norm"] = (boxes / np.array([375, 1242, 375, 1242], np.float32)).astype(np.float32)
    S["cam_p"] = KITTI_P2.copy()
    S["class_indices"] = np.ones((N, 1), np.int32)                  # 'Car' -> index+1 (obj_utils.py:1110-1127)
    S["mean_lwh"] = np.tile(np.array([[3.892, 1.619, 1.530]], np.float32), (N, 1))
    S["prop_cen_z_offset"] = np.full((N,), 2.17799973487854, np.float32)
    xc = (boxes[:, 1] + boxes[:, 3]) / 2
    S["est_view_angs"] = np.arctan2((xc - KITTI_P2[0, 2]) / KITTI_P2[0, 0], 1.0).astype(np.float32)
    z = rng.uniform(5, 45, N)
    x3 = z * np.tan(S["est_view_angs"]) - KITTI_P2[0, 3] / KITTI_P2[0, 0]
    y3 = rng.uniform(1.0, 2.0, N)
    lwh = S["mean_lwh"] + rng.standard_normal((N, 3)).astype(np.float32) * 0.2
    ry = rng.uniform(-np.pi, np.pi, N)
    S["boxes_3d"] = np.stack([x3, y3, z, lwh[:, 0], lwh[:, 1], lwh[:, 2], ry], 1).astype(np.float32)
    S["gt_alpha_bins"] = rng.randint(0, 12, N).astype(np.int32)
    S["gt_alpha_regs"] = (rng.standard_normal((N, 12)) * 0.2).astype(np.float32)
    valid_bins = np.zeros((N, 12), np.float32)
    valid_bins[np.arange(N), S["gt_alpha_bins"]] = 1.0
    valid_bins[np.arange(N), (S["gt_alpha_bins"] + 1) % 12] = 1.0
    S["gt_alpha_valid_bins"] = valid_bins
    S["gt_view_angs"] = (S["est_view_angs"] + rng.standard_normal(N) * 0.01).astype(np.float32)
    S["gt_inst_xyz_maps_local"] = rng.uniform(-2, 2, (N, CROP, CROP, 3)).astype(np.float32)
    g = S["gt_inst_xyz_maps_local"].copy()
    g[..., 2] += z[:, None, None]
    S["gt_inst_xyz_maps_global"] = g.astype(np.float32)
    S["gt_valid_mask_maps"] = (rng.rand(N, CROP, CROP, 1) < 0.6).astype(np.float32)
    return S
