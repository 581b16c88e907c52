"""CPU tests of the host-side mirrors: config surface, loss registry, parameter table, the
oracle's TF-op restatements against tiny hand-computed cases and the reference's own KATs."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_parses_and_validates():
    from monopsr_b200.core import config_utils
    cfg = config_utils.parse_yaml_config(os.path.join(ROOT, "configs", "monopsr_model_000.yaml"), data_dir="/tmp/mpb")
    assert cfg.config_name == "monopsr_model_000"
    assert cfg.dataset_config.num_boxes == 32 and cfg.model_config.net_type == "resnet101_4x_squash"
    assert cfg.train_config.optimizer.adam_optimizer.initial_learning_rate == 0.00008
    assert cfg.train_config.paths_config.checkpoint_dir.endswith("/outputs/monopsr_model_000/checkpoints")
    assert config_utils.validate_for_engine(cfg)
    cfg.model_config.net_type = "vgg"
    with pytest.raises(NotImplementedError):
        config_utils.validate_for_engine(cfg)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/monopsr/configs/monopsr_model_000.yaml"),
                    reason="reference checkout not present")
def test_reference_yaml_accepted_unchanged():
    from monopsr_b200.core import config_utils
    ref = config_utils.parse_yaml_config("/root/reference/src/monopsr/configs/monopsr_model_000.yaml", data_dir="/tmp/mpb")
    mine = config_utils.parse_yaml_config(os.path.join(ROOT, "configs", "monopsr_model_000.yaml"), data_dir="/tmp/mpb")
    assert config_utils.validate_for_engine(ref)

    def flat(o, pre=""):
        out = {}
        for k, v in o.__dict__.items():
            if hasattr(v, "__dict__"):
                out.update(flat(v, pre + k + "."))
            else:
                out[pre + k] = v
        return out
    assert flat(ref) == flat(mine)


def test_duplicate_keys_rejected(tmp_path):
    from monopsr_b200.core import config_utils
    p = tmp_path / "bad.yaml"
    p.write_text("a: 1\na: 2\n")
    with pytest.raises(Exception):
        config_utils.parse_yaml_config(str(p))


def test_loss_registry():
    from monopsr_b200.builders import loss_builder
    from monopsr_b200.core import losses_custom
    assert isinstance(loss_builder.build_loss("chamfer_dist"), losses_custom.ChamferDistance)
    assert isinstance(loss_builder.build_loss("emd"), losses_custom.EarthMoversDistance)
    assert isinstance(loss_builder.build_loss("smooth_l1"), loss_builder.FusedLoss)
    with pytest.raises(ValueError):
        loss_builder.build_loss("nope")
    with pytest.raises(NotImplementedError):
        loss_builder.build_loss("focal")


def test_param_table_counts_match_survey():
    from monopsr_b200.core import model_spec as ms
    T = ms.param_table()
    total = sum(int(np.prod(s)) for _, s, _ in T)
    trainable = sum(int(np.prod(s)) for _, s, k in T if k in ms.TRAINABLE_KINDS)
    assert 100.0e6 < trainable < 100.6e6           # ~100.2 M parameters with gradients
    assert total - trainable < 0.3e6
    conv = sum(int(np.prod(s)) for n, s, k in T if k == "weights" and n.startswith("FirstStage"))
    assert abs(conv - 54.9e6) < 0.3e6              # 2 x 27.45 M tower conv weights
    P = ms.init_params(0)
    assert set(P) == {n for n, _, _ in T}
    S = ms.synthetic_sample(0)
    assert S["rgb_crops"].shape == (32, 48, 48, 3) and S["full_img"].shape == (1, 160, 608, 3)


def test_oracle_smooth_l1_and_softmax_kats():
    """known answers of object_detection/core/losses_test.py:83-106 (smooth L1) and :488-520 (softmax CE)"""
    from oracle import network as onet
    pred = torch.tensor([[[2.5, 0, .4, 0], [0, 0, 0, 0], [0, 2.5, 0, .4]], [[3.5, 0, 0, 0], [0, .4, 0, .9], [0, 0, 1.5, 0]]],
                        dtype=torch.float64)
    tgt = torch.zeros_like(pred)
    w = torch.tensor([[2, 1, 1], [0, 3, 0]], dtype=torch.float64)
    loss = (onet.huber(pred - tgt).sum(2) * w).sum()
    assert abs(float(loss) - 7.695) < 1e-6
    logits = torch.tensor([[[-100, 100, -100], [100, -100, -100], [0, 0, -100], [-100, -100, 100]],
                           [[-100, 0, 0], [-100, 100, -100], [-100, 100, -100], [100, -100, -100]]], dtype=torch.float64)
    target = torch.tensor([[[0, 1, 0], [1, 0, 0], [1, 0, 0], [0, 0, 1]], [[0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0]]],
                          dtype=torch.float64)
    weights = torch.tensor([[1, 1, .5, 1], [1, 1, 1, 0]], dtype=torch.float64)
    ce = -(target * torch.log_softmax(logits, dim=2)).sum(2) * weights
    assert abs(float(ce.sum()) - (-1.5 * np.log(.5))) < 1e-6


def test_oracle_tf_op_restatements_small_cases():
    from oracle import network as onet
    # conv2d_same stride 2 on an even input equals SAME stride-1 conv subsampled (resnet_utils docstring)
    x = torch.randn(1, 8, 8, 3, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 4, dtype=torch.float64)
    a = onet.conv2d_same(x, w, 2)
    b = onet.conv_hwio(x, w)[:, ::2, ::2]
    assert torch.allclose(a, b, atol=1e-12)
    # 3x3/2 SAME max pool on an even input: windows 2o..2o+2 clipped at the edge
    x = torch.arange(16, dtype=torch.float64).reshape(1, 4, 4, 1)
    p = onet.max_pool_same_3x3_s2(x)[0, :, :, 0]
    assert p.tolist() == [[10., 11.], [14., 15.]]
    # crop_and_resize: identity box reproduces the image at matching size; outside -> 0
    img = torch.arange(12, dtype=torch.float64).reshape(1, 3, 4, 1)
    c = onet.crop_and_resize(img, torch.tensor([[0., 0., 1., 1.]], dtype=torch.float64), 3, 4)
    assert torch.allclose(c[0], img[0])
    c = onet.crop_and_resize(img, torch.tensor([[0., 0., 1.5, 1.]], dtype=torch.float64), 3, 4)
    assert float(c[0, 2].abs().sum()) == 0.0 and torch.allclose(c[0, 0], img[0, 0])
    # align_corners resize keeps the corner pixels
    r = onet.resize_bilinear_ac(img, 6, 8)
    assert float(r[0, 0, 0, 0]) == 0.0 and float(r[0, -1, -1, 0]) == 11.0
    # tf32 emulation helper: 10 mantissa bits, ties away from zero
    onet.EMULATE_TF32 = True
    try:
        q = onet.Q(torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -12, -1.0 - 2 ** -11], dtype=torch.float64))
    finally:
        onet.EMULATE_TF32 = False
    assert q.tolist() == [1.0 + 2 ** -10, 1.0, -1.0 - 2 ** -10]


def test_geometry_restatement_vs_numpy_twins():
    """the reference pins its TF geometry helpers to numpy twins (instance_utils_test.py:27-73);
    restate the numpy twins here (instance_utils.py:552-564,684-735) and compare with the oracle."""
    rng = np.random.RandomState(0)
    boxes = np.array([[100., 200., 180., 330.]])
    # expected uv map at pixel centres: linspace(start+half, stop-half, 48), meshgrid 'xy'
    v1, u1, v2, u2 = boxes[0]
    hu, hv = (u2 - u1) / 48 / 2, (v2 - v1) / 48 / 2
    gu, gv = np.linspace(u1 + hu, u2 - hu, 48), np.linspace(v1 + hv, v2 - hv, 48)
    U, V = np.meshgrid(gu, gv)
    lin = torch.arange(48, dtype=torch.float64) / 47.0
    grid_u = (u1 + hu) + ((u2 - hu) - (u1 + hu)) * lin
    assert np.allclose(grid_u.numpy(), gu) and np.allclose(U[5], gu) and np.allclose(V[:, 7], gv)
    # local -> global: rotate about y by the view angle, then translate
    pts = rng.randn(3, 10)
    ang, cen = 0.3, np.array([1.0, 2.0, 20.0])
    rot = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    glob = rot @ pts + cen[:, None]
    gx = np.cos(ang) * pts[0] + np.sin(ang) * pts[2] + cen[0]
    gz = -np.sin(ang) * pts[0] + np.cos(ang) * pts[2] + cen[2]
    assert np.allclose(glob[0], gx) and np.allclose(glob[2], gz)


def test_oracle_inference_mode_batch_norm():
    """is_training=False (validation / inference graphs, monopsr_model.py:139): moving statistics, eps 1e-3, beta only"""
    import torch
    from oracle import network as onet
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 6, 6, 8, generator=g, dtype=torch.float64) * 3 + 1
    P = {"s/moving_mean": torch.randn(8, generator=g, dtype=torch.float64),
         "s/moving_variance": torch.rand(8, generator=g, dtype=torch.float64) + 0.5,
         "s/beta": torch.randn(8, generator=g, dtype=torch.float64)}
    y, m, v = onet.infer_bn_relu(x, P, "s")
    ref = torch.relu((x - P["s/moving_mean"]) / torch.sqrt(P["s/moving_variance"] + 1e-3) + P["s/beta"])
    assert torch.allclose(y, ref, rtol=1e-12, atol=1e-12) and m is P["s/moving_mean"]
    # with the batch's own statistics as "moving" statistics the two modes coincide
    P["s/moving_mean"], P["s/moving_variance"] = x.mean((0, 1, 2)), x.var((0, 1, 2), unbiased=False)
    yt, _, _ = onet.train_bn_relu(x, P, "s")
    yi, _, _ = onet.infer_bn_relu(x, P, "s")
    assert torch.allclose(yt, yi, rtol=1e-10, atol=1e-10)


def test_net_builder_plugin_boundary():
    """extract_features: only resnet101_4x_squash, inputs by the reference's keys, the two feature maps back"""
    import types
    from monopsr_b200.builders import net_builder as NB

    class Eng(object):
        def set_inputs(self, S):
            self.S = S

        def forward(self, train=True, compute_losses=None, features_only=False):
            self.args = (train, features_only)
            return {"features_for_map": "MAP", "features_for_box_3d": "BOX"}
    eng = Eng()
    model = types.SimpleNamespace(engine=eng, boxes_2d_norm="B")
    out = NB.extract_features(model, "resnet101_4x_squash", None, {NB.NET_IN_RGB_CROP: "C", NB.NET_IN_FULL_IMG: "F"}, False)
    assert out == {NB.FEATURES_FOR_MAP: "MAP", NB.FEATURES_FOR_BOX_3D: "BOX"}
    assert eng.S == {"rgb_crops": "C", "full_img": "F", "boxes_2d_norm": "B"} and eng.args == (False, True)
    NB.extract_features(eng, "resnet101_4x_squash", None, {NB.NET_IN_RGB_CROP: "C", NB.NET_IN_FULL_IMG: "F"}, True)
    assert eng.args == (True, True) and "boxes_2d_norm" not in eng.S
    with pytest.raises(ValueError):
        NB.extract_features(model, "vgg16", None, {}, True)
    cfg = types.SimpleNamespace(net_type="resnet101_4x_squash",
                                net_config=types.SimpleNamespace(resnet101_4x_squash="EXTRACTOR"))
    assert NB.get_net_config(cfg) == "EXTRACTOR"


def test_oracle_crop_and_resize_against_grid_sample():
    """an independent implementation of the same sampling rule: for boxes inside the image, tf.image.crop_and_resize
    is bilinear sampling at x = x1 (W-1) + i (x2-x1)(W-1)/(cw-1), i.e. torch's grid_sample(align_corners=True)"""
    import torch.nn.functional as F
    from oracle import network as onet
    g = torch.Generator().manual_seed(3)
    img = torch.randn(1, 40, 152, 5, generator=g, dtype=torch.float64)
    lo = torch.rand(9, 2, generator=g, dtype=torch.float64) * 0.6
    boxes = torch.cat([lo, lo + 0.05 + torch.rand(9, 2, generator=g, dtype=torch.float64) * 0.35], dim=1)   # y1 x1 y2 x2 in [0,1]
    ch, cw = 24, 24
    got = onet.crop_and_resize(img, boxes, ch, cw)
    iy = torch.arange(ch, dtype=torch.float64) / (ch - 1)
    ix = torch.arange(cw, dtype=torch.float64) / (cw - 1)
    gy = (boxes[:, 0:1] + iy[None, :] * (boxes[:, 2:3] - boxes[:, 0:1])) * 2 - 1           # (N,ch) in [-1,1]
    gx = (boxes[:, 1:2] + ix[None, :] * (boxes[:, 3:4] - boxes[:, 1:2])) * 2 - 1
    grid = torch.stack([gx[:, None, :].expand(-1, ch, -1), gy[:, :, None].expand(-1, -1, cw)], dim=-1)
    want = F.grid_sample(img.permute(0, 3, 1, 2).expand(9, -1, -1, -1), grid, mode="bilinear", align_corners=True)
    assert torch.allclose(got, want.permute(0, 2, 3, 1), rtol=1e-10, atol=1e-12)


def test_oracle_bottleneck_and_stem_against_torchvision():
    """an independent implementation of the standard blocks: torchvision's Bottleneck (eval-mode batch norm = the frozen
    affine of the towers, eps 1e-5; 3x3 conv dilated by the atrous rate) and its 7x7/2 stem with padding 3, which is
    what conv2d_same pads explicitly (resnet_utils.py:111-122) -- fed with the same random weights as the oracle"""
    tv = pytest.importorskip("torchvision.models.resnet")
    from oracle import network as onet
    g = torch.Generator().manual_seed(11)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)

    def bn_params(P, scope, c):
        P[scope + "/gamma"], P[scope + "/beta"] = rnd(c) * 0.3 + 1, rnd(c) * 0.2
        P[scope + "/moving_mean"], P[scope + "/moving_variance"] = rnd(c) * 0.1, torch.rand(c, generator=g, dtype=torch.float64) + 0.5

    def load_bn(bn, P, scope):
        bn.weight.data, bn.bias.data = P[scope + "/gamma"].clone(), P[scope + "/beta"].clone()
        bn.running_mean.data, bn.running_var.data = P[scope + "/moving_mean"].clone(), P[scope + "/moving_variance"].clone()

    for cin, base, rate, proj in ((256, 64, 1, False), (256, 128, 2, True), (1024, 256, 4, False)):
        cout = base * 4
        P, s = {}, "enc/resnet_v1_101/blockX/unit_1"
        b = s + "/bottleneck_v1"
        shapes = {"conv1": (1, 1, cin, base), "conv2": (3, 3, base, base), "conv3": (1, 1, base, cout)}
        if proj:
            shapes["shortcut"] = (1, 1, cin, cout)
        for name, shp in shapes.items():
            P["%s/%s/weights" % (b, name)] = rnd(*shp) / (shp[0] * shp[1] * shp[2]) ** 0.5          # HWIO
            bn_params(P, "%s/%s/BatchNorm" % (b, name), shp[3])
        down = None
        if proj:
            down = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 1, bias=False), torch.nn.BatchNorm2d(cout))
        m = tv.Bottleneck(cin, base, stride=1, downsample=down, dilation=rate).double().eval()
        for conv, bn, name in ((m.conv1, m.bn1, "conv1"), (m.conv2, m.bn2, "conv2"), (m.conv3, m.bn3, "conv3")):
            conv.weight.data = P["%s/%s/weights" % (b, name)].permute(3, 2, 0, 1).clone()           # HWIO -> OIHW
            load_bn(bn, P, "%s/%s/BatchNorm" % (b, name))
        if proj:
            down[0].weight.data = P[b + "/shortcut/weights"].permute(3, 2, 0, 1).clone()
            load_bn(down[1], P, b + "/shortcut/BatchNorm")
        x = torch.relu(rnd(2, 12, 12, cin))
        with torch.no_grad():
            want = m(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            got = onet.bottleneck(x, P, s, cout, base, rate)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-9, atol=1e-9), (cin, base, rate)

    # stem: 7x7 stride 2 with explicit padding 3, frozen BN, ReLU (the pooling that follows differs: SAME vs padding 1)
    P = {"enc/resnet_v1_101/conv1/weights": rnd(7, 7, 3, 64) / 12.0}
    bn_params(P, "enc/resnet_v1_101/conv1/BatchNorm", 64)
    net = tv.ResNet(tv.Bottleneck, [1, 1, 1, 1]).double().eval()
    net.conv1.weight.data = P["enc/resnet_v1_101/conv1/weights"].permute(3, 2, 0, 1).clone()
    load_bn(net.bn1, P, "enc/resnet_v1_101/conv1/BatchNorm")
    x = rnd(2, 48, 48, 3) * 50
    with torch.no_grad():
        want = net.relu(net.bn1(net.conv1(x.permute(0, 3, 1, 2)))).permute(0, 2, 3, 1)
        got = torch.relu(onet.conv_bn(x, P, "enc/resnet_v1_101/conv1", 2))
    assert torch.allclose(got, want, rtol=1e-9, atol=1e-9)
    # SAME 3x3/2 max pooling on an even grid = windows starting at 0, 2, 4, ... clipped at the far edge
    pooled = onet.max_pool_same_3x3_s2(got)
    want_pool = torch.nn.functional.max_pool2d(got.permute(0, 3, 1, 2), 3, 2, ceil_mode=True).permute(0, 2, 3, 1)
    assert pooled.shape == (2, 12, 12, 64) and torch.equal(pooled, want_pool)


def test_oracle_batch_norm_and_huber_against_torch_functional():
    """slim.batch_norm defaults (no gamma, eps 1e-3, biased batch variance / moving statistics) and the Huber loss
    (delta 1) against torch's own functional implementations"""
    import torch.nn.functional as F
    from oracle import network as onet
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 6, 6, 16, generator=g, dtype=torch.float64) * 2 + 0.5
    P = {"bn/beta": torch.randn(16, generator=g, dtype=torch.float64),
         "bn/moving_mean": torch.randn(16, generator=g, dtype=torch.float64) * 0.3,
         "bn/moving_variance": torch.rand(16, generator=g, dtype=torch.float64) + 0.5}
    y, mean, var = onet.train_bn_relu(x, P, "bn")
    want = F.relu(F.batch_norm(x.permute(0, 3, 1, 2), None, None, None, P["bn/beta"], True, 0.0, 1e-3)).permute(0, 2, 3, 1)
    assert torch.allclose(y, want, rtol=1e-10, atol=1e-12)
    assert torch.allclose(mean, x.mean((0, 1, 2))) and torch.allclose(var, x.var((0, 1, 2), unbiased=False))
    y, _, _ = onet.infer_bn_relu(x, P, "bn")
    want = F.relu(F.batch_norm(x.permute(0, 3, 1, 2), P["bn/moving_mean"], P["bn/moving_variance"], None, P["bn/beta"],
                               False, 0.0, 1e-3)).permute(0, 2, 3, 1)
    assert torch.allclose(y, want, rtol=1e-10, atol=1e-12)
    d = torch.randn(1000, generator=g, dtype=torch.float64) * 2
    assert torch.allclose(onet.huber(d), F.smooth_l1_loss(d, torch.zeros_like(d), reduction="none", beta=1.0))
    assert torch.allclose(onet.huber(d), F.huber_loss(d, torch.zeros_like(d), reduction="none", delta=1.0))
