"""GPU tests of the rows either side of the path (inference-mode forward, checkpoint round trip, loader -> engine,
feature-extractor plugin, validation pass).  Written at the end of round 1, first run on a B200 in round 2
(profiles/r2_first_call_summary.txt); plain tests now -- no xfail markers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monopsr_b200.core import model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from oracle import network as onet  # noqa: E402


def test_inference_mode_forward_matches_oracle(cuda):
    """forward(train=False): decoder batch norm with the MOVING statistics (is_training=False graphs)"""
    P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
    eng = Engine(cuda, params=P)
    eng.set_inputs(S)
    eng.forward(train=False)
    o = {k: v.clone() for k, v in eng.outputs().items()}      # outputs() are views of the engine's buffers
    out, _ = onet.forward(onet.to_torch(P, torch.float64, cuda), onet.to_torch(S, torch.float64, cuda), train=False)
    for k in ("centroids", "lwh", "inst_xyz_map_local"):
        a, b = o[k].double().reshape(-1), out[k].reshape(-1)
        assert float((a - b).norm() / b.norm()) < 1e-3, k
    eng.forward(train=True)                       # and the two modes really differ on the decoder output
    assert not torch.allclose(eng.outputs()["inst_xyz_map_local"], o["inst_xyz_map_local"])


def test_checkpoint_roundtrip_through_tf_bundle(cuda, tmp_path):
    """Engine.save_checkpoint -> TensorFlow tensor bundle -> Engine.load_checkpoint restores variables and EMA shadows"""
    a = Engine(cuda, params=ms.init_params(3))
    a.ema.mul_(0.5)                                         # make the shadows differ from the variables
    prefix = str(tmp_path / "model.ckpt-11")
    a.save_checkpoint(prefix, global_step=11)
    b = Engine(cuda, params=ms.init_params(4))
    rep = b.load_checkpoint(prefix)
    assert not rep["missing"] and not rep["shape_mismatch"]
    assert torch.equal(a.params, b.params) and torch.equal(a.state, b.state)
    c = Engine(cuda, params=ms.init_params(4))
    c.load_checkpoint(prefix, use_ema=True)
    assert torch.equal(c.params, a.ema)


def test_kitti_loader_feeds_the_engine(cuda, tmp_path):
    """synthetic KITTI tree -> KittiDataset -> PrefetchLoader -> Engine.train_step: the raw image / depth map /
    instance masks of the sample are turned into crops, the resized image and the ground-truth maps on the GPU"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kitti_tree
    from monopsr_b200.core import targets
    from monopsr_b200.datasets import kitti_loader as KL
    dataset_dir, data_dir = kitti_tree.make_tree(str(tmp_path))
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), dataset_dir, {})
    ds = KL.KittiDataset(cfg, "train", data_dir=data_dir, rng=np.random.RandomState(0))
    eng = Engine(cuda, params=ms.init_params(0))
    with KL.PrefetchLoader(ds, max_samples=3) as loader:
        losses = []
        for sample, sample_dict in loader:
            eng.train_step(sample)
            losses.append(eng.losses()["total_loss"])
            want = targets.image_inputs(sample["rgb_image"], sample["boxes_2d_norm"], cuda)
            assert torch.equal(eng.inputs["rgb_crops"], want["rgb_crops"])
            assert torch.equal(eng.inputs["full_img"], want["full_img"])
            assert eng.inputs["gt_valid_mask_maps"].shape == (32, 48, 48, 1)
            assert 0 < float(eng.inputs["gt_valid_mask_maps"].mean()) < 1
    assert len(losses) == 3 and all(np.isfinite(losses))


def test_feature_extractor_plugin(cuda):
    """net_builder.extract_features returns the same feature maps a full forward pass produces"""
    from monopsr_b200.builders import net_builder as NB
    S = ms.synthetic_sample(0)
    eng = Engine(cuda, params=ms.init_params(0))
    eng.set_inputs(S)
    eng.forward(train=True)
    want_map, want_box = eng.dec[-1]["y"].clone().view(32, 48, 48, 128), eng.pooled.clone().view(32, 6, 6, 512)
    eng2 = Engine(cuda, params=ms.init_params(0))
    model = type("M", (), {"engine": eng2, "boxes_2d_norm": S["boxes_2d_norm"]})()
    f = NB.extract_features(model, "resnet101_4x_squash", None,
                            {NB.NET_IN_RGB_CROP: S["rgb_crops"], NB.NET_IN_FULL_IMG: S["full_img"]}, True)
    torch.cuda.synchronize()
    assert f[NB.FEATURES_FOR_MAP].shape == (32, 48, 48, 128) and f[NB.FEATURES_FOR_BOX_3D].shape == (32, 6, 6, 512)
    assert torch.equal(f[NB.FEATURES_FOR_BOX_3D], want_box)
    assert torch.allclose(f[NB.FEATURES_FOR_MAP], want_map, rtol=1e-4, atol=1e-5)      # (split-K RED order may differ)


def test_validation_pass_on_the_engine(cuda, tmp_path):
    """val split of the synthetic KITTI tree through the real engine: predictions, losses, per-object EMD / Chamfer
    metrics from the point-set ops, KITTI result files and the AP lines"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kitti_tree
    from monopsr_b200.core import evaluator as E
    from monopsr_b200.core import metrics as M
    from monopsr_b200.core import predictions as P
    from monopsr_b200.datasets import kitti_loader as KL
    dataset_dir, data_dir = kitti_tree.make_tree(str(tmp_path / "tree"))
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), dataset_dir,
                                     {"data_split": "val"})
    ds = KL.KittiDataset(cfg, "val", data_dir=data_dir, rng=np.random.RandomState(0))
    eng = Engine(cuda, params=ms.init_params(0))
    base = str(tmp_path / "predictions")
    dirs = {P.OUT_DIR_BOX_2D: base + "/b2", P.OUT_DIR_BOX_3D: base + "/b3", P.OUT_DIR_XYZ_MAP_LOCAL: base + "/xyz"}
    types = [P.KEY_INST_XYZ_MAP_LOCAL, P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG, P.KEY_ALPHA]
    ev = E.Evaluator(eng, types, dirs, train_val_test="val", log=lambda *a: None)
    with KL.PrefetchLoader(ds, epochs=1, workers=2) as loader:
        res = ev.run_checkpoint_once(None, loader)
    assert res["num_samples"] >= 2 and np.isfinite(res["mean_losses"]["total_loss"])
    for k in (M.METRIC_EMD, M.METRIC_CHAMFER, M.METRIC_CEN_Z_ERR, M.METRIC_DIM_ERR):
        assert len(res["metrics"][k]) > 0 and np.all(np.isfinite(res["metrics"][k])), k
    assert min(res["metrics"][M.METRIC_EMD]) >= 0 and min(res["metrics"][M.METRIC_CHAMFER]) >= 0
    out = ev.convert_and_evaluate(ds, base, 0, kitti_score_threshold=0.0)
    assert out and any(l.startswith("car_detection AP:") for l in out[0]["lines"])


def test_inputs_may_not_change_shape_after_capture(cuda):
    """the captured step holds raw pointers to the input buffers: a different geometry afterwards must fail loudly"""
    from monopsr_b200 import lib as mlib
    S = ms.synthetic_sample(0)
    eng = Engine(cuda, params=ms.init_params(0))
    eng.train_step(S)
    eng.train_step(S)                                   # same shapes: copied into the captured buffers
    bad = dict(S)
    bad["full_img"] = np.zeros((1, 80, 304, 3), np.float32)
    with pytest.raises(mlib.MpbError):
        eng.set_inputs(bad)


def test_training_resume_restores_optimizer_state(cuda, tmp_path):
    """save after two steps, resume into a fresh engine: variables, Adam moments and EMA shadows are back bit for bit
    (tf.train.Saver semantics; a resume that zeroed the moments would over-step ~3x on its first updates)"""
    S = ms.synthetic_sample(3)
    a = Engine(cuda, params=ms.init_params(3))
    a.train_step(S)
    a.train_step(S)
    torch.cuda.synchronize()
    prefix = str(tmp_path / "monopsr_model-00000002")
    a.save_checkpoint(prefix, global_step=2)
    b = Engine(cuda, params=ms.init_params(4))
    rep = b.load_checkpoint(prefix, resume=True)
    assert rep["global_step"] == 2 and len(rep["slots"]) == len(b.trainable_names)
    for x, y in ((a.params, b.params), (a.adam_m, b.adam_m), (a.adam_v, b.adam_v), (a.ema, b.ema), (a.state, b.state)):
        assert torch.equal(x, y)
    assert float(b.adam_v.abs().max()) > 0
    c = Engine(cuda, params=ms.init_params(4))
    c.load_checkpoint(prefix)                      # plain restore (evaluation): fresh optimizer state, raw variables
    assert torch.equal(c.params, a.params) and float(c.adam_m.abs().max()) == 0 and torch.equal(c.ema, c.params)
