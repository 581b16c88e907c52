"""KITTI loader (monopsr_b200/datasets/kitti_loader.py, augment.py, the filters / readers of kitti_formats.py) against
golden sample dicts produced by the reference's own, unmodified KittiDataset / kitti_aug / obj_utils on the synthetic
tree of tests/kitti_tree.py (tests/golden/make_kitti_loader_golden.py): same seeds -> same oversampling, jittered boxes,
noisy images, merged detections, epoch order.  Value-for-value (bit-exact) comparisons throughout."""
import os
import sys
import threading
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kitti_tree  # noqa: E402
from monopsr_b200.datasets import augment as A  # noqa: E402
from monopsr_b200.datasets import kitti_formats as K  # noqa: E402
from monopsr_b200.datasets import kitti_loader as KL  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "kitti_loader_golden.npz"))


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    return kitti_tree.make_tree(str(tmp_path_factory.mktemp("kitti")))


def _dataset(tree, case):
    mode, overrides = kitti_tree.CASES[case]
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), tree[0], overrides)
    return KL.DatasetBuilder.build_kitti_dataset(cfg, mode, data_dir=tree[1]), mode


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind in "US" or b.dtype.kind in "US":
        assert a.astype(str).tolist() == b.astype(str).tolist(), what
    else:
        assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
        assert np.array_equal(a, b), (what, np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


@pytest.mark.parametrize("case", list(kitti_tree.CASES))
def test_sample_dicts_equal_the_reference_loader(tree, case):
    np.random.seed(1234)
    ds, mode = _dataset(tree, case)
    samples = ds.get_sample_dict(np.arange(ds.num_samples))
    assert len(samples) == int(G[case + "/n"])
    n_none = 0
    for i, s in enumerate(samples):
        got = kitti_tree.summarize(s)
        want = {k.split("/", 2)[2]: G[k] for k in G.files if k.startswith("%s/%d/" % (case, i))}
        assert sorted(got) == sorted(want), (case, i)
        for k in want:
            _same(got[k], want[k], (case, i, k))
        n_none += s is None
    assert 0 < n_none < len(samples) or case == "test"      # the tree exercises the "no usable object" path
    full = [s for s in samples if s is not None]
    if ds.oversample:
        assert all(len(s[KL.SAMPLE_LABEL_BOXES_2D]) == ds.num_boxes for s in full)


@pytest.mark.parametrize("case", list(kitti_tree.CASES))
def test_epoch_bookkeeping_equals_the_reference(tree, case):
    np.random.seed(99)
    ds, mode = _dataset(tree, case)
    for j, (bs, n, idx, epochs) in enumerate(G[case + "/trace"]):
        batch = ds.next_batch(int(bs), shuffle=(mode == "train"))
        assert (len(batch), ds._index_in_epoch, ds.epochs_completed) == (n, idx, epochs)
        assert ["" if s is None else s[KL.SAMPLE_NAME] for s in batch] == G["%s/batch%d/names" % (case, j)].tolist()


def test_reference_unit_test_expectations(tree):
    """the assertions of the reference's kitti_dataset_test.py (split sizes adapted to this tree; invalid splits raise)"""
    cfg = KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN)
    cfg.dataset_dir = tree[0]
    for bad, mode in (("bad", "train"), ("training", "train"), ("validation", "val"), ("testing", "test")):
        cfg.data_split = bad
        with pytest.raises(ValueError):
            KL.KittiDataset(cfg, mode)
    cfg.data_split, cfg.data_split_dir = "train", "nowhere"
    with pytest.raises(ValueError):
        KL.KittiDataset(cfg, "train")
    cfg.dataset_dir = os.path.join(tree[0], "missing")
    with pytest.raises(FileNotFoundError):
        KL.KittiDataset(cfg, "train")
    for split, n in (("train", 7), ("val", 3), ("trainval", 7)):
        c = KL.DatasetBuilder.get_config_obj("kitti_obj_" + split)
        c.dataset_dir = tree[0]
        assert KL.KittiDataset(c, "train" if split != "val" else "val", data_dir=tree[1]).num_samples == n
    c = KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TEST)
    c.dataset_dir = tree[0]
    ds = KL.KittiDataset(c, "test", data_dir=tree[1])
    assert ds.num_samples == 2 and (c.data_split_dir, c.has_kitti_labels) == ("testing", False)
    batch = ds.next_batch(2)
    assert len(batch) == 2 and batch[0].get("label") is None and KL.SAMPLE_LABEL_BOXES_3D not in batch[0]
    with pytest.raises(ValueError):
        KL.DatasetBuilder.get_config_obj("kitti_obj_nope")
    c = KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN)
    c.dataset_dir, c.classes = tree[0], ["Car", "Pedestrian"]
    with pytest.raises(NotImplementedError):
        KL.KittiDataset(c, "train")
    # batch wrapping, as test_batch_wrapping
    ds, _ = _dataset(tree, "train_default")
    assert len(ds.next_batch(7)) == 7 and ds.epochs_completed == 1
    assert len(ds.next_batch(3)) == 3 and ds.epochs_completed == 1
    assert len(ds.next_batch(5)) == 5 and ds.epochs_completed == 2 and ds._index_in_epoch == 1


def test_jitter_needs_oversampling(tree):
    for jt in ("oversample", "oversample_gt"):
        mode, ov = "train", {"aug_config.box_jitter_type": jt, "oversample": False}
        cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), tree[0], ov)
        ds = KL.KittiDataset(cfg, mode, data_dir=tree[1])
        with pytest.raises(ValueError):
            ds.get_sample_dict([1])
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), tree[0],
                                     {"aug_config.box_jitter_type": "sideways"})
    with pytest.raises(ValueError):
        KL.KittiDataset(cfg, "train", data_dir=tree[1]).get_sample_dict([1])
    with pytest.raises(ValueError):
        KL.KittiDataset(cfg, "deploy", data_dir=tree[1]).get_sample_dict([1])


def test_private_generator_is_reproducible_and_leaves_the_global_one_alone(tree):
    np.random.seed(3)
    before = np.random.get_state()[1].copy()
    runs = []
    for _ in range(2):
        mode, ov = kitti_tree.CASES["train_all_noise"]
        cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), tree[0], ov)
        ds = KL.KittiDataset(cfg, mode, data_dir=tree[1], rng=np.random.RandomState(11))
        runs.append([s for s in ds.next_batch(7, shuffle=True) if s is not None])
    assert np.array_equal(np.random.get_state()[1], before)
    assert len(runs[0]) == len(runs[1]) > 0
    for a, b in zip(*runs):
        assert a[KL.SAMPLE_NAME] == b[KL.SAMPLE_NAME]
        assert np.array_equal(a[KL.SAMPLE_LABEL_BOXES_2D], b[KL.SAMPLE_LABEL_BOXES_2D])
        assert np.array_equal(a[KL.SAMPLE_IMAGE_INPUT], b[KL.SAMPLE_IMAGE_INPUT])


# ---------------------------------------------------------------------------------------------- stand-alone pieces
def test_filters_iou_merge_jitter_noise_flips(tree):
    labels = K.read_labels(os.path.join(tree[0], "training", "label_2"), "000008")
    for name, kw in (("hard", dict(difficulty=2)), ("easy", dict(difficulty=0)),
                     ("moderate_h60", dict(difficulty=1, box_2d_height=60)), ("occ2", dict(occlusion=2)),
                     ("trunc", dict(truncation=0.3)), ("depth", dict(depth_range=[5, 15])),
                     ("cars_all", dict(classes=["Car"], difficulty=3))):
        kept, mask = K.filter_labels(labels, **kw)
        assert mask.tolist() == G["filter/" + name].tolist() and len(kept) == int(mask.sum())
    assert K.Difficulty.from_string("hard") == K.Difficulty.HARD == 2 and K.Difficulty.to_string(3) == "all"
    with pytest.raises(KeyError):
        K.Difficulty.from_string("impossible")

    boxes = G["iou/boxes"]
    for b, want, want3 in zip(boxes, G["iou/values"], G["iou/values_rounded"]):
        _same(K.two_d_iou(b, boxes), want, "iou")
        _same(K.two_d_iou(b, boxes, decimals=3), want3, "iou, 3 decimals")
    assert not np.array_equal(G["iou/values"], G["iou/values_rounded"])
    assert K.two_d_iou(np.asarray([0., 0., 1., 1.]), np.asarray([[2., 2., 3., 3.]])).tolist() == [0.0]

    np.random.seed(7)
    jit = A.jitter_obj_boxes_2d(labels, 0.7, (kitti_tree.H, kitti_tree.W))
    _same(np.asarray([[o.x1, o.y1, o.x2, o.y2] for o in jit], np.float64), G["jitter/boxes"], "jitter")
    for o, j in zip(labels, jit):      # only the box moves, by a bounded amount, inside the image
        assert (j.type, j.alpha, j.ry, tuple(j.t)) == (o.type, o.alpha, o.ry, tuple(o.t))
        assert 0 <= j.x1 < j.x2 <= kitti_tree.W - 1 and 0 <= j.y1 < j.y2 <= kitti_tree.H - 1
        assert K.two_d_iou(np.asarray([j.x1, j.y1, j.x2, j.y2]), np.asarray([[o.x1, o.y1, o.x2, o.y2]]))[0] >= 0.7
    tiny = K.ObjectLabel()
    tiny.x1, tiny.y1, tiny.x2, tiny.y2 = 5.0, 5.0, 12.0, 40.0
    assert A.jitter_obj_boxes_2d([tiny], 0.7, (375, 1242))[0] == tiny      # under 10 px wide: left alone

    img = G["noise/image"]
    seen = set()
    for seed in range(12):
        np.random.seed(seed)
        got = A.apply_image_noise(img)
        _same(got, G["noise/%d" % seed], ("noise", seed))
        seen.add(bool(np.array_equal(got, img)))
    assert seen == {True, False}        # some seeds fire no effect, some do

    det = K.read_labels(os.path.join(tree[1], "detections/mscnn/kitti_fmt/val/merged_0.2_0.2_0.2/data"), "000008")
    for st in ("distance", "max", "min"):
        merged = K.merge_kitti_and_mscnn_obj_labels(labels, det, 0.7, default_score_type=st)
        _same(np.asarray([[o.x1, o.y1, o.x2, o.y2, o.score] for o in merged], np.float64), G["merge/" + st], st)
    assert labels[1].score == 0.0       # the inputs are not modified
    with pytest.raises(ValueError):
        K.merge_kitti_and_mscnn_obj_labels(labels, det, 0.7, default_score_type="median")

    _same(A.flip_boxes_3d(G["flip/boxes_3d_in"]), G["flip/boxes_3d"], "flip_boxes_3d")
    _same(A.flip_boxes_3d(G["flip/boxes_3d_in"], flip_ry=False), G["flip/boxes_3d_noflip"], "flip_boxes_3d noflip")
    _same(A.flip_stereo_calib_p2(G["flip/p2_in"], (375, 1242)), G["flip/p2"], "flip p2")
    fl = A.flip_label_in_3d_only(labels[1])
    _same(np.asarray([fl.ry, fl.t[0], fl.t[1], fl.t[2], fl.x1], np.float64), G["flip/label"], "flip label")
    # the reference's own kitti_aug_test.py case
    b = np.array([[1, 2, 3, 4, 5, 6, np.pi / 4], [1, 2, 3, 4, 5, 6, -np.pi / 4]])
    np.testing.assert_almost_equal(A.flip_boxes_3d(b), np.array([[-1, 2, 3, 4, 5, 6, 3 * np.pi / 4],
                                                                  [-1, 2, 3, 4, 5, 6, -3 * np.pi / 4]]))
    pts = np.arange(12.0).reshape(4, 3)
    assert np.array_equal(A.flip_points(pts)[:, 0], -pts[:, 0]) and np.array_equal(A.flip_points(pts)[:, 1:], pts[:, 1:])
    assert np.array_equal(A.flip_point_cloud(pts.T), A.flip_points(pts).T)
    assert np.array_equal(A.flip_image(img), img[:, ::-1]) and A.flip_ground_plane(np.array([1., 2, 3, 4])).tolist() == [-1, 2, 3, 4]


def test_depth_and_instance_readers(tmp_path):
    depth = np.array([[0.0, 0.05, 0.11, 0.5], [1.0, 12.34, 80.0, 255.9]], np.float32)
    path = str(tmp_path / "d.png")
    K.write_depth_map(path, depth)
    got = K.read_depth_map(path)
    assert got.dtype == np.float32 and got[0, 0] == 0 and got[0, 1] == 0        # under 10 cm -> 0
    np.testing.assert_allclose(got[0, 2:], depth[0, 2:], atol=1 / 256)
    np.testing.assert_allclose(got[1], depth[1], atol=1 / 256)
    inst = np.array([[255, 0, 0], [2, 255, 2]], np.uint8)
    masks = K.get_instance_mask_list(inst)
    assert masks.shape == (3, 2, 3) and masks[0].sum() == 2 and masks[1].sum() == 0 and masks[2].sum() == 2
    assert K.get_instance_mask_list(inst, 5).shape == (5, 2, 3)
    assert K.get_instance_mask_list(np.full((2, 2), 255, np.uint8)) == []
    with pytest.raises(FileNotFoundError):
        K.read_depth_map(str(tmp_path / "missing.png"))


# ---------------------------------------------------------------------------------------------- engine side
def test_engine_sample_has_the_engine_keys_and_dtypes(tree):
    np.random.seed(0)
    ds, mode = _dataset(tree, "train_default")
    s = [x for x in ds.get_sample_dict(np.arange(ds.num_samples)) if x is not None][0]
    e = KL.engine_sample(s, "train")
    n = ds.num_boxes
    shapes = dict(rgb_image=(kitti_tree.H, kitti_tree.W, 3), boxes_2d=(n, 4), boxes_2d_norm=(n, 4), cam_p=(3, 4),
                  class_indices=(n, 1), mean_lwh=(n, 3), prop_cen_z_offset=(n,), est_view_angs=(n,), boxes_3d=(n, 7),
                  gt_alphas=(n,), gt_alpha_bins=(n,), gt_alpha_regs=(n, 12), gt_alpha_valid_bins=(n, 12),
                  gt_view_angs=(n,), depth_map=(kitti_tree.H, kitti_tree.W),
                  instance_masks=(n, kitti_tree.H, kitti_tree.W))
    assert sorted(e) == sorted(shapes)
    for k, shp in shapes.items():
        assert e[k].shape == shp and e[k].flags["C_CONTIGUOUS"], k
        want = np.uint8 if k in ("rgb_image", "instance_masks") else np.int32 if k in ("class_indices", "gt_alpha_bins") \
            else np.float32
        assert e[k].dtype == want, (k, e[k].dtype)
    from monopsr_b200.core import model_spec as ms
    syn = ms.synthetic_sample(0)
    for k in e:
        if k in syn:        # same shapes / dtypes as the synthetic sample the engine is benchmarked with
            assert tuple(np.asarray(syn[k]).shape) == e[k].shape, k
    t = KL.engine_sample(_dataset(tree, "test")[0].next_batch(1)[0], "test")
    assert "boxes_3d" not in t and "depth_map" not in t and t["boxes_2d"].shape == (n, 4)
    with pytest.raises(ValueError):
        KL.engine_sample(s, "deploy")


def test_prefetch_loader_order_bound_errors_and_shutdown(tree):
    np.random.seed(5)
    ds, _ = _dataset(tree, "train_default")
    want = []
    while len(want) < 9:            # the synchronous loop of create_feed_dict: batch 1, shuffle, skip empty samples
        s = ds.next_batch(1, shuffle=True)[0]
        if s is not None:
            want.append((s[KL.SAMPLE_NAME], s[KL.SAMPLE_LABEL_BOXES_2D].copy()))
    np.random.seed(5)
    ds, _ = _dataset(tree, "train_default")
    n_threads = threading.active_count()
    with KL.PrefetchLoader(ds, depth=2, max_samples=9) as loader:
        got = [(s[KL.SAMPLE_NAME], e["boxes_2d"]) for e, s in loader]
        assert threading.active_count() == n_threads + 1 or not loader._thread.is_alive()
        with pytest.raises(StopIteration):
            next(loader)
    assert [g[0] for g in got] == [w[0] for w in want]
    for g, w in zip(got, want):
        assert np.array_equal(g[1], w[1])
    time.sleep(0.05)
    assert threading.active_count() == n_threads

    # an endless loader stops producing when closed while blocked on a full queue
    ds, _ = _dataset(tree, "val_kitti")
    loader = KL.PrefetchLoader(ds, depth=1)
    assert loader.shuffle is False
    e, s = next(loader)
    assert s[KL.SAMPLE_NAME] == "000008" and e["depth_map"].dtype == np.float32
    loader.close()
    assert not loader._thread.is_alive()

    # epochs=1: the evaluator's single pass ("while current_epoch == dataset.epochs_completed"), empty samples skipped
    ds, _ = _dataset(tree, "val_kitti")
    with KL.PrefetchLoader(ds, epochs=1) as loader:
        names = [s[KL.SAMPLE_NAME] for _, s in loader]
    # (the split ends with a sample without cars: as in the reference, skipping it runs into the next epoch and the
    # first sample is served a second time before the epoch counter is looked at again)
    assert names == ["000008", "000108", "000008"] and ds.epochs_completed == 1
    ds, _ = _dataset(tree, "test")
    with KL.PrefetchLoader(ds, epochs=2) as loader:
        assert [s[KL.SAMPLE_NAME] for _, s in loader] == ["000008", "000108"] * 2
    ds, _ = _dataset(tree, "train_plain")
    with KL.PrefetchLoader(ds, depth=1) as loader:         # trainer.train's sample_fn
        e = loader.sample_fn()
        assert e["rgb_image"].dtype == np.uint8 and e["boxes_2d"].shape[1] == 4

    # a failure in the producer surfaces in the consumer
    class Broken(object):
        train_val_test = "train"

        def next_batch(self, batch_size, shuffle):
            raise OSError("disk gone")
    with KL.PrefetchLoader(Broken()) as loader:
        with pytest.raises(OSError):
            next(loader)
        with pytest.raises(OSError):
            next(loader)


def test_read_ahead_changes_nothing_but_the_speed(tree):
    """workers > 0: same samples, same order, same values as the single-thread loader; the decoded files come from
    the read-ahead cache; a missing file fails where the direct read would"""
    runs = []
    for workers in (0, 3):
        np.random.seed(21)
        ds, _ = _dataset(tree, "train_all_noise")
        with KL.PrefetchLoader(ds, depth=2, max_samples=10, workers=workers) as loader:
            runs.append([(s[KL.SAMPLE_NAME], e["boxes_2d"], e["rgb_image"], e["depth_map"], e["instance_masks"])
                         for e, s in loader])
            if workers:
                assert loader.read_ahead.hits >= 10 and ds._read_ahead is loader.read_ahead
        assert ds._read_ahead is None
    assert len(runs[0]) == len(runs[1]) == 10
    for a, b in zip(*runs):
        assert a[0] == b[0]
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)
    ds, _ = _dataset(tree, "val_kitti")
    os.rename(ds.depth_dir + "/000108.png", ds.depth_dir + "/000108.png.away")
    try:
        with KL.PrefetchLoader(ds, epochs=1, workers=2) as loader:
            e, s = next(loader)
            assert s[KL.SAMPLE_NAME] == "000008"
            with pytest.raises(FileNotFoundError):
                next(loader)
    finally:
        os.rename(ds.depth_dir + "/000108.png.away", ds.depth_dir + "/000108.png")
