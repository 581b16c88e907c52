"""Generate tests/golden/graph_golden.npz: the reference's WHOLE training graph executed end to end on arrays --
MonoPSRModel.__init__ (placeholders, image preprocessing), build (ground-truth maps, crops, both ResNet-101 towers,
crop-and-resize, squash, decoder, FC stacks, box heads, projections) and loss, all unmodified, imported from
/root/reference/src -- through tests/golden/fake_tf_full.py (numeric TF-slim layers with real scoping, TF kernels
supplied by the oracle's primitives) on one seeded synthetic RAW sample (camera image, depth map, instance masks,
labels) and the seeded parameters of model_spec.init_params.  tests/test_graph_golden.py feeds the same sample to
oracle.network.forward / loss and compares every output and loss term: the WIRING of the oracle against the
reference's code.  Run from the repository root:  python tests/golden/make_graph_golden.py   (~1 min)"""
import os
import sys
import types

import numpy as np
import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def raw_sample(seed=3, n=32, H=188, W=621):
    """one KITTI-like raw sample at half resolution"""
    rs = np.random.RandomState(seed)
    img = np.repeat(np.repeat(rs.randint(0, 256, (H // 4, W // 9, 3)), 4, 0), 9, 1)[:H, :W].astype(np.float64)
    depth = np.repeat(np.repeat(rs.uniform(4, 50, (H // 4, W // 9)), 4, 0), 9, 1)[:H, :W]
    depth[rs.rand(H, W) < 0.15] = 0.0
    y1, x1 = rs.uniform(20, 90, n), rs.uniform(5, 420, n)
    b2 = np.column_stack([y1, x1, y1 + rs.uniform(30, 90, n), x1 + rs.uniform(40, 190, n)])
    masks = np.zeros((n, H, W))
    for i, b in enumerate(b2):
        a = np.rint(b).astype(int)
        masks[i, a[0] + 2:a[2] - 2, a[1] + 3:a[3] - 3] = 1.0
    b3 = np.column_stack([rs.uniform(-8, 8, n), rs.uniform(1.2, 1.9, n), rs.uniform(8, 40, n), rs.uniform(3, 4.5, n),
                          rs.uniform(1.5, 1.9, n), rs.uniform(1.4, 1.7, n), rs.uniform(-3, 3, n)])
    nb = 12
    return {
        "rgb_image": img, "depth_map": depth, "instance_masks": masks, "boxes_2d": b2,
        "boxes_2d_norm": b2 / np.array([H, W, H, W], np.float64),
        "cam_p": np.array([[721.5377, 0.0, 304.78, 44.85728], [0.0, 721.5377, 86.4, 0.2163791], [0.0, 0.0, 1.0, 0.002745884]]),
        "class_indices": np.ones((n, 1), np.int32), "mean_lwh": np.tile([3.892, 1.619, 1.530], (n, 1)),
        "prop_cen_z_offset": np.full(n, 2.17799973487854), "est_view_angs": rs.uniform(-0.6, 0.6, n),
        "boxes_3d": b3, "alphas": rs.uniform(-3, 3, n), "alpha_bins": rs.randint(0, nb, n), "alpha_regs": rs.randn(n, nb) * 0.2,
        "alpha_valid_bins": (rs.rand(n, nb) < 0.2).astype(np.float64), "view_angs": rs.uniform(-0.6, 0.6, n), "num_objs": 5,
    }


def main():
    import fake_tf_full as FT
    from monopsr_b200.core import model_spec as ms
    FT.install()
    sys.path.insert(0, "/root/reference/src")
    import monopsr.core.models.monopsr.monopsr_model as mm

    class _NP(object):                         # numpy >= 1.24 refuses the ragged list build() hands to np.asarray
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def asarray(x, *a, **k):
            try:
                return np.asarray(x, *a, **k)
            except ValueError:
                out = np.empty((len(x), len(x[0])), dtype=object)
                for i, row in enumerate(x):
                    for j, v in enumerate(row):
                        out[i, j] = v
                return out
    mm.np = _NP()
    cfg = yaml.safe_load(open("/root/reference/src/monopsr/configs/monopsr_model_000.yaml"))

    def obj(d):
        return types.SimpleNamespace(**{k: obj(v) for k, v in d.items()}) if isinstance(d, dict) else d
    S = raw_sample()
    FT.PARAMS.update(ms.init_params(0, randomize_bn=True))
    feeds = {"rgb_image": [S["rgb_image"]], "cam_p": [S["cam_p"]], "boxes_2d": [S["boxes_2d"]], "boxes_2d_norm": [S["boxes_2d_norm"]],
             "instance_masks": [S["instance_masks"]], "class_strs": [np.full((32, 1), "Car")], "class_indices": [S["class_indices"]],
             "mean_lwh": [S["mean_lwh"]], "prop_cen_z_offset": [S["prop_cen_z_offset"]], "est_view_angs": [S["est_view_angs"]],
             "num_objs": [S["num_objs"]], "depth_map": [S["depth_map"]], "boxes_3d": [S["boxes_3d"]], "alphas": [S["alphas"]],
             "alpha_bins": [S["alpha_bins"]], "alpha_regs": [S["alpha_regs"], S["alpha_valid_bins"]], "view_angs": [S["view_angs"]]}
    FT.FEEDS.update(feeds)
    dcfg = obj(cfg["dataset_config"])
    dataset = types.SimpleNamespace(num_boxes=dcfg.num_boxes, num_alpha_bins=dcfg.num_alpha_bins, centroid_type=dcfg.centroid_type,
                                    dataset_config=dcfg, classes_name="Car", classes=["Car"])
    save = {}
    for mode, prefix in (("train", ""), ("val", "val/")):          # 'val': is_training=False graph, still with losses
        FT.FEEDS.clear()
        FT.FEEDS.update({k: list(v) for k, v in feeds.items()})
        model = mm.MonoPSRModel(obj(cfg["model_config"]), mode, dataset)
        out, gt, _ = model.build()
        losses, total = model.loss(out, gt)
        out = out.dict if hasattr(out, "dict") else out
        for k, v in out.items():
            a = np.asarray(v)
            if a.dtype.kind == "f":
                save[prefix + "out/" + k] = a[:, ::6, ::6] if a.ndim == 4 else a
        for k, v in losses.items():
            save[prefix + "loss/" + k] = np.asarray(v, np.float64)
        save[prefix + "total"] = np.asarray(total, np.float64)
    save["created"] = np.asarray(sorted(set(FT.CREATED)))
    path = os.path.join(HERE, "graph_golden.npz")
    np.savez_compressed(path, **save)
    print({k: v.shape for k, v in save.items() if k.startswith("out/")})
    print({k[5:]: float(v) for k, v in save.items() if k.startswith("loss/")}, float(save["total"]), "val total", float(save["val/total"]))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
