"""Generate tests/golden/targets_tf_golden.npz: the reference's TENSORFLOW ground-truth target synthesis
(instance_utils.tf_instance_xyz_crop_from_depth_map with depth_map_utils.tf_depth_patch_to_pc_map and
transform_utils.tf_get_tr_mat), unmodified, executed on arrays through the numpy-backed TF stand-in, called per box
with view_norm True / False as MonoPSRModel.build does (monopsr_model.py:165-203).  The one TF kernel involved,
ResizeNearestNeighbor(align_corners=True), is restated in the stand-in from its TF 1.8 definition.
Run from the repository root:  python tests/golden/make_targets_tf_golden.py"""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_tf_numeric as F  # noqa: E402


def inputs():
    """a small scene: depth map with holes, four boxes with rectangular instance masks"""
    rs = np.random.RandomState(41)
    H, W = 120, 400
    depth = np.repeat(np.repeat(rs.uniform(4, 50, (H // 8, W // 8)), 8, 0), 8, 1).astype(np.float32)
    depth[rs.rand(H, W) < 0.2] = 0.0
    boxes_2d = np.array([[20.4, 30.6, 80.5, 130.2], [10.5, 200.5, 60.49, 290.51], [55.0, 150.0, 118.7, 260.3], [30.2, 300.9, 90.8, 395.1]], np.float32)
    masks = np.zeros((4, H, W), np.float32)
    for i, b in enumerate(boxes_2d):
        y1, x1, y2, x2 = np.rint(b).astype(int)
        masks[i, y1 + 3:y2 - 2, x1 + 4:x2 - 3] = 1.0
    boxes_3d = np.column_stack([rs.uniform(-8, 8, 4), rs.uniform(1.2, 1.9, 4), rs.uniform(8, 40, 4), rs.uniform(3, 4.5, 4),
                                rs.uniform(1.5, 1.9, 4), rs.uniform(1.4, 1.7, 4), rs.uniform(-3, 3, 4)]).astype(np.float32)
    view = rs.uniform(-0.6, 0.6, 4).astype(np.float32)
    cam_p = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791], [0.0, 0.0, 1.0, 0.002745884]], np.float32)
    return depth, masks, boxes_2d, boxes_3d, view, cam_p


def main():
    F.install()
    sys.path.insert(0, "/root/reference/src")
    from monopsr.datasets.kitti import instance_utils
    depth, masks, b2, b3, view, cam_p = inputs()
    T = F.t
    depth_batched = T(depth[None, :, :, None].astype(np.float64))            # expand_dims(pl_depth_map, 2) with a batch axis
    out = {}
    for name, view_norm, ctype, rot in (("local", True, "middle", True), ("global", False, "middle", True),
                                        ("local_bottom_norot", True, "bottom", False)):
        xyz, valid = [], []
        for i in range(len(b2)):
            x, v = instance_utils.tf_instance_xyz_crop_from_depth_map(
                i, T(b2.astype(np.float64)), T(b3.astype(np.float64)), T(masks.astype(np.float64)), depth_batched, [48, 48],
                T(view.astype(np.float64)), T(cam_p.astype(np.float64)), view_norm=view_norm, centroid_type=ctype,
                rotate_view=rot)
            xyz.append(np.asarray(x, np.float64)[0])
            valid.append(np.asarray(v, np.float64)[0])
        out["xyz_" + name], out["valid_" + name] = np.stack(xyz), np.stack(valid)
    path = os.path.join(HERE, "targets_tf_golden.npz")
    np.savez_compressed(path, **{k: v.astype(np.float32) for k, v in out.items()})
    print({k: v.shape for k, v in out.items()}, "valid fraction %.2f" % out["valid_local"].mean())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
