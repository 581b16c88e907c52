"""Generate tests/golden/kitti_formats_golden.npz: the reference's OWN parsers / encoders (cut out of their modules
with `ast`, which import tensorflow) run on three samples of the reference's KITTI test fixture
(src/monopsr/tests/datasets/Kitti/object/training; the three label/calib text files are copied to tests/golden/kitti)."""
import ast
import csv
import os
import types

import numpy as np

REF = "/root/reference/src/monopsr"
HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLES = ["000001", "000008", "000076"]


def cut(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    assert all(n in ns for n in names), [n for n in names if n not in ns]
    return types.SimpleNamespace(**{n: ns[n] for n in names})


def main():
    fc = types.SimpleNamespace(check_box_3d_format=lambda b: None, check_obj_label_format=lambda o: None)
    calib = cut(os.path.join(REF, "datasets/kitti/calib_utils.py"), ["FrameCalib", "read_frame_calib", "project_pc_to_image"],
                {"np": np, "csv": csv})
    obj = cut(os.path.join(REF, "datasets/kitti/obj_utils.py"),
              ["ObjectLabel", "read_labels", "filter_labels_by_class", "get_viewing_angle_box_2d",
               "get_viewing_angle_box_3d", "get_mean_lwh_and_std_dev", "class_str_to_index"],
              {"np": np, "os": os, "format_checker": fc, "calib_utils": calib})
    enc = cut(os.path.join(REF, "core/box_3d_encoder.py"), ["object_label_to_box_2d", "object_label_to_box_3d"],
              {"np": np, "fc": fc})
    ori = cut(os.path.join(REF, "core/orientation_encoder.py"), ["np_orientation_to_angle_bin"], {"np": np})
    inst = cut(os.path.join(REF, "datasets/kitti/instance_utils.py"), ["get_prop_cen_z_offset"], {"np": np})
    out = {}
    for s in SAMPLES:
        labels = obj.read_labels(os.path.join(HERE, "kitti/label_2"), s)
        c = calib.read_frame_calib(os.path.join(HERE, "kitti/calib", s + ".txt"))
        out[s + "_p2"], out[s + "_r0"], out[s + "_v2c"] = c.p2, c.r0_rect, c.velo_to_cam
        out[s + "_types"] = np.asarray([o.type for o in labels])
        out[s + "_raw"] = np.asarray([[o.truncation, o.occlusion, o.alpha, o.x1, o.y1, o.x2, o.y2, o.h, o.w, o.l,
                                       o.t[0], o.t[1], o.t[2], o.ry, o.score] for o in labels], np.float64)
        cars, mask = obj.filter_labels_by_class(labels, ["Car"])
        out[s + "_car_mask"] = np.asarray(mask)
        b2 = np.asarray([enc.object_label_to_box_2d(o) for o in cars])
        b3 = np.asarray([enc.object_label_to_box_3d(o) for o in cars])
        out[s + "_boxes_2d"], out[s + "_boxes_3d"] = b2, b3
        out[s + "_va2d"] = np.asarray([obj.get_viewing_angle_box_2d(b, c.p2) for b in b2])
        out[s + "_va3d"] = np.asarray([obj.get_viewing_angle_box_3d(b, c.p2) for b in b3])
        out[s + "_va3d_proj"] = np.asarray([obj.get_viewing_angle_box_3d(b, c.p2, version="projection") for b in b3])
        bins = [ori.np_orientation_to_angle_bin(o.alpha, 12, 0.0) for o in cars]
        out[s + "_bins"] = np.asarray([b[0] for b in bins])
        out[s + "_regs"] = np.asarray([b[1] for b in bins])
        out[s + "_valid"] = np.asarray([b[2] for b in bins])
    angs = np.linspace(-7, 7, 57)
    ov = [ori.np_orientation_to_angle_bin(a, 8, 0.2) for a in angs]
    out.update(ov_angles=angs, ov_bins=np.asarray([b[0] for b in ov]), ov_regs=np.asarray([b[1] for b in ov]),
               ov_valid=np.asarray([b[2] for b in ov]))
    out["mean_lwh"] = np.asarray([obj.get_mean_lwh_and_std_dev(c)[0] for c in ("Car", "Pedestrian", "Cyclist")])
    out["prop_off"] = np.asarray([inst.get_prop_cen_z_offset(c) for c in ("Car", "Pedestrian", "Cyclist")])
    np.savez_compressed(os.path.join(HERE, "kitti_formats_golden.npz"), **out)
    print("wrote kitti_formats_golden.npz:", {s: int(len(out[s + "_boxes_2d"])) for s in SAMPLES}, "cars")


if __name__ == "__main__":
    main()
