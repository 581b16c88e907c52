"""Generate tests/golden/predictions_golden.npz from the reference's OWN numpy functions (cut out of their source
files with `ast`, because the modules import tensorflow): np_angle_bin_to_orientation, compute_box_3d_corners,
compute_obj_label_corners_3d, project_pc_to_image, postprocess_cen_x, project_to_image_space, score_boxes."""
import ast
import os
import types

import numpy as np

REF = "/root/reference/src/monopsr"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return types.SimpleNamespace(**{n: ns[n] for n in names})


def main():
    calib_utils = cut(os.path.join(REF, "datasets/kitti/calib_utils.py"), ["project_pc_to_image"], {"np": np})
    obj_ns = {"np": np}
    obj_utils = cut(os.path.join(REF, "datasets/kitti/obj_utils.py"),
                    ["ObjectLabel", "compute_box_3d_corners", "compute_obj_label_corners_3d"], obj_ns)
    fc = types.SimpleNamespace(check_box_3d_format=lambda b: None)
    enc = cut(os.path.join(REF, "core/box_3d_encoder.py"), ["box_3d_to_object_label", "boxes_2d_to_iou_fmt"],
              {"np": np, "obj_utils": obj_utils, "fc": fc})
    proj = cut(os.path.join(REF, "core/box_3d_projector.py"), ["project_to_image_space"],
               {"np": np, "format_checker": fc, "box_3d_encoder": enc, "obj_utils": obj_utils, "calib_utils": calib_utils})
    inst = cut(os.path.join(REF, "datasets/kitti/instance_utils.py"), ["postprocess_cen_x"],
               {"np": np, "obj_utils": obj_utils, "calib_utils": calib_utils})
    ori = cut(os.path.join(REF, "core/orientation_encoder.py"), ["np_angle_bin_to_orientation"], {"np": np})
    cam_p = np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]])
    # score_boxes reads the calibration through the dataset: give it one
    ob = cut(os.path.join(REF, "core/models/monopsr/monopsr_output_builder.py"), ["score_boxes"],
             {"np": np, "box_3d_projector": proj, "box_3d_encoder": enc,
              "calib_utils": types.SimpleNamespace(get_frame_calib=lambda d, n: types.SimpleNamespace(p2=cam_p))})

    rng = np.random.RandomState(3)
    n = 24
    z = rng.uniform(4, 70, n)
    boxes_3d = np.stack([rng.uniform(-1, 1, n) * z * 0.6, rng.uniform(1.2, 2.0, n), z, rng.uniform(3.2, 4.6, n),
                         rng.uniform(1.4, 1.9, n), rng.uniform(1.3, 1.8, n), rng.uniform(-np.pi, np.pi, n)], 1)
    boxes_2d = []
    for b in boxes_3d:
        uv = calib_utils.project_pc_to_image(obj_utils.compute_box_3d_corners(b), cam_p)
        jit = rng.uniform(-6, 6, 4)
        boxes_2d.append([uv[1].min() + jit[0], uv[0].min() + jit[1], uv[1].max() + jit[2], uv[0].max() + jit[3]])
    boxes_2d = np.asarray(boxes_2d)
    scores = rng.uniform(0.1, 1.0, (n, 1))
    out = {"cam_p": cam_p, "boxes_3d": boxes_3d, "boxes_2d": boxes_2d, "scores": scores,
           "corners": np.stack([obj_utils.compute_box_3d_corners(b) for b in boxes_3d]),
           "cen_x": np.asarray([np.squeeze(inst.postprocess_cen_x(b2, b3, cam_p)) for b2, b3 in zip(boxes_2d, boxes_3d)]),
           "new_scores": ob.score_boxes(types.SimpleNamespace(calib_dir=""), "000000", (375, 1242), boxes_2d, boxes_3d,
                                        scores),
           "proj_trunc": np.asarray([(lambda r: np.full(4, np.nan) if r is None else r)(
               proj.project_to_image_space(b, cam_p, truncate=True, image_size=(1242, 375))) for b in boxes_3d])}
    bins, res = rng.randint(0, 12, 50), rng.uniform(-1.5, 1.5, 50)
    out.update(ang_bins=bins, ang_res=res,
               ang=np.asarray([ori.np_angle_bin_to_orientation(b, r, 12) for b, r in zip(bins, res)]))
    np.savez_compressed(os.path.join(HERE, "predictions_golden.npz"), **out)
    print("wrote predictions_golden.npz", {k: np.shape(v) for k, v in out.items()},
          "truncated/discarded:", int(np.isnan(out["proj_trunc"][:, 0]).sum()))


if __name__ == "__main__":
    main()
