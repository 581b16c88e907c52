"""Generate tests/golden/targets_golden.npz from the reference's OWN TF-free numpy functions (run in the build
container, where /root/reference exists; the fixture travels, the reference does not).

The reference modules import tensorflow at the top, so the functions are cut out of their source files with `ast`
and executed unmodified in a namespace that only holds numpy:
  depth_patch_to_pc_map        src/monopsr/datasets/kitti/depth_map_utils.py:52-126
  apply_view_norm_to_pc_map    src/monopsr/datasets/kitti/instance_utils.py:512-536
  np_get_tr_mat, pad_pc        src/monopsr/core/transform_utils.py
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/src/monopsr"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing


def main():
    tu = {"np": np}
    cut(os.path.join(REF, "core/transform_utils.py"), ["np_get_tr_mat", "pad_pc"], tu)
    transform_utils = types.SimpleNamespace(**{k: tu[k] for k in ("np_get_tr_mat", "pad_pc")})
    dm = {"np": np}
    cut(os.path.join(REF, "datasets/kitti/depth_map_utils.py"), ["depth_patch_to_pc_map"], dm)
    iu = {"np": np, "transform_utils": transform_utils}
    cut(os.path.join(REF, "datasets/kitti/instance_utils.py"), ["apply_view_norm_to_pc_map"], iu)

    rng = np.random.RandomState(7)
    n, roi = 6, (48, 48)
    cam_p = np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]],
                     np.float32)
    out = {"cam_p": cam_p}
    boxes, patches, pcs, vas, cens, valids, locs = [], [], [], [], [], [], []
    for i in range(n):
        h, w = rng.uniform(30, 150), rng.uniform(40, 250)
        y1, x1 = rng.uniform(100, 370 - h), rng.uniform(5, 1237 - w)
        box = np.array([y1, x1, y1 + h, x1 + w], np.float32)
        patch = rng.uniform(3, 60, roi).astype(np.float32)
        patch[rng.rand(*roi) < 0.3] = 0.0                       # masked-out pixels
        pc = dm["depth_patch_to_pc_map"](patch, box, cam_p, roi, round_box_2d=False, use_pixel_centres=True,
                                         use_corr_factors=False)
        valid = np.abs(patch) >= 0.1
        va = np.float32(rng.uniform(-0.7, 0.7))
        cen = np.array([rng.uniform(-10, 10), rng.uniform(0.5, 2), rng.uniform(5, 50)], np.float32)
        loc = iu["apply_view_norm_to_pc_map"](pc, valid, va, cen, roi)
        boxes.append(box); patches.append(patch); pcs.append(pc); vas.append(va); cens.append(cen)
        valids.append(valid); locs.append(loc)
    out.update(boxes=np.stack(boxes), patches=np.stack(patches), pc_maps=np.stack(pcs).astype(np.float64),
               view_angs=np.array(vas, np.float32), centroids=np.stack(cens), valid=np.stack(valids),
               xyz_local=np.stack(locs).astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "targets_golden.npz"), **out)
    print("wrote targets_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
