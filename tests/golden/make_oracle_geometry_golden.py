"""Generate tests/golden/oracle_geometry_golden.npz: the reference's TF-free numpy twins of the graph's geometry ops
(imported from /root/reference/src behind the tensorflow / pypng stub hook) on seeded inputs:
  inst_points_local_to_global   instance_utils.py:552-564     (twin of tf_inst_xyz_map_local_to_global :567-604)
  project_pc_to_image           calib_utils.py:245-260        (twin of tf_project_pc_to_image :263-280)
  get_exp_proj_uv_map           instance_utils.py:684-735     (twin of tf_get_exp_proj_uv_map :738-788)
  est_y_from_box_2d_and_depth   instance_utils.py:841-904     (twin of tf_est_y_from_box_2d_and_depth :907-953)
  np_get_tr_mat                 transform_utils.py:6-33
The reference's own tests (instance_utils_test.py:27-73, transform_utils_test.py:39-100) assert the TF ops equal these.
Run from the repository root:  python tests/golden/make_oracle_geometry_golden.py"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return self


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ("tensorflow", "png"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, "/root/reference/src")
    from monopsr.core import transform_utils
    from monopsr.datasets.kitti import calib_utils, instance_utils

    rs = np.random.RandomState(42)
    n, roi = 6, 48
    cam_p = calib_utils.get_frame_calib(os.path.join(HERE, "kitti", "calib"), "000008").p2
    xyz_local = (rs.randn(n, roi, roi, 3) * [1.8, 0.7, 0.9]).astype(np.float32).astype(np.float64)
    view = rs.uniform(-0.7, 0.7, n)
    cen = np.column_stack([rs.uniform(-12, 12, n), rs.uniform(0.8, 1.9, n), rs.uniform(6, 45, n)])
    v1, u1 = rs.uniform(120, 220, n), rs.uniform(0, 900, n)
    boxes_2d = np.column_stack([v1, u1, v1 + rs.uniform(25, 150, n), u1 + rs.uniform(30, 300, n)])
    depth = rs.uniform(5, 50, n)
    out = dict(cam_p=cam_p, xyz_local=xyz_local, view=view, cen=cen, boxes_2d=boxes_2d, depth=depth)
    glob, proj, exp_c, exp_tl, est_y = [], [], [], [], []
    for i in range(n):
        g = instance_utils.inst_points_local_to_global(xyz_local[i].reshape(-1, 3), view[i], cen[i])
        glob.append(g.reshape(roi, roi, 3))
        proj.append(calib_utils.project_pc_to_image(g.T, cam_p).T.reshape(roi, roi, 2))
        exp_c.append(instance_utils.get_exp_proj_uv_map(boxes_2d[i], (roi, roi), use_pixel_centres=True))
        exp_tl.append(instance_utils.get_exp_proj_uv_map(boxes_2d[i], (roi, roi)))
        est_y.append(instance_utils.est_y_from_box_2d_and_depth(cam_p, boxes_2d[i], depth[i], "middle", class_str="Car"))
    sub = (slice(None), slice(1, None, 5), slice(2, None, 5))        # every 5th pixel keeps the fixture small
    out.update(glob=np.asarray(glob)[sub], proj=np.asarray(proj)[sub], exp_centres=np.asarray(exp_c)[sub],
               exp_topleft=np.asarray(exp_tl)[sub], est_y=np.asarray(est_y))
    out["xyz_local"] = xyz_local.astype(np.float32)                    # (the functions ran on exactly these values)
    out["tr_mat"] = np.asarray([transform_utils.np_get_tr_mat(a, t) for a, t in zip(view, cen)])
    path = os.path.join(HERE, "oracle_geometry_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
