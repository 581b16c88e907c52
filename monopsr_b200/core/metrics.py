"""Validation metrics of one sample -- host-side mirror of MonoPSRModel.evaluate_predictions
(src/monopsr/core/models/monopsr/monopsr_model.py:1105-1221; names: core/constants.py:87-97).

  metric_emd / metric_chamfer      the two point-set ops on the valid-masked predicted and ground-truth local maps, per
                                   object, divided by the object's number of valid pixels (:1112-1170) -- computed on
                                   the GPU by core/losses_custom.point_set_metrics (csrc/approxmatch.cu, nn_distance.cu)
  metric_prop_cen_z_err, metric_cen_{x,y,z}_err      ground-truth centroid minus proposal / prediction (:1172-1192)
  metric_dim_err                   gt_dict['lwh_offs'] - pred offsets, where (sic) the graph's ground-truth offsets are
                                   gt_lwh - PREDICTED lwh (monopsr_output_builder.py:655-660)
  metric_view_ang_error            ground-truth minus estimated viewing angle (:1207-1217)
Only the first `num_objs` rows (the real objects; the rest of the 32 boxes are oversampled copies) are evaluated."""
import numpy as np

from . import predictions as P

METRIC_EMD, METRIC_CHAMFER = "metric_emd", "metric_chamfer"
METRIC_VIEW_ANG_ERR = "metric_view_ang_error"
METRIC_PROP_CEN_Z_ERR = "metric_prop_cen_z_err"
METRIC_CEN_X_ERR, METRIC_CEN_Y_ERR, METRIC_CEN_Z_ERR = "metric_cen_x_err", "metric_cen_y_err", "metric_cen_z_err"
METRIC_DIM_ERR = "metric_dim_err"


def gt_centroids(boxes_3d, centroid_type):
    """monopsr_model.py:262-277: box_3d carries the BOTTOM centre; 'middle' moves y up by half the height"""
    b3 = np.asarray(boxes_3d, np.float32)
    cen = b3[:, 0:3].copy()
    if centroid_type == "middle":
        cen[:, 1] = b3[:, 1] - b3[:, 5] / 2
    elif centroid_type != "bottom":
        raise ValueError("Invalid centroid type", centroid_type)
    return cen


def evaluate_predictions(outputs, sample, num_objs, output_types, centroid_type="middle", point_set=None):
    """outputs: the engine's output dict as numpy arrays; sample: the engine-side sample (boxes_3d, gt_view_angs);
    point_set: optional dict with metric_emd / metric_chamfer already computed on the device -> {name: ndarray}"""
    n = int(num_objs)
    m = {}
    if P.KEY_INST_XYZ_MAP_LOCAL in output_types and point_set is not None:
        for k in (METRIC_EMD, METRIC_CHAMFER):
            m[k] = np.asarray(point_set[k])[:n]
    b3 = np.asarray(sample["boxes_3d"], np.float32)
    if P.KEY_CENTROIDS in output_types:
        cens = gt_centroids(b3, centroid_type)[:n]
        m[METRIC_PROP_CEN_Z_ERR] = cens[:, 2:3] - np.asarray(outputs["prop_cen_z"]).reshape(-1, 1)[:n]
        err = cens - np.asarray(outputs[P.KEY_CENTROIDS])[:n]
        m[METRIC_CEN_X_ERR], m[METRIC_CEN_Y_ERR], m[METRIC_CEN_Z_ERR] = err[:, 0], err[:, 1], err[:, 2]
    if P.KEY_LWH in output_types:
        gt_offs = b3[:, 3:6] - np.asarray(outputs[P.KEY_LWH])
        m[METRIC_DIM_ERR] = (gt_offs - np.asarray(outputs[P.KEY_LWH + "_offs"]))[:n]
    if P.KEY_VIEW_ANG in output_types:
        gt_va = np.asarray(sample["gt_view_angs"], np.float32).reshape(-1, 1)
        m[METRIC_VIEW_ANG_ERR] = (gt_va - np.asarray(outputs[P.KEY_VIEW_ANG]).reshape(-1, 1))[:n]
    return m


def accumulate(metrics_lists, metrics):
    """the evaluator's bookkeeping (evaluator.py:268-281): values with a NaN are dropped, the rest flattened"""
    for k, v in metrics.items():
        v = np.asarray(v)
        if np.isnan(v).any():
            continue
        metrics_lists.setdefault(k, []).extend(np.reshape(v, (-1)).tolist())
    return metrics_lists
