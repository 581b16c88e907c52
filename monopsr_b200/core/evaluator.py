"""Inference / evaluation driver -- host-side mirror of the per-checkpoint loop of src/monopsr/core/evaluator.py
(run_checkpoint_once, :136-385) for the B200 engine: restore a checkpoint (EMA shadows, as the reference's
MovingAverageOptimizer swapping saver does), run every sample forward, format and save the predictions
(monopsr_model.py:960-1102 via core/predictions.py) and, in 'val' mode, collect the losses; then
(`convert_and_evaluate`, evaluator.py:336-372) convert the prediction files to KITTI result files and run the AP
evaluation (core/evaluator_utils.py, core/kitti_eval.py -- in process instead of the reference's compiled binary)."""
import os
import time

import numpy as np

from . import evaluator_utils
from . import metrics as M
from . import predictions as P


class Evaluator(object):
    def __init__(self, engine, output_types, output_dirs, train_val_test="val", centroid_type="middle",
                 post_process_cen_x=True, num_alpha_bins=12, log=print):
        if train_val_test not in ("val", "test"):
            raise ValueError("Invalid run mode", train_val_test)
        self.engine, self.output_types, self.output_dirs = engine, list(output_types), dict(output_dirs)
        self.mode, self.centroid_type, self.post_process_cen_x = train_val_test, centroid_type, post_process_cen_x
        self.num_alpha_bins, self.log = num_alpha_bins, log
        for d in self.output_dirs.values():
            os.makedirs(d, exist_ok=True)

    def predict(self, sample, sample_dict):
        """one sample: engine inputs `sample` (Engine.set_inputs keys) + the reference's sample_dict (constants.SAMPLE_*)
        -> formatted prediction dict"""
        eng = self.engine
        eng.set_inputs(sample)
        eng.forward(train=False, compute_losses=self.mode == "val")
        out = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v))
               for k, v in eng.outputs().items() if v is not None}
        if P.KEY_VALID_MASK_MAPS not in out:          # test mode: tf.ones mask (monopsr_model.py:218)
            out[P.KEY_VALID_MASK_MAPS] = np.ones(out[P.KEY_INST_XYZ_MAP_LOCAL].shape[:3] + (1,), np.float32)
        out[P.SAMPLE_LABEL_CLASS_INDICES] = np.asarray(sample["class_indices"])
        self.last_metrics = None
        if self.mode == "val" and "boxes_3d" in sample:
            self.last_metrics = M.evaluate_predictions(out, sample, sample_dict[P.SAMPLE_NUM_OBJS], self.output_types,
                                                       self.centroid_type, point_set=self._point_set_metrics(sample_dict))
        return P.format_predictions(out, sample_dict, output_types=self.output_types, train_val_test=self.mode,
                                    num_boxes=out[P.KEY_VALID_MASK_MAPS].shape[0], num_alpha_bins=self.num_alpha_bins,
                                    centroid_type=self.centroid_type, post_process_cen_x=self.post_process_cen_x)

    def _point_set_metrics(self, sample_dict):
        """metric_emd / metric_chamfer on the device tensors of the engine (None when the engine has none, e.g. a stub)"""
        eng = self.engine
        inputs = getattr(eng, "inputs", None)
        if P.KEY_INST_XYZ_MAP_LOCAL not in self.output_types or not inputs or "gt_inst_xyz_maps_local" not in inputs:
            return None
        from . import losses_custom
        r = losses_custom.point_set_metrics(eng.outputs()[P.KEY_INST_XYZ_MAP_LOCAL], inputs["gt_inst_xyz_maps_local"],
                                            inputs["gt_valid_mask_maps"], int(sample_dict[P.SAMPLE_NUM_OBJS]))
        return {k: v.detach().cpu().numpy() for k, v in r.items()}

    def run_checkpoint_once(self, checkpoint_prefix, samples, use_ema=False):
        """samples: iterable of (sample, sample_dict).  Returns {'num_samples', 'mean_losses' (val), 'metrics' (val:
        name -> list of per-object values, for evaluator_utils.save_metrics), 'seconds'}.
        use_ema=False is the reference's behaviour: its Evaluator restores with a plain tf.train.Saver() built on the
        eval graph (core/evaluator.py:125,144), i.e. the RAW variables -- the MovingAverageOptimizer shadows are
        saved but never consumed.  use_ema=True is an explicit opt-in to evaluate the averaged weights instead."""
        if checkpoint_prefix is not None:
            self.engine.load_checkpoint(checkpoint_prefix, use_ema=use_ema)
        t0, n, sums = time.time(), 0, {}
        metrics_lists = {}
        for sample, sample_dict in samples:
            pred = self.predict(sample, sample_dict)
            P.save_predictions(sample_dict[P.SAMPLE_NAME], pred, self.output_dirs, self.output_types)
            if self.mode == "val":
                for k, v in self.engine.losses().items():
                    sums[k] = sums.get(k, 0.0) + float(v)
                if self.last_metrics:
                    M.accumulate(metrics_lists, self.last_metrics)
            n += 1
            self.log("Step {}: {} / {}, Inference on sample {}".format(
                os.path.basename(str(checkpoint_prefix)), n, "?", sample_dict[P.SAMPLE_NAME]))
        return {"num_samples": n, "mean_losses": {k: v / max(n, 1) for k, v in sums.items()}, "metrics": metrics_lists,
                "seconds": time.time() - t0}

    def convert_and_evaluate(self, dataset, predictions_base_dir, global_step, kitti_score_threshold=0.1,
                             results_root=None, checkpoint_name="model", label_dir=None, already_evaluated_path=None):
        """After the epoch (evaluator.py:336-372): 'val' converts the 2-D detections (if box_2d is an output type) and
        the 3-D detections (if centroids is) and evaluates each conversion; 'test' does the 3-D part when the split
        has labels.  -> list of core.kitti_eval.evaluate results, in that order.  The reference appends the step to
        its list of evaluated checkpoints (already_evaluated_path) in 'val' mode."""
        label_dir = label_dir if label_dir is not None else getattr(dataset, "kitti_label_dir", None)
        results_root = results_root if results_root is not None else os.path.join(predictions_base_dir, "offline_eval")
        out = []

        def native():
            out.append(evaluator_utils.run_kitti_native_eval(
                checkpoint_name, dataset.data_split, kitti_score_threshold, global_step, label_dir, predictions_base_dir,
                results_root, log=self.log))

        quiet = None
        if self.mode == "val":
            if P.KEY_BOX_2D in self.output_types:
                evaluator_utils.save_predictions_box_2d_in_kitti_format(
                    kitti_score_threshold, dataset, predictions_base_dir, self.output_dirs[P.OUT_DIR_BOX_2D], global_step,
                    log=quiet)
                native()
            if P.KEY_CENTROIDS in self.output_types:
                evaluator_utils.save_predictions_box_3d_in_kitti_format(
                    kitti_score_threshold, dataset, predictions_base_dir, self.output_dirs[P.OUT_DIR_BOX_3D],
                    self.output_dirs[P.OUT_DIR_BOX_2D], global_step, log=quiet)
                native()
            if already_evaluated_path is not None:
                with open(already_evaluated_path, "ba") as f:
                    np.savetxt(f, [global_step], fmt="%d")
        elif dataset.has_kitti_labels:
            evaluator_utils.save_predictions_box_3d_in_kitti_format(
                kitti_score_threshold, dataset, predictions_base_dir, self.output_dirs[P.OUT_DIR_BOX_3D],
                self.output_dirs[P.OUT_DIR_BOX_2D], global_step, log=quiet)
            native()
        self.log("\nStep {}: Finished evaluation".format(global_step))
        return out
