"""Experiment drivers behind the reference's entry scripts (src/monopsr/experiments/run_training.py:74-89,
run_evaluation.py:47-92, run_inference.py:78-125) and the checkpoint bookkeeping of its Evaluator
(src/monopsr/core/evaluator.py:28-135 set-up, :136-205 output folders per step, :385-527 which checkpoints to run).

The TF graph / session / saver are replaced by one Engine; the dataset, the config object, the directory layout under
<exp_output_dir>/predictions, the 'evaluated_<split>.txt' list and the console lines are the reference's.
`engine_factory(device)` lets tests (and other callers) inject the engine; the default builds core.engine.Engine."""
import os
import time

import numpy as np

from . import config_utils
from . import evaluator as evaluator_mod
from . import evaluator_utils
from . import predictions as P
from . import trainer as trainer_mod
from ..datasets.kitti_loader import DatasetBuilder, KittiDataset, PrefetchLoader


LOADER_WORKERS = max(1, min(4, (os.cpu_count() or 2) // 2))      # PNG decode threads beside the ordered producer


def output_types_list(output_config):
    """MonoPSROutputBuilder.get_output_types_list (monopsr_output_builder.py:76-81)"""
    return sorted(k for k in output_config.__dict__.keys() if not k.startswith("__"))


def _default_engine(device, config=None):
    """the engine of an experiment: loss type of the xyz map and the train-op hyper-parameters from the yaml"""
    from .engine import Engine
    return Engine(device) if config is None else Engine.from_config(device, config)


def _make_engine(engine_factory, device, config):
    """engine_factory(device) or engine_factory(device, config) -- older injected factories take the device only"""
    import inspect
    try:
        takes_config = len(inspect.signature(engine_factory).parameters) >= 2
    except (TypeError, ValueError):
        takes_config = False
    return engine_factory(device, config) if takes_config else engine_factory(device)


def checkpoints_in(checkpoint_dir, model_type):
    """[(step, prefix)] of '<model_type>-<8 digits>.index' files, oldest first"""
    import glob
    import re
    found = []
    for path in glob.glob(os.path.join(checkpoint_dir, model_type + "-*.index")):
        m = re.search(r"-(\d+)\.index$", path)
        if m:
            found.append((int(m.group(1)), path[:-len(".index")]))
    return sorted(found)


class ExperimentEvaluator(object):
    """Evaluator(model, config, eval_mode, skip_evaluated_checkpoints, eval_wait_interval, do_kitti_native_eval)"""

    def __init__(self, engine, dataset, config, eval_mode="val", skip_evaluated_checkpoints=True, eval_wait_interval=30,
                 do_kitti_native_eval=True, log=print):
        if eval_mode not in ("val", "test"):
            raise ValueError("Evaluation mode can only be set to `val` or `test`")
        self.engine, self.dataset, self.config, self.eval_mode, self.log = engine, dataset, config, eval_mode, log
        self.dataset_config, self.model_config, self.train_config = config.dataset_config, config.model_config, config.train_config
        self.model_name = self.model_config.model_type
        self.checkpoint_dir = self.train_config.paths_config.checkpoint_dir
        if not os.path.exists(self.checkpoint_dir):
            raise ValueError("{} must have at least one checkpoint entry.".format(self.checkpoint_dir))
        self.skip_evaluated_checkpoints, self.eval_wait_interval = skip_evaluated_checkpoints, eval_wait_interval
        if do_kitti_native_eval and eval_mode == "test":
            raise ValueError("Cannot run native eval in test mode.")
        self.do_kitti_native_eval = do_kitti_native_eval
        self.predictions_base_dir = self.train_config.paths_config.pred_dir
        os.makedirs(self.predictions_base_dir, exist_ok=True)
        self.already_evaluated_path = self.predictions_base_dir + "/evaluated_{}.txt".format(self.dataset_config.data_split)
        self.output_types = output_types_list(self.model_config.output_config)

    # ------------------------------------------------------------------ one checkpoint
    def output_dirs(self, global_step):
        base, split = self.predictions_base_dir, self.dataset_config.data_split
        dirs = {}
        if P.KEY_CENTROIDS in self.output_types:
            dirs[P.OUT_DIR_BOX_3D] = base + "/predictions_{}/{}/{}".format(P.KEY_BOX_3D, split, global_step)
            dirs[P.OUT_DIR_BOX_2D] = base + "/predictions_{}/{}/{}".format(P.KEY_BOX_2D, split, global_step)
        if P.KEY_INST_XYZ_MAP_LOCAL in self.output_types:
            dirs[P.OUT_DIR_XYZ_MAP_LOCAL] = base + "/predictions_{}/{}/{}".format(P.KEY_INST_XYZ_MAP_LOCAL, split, global_step)
        for d in dirs.values():
            os.makedirs(d, exist_ok=True)
        return dirs

    def run_checkpoint_once(self, checkpoint_to_restore):
        global_step = int(checkpoint_to_restore[-8:])
        d = self.dataset_config
        ev = evaluator_mod.Evaluator(self.engine, self.output_types, self.output_dirs(global_step),
                                     train_val_test=self.eval_mode, centroid_type=d.centroid_type,
                                     post_process_cen_x=getattr(self.model_config, "post_process_cen_x", True),
                                     num_alpha_bins=d.num_alpha_bins, log=self.log)
        with PrefetchLoader(self.dataset, shuffle=False, epochs=1, workers=LOADER_WORKERS) as samples:
            res = ev.run_checkpoint_once(checkpoint_to_restore, samples)
        res["global_step"] = global_step
        if self.eval_mode == "val" and res.get("metrics"):
            evaluator_utils.save_metrics(
                os.path.join(self.predictions_base_dir, "offline_eval", "metrics", self.config.config_name,
                             self.dataset_config.data_split),
                self.dataset_config.data_split, global_step, res["metrics"],
                metrics_to_show=getattr(self.model_config, "metrics_to_show", ()) or ())
        if self.eval_mode == "val" and not self.do_kitti_native_eval:
            with open(self.already_evaluated_path, "ba") as f:
                np.savetxt(f, [global_step], fmt="%d")
        elif self.eval_mode == "val" or self.dataset.has_kitti_labels:
            res["kitti"] = ev.convert_and_evaluate(
                self.dataset, self.predictions_base_dir, global_step,
                kitti_score_threshold=self.train_config.kitti_score_threshold, checkpoint_name=self.config.config_name,
                already_evaluated_path=self.already_evaluated_path)
        return res

    # ------------------------------------------------------------------ which checkpoints
    def get_evaluated_ckpts(self):
        if os.path.exists(self.already_evaluated_path):
            return np.loadtxt(self.already_evaluated_path, delimiter=",").reshape(-1).astype(np.int32)
        return []

    def run_latest_checkpoints(self, ckpt_indices):
        """ckpt_indices: steps to evaluate (ints / digit strings), [-1] = the newest checkpoint"""
        ckpts = checkpoints_in(self.checkpoint_dir, self.model_name)
        by_step = {"%08d" % s: p for s, p in ckpts}
        out = []
        for idx in np.asarray(ckpt_indices).reshape(-1):
            if int(idx) == -1:
                if not ckpts:
                    raise ValueError("{} must have at least one checkpoint entry.".format(self.checkpoint_dir))
                out.append(self.run_checkpoint_once(ckpts[-1][1]))
            else:
                out.append(self.run_checkpoint_once(by_step[str(int(idx)).rjust(8, "0")]))
        return out

    def repeated_checkpoint_run(self, max_polls=None):
        """evaluate every checkpoint not evaluated yet, oldest first, and keep polling the directory every
        `eval_wait_interval` seconds until the one of step `max_iterations` has been done (or `max_polls` polls
        found nothing new -- the reference polls forever)"""
        done = set(int(s) for s in self.get_evaluated_ckpts()) if self.skip_evaluated_checkpoints else set()
        self.log("Starting evaluation at " + time.strftime("%Y-%m-%d-%H:%M:%S", time.gmtime()))
        last_step, polls, out = -1, 0, []
        while last_step < self.train_config.max_iterations:
            todo = [(s, p) for s, p in checkpoints_in(self.checkpoint_dir, self.model_name) if s not in done and s > last_step]
            if not todo:
                polls += 1
                if max_polls is not None and polls >= max_polls:
                    return out
                self.log("No new checkpoints found in {}. Will try again in {} seconds".format(
                    self.checkpoint_dir, self.eval_wait_interval))
                time.sleep(self.eval_wait_interval)
                continue
            for step, prefix in todo:
                out.append(self.run_checkpoint_once(prefix))
                last_step = step
        self.log("All checkpoints evaluated, exiting.")
        return out


# ---------------------------------------------------------------------------------------------- entry points
def data_parallel_setup(device="cuda:0"):
    """(rank, world, device).  Under torchrun (WORLD_SIZE > 1): one process per GPU, NCCL process group, this rank's
    device = cuda:LOCAL_RANK -- Engine.train_step then all-reduces the gradients (samples are
    independent, one per GPU and step, no other exchange).  Otherwise (0, 1, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, device
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    return dist.get_rank(), world, "cuda:%d" % local


def train(config, device="cuda:0", engine_factory=_default_engine, data_dir=None, pretrained_checkpoint=None, log=print,
          seed=None):
    """run_training.train: dataset in 'train' mode, prefetching loader, trainer.train.  Data parallel under torchrun:
    every rank walks its OWN shuffled order of the split (generator seeded with seed + rank; with seed=None rank 0 uses
    numpy's global generator as the reference does), rank 0 writes the checkpoints."""
    config_utils.validate_for_engine(config)
    rank, world, device = data_parallel_setup(device)
    rng = np.random if (seed is None and rank == 0) else np.random.RandomState((seed or 0) + rank)
    dataset = KittiDataset(config.dataset_config, "train", data_dir=data_dir, rng=rng)
    engine = _make_engine(engine_factory, device, config)
    with PrefetchLoader(dataset, shuffle=True, workers=LOADER_WORKERS) as loader:
        return trainer_mod.train(engine, config, loader.sample_fn, pretrained_checkpoint=pretrained_checkpoint, log=log,
                                 chief=rank == 0)


def evaluate(config, device="cuda:0", engine_factory=_default_engine, data_dir=None, max_polls=None, log=print):
    """run_evaluation.evaluate: 'val' mode on config.dataset_config.data_split, every checkpoint, native AP evaluation"""
    d = config.dataset_config
    if d.data_split == "test":
        d.data_split_dir, d.has_kitti_labels = "testing", False
    else:
        d.data_split_dir, d.has_kitti_labels = "training", True
    d.aug_list = []
    dataset = KittiDataset(d, "val", data_dir=data_dir)
    config_utils.validate_for_engine(config)
    ev = ExperimentEvaluator(_make_engine(engine_factory, device, config), dataset, config, eval_mode="val", skip_evaluated_checkpoints=True,
                             do_kitti_native_eval=True, log=log)
    return ev.repeated_checkpoint_run(max_polls=max_polls)


def inference(config, data_split, ckpt_indices, device="cuda:0", engine_factory=_default_engine, data_dir=None,
              max_polls=None, log=print):
    """run_inference.inference: 'test' mode (no labels used, no losses), selected checkpoints or 'all'"""
    d = config.dataset_config
    d.data_split = data_split
    if data_split == "test":
        d.data_split_dir, d.has_kitti_labels = "testing", False
    d.aug_config.box_jitter_type = None
    dataset = DatasetBuilder.build_kitti_dataset(d, train_val_test="test", data_dir=data_dir)
    every = isinstance(ckpt_indices, str) and ckpt_indices == "all"
    config_utils.validate_for_engine(config)
    ev = ExperimentEvaluator(_make_engine(engine_factory, device, config), dataset, config, eval_mode="test", skip_evaluated_checkpoints=every,
                             do_kitti_native_eval=False, log=log)
    return ev.repeated_checkpoint_run(max_polls=max_polls) if every else ev.run_latest_checkpoints(ckpt_indices)
