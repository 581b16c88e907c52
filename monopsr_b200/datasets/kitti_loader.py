"""KITTI sample loader for the B200 engine.

  KittiDataset        host-side mirror of src/monopsr/datasets/kitti/kitti_dataset.py:26-556 -- same config fields,
                      directory layout, filtering, oversampling, box jitter, sample_dict keys (core/constants.py:1-31)
                      and epoch / wrap-around bookkeeping of next_batch; random draws in the reference's order.
  DatasetBuilder      the preconfigured dataset configs of src/monopsr/builders/dataset_builder.py:10-94.
  engine_sample       sample_dict -> the arrays Engine.set_inputs takes (what MonoPSRModel.create_feed_dict +
                      the input part of MonoPSRModel.build do, monopsr_model.py:153-233,494-552).
  PrefetchLoader      B200-side addition: the reference builds each feed_dict synchronously between two sess.run calls
                      (monopsr_model.py:494-503); at 8.5 ms per step the PNG decode + label work (tens of ms) would
                      dominate, so samples are produced by a background thread into a bounded queue while the GPU runs.
                      One producer thread => the dataset's random stream is consumed in the same order as a
                      synchronous loop, so seeded runs stay reproducible; with workers > 0 the PNG decoding (the bulk
                      of the time, no randomness) of the next samples runs on extra threads (ReadAhead).
"""
import fnmatch
import os
import queue
import threading

import numpy as np
import yaml

from . import augment
from . import kitti_formats as K
from ..core.config_utils import config_dict_to_object

# sample_dict keys (src/monopsr/core/constants.py:1-31)
SAMPLE_IMAGE_INPUT = "sample_image_input"
SAMPLE_NUM_OBJS = "sample_num_objs"
SAMPLE_LABEL_BOXES_2D = "sample_label_boxes_2d"
SAMPLE_LABEL_BOXES_2D_NORM = "sample_label_boxes_2d_norm"
SAMPLE_LABEL_BOXES_3D = "sample_label_boxes_3d"
SAMPLE_INSTANCE_MASKS = "sample_instance_masks"
SAMPLE_ALPHAS = "sample_alphas"
SAMPLE_ALPHA_BINS = "sample_alpha_bins"
SAMPLE_ALPHA_REGS = "sample_alpha_regressions"
SAMPLE_ALPHA_VALID_BINS = "sample_alpha_valid_bins"
SAMPLE_PROP_CEN_Z_OFFSET = "sample_prop_cen_z_offset"
SAMPLE_VIEWING_ANGLES_2D = "sample_viewing_angles_2d"
SAMPLE_VIEWING_ANGLES_3D = "sample_viewing_angles_3d"
SAMPLE_LABEL_CLASS_STRS = "sample_label_class_strs"
SAMPLE_LABEL_CLASS_INDICES = "sample_label_class_indices"
SAMPLE_LABEL_SCORES = "sample_label_scores"
SAMPLE_DEPTH_MAP = "sample_depth_map"
SAMPLE_CAM_P = "sample_cam_p"
SAMPLE_NAME = "sample_name"
SAMPLE_MEAN_LWH = "sample_mean_lwh"

JITTER_MIN_IOU = 0.7


class Sample(object):
    def __init__(self, name, augs):
        self.name, self.augs = name, augs

    def __repr__(self):
        return "({}, augs: {})".format(self.name, self.augs)


class KittiDataset(object):
    """dataset_config: attribute object (the yaml's `dataset_config`, or DatasetBuilder.get_config_obj);
    data_dir: root of the MS-CNN detections tree (the reference's monopsr.data_dir());
    rng: source of the oversampling / augmentation / shuffle draws (default numpy's global generator)."""

    def __init__(self, dataset_config, train_val_test, data_dir=None, rng=np.random):
        c = self.dataset_config = dataset_config
        self.train_val_test = train_val_test
        self.rng = rng
        self.name = c.name
        self.data_split = c.data_split
        self.dataset_dir = os.path.expanduser(c.dataset_dir)
        self.num_boxes = c.num_boxes
        self.num_alpha_bins = c.num_alpha_bins
        self.alpha_bin_overlap = c.alpha_bin_overlap
        self.centroid_type = c.centroid_type
        self.cam_idx = 2
        self.classes = list(c.classes)
        self.num_classes = len(self.classes)
        if train_val_test in ("train", "val"):
            c.obj_filter_config.classes = self.classes
            self.obj_filter = K.ObjectFilter(c.obj_filter_config)
        else:       # inference keeps every detection of the classes
            self.obj_filter = K.ObjectFilter.create_obj_filter(
                classes=self.classes, difficulty=K.Difficulty.ALL, occlusion=None, truncation=None, box_2d_height=None,
                depth_range=None)
        self.has_kitti_labels = c.has_kitti_labels
        self.use_mscnn_detections = c.use_mscnn_detections
        self.mscnn_thr = c.mscnn_thr
        self.trend_data = "kitti"
        if self.num_classes > 1:
            raise NotImplementedError("Number of classes must be 1")
        self.classes_name = self.classes[0]
        if self.classes_name == "Car":
            self.mscnn_merge_min_iou = 0.7
        elif self.classes_name in ("Pedestrian", "Cyclist"):
            self.mscnn_merge_min_iou = 0.5

        if not os.path.exists(self.dataset_dir):
            raise FileNotFoundError("Dataset path does not exist: {}".format(self.dataset_dir))
        entries = os.listdir(self.dataset_dir)
        splits = [os.path.splitext(f)[0] for f in entries if fnmatch.fnmatch(f, "*.txt")]
        splits = [s for s in splits if s != "readme"]
        if self.data_split not in splits:
            raise ValueError("Invalid data split: {}, possible_splits: {}".format(self.data_split, splits))
        split_dirs = [d for d in entries if os.path.isdir(self.dataset_dir + "/" + d)]
        if c.data_split_dir not in split_dirs:
            raise ValueError("Invalid data split dir: {}, possible dirs: {}".format(c.data_split_dir, split_dirs))
        self.data_split_dir = self.dataset_dir + "/" + c.data_split_dir

        self.depth_version = c.depth_version
        self.instance_version = c.instance_version
        d = self.data_split_dir
        self.rgb_image_dir = d + "/image_" + str(self.cam_idx)
        self.image_2_dir, self.image_3_dir = d + "/image_2", d + "/image_3"
        self.calib_dir, self.disp_dir, self.planes_dir, self.velo_dir = d + "/calib", d + "/disparity", d + "/planes", \
            d + "/velodyne"
        self.depth_dir = d + "/depth_{}_{}".format(self.cam_idx, self.depth_version)
        self.instance_dir = d + "/instance_{}_{}".format(self.cam_idx, self.instance_version)
        self.data_dir = data_dir if data_dir is not None else os.path.join(os.getcwd(), "data")
        self.mscnn_label_dir = self.data_dir + "/detections/mscnn/kitti_fmt/{}/merged_{}/data".format(
            self.data_split, "_".join(map(str, self.mscnn_thr)))
        if self.has_kitti_labels:
            self.kitti_label_dir = d + "/label_2"

        self.oversample = c.oversample
        self.aug_config = c.aug_config
        names = self.load_sample_names(self.data_split)
        self.sample_list = np.asarray([Sample(n, []) for n in names])
        self.num_samples = len(self.sample_list)
        self.clusters, self.std_devs = [3.892, 1.619, 1.530], [0.440, 0.106, 0.138]
        self._index_in_epoch = 0
        self.epochs_completed = 0

    # ------------------------------------------------------------------ paths / lists
    def get_sample_names(self):
        return [s.name for s in self.sample_list]

    def get_rgb_image_path(self, sample_name):
        return self.rgb_image_dir + "/" + sample_name + ".png"

    def get_image_2_path(self, sample_name):
        return self.image_2_dir + "/" + sample_name + ".png"

    def get_image_3_path(self, sample_name):
        return self.image_3_dir + "/" + sample_name + ".png"

    def get_depth_map_path(self, sample_name):
        # accessor kept as in the reference; the loader itself reads <depth_dir>/<name>.png (obj_utils.py:532-539)
        return self.depth_dir + "/" + sample_name + "_left_depth.png"

    def get_velodyne_path(self, sample_name):
        return self.velo_dir + "/" + sample_name + ".bin"

    def get_cluster_info(self):
        return self.clusters, self.std_devs

    def load_sample_names(self, data_split):
        with open(self.dataset_dir + "/" + data_split + ".txt", "r") as f:
            return np.asarray(f.read().splitlines())

    # ------------------------------------------------------------------ one sample
    _READERS = {"rgb": K.read_rgb_image, "depth": K.read_depth_map, "instance": K.read_instance_image}

    def sample_files(self, sample_name):
        """(kind, path) of the image-like files one sample may read"""
        files = [("rgb", self.get_rgb_image_path(sample_name))]
        if self.train_val_test in ("train", "val"):
            files += [("instance", self.instance_dir + "/{}.png".format(sample_name)),
                      ("depth", self.depth_dir + "/{}.png".format(sample_name))]
        return files

    def _read(self, kind, path):
        ahead = getattr(self, "_read_ahead", None)       # set by PrefetchLoader(workers > 0)
        if ahead is not None:
            hit = ahead.take(kind, path)
            if hit is not None:
                return hit
        return self._READERS[kind](path)

    def _oversample_indices(self, num_objs):
        extra = self.rng.choice(num_objs, self.num_boxes - num_objs, replace=True)
        return np.hstack([np.arange(0, num_objs), extra])

    def _load_one(self, sample):
        name = sample.name
        rgb_image = self._read("rgb", self.get_rgb_image_path(name))
        image_shape = rgb_image.shape[0:2]
        image_input = rgb_image
        cam_p = K.read_frame_calib(self.calib_dir + "/{}.txt".format(name)).p2
        mode = self.train_val_test
        extra = {}
        if mode in ("train", "val"):
            kitti_labels = K._as_label_array(K.read_labels(self.kitti_label_dir, name))
            if self.use_mscnn_detections and mode == "val":
                detections = K.read_labels(self.mscnn_label_dir, name)
                labels = K._as_label_array(K.merge_kitti_and_mscnn_obj_labels(
                    kitti_labels, detections, min_iou=self.mscnn_merge_min_iou, default_score_type="distance"))
            else:
                labels = kitti_labels
            num_all = len(labels)
            labels, keep = K.apply_obj_filter(labels, self.obj_filter)
            num_objs = len(labels)
            if num_objs < 1:
                return None
            jitter = self.aug_config.box_jitter_type if mode == "train" else None
            if self.use_mscnn_detections or jitter == "oversample_gt":
                # (the reference filters the ground-truth labels only on the MS-CNN path and then fails with a
                # NameError for 'oversample_gt' without it; here that combination simply works)
                kitti_labels, _ = K.apply_obj_filter(kitti_labels, self.obj_filter)
                if len(kitti_labels) < 1:
                    return None
            instance_image = self._read("instance", self.instance_dir + "/{}.png".format(name))
            # one mask per KEPT label (pixel value = the label's index in the file); the reference expands all labels
            # first and filters afterwards -- same masks
            masks = np.asarray([instance_image == i for i in np.flatnonzero(keep)])
            if self.oversample:
                idx = self._oversample_indices(num_objs)
                labels, masks = labels[idx], masks[idx]
            if mode == "train":
                if self.aug_config.use_image_aug:
                    image_input = augment.apply_image_noise(rgb_image, self.rng)
                if jitter is not None:
                    if jitter in ("oversample", "oversample_gt") and not self.oversample:
                        raise ValueError("Must oversample object labels to use {} box jitter type".format(jitter))
                    if jitter == "oversample":          # the oversampled copies get jittered boxes
                        labels[num_objs:] = augment.jitter_obj_boxes_2d(labels[num_objs:], JITTER_MIN_IOU, image_shape,
                                                                        self.rng)
                    elif jitter == "oversample_gt":     # ... jittered copies of randomly drawn ground-truth boxes
                        pick = self.rng.choice(len(kitti_labels), self.num_boxes - num_objs, replace=True)
                        labels[num_objs:] = augment.jitter_obj_boxes_2d(kitti_labels[pick], JITTER_MIN_IOU,
                                                                        image_shape, self.rng)
                    elif jitter == "all":
                        labels = augment.jitter_obj_boxes_2d(labels, JITTER_MIN_IOU, image_shape, self.rng)
                    else:
                        raise ValueError("Invalid box_jitter_type", jitter)
            boxes_2d = K.boxes_2d_from_obj_labels(labels)
            boxes_3d = K.boxes_3d_from_obj_labels(labels)
            bins = [K.np_orientation_to_angle_bin(o.alpha, self.num_alpha_bins, self.alpha_bin_overlap) for o in labels]
            extra = {
                SAMPLE_LABEL_BOXES_3D: boxes_3d,
                SAMPLE_ALPHAS: np.asarray([o.alpha for o in labels], dtype=np.float32),
                SAMPLE_ALPHA_BINS: np.asarray([b[0] for b in bins]),
                SAMPLE_ALPHA_REGS: np.asarray([b[1] for b in bins]),
                SAMPLE_ALPHA_VALID_BINS: np.asarray([b[2] for b in bins]),
                SAMPLE_VIEWING_ANGLES_3D: np.asarray([K.get_viewing_angle_box_3d(b, cam_p) for b in boxes_3d],
                                                     dtype=np.float32),
                SAMPLE_INSTANCE_MASKS: masks,
                SAMPLE_DEPTH_MAP: self._read("depth", self.depth_dir + "/{}.png".format(name)),
            }
        elif mode == "test":
            labels = K.read_labels(self.mscnn_label_dir, name)
            if len(labels) < 1:
                return None
            labels, _ = K.apply_obj_filter(labels, self.obj_filter)
            num_objs = len(labels)
            if num_objs < 1:
                return None
            labels = labels[self._oversample_indices(num_objs)]
            boxes_2d = K.boxes_2d_from_obj_labels(labels)
        else:
            raise ValueError("Invalid run mode", mode)

        class_strs = [o.type for o in labels]
        sample_dict = {
            SAMPLE_NUM_OBJS: num_objs,
            SAMPLE_IMAGE_INPUT: image_input,
            SAMPLE_CAM_P: cam_p,
            SAMPLE_NAME: name,
            SAMPLE_LABEL_BOXES_2D_NORM: boxes_2d / np.tile(image_shape, 2),
            SAMPLE_LABEL_BOXES_2D: boxes_2d,
            SAMPLE_LABEL_SCORES: np.asarray([o.score for o in labels], np.float32),
            SAMPLE_LABEL_CLASS_STRS: np.expand_dims(class_strs, 1),
            SAMPLE_LABEL_CLASS_INDICES: np.expand_dims(
                np.asarray([K.class_str_to_index(s, self.classes) for s in class_strs], dtype=np.int32), axis=1),
            SAMPLE_MEAN_LWH: np.asarray([K.get_mean_lwh_and_std_dev(s)[0] for s in class_strs]),
            SAMPLE_PROP_CEN_Z_OFFSET: np.asarray([K.get_prop_cen_z_offset(s) for s in class_strs]),
            SAMPLE_VIEWING_ANGLES_2D: np.asarray([K.get_viewing_angle_box_2d(b, cam_p) for b in boxes_2d],
                                                 dtype=np.float32),
        }
        sample_dict.update(extra)
        return sample_dict

    def get_sample_dict(self, indices):
        """sample dicts for dataset.sample_list[indices]; None for a sample without a usable object"""
        return [self._load_one(self.sample_list[i]) for i in indices]

    # ------------------------------------------------------------------ batches
    def _shuffle_samples(self):
        perm = np.arange(self.num_samples)
        self.rng.shuffle(perm)
        self.sample_list = self.sample_list[perm]

    def next_indices(self, batch_size, shuffle):
        """The sample_list index ranges of the next batch, advancing the epoch bookkeeping (the first epoch is
        shuffled on its first batch; an epoch ends when start + batch_size >= num_samples, the list is reshuffled
        BETWEEN the tail of the old epoch and the head of the new one).  Yields ranges lazily because the reshuffle
        must happen after the tail has been read."""
        start = self._index_in_epoch
        if self.epochs_completed == 0 and start == 0 and shuffle:
            self._shuffle_samples()
        if start + batch_size >= self.num_samples:
            self.epochs_completed += 1
            rest = self.num_samples - start
            yield np.arange(start, self.num_samples)
            if shuffle:
                self._shuffle_samples()
            self._index_in_epoch = batch_size - rest
            yield np.arange(0, self._index_in_epoch)
        else:
            self._index_in_epoch += batch_size
            yield np.arange(start, self._index_in_epoch)

    def next_batch(self, batch_size, shuffle=False):
        batch = []
        for indices in self.next_indices(batch_size, shuffle):
            batch.extend(self.get_sample_dict(indices))
        return batch


class DatasetBuilder(object):
    """Preconfigured dataset configs (builders/dataset_builder.py:10-94)."""

    CONFIG_DEFAULTS = dict(
        dataset_type="kitti_obj", use_mscnn_detections=True, mscnn_thr=[0.2, 0.2, 0.2], batch_size=1, oversample=True,
        num_boxes=32, num_alpha_bins=8, alpha_bin_overlap=0.1, centroid_type="middle", classes=["Car"],
        obj_filter_config=dict(difficulty_str="hard", occlusion=None, truncation=0.3, box_2d_height=None,
                               depth_range=[5, 45]),
        aug_config=dict(use_image_aug=False, box_jitter_type="oversample"),
        name="kitti", dataset_dir="~/Kitti/object", data_split="train", data_split_dir="training",
        has_kitti_labels=True, depth_version="multiscale", instance_version="depth_2_multiscale")
    KITTI_TRAIN, KITTI_VAL, KITTI_TRAINVAL, KITTI_TEST = "kitti_obj_train", "kitti_obj_val", "kitti_obj_trainval", \
        "kitti_obj_test"

    @staticmethod
    def get_config_obj(dataset_type):
        cfg = config_dict_to_object(yaml.safe_load(yaml.safe_dump(DatasetBuilder.CONFIG_DEFAULTS)))
        if dataset_type == DatasetBuilder.KITTI_TRAIN:
            pass
        elif dataset_type == DatasetBuilder.KITTI_VAL:
            cfg.data_split = "val"
        elif dataset_type == DatasetBuilder.KITTI_TRAINVAL:
            cfg.data_split = "trainval"
        elif dataset_type == DatasetBuilder.KITTI_TEST:
            cfg.data_split, cfg.data_split_dir, cfg.has_kitti_labels = "test", "testing", False
        else:
            raise ValueError("Invalid dataset type", dataset_type)
        return cfg

    @staticmethod
    def build_kitti_dataset(dataset_config, train_val_test="train", **kw):
        if isinstance(dataset_config, str):
            dataset_config = DatasetBuilder.get_config_obj(dataset_config)
        return KittiDataset(dataset_config, train_val_test, **kw)


def _as_uint8(masks):
    """boolean (N,H,W) masks as uint8 without copying 15 MB when they are already contiguous"""
    m = np.asarray(masks)
    if m.dtype == np.bool_ and m.flags["C_CONTIGUOUS"]:
        return m.view(np.uint8)
    return np.ascontiguousarray(m, dtype=np.uint8)


def engine_sample(sample_dict, train_val_test="train"):
    """sample_dict -> the dict Engine.set_inputs takes: the placeholders of monopsr_model.py:494-552 under the engine's
    key names, fp32 / int32, with the RAW image, depth map and instance masks -- the crops, the resized full image and
    the ground-truth maps are produced on the GPU (core/targets.py: mpb_image_inputs, mpb_gt_xyz_from_depth)."""
    s = sample_dict
    f32 = lambda k: np.ascontiguousarray(s[k], dtype=np.float32)
    out = {
        "rgb_image": np.ascontiguousarray(s[SAMPLE_IMAGE_INPUT]),
        "boxes_2d": f32(SAMPLE_LABEL_BOXES_2D),
        "boxes_2d_norm": f32(SAMPLE_LABEL_BOXES_2D_NORM),
        "cam_p": f32(SAMPLE_CAM_P),
        "class_indices": np.ascontiguousarray(s[SAMPLE_LABEL_CLASS_INDICES], dtype=np.int32),
        "mean_lwh": f32(SAMPLE_MEAN_LWH),
        "prop_cen_z_offset": f32(SAMPLE_PROP_CEN_Z_OFFSET),
        "est_view_angs": f32(SAMPLE_VIEWING_ANGLES_2D),
    }
    if train_val_test in ("train", "val"):
        out.update({
            "boxes_3d": f32(SAMPLE_LABEL_BOXES_3D),
            "gt_alphas": f32(SAMPLE_ALPHAS),
            "gt_alpha_bins": np.ascontiguousarray(s[SAMPLE_ALPHA_BINS], dtype=np.int32),
            "gt_alpha_regs": f32(SAMPLE_ALPHA_REGS),
            "gt_alpha_valid_bins": f32(SAMPLE_ALPHA_VALID_BINS),
            "gt_view_angs": f32(SAMPLE_VIEWING_ANGLES_3D),
            "depth_map": f32(SAMPLE_DEPTH_MAP),
            "instance_masks": _as_uint8(s[SAMPLE_INSTANCE_MASKS]),
        })
    elif train_val_test != "test":
        raise ValueError("Invalid run mode", train_val_test)
    return out


def _cuda_available():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def _pinned(sample):
    """{name: ndarray} -> {name: page-locked torch tensor} (non-array values are passed through)"""
    import torch
    out = {}
    for k, v in sample.items():
        out[k] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if isinstance(v, np.ndarray) else v
    return out


class ReadAhead(object):
    """Decodes the PNGs of the next `lookahead` samples of the split on `workers` threads (cv2 releases the GIL) while
    the single producer thread does the ordered, random-stream-consuming part of the work.  Purely a cache keyed by
    (kind, path): a miss falls back to a direct read, a failed read re-raises where the direct read would have."""

    def __init__(self, dataset, workers=4, lookahead=8):
        from concurrent.futures import ThreadPoolExecutor
        self.dataset, self.lookahead = dataset, max(1, lookahead)
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="kitti-decode")
        self.pending = {}            # (kind, path) -> Future, insertion-ordered
        self.hits = self.misses = 0

    def schedule(self):
        ds = self.dataset
        start = ds._index_in_epoch
        for sample in ds.sample_list[start:start + self.lookahead]:
            for key in ds.sample_files(sample.name):
                if key not in self.pending:
                    self.pending[key] = self.pool.submit(ds._READERS[key[0]], key[1])
        while len(self.pending) > 6 * self.lookahead:           # reshuffled away or never used: forget the oldest
            self.pending.pop(next(iter(self.pending))).cancel()

    def take(self, kind, path):
        f = self.pending.pop((kind, path), None)
        if f is None:
            self.misses += 1
            return None
        self.hits += 1
        return f.result()

    def close(self):
        for f in self.pending.values():
            f.cancel()
        self.pending.clear()
        self.pool.shutdown(wait=True)


class PrefetchLoader(object):
    """Iterator over (engine_sample, sample_dict) pairs produced `depth` samples ahead of the consumer by one
    background thread (batch_size 1, empty samples skipped, as create_feed_dict).  The stream ends after `max_samples`
    samples and / or `epochs` passes over the split (epochs=1 is the evaluator's loop, core/evaluator.py:203-205:
    "while current_epoch == dataset.epochs_completed", including its quirk that skipping an empty LAST sample runs on
    into the next epoch and serves that epoch's first sample too); with neither it is endless (the training loop stops
    on its step count).  An exception in the producer is re-raised in the consumer; close() (or leaving the `with` block)
    stops the thread.  `sample_fn` is the callable core/trainer.train takes."""

    _END = object()

    def __init__(self, dataset, shuffle=None, depth=4, max_samples=None, epochs=None, convert=engine_sample, workers=0,
                 pin_memory=None):
        """workers > 0: decode the image files of upcoming samples on that many extra threads (ReadAhead).
        pin_memory (default: when CUDA is available): hand the engine sample over as page-locked torch tensors, so that
        Engine.set_inputs' non_blocking copies (17 MB per training sample, mostly the instance masks) really are
        asynchronous; torch's caching host allocator recycles the blocks once the copies that read them are done."""
        self.dataset = dataset
        self.pin_memory = _cuda_available() if pin_memory is None else bool(pin_memory)
        self.read_ahead = ReadAhead(dataset, workers, lookahead=2 * workers + depth) if workers > 0 else None
        dataset._read_ahead = self.read_ahead
        self.shuffle = (dataset.train_val_test == "train") if shuffle is None else shuffle
        self.max_samples, self.epochs = max_samples, epochs
        self.convert = convert
        self._q = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._produce, name="kitti-prefetch", daemon=True)
        self._thread.start()

    def _put(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _produce(self):
        try:
            ds, made = self.dataset, 0
            last_epoch = None if self.epochs is None else ds.epochs_completed + self.epochs
            while not self._stop.is_set() and (self.max_samples is None or made < self.max_samples) and \
                    (last_epoch is None or ds.epochs_completed < last_epoch):
                sample_dict = None
                while sample_dict is None and not self._stop.is_set():
                    if self.read_ahead is not None:
                        self.read_ahead.schedule()
                    sample_dict = ds.next_batch(batch_size=1, shuffle=self.shuffle)[0]
                if sample_dict is None:
                    return
                sample = self.convert(sample_dict, ds.train_val_test)
                if self.pin_memory:
                    sample = _pinned(sample)
                if not self._put((sample, sample_dict)):
                    return
                made += 1
            self._put(self._END)
        except BaseException as e:          # handed to the consumer
            self._put(e)

    def sample_fn(self):
        """next engine sample (trainer.train's `sample_fn`)"""
        return next(self)[0]

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            self._q.put(item)               # stay exhausted
            raise StopIteration
        if isinstance(item, BaseException):
            self._q.put(item)
            raise item
        return item

    def close(self):
        self._stop.set()
        while True:                         # unblock a producer waiting on a full queue
            try:
                self._q.get_nowait()
            except queue.Empty:
                break
        self._thread.join(timeout=5.0)
        if self.read_ahead is not None:
            self.dataset._read_ahead = None
            self.read_ahead.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
