"""yaml experiment configuration -- host-side mirror of the reference's config surface.

Reference: src/monopsr/core/config_utils.py:8-85 (duplicate-key guard, recursive conversion to
attribute objects, ``config_name`` = file stem, derived output directories).  The B200 engine
consumes ``monopsr_model_000.yaml`` unchanged; ``validate_for_engine`` states which values the
hand-written kernels implement (everything else raises -- there is no fallback graph).
"""
import os

import yaml


class ConfigObject(object):
    def __repr__(self):
        return "ConfigObject(%s)" % ", ".join(sorted(k for k in self.__dict__))


def config_dict_to_object(d):
    if isinstance(d, dict):
        o = ConfigObject()
        for k, v in d.items():
            setattr(o, k, config_dict_to_object(v))
        return o
    return d


class _NoDupLoader(yaml.SafeLoader):
    pass


def _no_duplicates_constructor(loader, node, deep=False):
    """yaml mappings with a repeated key are an error (config_utils.py:8-25)"""
    mapping = {}
    for key_node, value_node in node.value:
        key = loader.construct_object(key_node, deep=deep)
        if key in mapping:
            raise yaml.constructor.ConstructorError("while constructing a mapping", node.start_mark,
                                                    "found duplicate key (%s)" % key, key_node.start_mark)
        mapping[key] = loader.construct_object(value_node, deep=deep)
    return mapping


_NoDupLoader.add_constructor(yaml.resolver.BaseResolver.DEFAULT_MAPPING_TAG, _no_duplicates_constructor)


def parse_yaml_config(yaml_path, data_dir=None):
    """-> config object with the reference's derived fields (config_name, exp_output_dir,
    train_config.paths_config.{checkpoint_dir, logdir, pred_dir}).  Unlike the reference this does
    not create directories as a side effect."""
    with open(yaml_path, "r") as f:
        config_dict = yaml.load(f, Loader=_NoDupLoader)
    cfg = config_dict_to_object(config_dict)
    cfg.config_name = os.path.splitext(os.path.basename(yaml_path))[0]
    data_dir = data_dir or os.path.join(os.getcwd(), "data")
    cfg.exp_output_dir = data_dir + "/outputs/" + cfg.config_name
    paths = cfg.train_config.paths_config
    if paths.checkpoint_dir is None:
        paths.checkpoint_dir = cfg.exp_output_dir + "/checkpoints"
    else:
        paths.checkpoint_dir = os.path.expanduser(paths.checkpoint_dir)
    paths.logdir = cfg.exp_output_dir + "/logs"
    paths.pred_dir = cfg.exp_output_dir + "/predictions"
    return cfg


def validate_for_engine(cfg):
    """The values of monopsr_model_000.yaml that the sm_100a engine implements."""
    m, d = cfg.model_config, cfg.dataset_config
    errs = []

    def need(cond, msg):
        if not cond:
            errs.append(msg)

    need(m.net_type == "resnet101_4x_squash", "net_type must be resnet101_4x_squash")
    need(list(m.image_input_shape) == [320, 1216], "image_input_shape must be [320, 1216]")
    need(list(m.img_roi_size) == [48, 48] and list(m.map_roi_size) == [48, 48], "roi sizes must be 48x48")
    need(list(m.resized_full_img_shape) == [160, 608], "resized_full_img_shape must be [160, 608]")
    need(d.num_boxes == 32 and d.num_alpha_bins == 12, "num_boxes 32 / num_alpha_bins 12")
    need(list(d.classes) == ["Car"], "classes must be ['Car']")
    need(d.centroid_type == "middle", "centroid_type must be 'middle'")
    need(bool(m.rotate_view), "rotate_view must be True")
    for stack in ("proposal_fc_layers", "regression_fc_layers"):
        fc = getattr(m, stack, None)
        need(fc is not None and list(fc.layer_sizes) == [1024, 1024], "%s.layer_sizes must be [1024, 1024]" % stack)
        need(fc is not None and float(fc.dropout_keep_prob) == 1.0, "%s.dropout_keep_prob must be 1.0 (no dropout kernel)" % stack)
    oc = m.output_config
    expect = dict(inst_xyz_map_local="map", lwh="offset", alpha="dc", view_ang="est", cen_x="from_view_ang_and_z",
                  cen_y="offset", cen_z="offset", centroids="xyz", inst_xyz_map_global="projection",
                  inst_depth_map_global="map")
    for k, v in expect.items():
        need(getattr(oc, k, None) == v, "output_config.%s must be %r" % (k, v))
    lc = m.loss_config
    xyz = list(getattr(lc, "inst_xyz_map_local", []))
    need(len(xyz) >= 2 and xyz[0] in ("smooth_l1_nonzero", "chamfer_dist", "emd"),
         "loss_config.inst_xyz_map_local must be [smooth_l1_nonzero | chamfer_dist | emd, weight]")
    expect_l = dict(lwh=["smooth_l1", 1.0],
                    alpha_cls=["softmax", 0.3, 0.001], alpha_reg=["smooth_l1", 1.0], cen_y=["smooth_l1", 0.1],
                    cen_z=["smooth_l1", 0.1], inst_xyz_map_global=["smooth_l1_nonzero", 0.1],
                    inst_depth_map_global=["smooth_l1_nonzero", 10.0])
    for k, v in expect_l.items():
        need(list(getattr(lc, k, [])) == v, "loss_config.%s must be %r (fused in csrc/heads.cu)" % (k, v))
    opt = cfg.train_config.optimizer
    need(opt.optimizer_type == "adam_optimizer", "optimizer_type must be adam_optimizer")
    if errs:
        raise NotImplementedError("configuration not implemented by the B200 engine: " + "; ".join(errs))
    return True
