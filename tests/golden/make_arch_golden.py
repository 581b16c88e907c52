"""Generate tests/golden/arch_golden.json: the layer-by-layer record of the feature path of the network (rows a8-a11:
both ResNet-101 towers at output stride 4, crop-and-resize + pool of the full-image features, squash, map decoder) as
the REFERENCE'S OWN graph-building code constructs it -- monopsr/builders/net_builder.extract_features with
FasterRCNNResnet101FeatureExtractor, object_detection/nets/resnet_v1.py and resnet_utils.py, unmodified, imported from
/root/reference/src and executed against the recording stand-in for TensorFlow / TF-slim in tests/golden/fake_tf.py.
The yaml that configures it is the reference's configs/monopsr_model_000.yaml.
Run from the repository root:  python tests/golden/make_arch_golden.py"""
import json
import os
import sys
import types

import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_tf  # noqa: E402


def main():
    tf = fake_tf.install()
    sys.path.insert(0, "/root/reference/src")
    from monopsr.builders import net_builder
    from monopsr.core import constants

    cfg = yaml.safe_load(open("/root/reference/src/monopsr/configs/monopsr_model_000.yaml"))
    mc = cfg["model_config"]

    def obj(d):
        return types.SimpleNamespace(**{k: obj(v) for k, v in d.items()}) if isinstance(d, dict) else d
    model_config = obj(mc)
    n = cfg["dataset_config"]["num_boxes"]
    model = types.SimpleNamespace(is_training=True, pl_boxes_2d_norm=fake_tf.Tensor([n, 4]), num_boxes=n,
                                  map_roi_size=mc["map_roi_size"])
    h, w = mc["img_roi_size"]
    fh, fw = mc["resized_full_img_shape"]
    inputs = {constants.NET_IN_RGB_CROP: fake_tf.Tensor([n, h, w, 3]), constants.NET_IN_FULL_IMG: fake_tf.Tensor([1, fh, fw, 3])}
    feats = net_builder.extract_features(model, mc["net_type"], model_config, inputs, True)
    n_feature_ops = len(fake_tf.RECORD)

    # ---- the learned layers of the output builder (rows a12-a14), driven as MonoPSRModel.build drives them
    # (monopsr_model.py:295-392): placeholders are shape-only tensors tagged with their names
    from monopsr.core.models.monopsr.monopsr_output_builder import MonoPSROutputBuilder
    T = fake_tf.Tensor
    for k, v in feats.items():
        v.tag = k
    dataset_config = obj(cfg["dataset_config"])
    ob = MonoPSROutputBuilder(model_config.output_config, model_config, dataset_config, feats, n, mc["map_roi_size"],
                              T([3, 4], tag="cam_p"), "train")
    with tf.variable_scope("output"):
        ob.add_inst_xyz_maps_local(gt_inst_xyz_maps_local=T([n, 48, 48, 3], tag="gt_xyz"))
        boxes_2d, view = T([n, 4], tag="boxes_2d"), T([n, 1], tag="est_view_angs")
        cls = T([n, 1], tag="class_indices")
        ob.add_proposal_fc_features(boxes_2d=boxes_2d, view_angs=view, class_indices=cls, image_shape=mc["image_input_shape"])
        f1 = ob.get_proposal_fc_features()
        ob.add_lwh_output(features_to_use=f1, est_lwh=T([n, 3], tag="mean_lwh"), gt_lwh=T([n, 3], tag="gt_lwh"))
        ob.add_alpha_output(features_to_use=f1, gt_alpha=T([n, 1]), gt_alpha_dc=[T([n, 1]), T([n, 12])])
        od = ob.get_output_dict()
        prop_y, prop_z = T([n, 1], tag="prop_cen_y"), T([n, 1], tag="prop_cen_z")
        ob.add_regression_fc_features(
            boxes_2d=boxes_2d, view_angs=view, class_indices=cls, image_shape=mc["image_input_shape"],
            est_lwh_off=od[constants.KEY_LWH + "_offs"], est_alpha_bins=od[constants.KEY_ALPHA_BINS],
            est_alpha_regs=od[constants.KEY_ALPHA_REGS], prop_cen_y=prop_y, prop_cen_z=prop_z,
            max_depth=cfg["dataset_config"]["obj_filter_config"]["depth_range"][1])
        f2 = ob.get_regression_fc_features()
        ob.add_cen_y_output(output_key=constants.KEY_CEN_Y, features_in=f2, prop_cen_y=prop_y, gt_cen_y=T([n, 1]))
        ob.add_cen_z_output(output_key=constants.KEY_CEN_Z, features_in=f2, prop_cen_z=prop_z, gt_cen_z=T([n, 1]))

    n_layer_ops = len(fake_tf.RECORD)

    # ---- the train-op (row a18): the optimizer the reference's optimizer_builder constructs from the yaml, and the
    # gradient clipping its trainer asks slim.learning.create_train_op for (a literal in core/trainer.py, read with ast)
    import ast
    from monopsr.builders import optimizer_builder
    optimizer_builder.build(obj(cfg["train_config"]["optimizer"]), set(), fake_tf.Tensor([], tag="global_step"))
    clip = None
    for node in ast.walk(ast.parse(open("/root/reference/src/monopsr/core/trainer.py").read())):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "create_train_op":
            clip = {k.arg: ast.literal_eval(k.value) for k in node.keywords if k.arg == "clip_gradient_norm"}
    train_op = {"record": fake_tf.RECORD[n_layer_ops:], "create_train_op": clip,
                "max_iterations": cfg["train_config"]["max_iterations"]}
    del fake_tf.RECORD[n_layer_ops:]

    out = {"net_type": mc["net_type"], "record": fake_tf.RECORD, "n_feature_ops": n_feature_ops, "train_op": train_op,
           "features": {k: v.get_shape().as_list() for k, v in feats.items()},
           "extractor_config": mc["net_config"][mc["net_type"]]}
    path = os.path.join(HERE, "arch_golden.json")
    json.dump(out, open(path, "w"))
    for r in fake_tf.RECORD[n_feature_ops:]:
        print({k: v for k, v in r.items() if k not in ("in_shape", "in_shapes")})
    convs = [r for r in fake_tf.RECORD[:n_feature_ops] if r["op"] == "conv2d"]
    print("wrote", path, os.path.getsize(path), "bytes;", len(fake_tf.RECORD), "ops,", len(convs), "convolutions")
    print({k: v for k, v in out["features"].items()})
    print(train_op)
    for r in convs[:4] + convs[-5:]:
        print(r["scope"], r["kernel"], "s%d r%d" % (r["stride"], r["rate"]), r["padding"], r["cin"], "->", r["cout"], r["activation"],
              "bn" if r["batch_norm"] else "bias" if r["bias"] else "-", r["out_shape"])


if __name__ == "__main__":
    main()
