"""Generate tests/golden/restore_golden.json: which model variable is restored from which checkpoint variable when the
reference initialises from the pre-trained object-detection-API checkpoint -- its OWN code
(monopsr/core/checkpoint_utils.restore_obj_detection_api_weights, MonoPSRModel.get_variable_restore_map,
object_detection/utils/variables_helper.get_variables_available_in_checkpoint), unmodified, executed against a stand-in
for the few TF calls it makes: tf.global_variables() lists the model's variables (names / shapes of the recorded
architecture, tests/golden/arch_golden.json, block4 included), tf.train.NewCheckpointReader lists a synthetic detection
checkpoint (FirstStageFeatureExtractor/... for conv1 + block1-3, SecondStageFeatureExtractor/... for block4, RPN /
box-predictor variables, one variable with a foreign shape, momentum slots), tf.train.Saver records what it is given.
Run from the repository root:  python tests/golden/make_restore_golden.py"""
import json
import os
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_tf_numeric as F  # noqa: E402


class Var(object):
    def __init__(self, name, shape):
        self.op = types.SimpleNamespace(name=name)
        self.shape = types.SimpleNamespace(as_list=lambda s=list(shape): list(s))


def model_variables():
    """every variable the recorded layers create (+ global_step), block4 of both towers included"""
    arch = json.load(open(os.path.join(HERE, "arch_golden.json")))
    out = {}
    for r in arch["record"]:
        if r["op"] == "conv2d":
            out[r["scope"] + "/weights"] = [r["kernel"][0], r["kernel"][1], r["cin"], r["cout"]]
            if r["batch_norm"]:
                for n in ["beta", "moving_mean", "moving_variance"] + (["gamma"] if r["batch_norm"]["scale"] else []):
                    out[r["scope"] + "/BatchNorm/" + n] = [r["cout"]]
            elif r["bias"]:
                out[r["scope"] + "/biases"] = [r["cout"]]
        elif r["op"] == "fully_connected":
            out[r["scope"] + "/weights"], out[r["scope"] + "/biases"] = [r["cin"], r["cout"]], [r["cout"]]
    return out


def detection_checkpoint(model_vars):
    """name -> shape of a faster_rcnn_resnet101 checkpoint: the first-stage extractor holds conv1 + block1-3, the
    second stage block4"""
    ck = {"global_step": []}
    enc = "FirstStageFeatureExtractor_full/"
    for name, shape in model_vars.items():
        if name.startswith(enc):
            rest = name[len(enc):]
            stage = "SecondStageFeatureExtractor/" if "/block4/" in rest else "FirstStageFeatureExtractor/"
            ck[stage + rest] = list(shape)
            ck[stage + rest + "/Momentum"] = list(shape)
    ck["FirstStageFeatureExtractor/resnet_v1_101/conv1/weights"] = [7, 7, 3, 32]          # a foreign shape: must be skipped
    ck["Conv/weights"], ck["FirstStageBoxPredictor/ClassPredictor/weights"] = [3, 3, 1024, 512], [1, 1, 512, 24]
    return ck


def main():
    tf = F.install()
    sys.path.insert(0, "/root/reference/src")
    mv = model_variables()
    ck = detection_checkpoint(mv)
    savers = []
    tf.global_variables = lambda: [Var(n, s) for n, s in mv.items()]
    tf.GraphKeys = types.SimpleNamespace(GLOBAL_STEP="global_step", TRAINABLE_VARIABLES="trainable_variables")
    tf.train = types.SimpleNamespace(
        NewCheckpointReader=lambda path: types.SimpleNamespace(get_variable_to_shape_map=lambda: dict(ck)),
        Saver=lambda var_map=None, **k: savers.append(var_map) or types.SimpleNamespace(restore=lambda sess, path: None))
    from monopsr.core import checkpoint_utils
    import monopsr.core.models.monopsr.monopsr_model as mm
    mm.tf.global_variables = tf.global_variables
    mm.slim.get_or_create_global_step = lambda: Var("global_step", [])
    mm.tf.contrib.framework.filter_variables = None
    framework = types.SimpleNamespace(filter_variables=lambda vs, include_patterns=None, **k: list(vs) if include_patterns is None
                                      else [v for v in vs if any(p in v.op.name for p in include_patterns)])
    mm.tf.contrib = types.SimpleNamespace(framework=framework, slim=mm.slim)
    me = types.SimpleNamespace(net_type="resnet101_4x_squash")
    me.get_variable_restore_map = lambda **kw: mm.MonoPSRModel.get_variable_restore_map(me, **kw)
    checkpoint_utils.tf.train = tf.train
    checkpoint_utils.variables_helper.tf.train = tf.train
    checkpoint_utils.variables_helper.tf.GraphKeys = tf.GraphKeys
    import logging
    logging.disable(logging.WARNING)
    checkpoint_utils.restore_obj_detection_api_weights(None, me, "/ckpt/model.ckpt")
    assert len(savers) == 2
    # [full tower, crop tower]; the variable / checkpoint listings are rebuilt by the test with the functions above
    restored = []
    for s in savers:
        pairs = {ck_name: v.op.name for ck_name, v in s.items()}
        # compact form: every pair is <ckpt_prefix><suffix> -> <model_prefix><suffix>; anything else is kept verbatim
        cp, mp = "FirstStageFeatureExtractor/", sorted(pairs.values())[0].split("/", 1)[0] + "/"
        regular = sorted(c[len(cp):] for c, m in pairs.items() if c.startswith(cp) and m == mp + c[len(cp):])
        other = {c: m for c, m in pairs.items() if not (c.startswith(cp) and m == mp + c[len(cp):])}
        restored.append({"ckpt_prefix": cp, "model_prefix": mp, "suffixes": regular, "other": other})
    out = {"restored": restored}
    path = os.path.join(HERE, "restore_golden.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes;", [len(s["suffixes"]) for s in out["restored"]], "variables restored per saver")


if __name__ == "__main__":
    main()
