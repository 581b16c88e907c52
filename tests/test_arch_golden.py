"""The architecture of the feature path (rows a8-a11) against the layer-by-layer record of the REFERENCE'S OWN
graph-building code (net_builder.extract_features -> FasterRCNNResnet101FeatureExtractor -> resnet_v1_101 ->
stack_blocks_dense, all unmodified), executed against a recording stand-in for TF / TF-slim
(tests/golden/fake_tf.py, tests/golden/make_arch_golden.py).  This pins, to the reference's code rather than to a reading
of it: variable scopes and shapes (what checkpoints map onto), kernel sizes, strides, atrous rates, paddings, which
units project their shortcut, where ReLU and batch norm sit and with which epsilon / scale flags, and every
intermediate tensor shape.  What it cannot pin is the arithmetic of the TF kernels themselves."""
import json
import os

import numpy as np

from monopsr_b200.core import model_spec as ms

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "arch_golden.json")))
REC = G["record"][:G["n_feature_ops"]]           # the feature path; the heads' record follows (HEADS below)
CONVS = [r for r in REC if r["op"] == "conv2d"]


def _tower(enc):
    return [r for r in CONVS if r["scope"].startswith(enc + "/")]


def test_towers_layer_for_layer():
    for enc, in_shape in zip(ms.ENCODERS, ([32, 48, 48, 3], [1, 160, 608, 3])):
        ref = _tower(enc)
        assert ref[0]["in_shape"][0] == in_shape[0]
        consumed = [r for r in ref if "/block4/" not in r["scope"]]          # block4 is built but never consumed
        mine = ms.conv_layers(enc)
        assert [r["scope"] for r in consumed] == [m[0] for m in mine]
        for r, (scope, k, cin, cout, rate) in zip(consumed, mine):
            assert r["kernel"] == [k, k] and (r["cin"], r["cout"]) == (cin, cout), scope
            assert r["rate"] == rate, (scope, r["rate"])
            leaf = scope.rsplit("/", 1)[1]
            if scope.endswith("resnet_v1_101/conv1"):                        # stem: explicit pad 3 + VALID, stride 2
                assert (r["stride"], r["padding"], r["activation"]) == (2, "VALID", "relu")
            else:                                                            # output stride reached: everything stride 1
                assert (r["stride"], r["padding"]) == (1, "SAME"), scope
                assert r["activation"] == ("relu" if leaf in ("conv1", "conv2") else None), scope
            bn = r["batch_norm"]                                             # frozen affine: eps 1e-5, gamma present
            assert bn and bn["epsilon"] == 1e-5 and bn["scale"] is True and bn["is_training"] is False and not r["bias"]
        # spatial sizes: /2 stem, /2 SAME pool, then constant; block3 output has 1024 channels
        h, w = in_shape[1] // 4, in_shape[2] // 4
        assert all(r["out_shape"][1:3] == [h, w] for r in consumed[1:])
        assert consumed[-1]["out_shape"] == [in_shape[0], h, w, 1024]
        b4 = [r for r in ref if "/block4/" in r["scope"]]
        assert len(b4) == 10 and all(r["rate"] == (8 if r["kernel"] == [3, 3] else 1) for r in b4)


def test_stem_padding_and_pool():
    pads = [r for r in REC if r["op"] == "pad"]
    assert len(pads) == 2 and all(p["paddings"] == [[0, 0], [3, 3], [3, 3], [0, 0]] for p in pads)
    pools = [r for r in REC if r["op"] == "max_pool2d" and r["scope"].endswith("/pool1")]
    assert len(pools) == 2 and all((p["kernel"], p["stride"], p["padding"]) == ([3, 3], 2, "SAME") for p in pools)
    assert pools[0]["out_shape"] == [32, 12, 12, 64] and pools[1]["out_shape"] == [1, 40, 152, 64]
    # no subsampling pools inside the blocks (all strides absorbed into atrous rates)
    assert not [r for r in REC if r["op"] == "max_pool2d" and r["kernel"] == [1, 1]]


def test_crop_squash_and_decoder():
    crop = [r for r in REC if r["op"] == "crop_and_resize"]
    assert len(crop) == 1 and crop[0]["in_shape"] == [1, 40, 152, 1024] and crop[0]["out_shape"] == [32, 24, 24, 1024]
    cat = [r for r in REC if r["op"] == "concat"][0]
    assert cat["in_shapes"] == [[32, 12, 12, 1024], [32, 12, 12, 1024]] and cat["axis"] == 3       # crop tower first
    sq = [r for r in CONVS if r["scope"] == "squash/1x1_conv"][0]
    assert (sq["kernel"], sq["cin"], sq["cout"], sq["activation"], sq["bias"], sq["batch_norm"]) == \
        ([1, 1], 2048, 512, "relu", True, None)
    rs = [r for r in REC if r["op"] == "resize_images"]
    assert [(r["size"], r["align_corners"]) for r in rs] == [([24, 24], True), ([48, 48], True)]
    dec = [r for r in CONVS if r["scope"].startswith("map_decoder/")]
    want = [("map_decoder/conv2/conv2_1", 512, 256, 24), ("map_decoder/conv2/conv2_2", 256, 256, 24),
            ("map_decoder/conv3/conv3_1", 256, 128, 48), ("map_decoder/conv3/conv3_2", 128, 128, 48)]
    assert [(r["scope"], r["cin"], r["cout"], r["out_shape"][1]) for r in dec] == want
    for r in dec:       # slim.batch_norm defaults: no gamma, eps 1e-3, decay 0.999, batch statistics while training
        assert (r["kernel"], r["stride"], r["rate"], r["padding"], r["activation"]) == ([3, 3], 1, 1, "SAME", "relu")
        assert r["batch_norm"] == {"decay": 0.999, "center": True, "scale": False, "epsilon": 0.001, "is_training": True}
        assert not r["bias"]
    pool2 = [r for r in REC if r["op"] == "max_pool2d" and r["kernel"] == [2, 2]]
    assert [p["out_shape"] for p in pool2] == [[32, 12, 12, 1024], [32, 6, 6, 512]]
    assert G["features"]["features_for_map"] == [32, 48, 48, 128] and G["features"]["features_for_box_3d"] == [32, 6, 6, 512]


def test_parameter_table_matches_the_recorded_variables():
    """every variable the recorded layers create (weights HWIO; gamma only where scale=True; beta + moving statistics;
    biases where there is no normaliser) is in model_spec.param_table with that shape -- block4 excepted, which the
    reference creates but never uses and which therefore has no kernel, no gradient and no all-reduce here"""
    table = {n: tuple(s) for n, s, _ in ms.param_table()}
    want = {}
    for r in CONVS:
        if "/block4/" in r["scope"]:
            continue
        want[r["scope"] + "/weights"] = (r["kernel"][0], r["kernel"][1], r["cin"], r["cout"])
        if r["batch_norm"]:
            names = ["beta", "moving_mean", "moving_variance"] + (["gamma"] if r["batch_norm"]["scale"] else [])
            for n in names:
                want[r["scope"] + "/BatchNorm/" + n] = (r["cout"],)
        elif r["bias"]:
            want[r["scope"] + "/biases"] = (r["cout"],)
    feature_part = {n: s for n, s in table.items() if not n.startswith("output/")}
    assert feature_part == want
    n_tower = sum(int(np.prod(s)) for n, s in want.items() if n.startswith(ms.ENCODERS[0]) and n.endswith("/weights"))
    assert 27.0e6 < n_tower < 28.0e6          # SURVEY: 27.45 M conv parameters per encoder up to block3


def test_oracle_uses_the_recorded_constants():
    from oracle import network as onet
    assert onet.BN_EPS_RESNET == 1e-5 and onet.BN_EPS_DECODER == 1e-3
    assert [(n, b, u) for n, b, u in onet.BLOCKS] == [(n, b, u) for n, b, u, _ in ms.BLOCKS]
    rates = {}
    for r in _tower(ms.ENCODERS[0]):
        if r["kernel"] == [3, 3] and "/block4/" not in r["scope"]:
            rates.setdefault(r["scope"].split("/")[2], set()).add(r["rate"])
    assert rates == {"block1": {1}, "block2": {2}, "block3": {4}} == {n: {r} for n, _, _, r in ms.BLOCKS}


HEADS = G["record"][G["n_feature_ops"]:]


def test_heads_layer_for_layer():
    """the learned layers of the output builder as its own code creates them (MonoPSROutputBuilder driven like
    MonoPSRModel.build drives it): names, widths, activations, and the ORDER of the concatenated FC inputs"""
    table = {n: tuple(s) for n, s, _ in ms.param_table()}
    want = {}
    for r in HEADS:
        if r["op"] == "conv2d":
            assert (r["scope"], r["kernel"], r["cin"], r["cout"], r["activation"], r["bias"], r["batch_norm"]) == \
                ("output/inst_xyz_map_local/inst_xyz_map_local", [3, 3], 128, 3, None, True, None)
            want[r["scope"] + "/weights"], want[r["scope"] + "/biases"] = (3, 3, 128, 3), (3,)
        elif r["op"] == "fully_connected":
            want[r["scope"] + "/weights"], want[r["scope"] + "/biases"] = (r["cin"], r["cout"]), (r["cout"],)
            hidden = r["scope"].rsplit("/", 1)[1] in ("img_fc", "fc0", "fc1")
            assert r["activation"] == ("relu" if hidden else None) and r["bias"], r["scope"]
        elif r["op"] == "dropout":
            assert r["keep_prob"] == 1.0            # the engine has no dropout: the config must keep it off
    assert {n: s for n, s in table.items() if n.startswith("output/")} == want
    cats = [r for r in HEADS if r["op"] == "concat"]
    widths = lambda r: [s[1] for s in r["in_shapes"]]
    assert cats[0]["in_tags"] == ["output/proposal_fc/proposal_fc/img_fc", "boxes_2d", "boxes_2d", "est_view_angs",
                                  "one_hot(class_indices)", "cam_p"]
    assert widths(cats[0]) == [1024, 4, 1, 1, 1, 12] and cats[0]["out_shape"] == [32, 1043]
    assert cats[1]["in_tags"] == ["output/regression_fc/regression_fc/img_fc", "boxes_2d", "boxes_2d", "est_view_angs",
                                  "one_hot(class_indices)", "output/lwh/lwh", "output/alpha", "output/alpha",
                                  "prop_cen_y", "prop_cen_z"]
    assert widths(cats[1]) == [1024, 4, 1, 1, 1, 3, 12, 12, 1, 1] and cats[1]["out_shape"] == [32, 1060]
    flat = [r for r in HEADS if r["op"] == "flatten"]
    assert all(r["in_shape"] == [32, 6, 6, 512] and r["out_shape"] == [32, 18432] for r in flat) and len(flat) == 2
    # heads hang off fc1 of their stack
    by = {r["scope"]: r for r in HEADS if r["op"] == "fully_connected"}
    assert by["output/lwh/lwh"]["in_tag"] == by["output/alpha"]["in_tag"] == "output/proposal_fc/proposal_fc/fc1"
    assert by["output/cen_y/cen_y"]["in_tag"] == by["output/cen_z_offs/cen_z"]["in_tag"] == "output/regression_fc/regression_fc/fc1"


def test_oracle_concat_order_matches_the_record():
    """perturb one placeholder at a time and see which columns of the first FC layer's input move: the column blocks
    must sit where the reference's concat puts them (checkpoint weights are tied to these positions)"""
    import torch
    from oracle import network as onet
    captured = {}
    orig_fc = onet.fc

    def spy(x, P, scope, relu=True):
        if scope.endswith("/fc0"):
            captured[scope] = x.detach().clone()
        return orig_fc(x, P, scope, relu)
    P = {n: torch.zeros(s, dtype=torch.float64) for n, s, _ in ms.param_table() if n.startswith("output/")}
    g = torch.Generator().manual_seed(0)
    for n in P:
        P[n] = torch.randn(P[n].shape, generator=g, dtype=torch.float64) * 0.01
    # only the head part of the oracle is needed: box_heads on hand-made pooled features
    S = {k: torch.as_tensor(np.asarray(v)).double() if np.asarray(v).dtype.kind == "f" else torch.as_tensor(np.asarray(v))
         for k, v in ms.synthetic_sample(0).items()}
    pooled = torch.randn(32, 6, 6, 512, generator=g, dtype=torch.float64)

    def run(S):
        captured.clear()
        onet.fc = spy
        try:
            onet.box_heads(P, S, pooled)
        finally:
            onet.fc = orig_fc
        return {k: v.clone() for k, v in captured.items()}
    base = run(S)
    p0, r0 = base["output/proposal_fc/proposal_fc/fc0"], base["output/regression_fc/regression_fc/fc0"]
    assert p0.shape == (32, 1043) and r0.shape == (32, 1060)

    def moved(key, delta):
        S2 = dict(S)
        S2[key] = S[key] + delta
        out = run(S2)
        a = (out["output/proposal_fc/proposal_fc/fc0"] - p0).abs().sum(0) > 0
        b = (out["output/regression_fc/regression_fc/fc0"] - r0).abs().sum(0) > 0
        return set(torch.nonzero(a).flatten().tolist()), set(torch.nonzero(b).flatten().tolist())
    pa, ra = moved("est_view_angs", 0.1)
    assert pa == {1029} and 1029 in ra                           # [img_fc 0..1023 | box 1024..1027 | h 1028 | view 1029 | ...]
    pa, ra = moved("boxes_2d", 1.0)
    assert {1024, 1025, 1026, 1027} <= pa and 1029 not in pa
    pa, ra = moved("cam_p", 0.001)
    assert set(range(1031, 1043)) & pa and not (pa & {1029, 1030})


def test_train_op_constants():
    """the optimizer the reference's optimizer_builder constructs from monopsr_model_000.yaml (recorded calls) and the
    clipping its trainer requests, against the constants of the engine's train-op"""
    from monopsr_b200.core import engine as E
    t = G["train_op"]
    by = {r["op"]: r for r in t["record"]}
    lr = by["exponential_decay"]
    assert (lr["learning_rate"], lr["decay_steps"], lr["decay_rate"], lr["staircase"]) == \
        (E.LR_INITIAL, E.LR_DECAY_STEPS, E.LR_DECAY_FACTOR, True)
    adam = by["AdamOptimizer"]
    assert adam["passed"] == [] and adam["learning_rate"] == lr            # only the schedule is passed: TF defaults
    assert adam["config"] == {"beta1": E.ADAM_BETA1, "beta2": E.ADAM_BETA2, "epsilon": E.ADAM_EPSILON}
    ema = by["MovingAverageOptimizer"]
    assert ema["average_decay"] == E.EMA_DECAY and ema["num_updates"] is None and ema["wraps"] == "AdamOptimizer"
    assert t["create_train_op"] == {"clip_gradient_norm": E.CLIP_GRADIENT_NORM}
    e = E.Engine.__new__(E.Engine)
    for step in (0, 1, 9999, 10000, 25000, t["max_iterations"]):
        assert e.learning_rate(step) == lr["learning_rate"] * lr["decay_rate"] ** (step // lr["decay_steps"])
