"""Debug/validation driver: product engine vs the fp64 oracle on one synthetic sample (GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200.core import model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from oracle import network as onet  # noqa: E402


def rel(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30)), float((a - b).norm() / (b.norm() + 1e-30))


def main():
    dev = torch.device("cuda:0")
    P = ms.init_params(0, randomize_bn=True)
    S = ms.synthetic_sample(0)
    eng = Engine(dev, params=P)
    eng.set_inputs(S)
    t = time.time()
    eng.forward(train=True)
    torch.cuda.synchronize()
    print("engine fwd (eager, first) %.3fs" % (time.time() - t))
    onet.EMULATE_TF32 = "--emu" in sys.argv
    Pt = onet.to_torch(P, torch.float64, dev)
    for v in Pt.values():
        v.requires_grad_(v.dtype == torch.float64)
    St = onet.to_torch(S, torch.float64, dev)
    out, aux = onet.forward(Pt, St, train=True)
    L, tot = onet.loss(out, St)
    Tc = eng.towers[ms.ENCODERS[0]]
    Tf = eng.towers[ms.ENCODERS[1]]
    print("crop_feat", rel(eng.concat[:, :1024].reshape(-1), aux["crop_feat"].reshape(-1)))
    print("full_feat", rel(Tf["units"][-1]["o"].reshape(-1), aux["full_feat"].reshape(-1)))
    print("concat", rel(eng.concat.reshape(-1), aux["concat"].reshape(-1)))
    print("squashed", rel(eng.squashed.reshape(-1), aux["features_squashed"].reshape(-1)))
    print("map_features", rel(eng.dec[3]["y"].reshape(-1), aux["map_features"].reshape(-1)))
    o = eng.outputs()
    for k in out:
        if k in o and o[k] is not None:
            print("%-24s max-rel %.3e  l2-rel %.3e" % ((k,) + rel(o[k].reshape(-1), out[k].reshape(-1))))
    el = eng.losses()
    for k, v in L.items():
        print("loss %-24s %.6f vs %.6f" % (k, el[k], float(v)))
    print("total", el["total_loss"], float(tot))
    if "--bwd" in sys.argv:
        tot.backward()
        eng.backward()
        torch.cuda.synchronize()
        G = eng.export_grads()
        worst = []
        for n in eng.trainable_names:
            ref = Pt[n].grad
            if ref is None:
                continue
            r = rel(torch.from_numpy(G[n]).to(dev), ref)
            worst.append((r[1], r[0], n))
        worst.sort(reverse=True)
        print("grad l2-rel: median %.3e" % np.median([w[0] for w in worst]))
        for w in worst[:25]:
            print("  %.3e %.3e %s" % w)
        for key in ("output/alpha/weights", "squash/1x1_conv/weights", "map_decoder/conv2/conv2_1/weights",
                    ms.ENCODERS[0] + "/resnet_v1_101/conv1/weights", ms.ENCODERS[1] + "/resnet_v1_101/conv1/weights",
                    ms.ENCODERS[0] + "/resnet_v1_101/block3/unit_23/bottleneck_v1/conv3/BatchNorm/gamma",
                    ms.ENCODERS[1] + "/resnet_v1_101/block1/unit_1/bottleneck_v1/shortcut/BatchNorm/beta"):
            print("  sel %-90s %s" % (key, rel(torch.from_numpy(G[key]).to(dev), Pt[key].grad)))
    if "--step" in sys.argv:
        for i in range(3):
            t = time.time()
            eng.train_step(S)
            torch.cuda.synchronize()
            print("train_step %d: %.4fs loss %.4f launches %s" % (i, time.time() - t, eng.losses()["total_loss"],
                                                                   getattr(eng, "launches_per_step", None)))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(10):
            eng.train_step(S)
        e.record()
        torch.cuda.synchronize()
        print("ms/step %.3f  crops/s %.1f" % (s.elapsed_time(e) / 10, 32 / (s.elapsed_time(e) / 10 / 1e3)))


if __name__ == "__main__":
    main()
