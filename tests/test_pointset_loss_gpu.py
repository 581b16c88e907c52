"""ChamferDistance / EarthMoversDistance on the CUDA ops: as stand-alone loss classes and metrics
and as the TRAINING loss of the local xyz map inside the engine step (loss_config.inst_xyz_map_local = [chamfer_dist |
emd, w]; loss_builder.py:32-36,60-84, losses_custom.py:135-198, monopsr_model.py:580-586,1112-1170)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monopsr_b200.core import losses_custom, model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from monopsr_b200.tf_ops.approxmatch import tf_approxmatch  # noqa: E402
from oracle import network as onet  # noqa: E402
from oracle import tfops  # noqa: E402


def l2rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _maps(cuda, B=32, seed=3):
    g = torch.Generator(device="cpu").manual_seed(seed)
    pred = (torch.rand(B, 48, 48, 3, generator=g) * 4 - 2).to(cuda)
    gt = (torch.rand(B, 48, 48, 3, generator=g) * 4 - 2).to(cuda)
    mask = (torch.rand(B, 48, 48, 1, generator=g) < 0.6).float().to(cuda)
    return pred, gt, mask


def test_loss_classes_and_metrics_on_the_cuda_ops_vs_oracle(cuda):
    """masked (32,48,48,3) sample: the classes / metrics on the sm_100a ops against the CPU oracle of the ops with the
    same assembly (mask both clouds, (B,2304,3), sum / B; per-object / number of valid pixels)"""
    pred, gt, mask = _maps(cuda)
    B = pred.shape[0]
    p = (pred * mask).reshape(B, -1, 3).cpu().numpy()
    t = (gt * mask).reshape(B, -1, 3).cpu().numpy()
    d1, _, d2, _ = tfops.nn_distance(p, t, "gpu")
    ch_ref = (d1.sum(1) + d2.sum(1))
    got = losses_custom.ChamferDistance()(pred, gt, weights=mask)
    assert abs(float(got) - ch_ref.sum() / B) < 1e-4 * ch_ref.sum() / B
    # EMD on 4 objects (the CPU oracle of approx_match at 2304 x 2304 takes seconds per element)
    sub = slice(0, 4)
    m_ref = tfops.approx_match(p[sub], t[sub], "gpu")
    c_ref = tfops.match_cost(p[sub], t[sub], m_ref, "gpu")
    got = losses_custom.EarthMoversDistance()(pred[sub], gt[sub], weights=mask[sub])
    assert abs(float(got) - c_ref.sum() / 4) < 1e-3 * c_ref.sum() / 4
    met = losses_custom.point_set_metrics(pred[sub], gt[sub], mask[sub], 3)
    nv = mask[sub].sum(dim=(1, 2, 3)).cpu().numpy()
    np.testing.assert_allclose(met["metric_chamfer"].cpu().numpy(), ch_ref[:3] / nv[:3], rtol=1e-4)
    np.testing.assert_allclose(met["metric_emd"].cpu().numpy(), c_ref[:3] / nv[:3], rtol=1e-3)


def test_loss_classes_backpropagate_through_the_op_gradients(cuda):
    pred, gt, mask = _maps(cuda, B=2, seed=4)
    for cls in (losses_custom.ChamferDistance, losses_custom.EarthMoversDistance):
        x = pred.clone().requires_grad_()
        cls()(x, gt, weights=mask).backward()
        xr = pred.double().clone().requires_grad_()
        if cls is losses_custom.ChamferDistance:
            ref = onet.chamfer_loss(xr, gt.double(), mask.double())
        else:
            ref = onet.emd_loss(xr, gt.double(), mask.double(), lambda p, t: tf_approxmatch.approx_match(p.float(), t.float()))
        ref.backward()
        assert l2rel(x.grad, xr.grad) < 2e-3, cls.__name__
        assert float((x.grad * (1 - mask)).abs().max()) == 0.0       # masked pixels get no gradient


@pytest.mark.parametrize("kind,weight", [("chamfer_dist", 10.0), ("emd", 1.0)])
def test_point_set_loss_as_the_training_loss_of_the_engine(cuda, kind, weight):
    """loss value, total, and the gradients that flow from it: d_xyz feeds the decoder and both towers"""
    P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
    eng = Engine(cuda, params=P, xyz_loss=(kind, weight))
    eng.set_inputs(S)
    eng.forward(train=True)
    eng.backward()
    torch.cuda.synchronize()
    Pt = onet.to_torch(P, torch.float64, cuda)
    for v in Pt.values():
        v.requires_grad_(v.dtype == torch.float64)
    St = onet.to_torch(S, torch.float64, cuda)
    out, aux = onet.forward(Pt, St, train=True)
    L, tot = onet.loss(out, St, xyz_loss=(kind, weight),
                       match_fn=lambda p, t: tf_approxmatch.approx_match(p.float().contiguous(), t.float().contiguous()))
    tot.backward()
    el = eng.losses()
    assert abs(el["inst_xyz_map_local"] - float(L["inst_xyz_map_local"])) < 2e-3 * abs(float(L["inst_xyz_map_local"]))
    assert abs(el["total_loss"] - float(tot)) < 2e-3 * float(tot)
    G = eng.export_grads()
    for n, tol in (("output/inst_xyz_map_local/inst_xyz_map_local/weights", 2e-2),
                   ("map_decoder/conv3/conv3_2/weights", 5e-2), ("squash/1x1_conv/weights", 8e-2)):
        assert l2rel(torch.from_numpy(G[n]).to(cuda), Pt[n].grad) < tol, n
    # and the step trains (graph capture included)
    e2 = Engine(cuda, params=ms.init_params(1), xyz_loss=(kind, weight))
    ls = []
    for _ in range(4):
        e2.train_step(ms.synthetic_sample(1))
        ls.append(e2.losses()["total_loss"])
    assert all(np.isfinite(ls))


def test_engine_from_yaml_config_with_a_point_set_loss(cuda):
    import os
    from monopsr_b200.core import config_utils
    cfg = config_utils.parse_yaml_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "configs", "monopsr_model_000.yaml"))
    cfg.model_config.loss_config.inst_xyz_map_local = ["chamfer_dist", 5.0]
    cfg.train_config.optimizer.adam_optimizer.initial_learning_rate = 1e-4
    eng = Engine.from_config(cuda, cfg)
    assert eng.xyz_loss_type == "chamfer_dist" and eng.xyz_loss_weight == 5.0 and eng.learning_rate(0) == 1e-4
    cfg.model_config.loss_config.inst_xyz_map_local = ["berHu", 5.0]
    with pytest.raises(NotImplementedError):
        Engine.from_config(cuda, cfg)
