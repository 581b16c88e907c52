"""GPU end-to-end tests of the network engine against the fp64 restatement (oracle/network.py).

Forward bar = the north star's: EVERY forward output within 1e-3 relative (l2) of the fp32 reference graph, in the
engine's DEFAULT precision (measured on a B200: <= 1e-4, profiles/r2_notes.md).  The single-pass tf32 FAST mode
(precision="tf32") is checked against its documented envelope only -- it misses the bar on the decoder maps.
Gradients come from a single-pass tf32 backward on that accurate forward pass: 4e-3 median, < 1e-2 at the stems
(100 layers of backward); a wiring error shows up as O(1).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monopsr_b200.core import model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from oracle import network as onet  # noqa: E402


def l2rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def setup(cuda):
    P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
    eng = Engine(cuda, params=P)
    eng.set_inputs(S)
    eng.forward(train=True)
    Pt = onet.to_torch(P, torch.float64, cuda)
    for v in Pt.values():
        v.requires_grad_(v.dtype == torch.float64)
    St = onet.to_torch(S, torch.float64, cuda)
    out, aux = onet.forward(Pt, St, train=True)
    L, tot = onet.loss(out, St)
    tot.backward()
    eng.backward()
    torch.cuda.synchronize()
    return eng, P, S, Pt, out, aux, L, tot


FORWARD_KEYS = ("centroids", "lwh", "cen_x", "cen_y", "cen_z", "prop_cen_z", "inst_depth_map_global", "proj_err_norm",
                "alpha_bins", "alpha_regs", "lwh_offs", "cen_y_offs", "cen_z_offs", "inst_xyz_map_local")
PARITY_TOL = 1e-3          # BASELINE.json north_star: "within 1e-3 relative fp32"


def test_forward_outputs_meet_the_parity_bar(setup):
    eng, P, S, Pt, out, aux, L, tot = setup
    assert eng.precision != "tf32"       # the default precision is a parity-conforming one
    o = eng.outputs()
    for k in FORWARD_KEYS:
        e = l2rel(o[k], out[k].detach())
        assert e < PARITY_TOL, (k, e)
    assert l2rel(eng.concat[:, :1024], aux["crop_feat"].detach()) < PARITY_TOL
    assert l2rel(eng.squashed, aux["features_squashed"].detach()) < PARITY_TOL
    assert l2rel(eng.dec[-1]["y"], aux["map_features"].detach()) < PARITY_TOL
    el = eng.losses()
    for k, v in L.items():
        assert abs(el[k] - float(v)) < 1e-3 * max(1.0, abs(float(v))), (k, el[k], float(v))
    assert abs(el["total_loss"] - float(tot)) < 1e-3 * float(tot)


def test_fast_mode_stays_inside_its_documented_envelope(cuda, setup):
    """precision="tf32" (single pass): box outputs ~1e-3, decoder maps a few 1e-3 -- NOT parity-conforming"""
    _, P, S, Pt, out, aux, L, tot = setup
    eng = Engine(cuda, params=P, precision="tf32")
    eng.set_inputs(S)
    eng.forward(train=True)
    torch.cuda.synchronize()
    o = eng.outputs()
    tol = {"centroids": 1.5e-3, "lwh": 1.5e-3, "cen_z": 1.5e-3, "inst_depth_map_global": 2e-3, "proj_err_norm": 4e-3,
           "alpha_bins": 5e-3, "alpha_regs": 5e-3, "cen_z_offs": 8e-3, "inst_xyz_map_local": 1.2e-2}
    for k, t in tol.items():
        e = l2rel(o[k], out[k].detach())
        assert e < t, (k, e)


def test_gradients_all_parameters(setup):
    eng, P, S, Pt, out, aux, L, tot = setup
    G = eng.export_grads()
    errs = {}
    for n in eng.trainable_names:
        errs[n] = l2rel(torch.from_numpy(G[n]).to(eng.dev), Pt[n].grad)
    vals = np.array(list(errs.values()))
    # measured on a B200 in the default precision (profiles/r2_check_network_h3.txt): median 4.2e-3, worst 8.5e-3 (the
    # stems, after ~100 layers of single-pass tf32 backward on an accurate forward pass).  Round 1's tf32 forward gave
    # 1e-2 / 1e-1 and this test allowed 4e-2 / 0.25.
    assert np.median(vals) < 1e-2, np.median(vals)
    worst = max(errs, key=errs.get)
    assert errs[worst] < 3e-2, (worst, errs[worst])
    # late layers (few tf32 roundings between loss and parameter) are tight
    for n in ("output/alpha/weights", "output/lwh/lwh/weights", "output/cen_y/cen_y/weights",
              "output/regression_fc/regression_fc/fc1/weights", "output/inst_xyz_map_local/inst_xyz_map_local/weights"):
        assert errs[n] < 8e-3, (n, errs[n])


def test_bn_moving_statistics_updated(setup):
    eng, P, S, Pt, out, aux, L, tot = setup
    for sc, (m, v) in aux["bn_stats"].items():
        mm = eng.view(sc + "/BatchNorm/moving_mean")
        ref = torch.as_tensor(P[sc + "/BatchNorm/moving_mean"]).to(eng.dev).double() * 0.999 + m.detach() * 0.001
        assert l2rel(mm, ref) < 1e-2


def test_graph_replay_equals_eager_and_trains(cuda):
    P, S = ms.init_params(1), ms.synthetic_sample(1)
    e1, e2 = Engine(cuda, params=P), Engine(cuda, params=P)
    for e in (e1, e2):
        e.set_inputs(S)
    for step in range(2):
        e1.set_hyper(step)
        e1.train_step_eager()
        e2.train_step()
    torch.cuda.synchronize()
    # split-K / scatter atomics make the fp32 summation order vary run to run, and Adam's first steps
    # are sign-like (m/sqrt(v) = +-1), so near-zero gradients may move a weight by +-lr either way
    assert l2rel(e1.params, e2.params) < 5e-3
    l0 = None
    losses = []
    for step in range(6):
        e2.train_step()
        losses.append(e2.losses()["total_loss"])
    assert all(np.isfinite(losses))
    assert e2.launches_per_step > 500

