"""Where does the tf32 forward error come from?  Pure-oracle study (torch fp64 on the GPU):
compare the plain fp64 graph with tf32-emulating variants."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200.core import model_spec as ms
from oracle import network as onet
dev = torch.device(os.environ.get("MPB_STUDY_DEV", "cuda:0"))
P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
Pt, St = onet.to_torch(P, torch.float64, dev), onet.to_torch(S, torch.float64, dev)
def run():
    with torch.no_grad():
        out, aux = onet.forward(Pt, St, train=True)
    return out, aux
ref, raux = run()
def rel(a, b): return float((a - b).norm() / b.norm())
for name, emu, resfull in (("tf32 everywhere (product today)", True, False), ("tf32 GEMM inputs, fp32 residual stream", True, True)):
    onet.EMULATE_TF32, onet.RESIDUAL_FULL = emu, resfull
    out, aux = run()
    print(name)
    print("   crop_feat %.2e  full_feat %.2e  squashed %.2e  map_features %.2e" % (
        rel(aux["crop_feat"], raux["crop_feat"]), rel(aux["full_feat"], raux["full_feat"]),
        rel(aux["features_squashed"], raux["features_squashed"]), rel(aux["map_features"], raux["map_features"])))
    print("   " + "  ".join("%s %.2e" % (k, rel(out[k], ref[k])) for k in ("inst_xyz_map_local", "centroids", "lwh", "alpha_bins", "alpha_regs", "cen_z_offs", "cen_y_offs", "proj_err_norm")))
onet.EMULATE_TF32 = onet.RESIDUAL_FULL = False

# ---- which part amplifies?  towers vs the rest (squash + decoder + heads)
orig_block3 = onet.resnet101_block3
def make(tower_emu, rest_emu, resfull):
    def block3(x, P_, scope):
        onet.EMULATE_TF32, onet.RESIDUAL_FULL = tower_emu, resfull
        y = orig_block3(x, P_, scope)
        onet.EMULATE_TF32 = rest_emu
        return y
    return block3
for name, te, re_, rf in (("towers tf32(+fp32 residual), rest exact", True, False, True), ("towers exact, rest tf32", False, True, False)):
    onet.resnet101_block3 = make(te, re_, rf)
    onet.EMULATE_TF32 = re_
    out, aux = run()
    print(name)
    print("   squashed %.2e  map_features %.2e  xyz %.2e  alpha %.2e  cen_z_offs %.2e" % (
        rel(aux["features_squashed"], raux["features_squashed"]), rel(aux["map_features"], raux["map_features"]),
        rel(out["inst_xyz_map_local"], ref["inst_xyz_map_local"]), rel(out["alpha_bins"], ref["alpha_bins"]), rel(out["cen_z_offs"], ref["cen_z_offs"])))
onet.resnet101_block3 = orig_block3
onet.EMULATE_TF32 = onet.RESIDUAL_FULL = False
# per decoder layer error growth in the all-tf32 setting is printed by comparing bn_stats
