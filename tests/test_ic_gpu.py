"""IC tracker (TrustRegionBase drop-in) against fixtures made by the reference's TrustRegionBase with
mEstimator None / MultiScale2w CNN and solver Direct-Nodamping / Direct-ResVol, and against the oracle."""
import pytest
import torch
import torch.nn as nn

from deep_prob_feature_track_b200 import algorithms as A
from helpers import TOL_POSE, ConvMEstimator, damping_mlp, frob_rel, level_inputs, load_golden
from oracle import ic_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(name, g):
    solver = A.DirectSolverNet("Direct-ResVol" if name != "ic_plain" else "Direct-Nodamping")
    if solver.net is not None:
        ref = damping_mlp(g)
        for i in range(3):
            solver.net[i][0].weight.data = ref[2 * i].weight.data.clone()
            solver.net[i][0].bias.data = ref[2 * i].bias.data.clone()
    mest = ConvMEstimator(g).to(DEV) if name == "ic_deepic" else None
    return A.TrustRegionBase(max_iter=int(g["flags"][3]), mEst_func=mest, solver_func=solver).to(DEV).eval()


@pytest.mark.parametrize("name", ["ic_plain", "ic_resvol", "ic_deepic"])
def test_ic_level_matches_reference(name):
    g = load_golden(name)
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}
    mod = build(name, g)
    with torch.no_grad():
        (R, t), w = mod([g["R0"].to(DEV), g["t0"].to(DEV)], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"],
                        wPrior=g["wprior"].to(DEV))
        loss = mod.forward_residuals([g["R0"].to(DEV), g["t0"].to(DEV)], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"],
                                     lv["K"], wPrior=g["wprior"].to(DEV))
    assert (R.cpu() - g["R_out"]).abs().max() < TOL_POSE, (R.cpu() - g["R_out"]).abs().max()
    assert (t.cpu() - g["t_out"]).abs().max() < TOL_POSE, (t.cpu() - g["t_out"]).abs().max()
    assert frob_rel(w.cpu(), g["weights"]) < 1e-4
    assert frob_rel(loss.cpu(), g["res_loss"]) < 1e-4


def test_ic_pieces_against_oracle():
    """Residual map + mask (bit-exact), J^T W J and J^T W r of the split entry points."""
    g = load_golden("ic_plain")
    lv = level_inputs(g)
    B, C, H, W = lv["x0"].shape
    lvl = A._IcLevel(*(lv[k].to(DEV) for k in ("x0", "x1", "invD0", "invD1", "K")))
    rows = A.pack_pose((g["R0"], g["t0"])).to(DEV)
    r, occ = lvl.residual(rows, first=True)
    px, py = O.pixel_rays(lv["K"], H, W)
    r_o, occ_o = O.ic_residual(g["R0"], g["t0"], lv["invD0"], lv["invD1"], lv["x0"], lv["x1"], px, py, lv["K"])
    assert torch.equal(occ.cpu(), occ_o)
    assert torch.equal(r.cpu(), r_o)          # same rounding chain as the oracle's explicit sampler
    gen = torch.Generator().manual_seed(0)
    w = torch.rand((B, C, H, W), generator=gen) + 0.5
    gx, gy = O.sobel_unit(lv["x0"])
    Ju, Jv = O.warp_rows(lv["invD0"], lv["K"], px, py)
    J = (gx.reshape(B, C, -1, 1) * Ju.view(B, 1, -1, 6) + gy.reshape(B, C, -1, 1) * Jv.view(B, 1, -1, 6)).reshape(B, -1, 6)
    A_o = torch.bmm(J.transpose(1, 2), w.reshape(B, -1, 1) * J)
    b_o = torch.bmm(J.transpose(1, 2), (w * r_o).reshape(B, -1, 1))
    A_c = A._tri_to_full(lvl.normal_matrix(w.to(DEV))).cpu()
    b_c = lvl.rhs(w.to(DEV), rows.unsqueeze(0))[0].cpu()
    assert frob_rel(A_c, A_o) < 1e-4
    assert frob_rel(b_c.unsqueeze(2), b_o) < 1e-4


def test_ic_forward_raises_on_a_system_that_is_not_positive_definite():
    """Negative M-estimator weights make J^T W J negative definite: Cholesky has no pivot, the device status word says
    so and forward raises instead of returning NaN poses (the U_IC path's raise_if_bad; the reference's torch.inverse
    raises on a singular system, alg:2085-2092)."""
    g = load_golden("ic_plain")
    lv = {k: v.to(DEV) for k, v in level_inputs(g).items()}

    class Negative(nn.Module):
        def forward(self, r, x0, x1, wPrior):
            return -torch.ones_like(r)

    mod = A.TrustRegionBase(max_iter=1, mEst_func=Negative(), solver_func=A.DirectSolverNet("Direct-Nodamping")).to(DEV).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="not positive definite"):
        mod([g["R0"].to(DEV), g["t0"].to(DEV)], lv["x0"], lv["x1"], lv["invD0"], lv["invD1"], lv["K"])
