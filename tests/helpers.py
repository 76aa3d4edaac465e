"""Shared test helpers: fixture loading, tolerances, small nets rebuilt from fixtures."""
import os

import numpy as np
import torch
import torch.nn as nn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated fp32 tolerances (BASELINE.json north_star / SURVEY.md section 8c)
TOL_POSE = 1e-5      # absolute on R entries and t at EVERY iteration of every level: the coarse levels, poorly
                     # conditioned, reach 5e-6 against the oracle and contract afterwards.  FINAL poses are held to
                     # 1e-6 absolute and TOL_TWIST_REL where a test has them (measured ~1e-7 = the reference's own
                     # fp32-vs-fp64 gap)
TOL_TWIST_REL = 1e-4 # relative on the twist of the estimated motion against the fp32 oracle.  north_star's example is
                     # 1e-5, but the reference's OWN fp32 run differs from its fp64 run by 3e-5 on these inputs (twists
                     # of ~0.02, absolute gap 5e-7): two fp32 evaluations cannot agree below that.  The meaningful gate
                     # is check_not_worse_than_fp32_reference below: against the fp64 oracle, this implementation must
                     # be at least as close as the fp32 oracle is.
TOL_SYS = 1e-4       # Frobenius-relative on J^T W J and J^T W r
TOL_GRAD = 1e-3      # Frobenius-relative on gradients through the unrolled solve


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def frob_rel(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def twist_of(R, t):
    """(B,6) twist [rotation vector, translation] of poses R (B,3,3), t (B,3) (fp64 log map)."""
    R = R.double()
    cos = ((R.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2).clamp(-1, 1)
    th = torch.acos(cos)
    w = torch.stack((R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]), dim=1) / 2
    scale = torch.where(th > 1e-8, th / torch.sin(th).clamp_min(1e-30), torch.ones_like(th))
    return torch.cat((w * scale[:, None], t.double().reshape(-1, 3)), dim=1)


def twist_rel_err(R, t, R_ref, t_ref):
    """max over the batch of |xi - xi_ref| / |xi_ref| (the north_star's "relative on twist")."""
    a, b = twist_of(R, t), twist_of(R_ref, t_ref)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-12)).max().item()


def check_not_worse_than_fp32_reference(R, t, pose32, pose64, slack=1.5):
    """(R, t): this implementation; pose32 / pose64: the oracle run in fp32 and in fp64 on the same inputs.  The error
    against the fp64 result may not exceed the fp32 reference's own error (times ``slack``)."""
    ours = twist_rel_err(R, t, pose64[0], pose64[1])
    ref = twist_rel_err(pose32[0], pose32[1], pose64[0], pose64[1])
    assert ours <= slack * ref + 1e-7, f"twist error vs fp64: ours {ours:.2e}, fp32 reference {ref:.2e}"
    return ours, ref


def level_inputs(g, prefix="in_"):
    keys = ("x0", "x1", "s0", "s1", "invD0", "invD1", "K", "depth0", "depth1")
    return {k: g[prefix + k] for k in keys if (prefix + k) in g}


def damping_mlp(g):
    """DirectSolverNet's 96->128->256->6 ReLU MLP (reference algorithms.py:1834-1842)."""
    sd = {k[len("solver__"):]: v for k, v in g.items() if k.startswith("solver__")}
    lin = [nn.Linear(96, 128), nn.Linear(128, 256), nn.Linear(256, 6)]
    for i, l in enumerate(lin):
        l.weight.data = sd[f"net.{i}.0.weight"].clone()
        l.bias.data = sd[f"net.{i}.0.bias"].clone()
    net = nn.Sequential(lin[0], nn.ReLU(), lin[1], nn.ReLU(), lin[2], nn.ReLU()).eval()
    return net


class ConvMEstimator(nn.Module):
    """DeepRobustEstimator('MultiScale2w') (reference algorithms.py:1432-1478) rebuilt from a
    fixture's state dict: 4 x (dilated conv + BN + ELU), sigmoid; input |r|, x0, x1, upsampled prior."""

    def __init__(self, g):
        super().__init__()
        sd = {k[len("mest__net."):]: v for k, v in g.items() if k.startswith("mest__net.")}
        layers = []
        for i, (cin, cout, dil) in enumerate(((4, 16, 1), (16, 32, 2), (32, 64, 4), (64, 1, 1))):
            conv = nn.Conv2d(cin, cout, 3, padding=dil, dilation=dil, bias=False)
            bn = nn.BatchNorm2d(cout)
            conv.weight.data = sd[f"{i}.0.weight"].clone()
            for k in ("weight", "bias", "running_mean", "running_var"):
                getattr(bn, k).data = sd[f"{i}.1.{k}"].clone()
            layers += [conv, bn, nn.ELU()]
        self.net = nn.Sequential(*layers, nn.Sigmoid()).eval()

    def forward(self, residual, x0, x1, ws=None):
        H, W = residual.shape[2:]
        wl = nn.functional.interpolate(ws, (H, W), mode="bilinear", align_corners=True)
        return self.net(torch.cat((residual.abs(), x0, x1, wl), dim=1))
