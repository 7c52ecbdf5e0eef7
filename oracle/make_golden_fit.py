"""Generate tests/golden/fit_loss.npz by EXECUTING the reference's loss building blocks:
GMoF (lib/Gen_SMPLH/util.py:60-71), SMPLifyAnglePrior / L2Prior (lib/Gen_SMPLH/prior.py:53-106) and
PerspectiveCamera.forward (lib/Gen_SMPLH/camera.py:52-117).  camera.py imports
smplx.lbs.transform_mat, which is absent offline: a stand-in with upstream's published definition
([R | t] padded with the row 0 0 0 1) is injected for that one helper only.
Runs only in the build container.  Usage:  python oracle/make_golden_fit.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import _load_ref_module  # noqa: E402


def main():
    lbs = types.ModuleType("smplx.lbs")
    lbs.transform_mat = lambda R, t: torch.cat([F.pad(R, [0, 0, 0, 1]), F.pad(t, [0, 0, 0, 1], value=1)], dim=2)
    pkg = types.ModuleType("smplx")
    pkg.lbs = lbs
    sys.modules.setdefault("smplx", pkg)
    sys.modules.setdefault("smplx.lbs", lbs)
    util = _load_ref_module("ref_util", "lib/Gen_SMPLH/util.py")
    prior = _load_ref_module("ref_prior", "lib/Gen_SMPLH/prior.py")
    camera = _load_ref_module("ref_camera", "lib/Gen_SMPLH/camera.py")

    rng = np.random.default_rng(33)
    B, Jn = 3, 67
    dt = torch.float64
    joints = torch.tensor(rng.standard_normal((B, Jn, 3)) * 0.4, dtype=dt)
    rot = torch.tensor(np.stack([np.linalg.qr(rng.standard_normal((3, 3)))[0] for _ in range(B)]), dtype=dt)
    trans = torch.tensor(rng.standard_normal((B, 3)) * 0.1 + np.array([0.0, 0.0, 6.0]), dtype=dt)
    center = torch.tensor(rng.standard_normal((B, 2)) * 5 + 256, dtype=dt)
    fx, fy = 5000.0, 4800.0
    cam = camera.PerspectiveCamera(rotation=rot, translation=trans, focal_length_x=fx, focal_length_y=fy,
                                   batch_size=B, center=center, dtype=dt)
    proj = cam(joints).detach()
    gt = proj + torch.tensor(rng.standard_normal((B, Jn, 2)) * 40.0, dtype=dt)
    conf = torch.tensor(rng.random((B, Jn)), dtype=dt)
    rho = 100.0
    rob = util.GMoF(rho=rho)
    joint_diff = rob(gt - proj)
    data_weight = 0.7
    joint_loss = [float(torch.sum(conf[b:b + 1].unsqueeze(-1) ** 2 * joint_diff[b:b + 1]) * data_weight ** 2) for b in range(B)]
    init_loss = [float(torch.sum(torch.pow(gt[b] - proj[b], 2)) * data_weight ** 2) for b in range(B)]

    body_pose = torch.tensor(rng.standard_normal((B, 63)) * 0.4, dtype=dt)
    betas = torch.tensor(rng.standard_normal((B, 10)), dtype=dt)
    emb = torch.tensor(rng.standard_normal((B, 32)), dtype=dt)
    lh = torch.tensor(rng.standard_normal((B, 12)), dtype=dt)
    ap = prior.SMPLifyAnglePrior(dtype=dt)
    l2 = prior.L2Prior()
    angle = ap(body_pose).numpy()                                           # (B,4)
    l2_betas = [float(l2(betas[b:b + 1])) for b in range(B)]
    l2_emb = [float(emb[b:b + 1].pow(2).sum()) for b in range(B)]
    l2_lh = [float(l2(lh[b:b + 1])) for b in range(B)]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fit_loss.npz"),
                        joints=joints.numpy(), rotation=rot.numpy(), translation=trans.numpy(), center=center.numpy(),
                        focal=np.array([fx, fy]), projected=proj.numpy(), gt=gt.numpy(), conf=conf.numpy(), rho=rho,
                        data_weight=data_weight, gmof=joint_diff.numpy(), joint_loss=np.array(joint_loss),
                        init_loss=np.array(init_loss), body_pose=body_pose.numpy(), betas=betas.numpy(),
                        embedding=emb.numpy(), lhand=lh.numpy(), angle_prior=angle, l2_betas=np.array(l2_betas),
                        l2_embedding=np.array(l2_emb), l2_lhand=np.array(l2_lh))
    print("fit_loss golden written: joint_loss", joint_loss, "angle", angle[0])


if __name__ == "__main__":
    main()
