"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, differentiable) of the fitting loss around the
body model, used by tests/ as the checker of the CUDA loss kernels.  Never imported by the product.

  perspective_project  lib/Gen_SMPLH/camera.py:93-117 (PerspectiveCamera.forward; transform_mat of
                       upstream smplx.lbs = [R | t; 0 0 0 1])
  gmof                 lib/Gen_SMPLH/util.py:60-71
  angle_prior          lib/Gen_SMPLH/prior.py:53-97 (SMPLifyAnglePrior, with_global_pose=False)
  smplify_loss         lib/Gen_SMPLH/fitting.py:365-449 (joint term + shape / pose / bending / hand
                       priors; interpenetration, face and jaw terms are outside this path)
  camera_init_loss     lib/Gen_SMPLH/fitting.py:486-503 (joint term)

Pinned by tests/golden/fit_loss.npz, made by oracle/make_golden_fit.py which executes the reference's
GMoF, SMPLifyAnglePrior, L2Prior and PerspectiveCamera (the latter with a two-line stand-in for the
absent smplx.lbs.transform_mat).
"""
import torch


def perspective_project(points, rotation, translation, focal, center):
    pc = torch.einsum("bki,bji->bjk", rotation, points) + translation[:, None, :]
    img = pc[:, :, :2] / pc[:, :, 2:3]
    return img * focal[:, None, :] + center[:, None, :]


def gmof(residual, rho):
    sq = residual ** 2
    return rho ** 2 * sq / (sq + rho ** 2)


def angle_prior(body_pose):
    idx = torch.tensor([55, 58, 12, 15]) - 3
    signs = torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=body_pose.dtype)
    return torch.exp(body_pose[:, idx] * signs).pow(2)


def data_term(joints, rotation, translation, focal, center, gt, weights, rho, data_weight):
    """(B,) per-body joint loss of SMPLifyLoss (rho > 0) or SMPLifyCameraInitLoss (rho <= 0)."""
    proj = perspective_project(joints, rotation, translation, focal, center)
    r = gt - proj
    d = gmof(r, rho) if rho > 0 else r ** 2
    w = torch.ones_like(d[..., 0]) if weights is None else weights.expand(d.shape[0], -1)
    return (w.unsqueeze(-1) ** 2 * d).sum(dim=(1, 2)) * data_weight ** 2


def prior_term(betas=None, pose_embedding=None, body_pose=None, lhand=None, rhand=None, shape_weight=0.0,
               body_pose_weight=0.0, bending_prior_weight=0.0, hand_prior_weight=0.0):
    """(B,) per-body prior loss."""
    ref = next(t for t in (betas, pose_embedding, body_pose, lhand, rhand) if t is not None)
    out = torch.zeros(ref.shape[0], dtype=ref.dtype)
    if betas is not None:
        out = out + betas.pow(2).sum(1) * shape_weight ** 2
    if pose_embedding is not None:
        out = out + pose_embedding.pow(2).sum(1) * body_pose_weight ** 2
    elif body_pose is not None:
        out = out + body_pose.pow(2).sum(1) * body_pose_weight ** 2
    if body_pose is not None:
        out = out + angle_prior(body_pose).sum(1) * bending_prior_weight
    for h in (lhand, rhand):
        if h is not None:
            out = out + h.pow(2).sum(1) * hand_prior_weight ** 2
    return out
