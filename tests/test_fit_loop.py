"""The caller of the hot path, end to end: a batched re-implementation of the
closure of lib/Gen_SMPLH/fit_single_frame.py:283-470 -- body model forward, camera projection +
robust data term + priors, backward, optimiser step -- entirely on the CUDA path, for a batch of
bodies at once (the reference fits one body per process, fit_single_frame.py:97)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batched_keypoint_fitting_converges():
    import smplk
    from smplk import synthetic
    from smplk.body_models import SMPLH
    from smplk.fitting import PerspectiveCamera, SMPLifyLoss
    B = 48
    dev = "cuda:0"
    m = synthetic.make_model("smplh", seed=11)
    rng = np.random.default_rng(2)
    gt = dict(betas=rng.standard_normal((B, 16)) * 0.5, global_orient=rng.standard_normal((B, 3)) * 0.2,
              body_pose=rng.standard_normal((B, 63)) * 0.25, left_hand_pose=rng.standard_normal((B, 12)) * 0.3,
              right_hand_pose=rng.standard_normal((B, 12)) * 0.3)
    cam = PerspectiveCamera(translation=np.tile([[0.0, 0.0, 10.0]], (B, 1)), batch_size=B,
                            center=np.tile([[512.0, 512.0]], (B, 1)))
    cam.translation.requires_grad_(False)
    # "detections": projected joints of the ground-truth bodies
    target = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)
    target.reset_params(**gt)
    with torch.no_grad():
        tj = target(return_verts=False).joints
        pc = tj + cam.translation[:, None, :]
        gt2d = pc[..., :2] / pc[..., 2:3] * cam.focal[:, None, :] + cam.center[:, None, :]
    conf = torch.ones(B, tj.shape[1], device=dev)
    jw = torch.ones(B, tj.shape[1], device=dev)

    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)      # starts at the mean pose
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, shape_weight=0.5, bending_prior_weight=0.0, hand_prior_weight=0.1)
    opt = torch.optim.Adam(mod.parameters(), lr=0.02)

    def closure():
        opt.zero_grad()
        out = mod(return_verts=True, return_full_pose=True)
        loss = loss_fn(out, cam, gt2d, conf, joint_weights=jw)
        loss.backward()
        return loss

    def pixel_error():
        with torch.no_grad():
            j = mod(return_verts=False).joints
            pc = j + cam.translation[:, None, :]
            px = pc[..., :2] / pc[..., 2:3] * cam.focal[:, None, :] + cam.center[:, None, :]
            return (px - gt2d).norm(dim=-1).mean(dim=1)

    err0 = pixel_error()
    first = float(closure())
    for _ in range(150):
        opt.step(closure)
    last = float(closure())
    assert np.isfinite(last) and last < 0.02 * first, (first, last)
    # every body improved, not just the sum: per-body reprojection error in pixels
    err = pixel_error()
    assert float((err / err0).max()) < 0.25, (float(err0.mean()), float(err.mean()), float((err / err0).max()))
