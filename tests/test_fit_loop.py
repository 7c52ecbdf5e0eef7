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


def test_graphed_closure_matches_eager_batch1():
    """The reference's fitting loop runs one body per process (fit_single_frame.py:97): the closure
    replayed as a CUDA graph gives the same loss, gradients and optimiser trajectory as the eager one."""
    import smplk
    from smplk import synthetic
    from smplk.body_models import SMPLH
    from smplk.fitting import GraphedClosure, PerspectiveCamera, SMPLifyLoss
    dev = "cuda:0"
    m = synthetic.make_model("smplh", seed=5)
    rng = np.random.default_rng(9)
    cam = PerspectiveCamera(translation=np.array([[0.1, -0.2, 8.0]]), batch_size=1, center=np.array([[480.0, 270.0]]))
    cam.translation.requires_grad_(False)
    target = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=1).to(dev)
    target.reset_params(betas=rng.standard_normal((1, 16)) * 0.5, body_pose=rng.standard_normal((1, 63)) * 0.25,
                        global_orient=rng.standard_normal((1, 3)) * 0.2)
    with torch.no_grad():
        pc = target(return_verts=False).joints + cam.translation[:, None, :]
        gt2d = pc[..., :2] / pc[..., 2:3] * cam.focal[:, None, :] + cam.center[:, None, :]
    conf = torch.ones(1, gt2d.shape[1], device=dev)
    jw = torch.ones(1, gt2d.shape[1], device=dev)
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, shape_weight=0.5, hand_prior_weight=0.1)

    def run(graphed, steps=40):
        mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=1).to(dev)
        opt = torch.optim.Adam(mod.parameters(), lr=0.02)
        fn = lambda: loss_fn(mod(return_verts=True, return_full_pose=True), cam, gt2d, conf, joint_weights=jw)
        if graphed:
            closure = GraphedClosure(fn, mod.parameters())
        else:
            def closure():
                opt.zero_grad()
                loss = fn()
                loss.backward()
                return loss
        first = float(closure())
        g0 = {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}
        for _ in range(steps):
            opt.zero_grad()
            opt.step(closure)
        return first, g0, float(closure()), {n: p.detach().clone() for n, p in mod.named_parameters()}

    f_e, g_e, l_e, p_e = run(False)
    f_g, g_g, l_g, p_g = run(True)
    assert abs(f_e - f_g) <= 1e-6 * abs(f_e)
    assert set(g_e) == set(g_g)
    for n in g_e:
        assert float((g_e[n] - g_g[n]).abs().max()) <= 1e-5 * max(1.0, float(g_e[n].abs().max())), n
    assert l_g < 0.2 * f_g
    assert abs(l_e - l_g) <= 1e-3 * abs(l_e) + 1e-6
    for n in p_e:
        assert float((p_e[n] - p_g[n]).abs().max()) <= 1e-3, n
