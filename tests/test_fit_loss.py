"""Fitting loss around the body model.
CPU: oracle/fit_oracle.py reproduces the outputs of the reference's own GMoF / SMPLifyAnglePrior /
L2Prior / PerspectiveCamera (tests/golden/fit_loss.npz, made by oracle/make_golden_fit.py).
GPU: the fused loss+gradient kernels against the oracle: loss relative error <= 1e-5, gradients
relative error <= 1e-4 (the BASELINE gradient tolerance)."""
import os

import numpy as np
import pytest
import torch

from oracle import fit_oracle as FO


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "fit_loss.npz"))


def _cam(g, dtype=torch.float64):
    t = lambda k: torch.tensor(g[k], dtype=dtype)   # noqa: E731
    B = g["joints"].shape[0]
    return t("joints"), t("rotation"), t("translation"), t("focal")[None].repeat(B, 1), t("center")


def test_oracle_matches_reference_modules(g):
    joints, rot, tr, focal, center = _cam(g)
    proj = FO.perspective_project(joints, rot, tr, focal, center)
    assert np.abs(proj.numpy() - g["projected"]).max() < 1e-9
    gt, conf = torch.tensor(g["gt"]), torch.tensor(g["conf"])
    assert np.abs(FO.gmof(gt - proj, float(g["rho"])).numpy() - g["gmof"]).max() < 1e-9
    jl = FO.data_term(joints, rot, tr, focal, center, gt, conf, float(g["rho"]), float(g["data_weight"]))
    assert np.abs(jl.numpy() / g["joint_loss"] - 1).max() < 1e-12
    il = FO.data_term(joints, rot, tr, focal, center, gt, None, 0.0, float(g["data_weight"]))
    assert np.abs(il.numpy() / g["init_loss"] - 1).max() < 1e-12
    bp = torch.tensor(g["body_pose"])
    assert np.abs(FO.angle_prior(bp).numpy() - g["angle_prior"]).max() < 1e-12
    pr = FO.prior_term(betas=torch.tensor(g["betas"]), pose_embedding=torch.tensor(g["embedding"]), body_pose=bp,
                       lhand=torch.tensor(g["lhand"]), shape_weight=2.0, body_pose_weight=3.0,
                       bending_prior_weight=5.0, hand_prior_weight=0.5)
    want = 4.0 * g["l2_betas"] + 9.0 * g["l2_embedding"] + 5.0 * g["angle_prior"].sum(1) + 0.25 * g["l2_lhand"]
    assert np.abs(pr.numpy() / want - 1).max() < 1e-12


def _rel(a, b):
    return float((a.detach().double().cpu() - b.detach()).abs().max() / b.detach().abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("rho,shared_cam", [(100.0, False), (0.0, False), (100.0, True)])
def test_reprojection_kernel_matches_oracle(g, rho, shared_cam):
    from smplk.fitting import reprojection_loss
    joints, rot, tr, focal, center = _cam(g)
    if shared_cam:
        rot, tr, focal, center = rot[:1], tr[:1], focal[:1], center[:1]
    gt, conf = torch.tensor(g["gt"]), torch.tensor(g["conf"])
    dw = float(g["data_weight"])
    j64 = joints.clone().requires_grad_(True)
    t64 = tr.clone().requires_grad_(True)
    B = joints.shape[0]
    ex = lambda x: x.expand(B, *x.shape[1:])    # noqa: E731
    ref = FO.data_term(j64, ex(rot), ex(t64), ex(focal), ex(center), gt, conf if rho > 0 else None, rho, dw)
    cw = torch.tensor([1.0, 0.5, 2.0], dtype=torch.float64)                  # per-body cotangents
    (ref * cw).sum().backward()
    dev = "cuda:0"
    jg = joints.float().to(dev).requires_grad_(True)
    tg = tr.float().to(dev).requires_grad_(True)
    out = reprojection_loss(jg, rot.float().to(dev), tg, focal.float().to(dev), center.float().to(dev),
                            gt.float().to(dev), conf.float().to(dev) if rho > 0 else None, rho, dw)
    (out * cw.float().to(dev)).sum().backward()
    assert _rel(out, ref.detach()) <= 1e-5
    if rho > 0 and not shared_cam:
        assert np.abs(out.detach().double().cpu().numpy() / g["joint_loss"] - 1).max() <= 1e-5    # the reference's numbers
    assert _rel(jg.grad, j64.grad) <= 1e-4 and _rel(tg.grad, t64.grad) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("vposer", [True, False])
def test_prior_kernel_matches_oracle(g, vposer):
    from smplk.fitting import fit_priors
    names = ("betas", "embedding", "body_pose", "lhand")
    cpu = {k: torch.tensor(g[k]).requires_grad_(True) for k in names}
    rh = (cpu["lhand"].detach() * -0.5).requires_grad_(True)
    w = dict(shape_weight=2.0, body_pose_weight=3.0, bending_prior_weight=5.0, hand_prior_weight=0.5)
    ref = FO.prior_term(betas=cpu["betas"], pose_embedding=cpu["embedding"] if vposer else None,
                        body_pose=cpu["body_pose"], lhand=cpu["lhand"], rhand=rh, **w)
    ref.sum().backward()
    dev = "cuda:0"
    gpu = {k: torch.tensor(g[k], dtype=torch.float32, device=dev).requires_grad_(True) for k in names}
    rhg = rh.detach().float().to(dev).requires_grad_(True)
    out = fit_priors(betas=gpu["betas"], pose_embedding=gpu["embedding"] if vposer else None,
                     body_pose=gpu["body_pose"], left_hand_pose=gpu["lhand"], right_hand_pose=rhg, **w)
    out.sum().backward()
    assert _rel(out, ref.detach()) <= 1e-5
    for k in names:
        if k == "embedding" and not vposer:
            continue
        assert _rel(gpu[k].grad, cpu[k].grad) <= 1e-4, k
    assert _rel(rhg.grad, rh.grad) <= 1e-4


@pytest.mark.gpu
def test_batched_fitting_step_through_body_model_and_loss():
    """A whole closure evaluation for a batch of 64 bodies: SMPLH forward (CUDA) -> SMPLifyLoss (CUDA)
    -> backward, against the float64 oracle of both (the reference is limited to batch_size == 1)."""
    import smplk
    from smplk import synthetic
    from smplk.body_models import SMPLH
    from smplk.fitting import PerspectiveCamera, SMPLifyLoss
    from oracle import smpl_oracle as O
    B = 64
    m = synthetic.make_model("smplh", seed=3)
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B, create_transl=True).to("cuda:0")
    rng = np.random.default_rng(1)
    vals = dict(betas=rng.standard_normal((B, 16)) * 0.5, global_orient=rng.standard_normal((B, 3)) * 0.2,
                body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)) * 0.5,
                right_hand_pose=rng.standard_normal((B, 12)) * 0.5, transl=rng.standard_normal((B, 3)) * 0.1)
    mod.reset_params(**vals)
    cam = PerspectiveCamera(translation=np.tile([[0.0, 0.0, 8.0]], (B, 1)), batch_size=B,
                            center=np.tile([[256.0, 256.0]], (B, 1)))
    out = mod(return_verts=True, return_full_pose=True)
    Jn = out.joints.shape[1]
    gt = torch.tensor(rng.standard_normal((B, Jn, 2)) * 30 + 256, dtype=torch.float32, device="cuda:0")
    conf = torch.tensor(rng.random((B, Jn)), dtype=torch.float32, device="cuda:0")
    jw = torch.ones(B, Jn, device="cuda:0")
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, body_pose_weight=0.0, shape_weight=5.0,
                          bending_prior_weight=3.17, hand_prior_weight=4.0)
    loss = loss_fn(out, cam, gt, conf, joint_weights=jw)
    loss.backward()
    # oracle
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in vals.items()}
    ro = om.forward(t["betas"], t["global_orient"], t["body_pose"], t["left_hand_pose"], t["right_hand_pose"],
                    transl=t["transl"])
    ct = torch.tensor(np.tile([[0.0, 0.0, 8.0]], (B, 1)), requires_grad=True)
    eye = torch.eye(3, dtype=torch.float64).repeat(B, 1, 1)
    data = FO.data_term(ro.joints, eye, ct, torch.full((B, 2), 5000.0, dtype=torch.float64),
                        torch.full((B, 2), 256.0, dtype=torch.float64), gt.double().cpu(), (jw * conf).double().cpu(), 100.0, 1.0)
    # upstream hands the PROJECTED 45-D hand poses to the prior (lib/Gen_SMPLH/fitting.py:399-413)
    pri = FO.prior_term(betas=t["betas"], body_pose=ro.full_pose[:, 3:66], lhand=ro.left_hand_pose,
                        rhand=ro.right_hand_pose, shape_weight=5.0, bending_prior_weight=3.17, hand_prior_weight=4.0)
    ref = data.sum() + pri.sum()
    ref.backward()
    assert abs(float(loss) / float(ref) - 1) <= 1e-5
    for name in ("betas", "body_pose", "global_orient", "left_hand_pose", "transl"):
        assert _rel(getattr(mod, name).grad, t[name].grad) <= 1e-4, name
    assert _rel(cam.translation.grad, ct.grad) <= 1e-4


@pytest.mark.gpu
def test_smplify_loss_with_shared_betas_row():
    """the body modules allow ONE betas row for B poses (upstream lbs batch = max(...)); the prior
    kernel must then see B rows (broadcast) and the betas gradient is the sum over the bodies."""
    import smplk
    from smplk import synthetic
    from smplk.body_models import SMPLH
    from smplk.fitting import PerspectiveCamera, SMPLifyLoss
    from oracle import smpl_oracle as O
    B = 5
    dev = "cuda:0"
    m = synthetic.make_model("smplh", seed=3)
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=1, create_transl=False,
                create_global_orient=False, create_body_pose=False, create_left_hand_pose=False,
                create_right_hand_pose=False).to(dev)
    rng = np.random.default_rng(9)
    with torch.no_grad():
        mod.betas.copy_(torch.tensor(rng.standard_normal((1, 16)) * 0.5, dtype=torch.float32))
    vals = dict(global_orient=rng.standard_normal((B, 3)) * 0.2, body_pose=rng.standard_normal((B, 63)) * 0.3,
                left_hand_pose=rng.standard_normal((B, 12)) * 0.5, right_hand_pose=rng.standard_normal((B, 12)) * 0.5,
                transl=rng.standard_normal((B, 3)) * 0.1)
    g = {k: torch.tensor(v, dtype=torch.float32, device=dev, requires_grad=True) for k, v in vals.items()}
    out = mod(return_verts=True, return_full_pose=True, **g)
    assert out.betas.shape == (1, 16) and out.joints.shape[0] == B
    cam = PerspectiveCamera(translation=np.tile([[0.0, 0.0, 8.0]], (B, 1)), batch_size=B,
                            center=np.tile([[256.0, 256.0]], (B, 1)))
    Jn = out.joints.shape[1]
    gt = torch.tensor(rng.standard_normal((B, Jn, 2)) * 30 + 256, dtype=torch.float32, device=dev)
    conf = torch.tensor(rng.random((B, Jn)), dtype=torch.float32, device=dev)
    jw = torch.ones(B, Jn, device=dev)
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, shape_weight=5.0, bending_prior_weight=3.17, hand_prior_weight=4.0)
    loss = loss_fn(out, cam, gt, conf, joint_weights=jw)
    loss.backward()
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    ob = mod.betas.detach().double().cpu().requires_grad_(True)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in vals.items()}
    ro = om.forward(ob.expand(B, -1), t["global_orient"], t["body_pose"], t["left_hand_pose"], t["right_hand_pose"],
                    transl=t["transl"])
    ct = torch.tensor(np.tile([[0.0, 0.0, 8.0]], (B, 1)), requires_grad=True)
    eye = torch.eye(3, dtype=torch.float64).repeat(B, 1, 1)
    data = FO.data_term(ro.joints, eye, ct, torch.full((B, 2), 5000.0, dtype=torch.float64),
                        torch.full((B, 2), 256.0, dtype=torch.float64), gt.double().cpu(), (jw * conf).double().cpu(), 100.0, 1.0)
    pri = FO.prior_term(betas=ob.expand(B, -1), body_pose=ro.full_pose[:, 3:66], lhand=ro.left_hand_pose,
                        rhand=ro.right_hand_pose, shape_weight=5.0, bending_prior_weight=3.17, hand_prior_weight=4.0)
    ref = data.sum() + pri.sum()
    ref.backward()
    assert abs(float(loss) / float(ref) - 1) <= 1e-5
    assert mod.betas.grad.shape == (1, 16) and _rel(mod.betas.grad, ob.grad) <= 1e-4
    for name in ("body_pose", "global_orient", "left_hand_pose", "transl"):
        assert _rel(g[name].grad, t[name].grad) <= 1e-4, name
