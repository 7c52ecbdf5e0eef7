"""Keypoint fitting with the drop-in modules -- the loop of lib/Gen_SMPLH/fit_single_frame.py:283-470
(body model -> camera projection -> robust data term + priors -> backward -> optimiser step), for a
batch of bodies at once and with the closure replayed as a CUDA graph.

    python examples/fit_keypoints.py [batch] [iterations]

Synthetic model and synthetic "detections" (projected joints of random ground-truth bodies): there
is no network for the real SMPL-H files, and the path only needs tensors of the canonical shapes.
With a real model: SMPLH(model_path=".../SMPLH_neutral.pkl", ...), as lib/gen_smplh.py:75-90 does.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk  # noqa: E402
from smplk import synthetic  # noqa: E402
from smplk.body_models import SMPLH  # noqa: E402
from smplk.fitting import GraphedClosure, PerspectiveCamera, SMPLifyLoss  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = "cuda:0"
    model = synthetic.make_model("smplh", seed=11)
    rng = np.random.default_rng(2)
    cam = PerspectiveCamera(translation=np.tile([[0.0, 0.0, 10.0]], (B, 1)), batch_size=B,
                            center=np.tile([[512.0, 512.0]], (B, 1)))
    cam.translation.requires_grad_(False)

    def project(joints):
        pc = joints + cam.translation[:, None, :]
        return pc[..., :2] / pc[..., 2:3] * cam.focal[:, None, :] + cam.center[:, None, :]

    truth = SMPLH(model=model, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)
    truth.reset_params(betas=rng.standard_normal((B, 16)) * 0.5, global_orient=rng.standard_normal((B, 3)) * 0.2,
                       body_pose=rng.standard_normal((B, 63)) * 0.25,
                       left_hand_pose=rng.standard_normal((B, 12)) * 0.3,
                       right_hand_pose=rng.standard_normal((B, 12)) * 0.3)
    with torch.no_grad():
        gt2d = project(truth(return_verts=False).joints)
    conf = torch.ones(B, gt2d.shape[1], device=dev)

    body = SMPLH(model=model, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)   # starts at the mean pose
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, shape_weight=0.5, hand_prior_weight=0.1)
    opt = torch.optim.Adam(body.parameters(), lr=0.02)
    # return_verts=False: only the vertex-pick joints' vertices are blended and skinned
    closure = GraphedClosure(lambda: loss_fn(body(return_verts=False, return_full_pose=True), cam, gt2d, conf,
                                             joint_weights=conf), body.parameters())

    def pixel_error():
        with torch.no_grad():
            return (project(body(return_verts=False).joints) - gt2d).norm(dim=-1).mean(dim=1)

    e0 = pixel_error()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        opt.step(closure)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e1 = pixel_error()
    print("bodies %d, %d Adam steps in %.1f ms (%.3f ms per step, closure replayed as a CUDA graph)" % (
        B, iters, dt * 1e3, dt * 1e3 / iters))
    print("mean reprojection error: %.1f px -> %.2f px (worst body %.2f px)" % (
        float(e0.mean()), float(e1.mean()), float(e1.max())))
    verts = body(return_verts=True).vertices            # the fitted meshes, (B, 6890, 3)
    print("fitted vertices", tuple(verts.shape), "library", os.path.basename(smplk._lib.LIB_PATH))


if __name__ == "__main__":
    main()
