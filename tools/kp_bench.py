"""Keypoint-fitting closure (the reference's fit_single_frame closure: SMPL-H forward, camera
projection + GMoF data term + priors, backward) timed eager and as a replayed CUDA graph.
Usage: [SMPLK_SPARSE_PICKS=0] [RETURN_VERTS=0] python tools/kp_bench.py [B ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import synthetic
from smplk.body_models import SMPLH
from smplk.fitting import GraphedClosure, PerspectiveCamera, SMPLifyLoss

dev = "cuda:0"
m = synthetic.make_model("smplh", seed=0)
for B in [int(x) for x in sys.argv[1:]] or [1, 64, 1024]:
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)
    if os.environ.get("SMPLK_SPARSE_PICKS") == "0":      # dense vertex backward for joints-only losses (handle option)
        mod._dm_kwargs["options"] = {"sparse_picks": 0}
    cam = PerspectiveCamera(translation=np.tile([[0.0, 0.0, 10.0]], (B, 1)), batch_size=B,
                            center=np.tile([[512.0, 512.0]], (B, 1)))
    cam.translation.requires_grad_(False)
    nj = mod(return_verts=False).joints.shape[1]
    gt2d = torch.rand(B, nj, 2, device=dev) * 1024
    conf = torch.ones(B, nj, device=dev)
    loss_fn = SMPLifyLoss(rho=100.0, data_weight=1.0, shape_weight=0.5, hand_prior_weight=0.1)
    rv = os.environ.get("RETURN_VERTS", "1") != "0"
    fn = lambda: loss_fn(mod(return_verts=rv, return_full_pose=True), cam, gt2d, conf, joint_weights=conf)

    def eager():
        mod.zero_grad()
        fn().backward()

    def timeit(f, n=50):
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_e = timeit(eager)
    g = GraphedClosure(fn, mod.parameters())
    t_g = timeit(g)
    print("keypoint closure B=%d sparse_picks=%s return_verts=%s: eager %.4f ms, graph %.4f ms" % (
        B, os.environ.get("SMPLK_SPARSE_PICKS", "1"), rv, t_e, t_g))
