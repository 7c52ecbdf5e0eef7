"""Small end-to-end workload touching every kernel family (written for `compute-sanitizer --tool memcheck`, which is
closed on this pool -- it then serves as a plain crash / launch-error screen):
once at ragged sizes -- fused and two-kernel forward, fitting step (skin_fit_l2, dA_seg, backward GEMM, pose
backward), joints-only sparse paths, rigged-mesh replay, the twins' compute_R_G / do_skinning / inverse, the
chunked host-buffer forward.   Usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk  # noqa: E402
from smplk import synthetic  # noqa: E402
from smplk.body_models import SMPLH, body_model_apply, fit_vertex_l2  # noqa: E402

dev = torch.device("cuda:0")
m = synthetic.make_model("smplh", seed=3)
t = lambda a, g=False: torch.tensor(np.asarray(a, np.float32), device=dev, requires_grad=g)
dm = smplk.DeviceModel(m, device=0, extra_vertex_ids=m["extra_vertex_ids"], regressor_posed=m["J_regressor_extra"])
for B in (131, 300):
    b, p, tr = synthetic.make_inputs(m, B, seed=B)
    v, j, jr, fp = body_model_apply(dm, t(b), t(p), transl=t(tr), want_regressed=True)          # fused forward
    tb, tp, tt = t(b, True), t(p, True), t(tr, True)
    v, j, jr, fp = body_model_apply(dm, tb, tp, transl=tt, want_regressed=True)                 # two-kernel forward
    (v.sum() + j.sum() + jr.sum()).backward()                                                   # dense backward + scatter
    tb.grad = tp.grad = tt.grad = None
    j2 = body_model_apply(dm, tb, tp, transl=tt, want_verts=False)[1]                            # sparse forward / backward
    j2.sum().backward()
dm2 = smplk.DeviceModel(m, device=0)
for B in (1, 130, 257):
    b, p, tr = synthetic.make_inputs(m, B, seed=B)
    tb, tp, tt = t(b, True), t(p, True), t(tr, True)
    fit_vertex_l2(dm2, tb, tp, torch.zeros(B, 6890, 3, device=dev), transl=tt, reduce="sum").backward()
mod = SMPLH(model=m, use_pca=True, num_pca_comps=6, batch_size=5).to(dev)
with torch.no_grad():
    mod(return_verts=True)
mod.vertex_l2(torch.zeros(5, 6890, 3, device=dev)).sum().backward()
rig = synthetic.make_rigged_mesh(3001, seed=9)
rm = smplk.RecoverModel(rig)
rm.replay(np.random.default_rng(0).standard_normal((300, 72)) * 0.3, np.zeros((300, 3)))
G = rm.compute_R_G()
rm.do_skinning(G)
tw = smplk.SMPLHModel(synthetic.make_model("smplh", num_betas=10, seed=7))
tw.set_params(pose=np.random.default_rng(1).standard_normal((52, 3)) * 0.3, beta=np.ones(10) * 0.1, trans=np.ones(3))
G = tw.compute_R_G()
tw.do_skinning(G)
tw.inverse()
tw.gen_J_3d()
tw.forward_batch(np.random.default_rng(2).standard_normal((2500, 156)) * 0.2, np.zeros(10), np.zeros((2500, 3)))   # 3 host chunks
torch.cuda.synchronize()
print("sanitize target ok, launches", smplk._lib.launch_count())
