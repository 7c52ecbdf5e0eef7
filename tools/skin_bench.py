"""Streaming skinning kernels in isolation (CUDA events inside the library):
  * two-kernel forward, batch 4096 (SAVE_FOR_BACKWARD): blend GEMM + skinning kernel -> GB/s of 167,856 B/body
  * rigged-mesh replay (LBS only, 24 joints) at Nv in {6890, 50000, 200000}: vertex-write GB/s
  * fitting step, batch 1024: skin_fit_l2 / dA / ...
Usage: [SMPLK_LIB=variant.so] [SMPLK_SKIN_G8=0] python tools/skin_bench.py"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic
from smplk.body_models import fit_vertex_l2

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(dev)
tag = "%s G8=%s" % (os.path.basename(_lib.LIB_PATH), os.environ.get("SMPLK_SKIN_G8", "-"))


def fwd(dm, betas, pose, transl, verts, ws, flags):
    a = _lib.ForwardArgs()
    a.batch, a.flags = pose.shape[0], flags
    a.betas, a.betas_batch = (ctypes.c_void_p(betas.data_ptr()) if betas is not None else None), (betas.shape[0] if betas is not None else 1)
    a.pose, a.transl = ctypes.c_void_p(pose.data_ptr()), ctypes.c_void_p(transl.data_ptr())
    a.verts = ctypes.c_void_p(verts.data_ptr())
    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    a.stream = ctypes.c_void_p(stream.cuda_stream)
    dm.forward(a)


def prof(dm, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dm.profile_enable(True)
    dm.profile_read()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    p = dm.profile_read()
    dm.profile_enable(False)
    return {k: v[0] / v[1] for k, v in p.items() if v[1]}


model = synthetic.make_model("smplh", seed=0)
dm = smplk.DeviceModel(model, device=0)
B = 4096
b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
verts = torch.empty(B, dm.V, 3, device=dev)
fl = _lib.FLAG_SAVE_FOR_BACKWARD
ws = torch.empty(dm.workspace_bytes(B, fl), device=dev, dtype=torch.uint8)
k = prof(dm, lambda: fwd(dm, b, p, t, verts, ws, fl))
print("%s | two-kernel fwd B=4096: skin %.4f ms = %.0f GB/s (%.3f of 6548.8) | %s" % (
    tag, k["skin"], 167856 * B / k["skin"] / 1e6, 167856 * B / k["skin"] / 1e6 / 6548.8,
    " ".join("%s=%.4f" % kv for kv in k.items())), flush=True)
del verts, ws
for nv in (6890, 50000, 200000):
    rdm = smplk.DeviceModel(synthetic.make_rigged_mesh(nv, seed=13), device=0, lbs_only=True)
    ck = min(8192, max(256, int(8e9 // (nv * 12))))
    pose = torch.randn(ck, 72, device=dev) * 0.3
    tr = torch.randn(ck, 3, device=dev)
    vr = torch.empty(ck, nv, 3, device=dev)
    wsr = torch.empty(rdm.workspace_bytes(ck, 0), device=dev, dtype=torch.uint8)
    k = prof(rdm, lambda: fwd(rdm, None, pose, tr, vr, wsr, 0), n=10)
    gbs = ck * nv * 12 / k["skin"] / 1e6
    print("%s | LBS-only nv=%d frames=%d: skin %.4f ms = %.0f GB/s written (%.3f of 6548.8) pose %.4f" % (
        tag, nv, ck, k["skin"], gbs, gbs / 6548.8, k["pose_fwd"]), flush=True)
    del vr, wsr, rdm
    torch.cuda.empty_cache()
Bf = 1024
bb, pb, tb = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, Bf, seed=99))
target = torch.randn(Bf, dm.V, 3, device=dev)


def fit():
    for x in (bb, pb, tb):
        x.grad = None
    fit_vertex_l2(dm, bb, pb, target, transl=tb, reduce="sum").backward()


k = prof(dm, fit)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    fit()
e1.record()
torch.cuda.synchronize()
print("%s | fit step B=1024: %.4f ms eager | %s" % (tag, e0.elapsed_time(e1) / 30, " ".join("%s=%.4f" % kv for kv in k.items())), flush=True)
