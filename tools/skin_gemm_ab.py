"""Skinning pass of the two-kernel forward (SAVE_FOR_BACKWARD: blend GEMM -> v_posed -> skinning): transform blend on the
tensor cores (handle option skin_gemm, kSkin instance of lbs_replay_gemm_kernel) against the streaming skinning kernel.
Max |difference| of the outputs and CUDA-event timings.  Usage: python tools/skin_gemm_ab.py [B ...]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(dev)


def fwd(dm, betas, pose, transl, verts, ws, flags):
    a = _lib.ForwardArgs()
    a.batch, a.flags = pose.shape[0], flags
    a.betas, a.betas_batch = ctypes.c_void_p(betas.data_ptr()), betas.shape[0]
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


sizes = [int(x) for x in sys.argv[1:]] or [4096, 1024, 333]
for kind, nv in (("smplh", 6890), ("smpl", 6890), ("smplh", 5003)):
    model = synthetic.make_model(kind, seed=0, num_verts=nv)
    dms = {k: smplk.DeviceModel(model, device=0, options={"skin_gemm": k}) for k in (1, 0)}
    for B in sizes:
        b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
        fl = _lib.FLAG_SAVE_FOR_BACKWARD
        outs, times = {}, {}
        for k, dm in dms.items():
            verts = torch.full((B, nv, 3), float("nan"), device=dev)
            ws = torch.empty(dm.workspace_bytes(B, fl), device=dev, dtype=torch.uint8)
            times[k] = prof(dm, lambda: fwd(dm, b, p, t, verts, ws, fl))
            outs[k] = verts
        byt = (dms[1].Npad if hasattr(dms[1], "Npad") else 0)
        gb = B * nv * 24 / 1e6
        t1, t0 = times[1], times[0]
        print("%s V=%d B=%d: gemm skin %.4f (+operand %.4f) ms = %.0f GB/s (%.3f) | streaming skin %.4f ms = %.0f GB/s (%.3f) | max diff %.3g nan %d" % (
            kind, nv, B, t1["skin"], t1.get("transpose", 0.0), gb / t1["skin"], gb / t1["skin"] / 6548.8, t0["skin"], gb / t0["skin"],
            gb / t0["skin"] / 6548.8, float((outs[0] - outs[1]).abs().max()), int(torch.isnan(outs[1]).sum())), flush=True)
        del outs, verts, ws
        torch.cuda.empty_cache()
