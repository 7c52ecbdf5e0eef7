"""Rigged-mesh replay (LBS only): tensor-core GEMM kernel against the streaming skinning kernel (handle option
replay_gemm): max |difference| of the two outputs, error of both against a float64 evaluation of the same
transforms on sampled frames, and CUDA-event timings.  Usage: python tools/replay_ab.py [nv ...]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(dev)


def fwd(dm, pose, transl, verts, ws):
    a = _lib.ForwardArgs()
    a.batch, a.flags = pose.shape[0], 0
    a.betas, a.betas_batch = None, 1
    a.pose = ctypes.c_void_p(pose.data_ptr())
    a.transl = ctypes.c_void_p(transl.data_ptr()) if transl is not None else None
    a.verts = ctypes.c_void_p(verts.data_ptr())
    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    a.stream = ctypes.c_void_p(stream.cuda_stream)
    dm.forward(a)


def prof(dm, fn, n=10):
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


sizes = [int(x) for x in sys.argv[1:]] or [6890, 50000, 200000, 5003]
for nv in sizes:
    mesh = synthetic.make_rigged_mesh(nv, seed=13)
    dms = {k: smplk.DeviceModel(mesh, device=0, lbs_only=True, options={"replay_gemm": k}) for k in (1, 0)}
    for ck in (min(8192, max(256, int(8e9 // (nv * 12)))), 333):
        pose = torch.randn(ck, 72, device=dev) * 0.4
        tr = torch.randn(ck, 3, device=dev) * 2.0
        outs, times = {}, {}
        for k, dm in dms.items():
            vr = torch.full((ck, nv, 3), float("nan"), device=dev)
            wsr = torch.empty(dm.workspace_bytes(ck, 0), device=dev, dtype=torch.uint8)
            t = prof(dm, lambda: fwd(dm, pose, tr, vr, wsr))
            times[k] = t["skin"]
            outs[k] = vr
        diff = float((outs[0] - outs[1]).abs().max())
        nan = int(torch.isnan(outs[1]).sum())
        gb = ck * nv * 12 / 1e6
        print("nv=%d frames=%d: gemm %.4f ms = %.0f GB/s (%.3f) | streaming %.4f ms = %.0f GB/s (%.3f) | max |gemm - streaming| %.3g, nan %d" % (
            nv, ck, times[1], gb / times[1], gb / times[1] / 6548.8, times[0], gb / times[0], gb / times[0] / 6548.8, diff, nan), flush=True)
        del outs, vr, wsr
        torch.cuda.empty_cache()
