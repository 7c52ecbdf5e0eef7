"""Rigged-mesh replay of a whole sequence (BASELINE config 5, LBS only) with programmatic dependent launch on / off:
pose kernel -> operand pass -> replay GEMM per 8192-frame chunk, back-to-back smplk_forward calls, CUDA-event time,
outputs compared bitwise.  Usage: python tools/replay_pdl_ab.py [nv ...]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(dev)
for nv in [int(x) for x in sys.argv[1:]] or [6890, 50000]:
    mesh = synthetic.make_rigged_mesh(nv, seed=13)
    F = min(32768, max(256, int(12e9 // (nv * 12))))           # several 8192-frame chunks per call, <= 12 GB of output
    pose = torch.randn(F, 72, device=dev) * 0.4
    tr = torch.randn(F, 3, device=dev)
    outs = {}
    for val in (1, 0, 1, 0):
        dm = smplk.DeviceModel(mesh, device=0, lbs_only=True, options={"pdl": val})
        verts = torch.empty(F, nv, 3, device=dev)
        ws = torch.empty(dm.workspace_bytes(F, 0), device=dev, dtype=torch.uint8)
        a = _lib.ForwardArgs()
        a.batch, a.flags = F, 0
        a.betas, a.betas_batch = None, 1
        a.pose, a.transl = ctypes.c_void_p(pose.data_ptr()), ctypes.c_void_p(tr.data_ptr())
        a.verts = ctypes.c_void_p(verts.data_ptr())
        a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
        a.stream = ctypes.c_void_p(stream.cuda_stream)
        for _ in range(2):
            dm.forward(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            dm.forward(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        outs[val] = verts
        print("pdl=%d nv=%d frames=%d  %.4f ms  %.2f M frames/s  %.0f GB/s of vertex writes" % (
            val, nv, F, ms, F / ms / 1e3, F * nv * 12 / ms / 1e6), flush=True)
        del ws
    print("nv=%d outputs bitwise equal across pdl: %s" % (nv, torch.equal(outs[0], outs[1])), flush=True)
    del outs, verts
