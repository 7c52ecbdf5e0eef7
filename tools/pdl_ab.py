"""Programmatic dependent launch (handle option `pdl`) A/B on the forward step (BASELINE config 2): back-to-back
smplk_forward calls on one stream, CUDA-event time per step with pdl = 1 and 0, outputs compared bitwise.
Usage: python tools/pdl_ab.py [B ...]        (the fitting step: python tools/fit_ab.py 1024 pdl)"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(dev)
model = synthetic.make_model("smplh", seed=0)
for B in [int(x) for x in sys.argv[1:]] or [4096, 1024, 16384]:
    b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
    outs = {}
    for val in (1, 0, 1, 0):
        dm = smplk.DeviceModel(model, device=0, options={"pdl": val})
        verts = torch.empty(B, dm.V, 3, device=dev)
        joints = torch.empty(B, dm.J + dm.E, 3, device=dev)
        ws = torch.empty(dm.workspace_bytes(B, 0), device=dev, dtype=torch.uint8)
        a = _lib.ForwardArgs()
        a.batch, a.flags = B, 0
        a.betas, a.betas_batch = ctypes.c_void_p(b.data_ptr()), B
        a.pose, a.transl = ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(t.data_ptr())
        a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
        a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
        a.stream = ctypes.c_void_p(stream.cuda_stream)
        for _ in range(5):
            dm.forward(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(40):
            dm.forward(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        outs[val] = (verts.clone(), joints.clone())
        print("pdl=%d B=%d  %.4f ms/step  %.2f M meshes/s" % (val, B, ms, B / ms / 1e3), flush=True)
    print("B=%d outputs bitwise equal across pdl: %s" % (
        B, all(torch.equal(x, y) for x, y in zip(outs[0], outs[1]))), flush=True)
