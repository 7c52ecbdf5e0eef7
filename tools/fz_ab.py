"""Fused forward kernel: TMA tensor stores against per-lane stores (handle option fused_tma_out), bitwise
comparison of the two outputs and CUDA-event timings.  Usage: python tools/fz_ab.py [B ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import synthetic
from smplk.body_models import body_model_apply

dev = torch.device("cuda:0")
sizes = [int(x) for x in sys.argv[1:]] or [4096, 4097, 300, 16384]
for kind, nv in (("smplh", 6890), ("smpl", 6890), ("smplh", 5002)):
    model = synthetic.make_model(kind, seed=0, num_verts=nv)
    dms = {k: smplk.DeviceModel(model, device=0, options={"fused_tma_out": k}) for k in (1, 0)}
    for B in sizes if nv == 6890 and kind == "smplh" else sizes[:3]:
        b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
        outs, times = {}, {}
        for k, dm in dms.items():
            dm.profile_enable(False)
            for _ in range(3):
                v = body_model_apply(dm, b, p, transl=t)[0]
            torch.cuda.synchronize()
            dm.profile_enable(True); dm.profile_read()
            for _ in range(20):
                v = body_model_apply(dm, b, p, transl=t)[0]
            torch.cuda.synchronize()
            pr = dm.profile_read()
            times[k] = pr["blend_skin_fused"][0] / max(pr["blend_skin_fused"][1], 1)
            outs[k] = v.clone()
        same = bool(torch.equal(outs[0], outs[1]))
        print("%s V=%d B=%d  tma %.4f ms  lanes %.4f ms  bitwise equal: %s  maxdiff %.3g" % (
            kind, nv, B, times[1], times[0], same, float((outs[0] - outs[1]).abs().max())), flush=True)
