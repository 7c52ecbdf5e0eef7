"""One-node fitting steps (BASELINE config 3) for profiling: python tools/fit_step.py [B] [steps].
Used under `ncu --metrics gpu__time_duration.sum` for the launch list in profiles/."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import synthetic
from smplk.body_models import fit_vertex_l2

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
model = synthetic.make_model("smplh", seed=0)
dm = smplk.DeviceModel(model, device=0)
b, p, t = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, B, seed=1))
tgt = torch.randn(B, dm.V, 3, device=dev)
for _ in range(steps):
    for x in (b, p, t):
        x.grad = None
    fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()
torch.cuda.synchronize()
print("ok", float(b.grad.abs().max()))
