"""Per-kernel timing of the forward (and optionally backward) at a given batch, using the
library's own CUDA-event hooks.  Usage: [SMPLK_LIB=variant.so] python tools/kbench.py [B] [bwd]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import _lib, synthetic
from smplk.body_models import body_model_apply

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bwd = len(sys.argv) > 2 and sys.argv[2] in ("bwd", "fit")
fit = len(sys.argv) > 2 and sys.argv[2] == "fit"     # body model + vertex-L2 loss as one autograd node
kind = os.environ.get("KIND", "smplh")
dev = torch.device("cuda:0")
model = synthetic.make_model(kind, seed=0)
dm = smplk.DeviceModel(model, device=0)
b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
if bwd:
    b.requires_grad_(True); p.requires_grad_(True); t.requires_grad_(True)
tgt = torch.randn(B, dm.V, 3, device=dev) if bwd else None
def step():
    if fit:
        from smplk.body_models import fit_vertex_l2
        fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()
        return
    v = body_model_apply(dm, b, p, transl=t)[0]
    if bwd:
        ((v - tgt) ** 2).sum().backward()
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / n
dm.profile_enable(True)
for _ in range(n): step()
torch.cuda.synchronize()
pr = dm.profile_read()
print("%s B=%d %s total %.4f ms (%.2f M/s) | " % (os.path.basename(_lib.LIB_PATH), B, "fwd+bwd" if bwd else "fwd", tot, B / tot / 1e3)
      + " ".join("%s=%.4f" % (k, v[0] / v[1]) for k, v in pr.items() if v[1]))
