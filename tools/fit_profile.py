"""Per-kernel CUDA-event times of the one-node fitting step (BASELINE config 3) from the library's own profile
counters: python tools/fit_profile.py [B]   (SMPLK_LIB=<variant .so> selects another build)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import synthetic
from smplk.body_models import fit_vertex_l2

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
model = synthetic.make_model("smplh", seed=0)
dm = smplk.DeviceModel(model, device=0)
b, p, t = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, B, seed=1))
tgt = torch.randn(B, dm.V, 3, device=dev) * 0.01
with torch.no_grad():
    tgt += smplk.body_models.body_model_apply(dm, b + 0.05, p + 0.01, transl=t)[0]


def step():
    for x in (b, p, t):
        x.grad = None
    fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
dm.profile_enable(True)
dm.profile_read()
for _ in range(30):
    step()
torch.cuda.synchronize()
pr = dm.profile_read()
print("lib", os.environ.get("SMPLK_LIB", "libsmplk.so"), "B", B)
for k, (ms, n) in sorted(pr.items()):
    if n:
        print("  %-22s %.4f ms x %d" % (k, ms / n, n))
