"""Tuning aid: per-tile timeline of the fused blend+skinning kernel (CTA 0), printed to stderr.

Builds a variant of the library with -DSMPLK_FZ_TIMELINE=1 (clock64 stamps around the MMA issue, the
accumulator waits / releases and the chunk phases of the epilogue warps) and runs one forward.
Usage (GPU box):  python tools/fz_timeline.py [B]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "3d-human-body-reconstruction_b200")
VARIANT = os.path.join(PKG, "libsmplk_timeline.so")
if not os.path.exists(VARIANT) or os.path.getmtime(VARIANT) < os.path.getmtime(os.path.join(PKG, "csrc", "blend_skin_fused.cuh")):
    subprocess.check_call(["nvcc", "-DSMPLK_FZ_TIMELINE=1", "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                           "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", VARIANT,
                           os.path.join(PKG, "csrc", "smplk_api.cu")])
os.environ["SMPLK_LIB"] = VARIANT

import torch  # noqa: E402
import smplk  # noqa: E402
from smplk import synthetic  # noqa: E402
from smplk.body_models import body_model_apply  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = synthetic.make_model("smplh", seed=0)
dm = smplk.DeviceModel(model, device=0)
dev = torch.device("cuda:0")
b, p, t = (torch.tensor(x, device=dev) for x in synthetic.make_inputs(model, B, seed=1))
for _ in range(3):
    body_model_apply(dm, b, p, transl=t)
torch.cuda.synchronize()
os.environ["SMPLK_FZ_DEBUG"] = "1"
body_model_apply(dm, b, p, transl=t)
torch.cuda.synchronize()
