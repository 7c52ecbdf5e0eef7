"""Short, single-purpose workloads to put under ncu (one kernel family per run; see profiles/README.md).

  python tools/ncu_targets.py fused      # config 2 forward, batch 4096: pose_forward_block + blend_skin_fused
  python tools/ncu_targets.py twokernel  # same batch with SAVE_FOR_BACKWARD: blend_tcgen05_2cta + skin_grouped8
  python tools/ncu_targets.py fit        # config 3 fitting step, batch 1024: skin_fit_l2, dA, backward GEMM, pose backward
  python tools/ncu_targets.py lbs [nv]   # config 5 rigged-mesh replay (LBS only): skinning of the shared template
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import smplk  # noqa: E402
from smplk import _lib, synthetic  # noqa: E402
from smplk.body_models import fit_vertex_l2  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fused"
    reps = int(os.environ.get("REPS", "3"))
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev)
    if what == "lbs":
        nv = int(sys.argv[2]) if len(sys.argv) > 2 else 6890
        dm = smplk.DeviceModel(synthetic.make_rigged_mesh(nv, seed=13), device=0, lbs_only=True)
        B = max(256, min(16384, int(8e9 // (nv * 12))))
        pose = torch.randn(B, 72, device=dev) * 0.3
        betas = None
    else:
        model = synthetic.make_model("smplh", seed=0)
        dm = smplk.DeviceModel(model, device=0)
        B = 1024 if what == "fit" else 4096
        b, p, t = synthetic.make_inputs(model, B, seed=1)
        betas, pose = torch.tensor(b, device=dev), torch.tensor(p, device=dev)
    transl = torch.randn(B, 3, device=dev)
    if what == "fit":
        tb, tp, tt = (x.clone().requires_grad_(True) for x in (betas, pose, transl))
        target = torch.randn(B, dm.V, 3, device=dev)
        for _ in range(reps):
            for x in (tb, tp, tt):
                x.grad = None
            fit_vertex_l2(dm, tb, tp, target, transl=tt, reduce="sum").backward()
    else:
        flags = _lib.FLAG_SAVE_FOR_BACKWARD if what == "twokernel" else 0
        verts = torch.empty(B, dm.V, 3, device=dev)
        joints = torch.empty(B, dm.J + dm.E, 3, device=dev)
        ws = torch.empty(dm.workspace_bytes(B, flags), device=dev, dtype=torch.uint8)
        for _ in range(reps):
            a = _lib.ForwardArgs()
            a.batch, a.flags = B, flags
            if betas is not None:
                a.betas, a.betas_batch = ctypes.c_void_p(betas.data_ptr()), B
            else:
                a.betas_batch = 1
            a.pose, a.transl = ctypes.c_void_p(pose.data_ptr()), ctypes.c_void_p(transl.data_ptr())
            a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
            a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
            a.stream = ctypes.c_void_p(stream.cuda_stream)
            dm.forward(a)
    torch.cuda.synchronize(dev)
    print("ok", what, "B =", B, "launches", _lib.launch_count())


if __name__ == "__main__":
    main()
