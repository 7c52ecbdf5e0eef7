"""Stage-by-stage GPU bring-up check against the oracle (run each mode in its own process so a
faulting kernel cannot poison the others):  python tools/gpu_check.py {simt|tc|bwd|lbs|twins}"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk  # noqa: E402
from smplk import _lib, synthetic  # noqa: E402
from smplk.body_models import body_model_apply  # noqa: E402
from oracle import smpl_oracle as O  # noqa: E402


def intermediates(dm, B, flags, ws):
    lay = dm.workspace_layout(B, flags)
    Npad = (3 * dm.V + 255) // 256 * 256
    A = ws[lay["A"]:lay["A"] + B * dm.J * 48].view(torch.float32).view(B, dm.J, 3, 4)
    vp = ws[lay["v_posed"]:lay["v_posed"] + B * Npad * 4].view(torch.float32).view(B, Npad)[:, :3 * dm.V]
    return A, vp.reshape(B, dm.V, 3)


def run_forward(dm, betas, pose, transl, flags):
    import ctypes
    B = pose.shape[0]
    dev = pose.device
    verts = torch.empty(B, dm.V, 3, device=dev)
    joints = torch.empty(B, dm.J + dm.E, 3, device=dev)
    wsb = dm.workspace_bytes(B, flags)
    ws = torch.zeros(wsb, device=dev, dtype=torch.uint8)
    a = _lib.ForwardArgs()
    a.batch, a.flags = B, flags
    a.betas, a.betas_batch = ctypes.c_void_p(betas.data_ptr()), betas.shape[0]
    a.pose = ctypes.c_void_p(pose.data_ptr())
    a.transl = ctypes.c_void_p(transl.data_ptr())
    a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), wsb
    a.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    dm.forward(a)
    torch.cuda.synchronize()
    return verts, joints, ws


def check_forward(kind, B, flags, label, dense=False):
    model = synthetic.make_model(kind, seed=3, dense_weights=dense)
    dm = smplk.DeviceModel(model, device=0, extra_vertex_ids=model["extra_vertex_ids"])
    print(label, "ell_k", dm.info.max_weights_per_vertex, "tcgen05", dm.info.has_tcgen05_path, flush=True)
    betas, pose, transl = synthetic.make_inputs(model, B, seed=5)
    if B >= 4:
        pose[1] = 0.0
        pose[2] *= 1e-6
        pose[3] *= 8.0
    om = O.TorchOracleModel(model, dtype=torch.float64)
    ref = om.forward_full_pose(torch.tensor(betas, dtype=torch.float64), torch.tensor(pose, dtype=torch.float64),
                               torch.tensor(transl, dtype=torch.float64))
    tb, tp, tt = (torch.tensor(x, device="cuda") for x in (betas, pose, transl))
    verts, joints, ws = run_forward(dm, tb, tp, tt, flags | _lib.FLAG_SAVE_FOR_BACKWARD)
    A, vp = intermediates(dm, B, flags | _lib.FLAG_SAVE_FOR_BACKWARD, ws)
    eA = (A.double().cpu() - ref.A[:, :, :3, :]).abs().max().item()
    evp = (vp.double().cpu() - ref.v_posed).abs().max().item()
    ev = (verts.double().cpu() - ref.vertices).abs().max().item()
    ej = (joints[:, :dm.J].double().cpu() - ref.joints).abs().max().item()
    print("%s B=%d: max|dA|=%.3e max|dv_posed|=%.3e max|dverts|=%.3e max|djoints|=%.3e  (|v|max %.2f)"
          % (label, B, eA, evp, ev, ej, ref.vertices.abs().max().item()), flush=True)
    if evp > 1e-5:
        d = (vp.double().cpu() - ref.v_posed).abs()
        bad = (d > 1e-5).nonzero()
        print("  v_posed mismatches:", bad.shape[0], "first:", bad[:8].tolist(), flush=True)
        print("  got", vp[0, :3].tolist(), "ref", ref.v_posed[0, :3].tolist())
    return max(eA, evp, ev, ej)


def main():
    mode = sys.argv[1]
    torch.cuda.init()
    print(torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    if mode == "simt":
        worst = max(worst, check_forward("smplh", 5, _lib.FLAG_BLEND_SIMT, "simt/smplh"))
        worst = max(worst, check_forward("smpl", 37, _lib.FLAG_BLEND_SIMT, "simt/smpl"))
        worst = max(worst, check_forward("smplh", 9, _lib.FLAG_BLEND_SIMT, "simt/smplh-denseW", dense=True))
    elif mode == "tc":
        worst = max(worst, check_forward("smplh", 128, _lib.FLAG_BLEND_TF32, "tcgen05-tf32/smplh"))
        worst = max(worst, check_forward("smplh", 2500, _lib.FLAG_BLEND_TF32, "tcgen05-tf32/smplh"))
        worst = max(worst, check_forward("smplh", 128, _lib.FLAG_BLEND_TCGEN05, "tcgen05/smplh"))
        worst = max(worst, check_forward("smplh", 300, _lib.FLAG_BLEND_TCGEN05, "tcgen05/smplh"))
        worst = max(worst, check_forward("smpl", 1, _lib.FLAG_BLEND_TCGEN05, "tcgen05/smpl"))
        worst = max(worst, check_forward("smplh", 2500, _lib.FLAG_BLEND_TCGEN05, "tcgen05/smplh"))
    elif mode == "bwd":
        for kind, B, pca in (("smplh", 6, False), ("smpl", 130, False), ("smplh", 40, True)):
            model = synthetic.make_model(kind, seed=4)
            om = O.TorchOracleModel(model, dtype=torch.float64, num_pca_comps=12)
            dm = smplk.DeviceModel(model, device=0, num_pca_comps=12 if pca else 0,
                                   extra_vertex_ids=model["extra_vertex_ids"])
            betas, pose, transl = synthetic.make_inputs(model, B, seed=6)
            b2, p2, t2 = synthetic.make_inputs(model, B, seed=7)
            tgt = om.forward_full_pose(*(torch.tensor(x, dtype=torch.float64) for x in (b2, p2, t2))).vertices
            loss, gb, gp, gt = O.torch_vertex_l2_grads(
                om, *(torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)), tgt)
            tb, tp, tt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (betas, pose, transl))
            v, j, _, _ = body_model_apply(dm, tb, tp, transl=tt)
            l = ((v - tgt.float().cuda()) ** 2).sum()
            l.backward()
            torch.cuda.synchronize()
            for name, g, r in (("betas", tb.grad, gb), ("pose", tp.grad, gp), ("transl", tt.grad, gt)):
                err = (g.double().cpu() - r).abs().max().item()
                rel = err / r.abs().max().item()
                print("bwd %s B=%d d_%s: max abs err %.3e rel %.3e" % (kind, B, name, err, rel), flush=True)
                worst = max(worst, rel * 1e-1)
    elif mode == "lbs":
        rig = synthetic.make_rigged_mesh(5001, seed=2)
        rm = smplk.RecoverModel(rig)
        rng = np.random.default_rng(0)
        poses = rng.standard_normal((20, 72)) * 0.4
        trans = rng.standard_normal((20, 3))
        out = rm.replay(poses, trans)
        for i in (0, 7, 19):
            ref = O.np_lbs_only(rig, poses[i], trans[i])["verts"]
            e = np.abs(out[i] - ref).max()
            worst = max(worst, e)
            print("lbs frame", i, "err %.3e" % e, flush=True)
        v = rm.set_params(pose=poses[3].reshape(24, 3).copy(), trans=trans[3])
        print("set_params err %.3e" % np.abs(v - O.np_lbs_only(rig, poses[3], trans[3])["verts"]).max())
    print("WORST", worst, flush=True)
    sys.exit(0 if worst < 1e-5 else 1)


if __name__ == "__main__":
    main()
