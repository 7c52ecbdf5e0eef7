"""GPU parity at the BASELINE configurations' own sizes:

  * the inference call of the drop-in module under torch.no_grad() (lib/Gen_SMPLH/fitting.py:82) runs
    the fused blend+skinning kernel -- asserted through the library's per-kernel launch counters;
  * config 3: the one-node fitting step `fit_vertex_l2` (skin_fit_l2_kernel + bf16 two-term backward
    GEMM) against the float64 autograd oracle at B = 1024 and B = 129, with upstream gradients and
    residuals spanning many decades across the batch; bound 1e-4 relative PER BODY;
  * config 5: a 100,000-frame SMPL-H sequence tiled from a real AMASS clip (156-D with hands,
    trans - trans[0], one betas row), spot-checked on both sides of every 8192-frame chunk boundary.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import smplk
from smplk import _lib, clips, synthetic
from smplk.body_models import SMPLH, body_model_apply, fit_vertex_l2
from oracle import smpl_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
GRAD_RTOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def smplh_model():
    return synthetic.make_model("smplh", seed=21)


def _t(x, dev, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device=dev, requires_grad=grad)


def test_parameter_driven_module_call_under_no_grad_runs_the_fused_kernel(dev, smplh_model):
    """nn.Parameters keep requires_grad=True under torch.no_grad(); the inference forward must still
    take pose_forward_block + blend_skin_fused (2 kernels + the vertex-pick gather) and not the
    SAVE_FOR_BACKWARD two-kernel path with its whole-batch v_posed workspace."""
    m = smplh_model
    B = 300
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B, create_transl=True).to(dev)
    rng = np.random.default_rng(4)
    vals = dict(betas=rng.standard_normal((B, 16)), global_orient=rng.standard_normal((B, 3)) * 0.3,
                body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)),
                right_hand_pose=rng.standard_normal((B, 12)), transl=rng.standard_normal((B, 3)))
    mod.reset_params(**vals)
    assert all(p.requires_grad for p in mod.parameters())
    dm = mod.device_model(dev)
    dm.profile_enable(True)
    dm.profile_read()
    with torch.no_grad():
        out = mod(return_verts=True)
    torch.cuda.synchronize()
    prof = dm.profile_read()
    assert prof["blend_skin_fused"][1] == 1 and prof["pose_fwd"][1] == 1
    assert prof["skin"][1] == 0 and prof["blend_tcgen05"][1] == 0
    # with gradients enabled the same call keeps v_posed for the backward: two-kernel path
    out_g = mod(return_verts=True)
    torch.cuda.synchronize()
    prof = dm.profile_read()
    dm.profile_enable(False)
    assert prof["blend_skin_fused"][1] == 0 and prof["blend_tcgen05"][1] == 1 and prof["skin"][1] == 1
    assert out_g.vertices.requires_grad and not out.vertices.requires_grad
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    t = {k: torch.tensor(v) for k, v in vals.items()}
    ref = om.forward(t["betas"], t["global_orient"], t["body_pose"], t["left_hand_pose"], t["right_hand_pose"],
                     transl=t["transl"])
    for o in (out, out_g):
        assert float((o.vertices.detach().double().cpu() - ref.vertices).abs().max()) <= TOL
        assert float((o.joints.detach().double().cpu() - ref.joints).abs().max()) <= TOL
    # upstream returns the PCA-projected 45-D hand poses
    assert out.left_hand_pose.shape == (B, 45)
    assert float((out.left_hand_pose.double().cpu() - ref.left_hand_pose).abs().max()) <= 1e-5
    assert float((out.right_hand_pose.double().cpu() - ref.right_hand_pose).abs().max()) <= 1e-5


def _oracle_fit_grads(m, betas, pose, transl, target, d_loss, chunk=128):
    """float64 autograd of sum_b d_loss[b] * ||V_b - V*_b||^2, in chunks of bodies (they are independent)."""
    om = O.TorchOracleModel(m, dtype=torch.float64)
    B = pose.shape[0]
    loss = np.zeros(B)
    gb, gp, gt = np.zeros_like(betas, np.float64), np.zeros_like(pose, np.float64), np.zeros_like(transl, np.float64)
    for c0 in range(0, B, chunk):
        s = slice(c0, min(B, c0 + chunk))
        tb, tp, tt = (torch.tensor(np.asarray(x[s], np.float64), requires_grad=True) for x in (betas, pose, transl))
        out = om.forward_full_pose(tb, tp, tt)
        per = ((out.vertices - torch.tensor(np.asarray(target[s], np.float64))) ** 2).sum(dim=(1, 2))
        (per * torch.tensor(d_loss[s])).sum().backward()
        loss[s] = per.detach().numpy()
        gb[s], gp[s], gt[s] = tb.grad.numpy(), tp.grad.numpy(), tt.grad.numpy()
    return loss, gb, gp, gt


@pytest.mark.parametrize("B", [1024, 129])
def test_fit_vertex_l2_matches_float64_oracle_at_config3_size(dev, smplh_model, B):
    """BASELINE config 3 at its own batch (1024) and at a ragged one: loss and gradients of the one-node
    fitting step against float64 autograd.  Residual magnitudes span 5 decades across the bodies
    (0.05 .. 1e4 m) and the upstream d_loss 16 decades; one body has d_loss = 0.  The d_v_posed rows of
    this path are bf16 two-term splits (16 mantissa bits): the bound is 1e-4 relative per body."""
    m = smplh_model
    dm = smplk.DeviceModel(m, device=0)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=300 + B)
    rng = np.random.default_rng(B)
    # V* = the forward of perturbed parameters: a coherent residual of 0.05 .. 1 m,
    # plus per-body noise of 1e-3 .. 1e4 m.  (Pure noise residuals of ~1e-2 m are NOT a fair fp32 test:
    # their true parameter gradient cancels like sqrt(V) while the 2e-7 m fp32 error of a limb's
    # transform is coherent over its vertices -- 4e-4 relative on such a body for ANY fp32 path.)
    om64 = O.TorchOracleModel(m, dtype=torch.float64)
    s_b = 10.0 ** rng.uniform(-0.5, 0.0, size=(B, 1))
    pb = betas + rng.standard_normal(betas.shape) * s_b
    pp = pose + rng.standard_normal(pose.shape) * 0.3 * s_b
    pt = transl + rng.standard_normal(transl.shape) * 10.0 ** rng.uniform(-1, 0.5, size=(B, 1))
    with torch.no_grad():
        parts = [om64.forward_full_pose(*[torch.tensor(np.asarray(x[c:c + 256], np.float64)) for x in (pb, pp, pt)]).vertices
                 for c in range(0, B, 256)]
    v_tgt = torch.cat(parts).numpy()
    res_scale = 10.0 ** rng.uniform(-3, 4, size=B)
    target = (v_tgt + rng.standard_normal(v_tgt.shape) * res_scale[:, None, None]).astype(np.float32)
    d_loss = 10.0 ** rng.uniform(-8, 8, size=B)
    d_loss[5] = 0.0
    tb, tp, tt = (_t(x, dev, True) for x in (betas, pose, transl))
    dm.profile_enable(True)
    dm.profile_read()
    loss = fit_vertex_l2(dm, tb, tp, _t(target, dev), transl=tt)
    (loss * _t(d_loss, dev)).sum().backward()
    torch.cuda.synchronize()
    prof = dm.profile_read()
    dm.profile_enable(False)
    assert prof["skin"][1] == 1 and prof["skin_bwd"][1] == 0 and prof["blend_bwd"][1] == 1   # the fused fitting kernel ran
    want_loss, gb, gp, gt = _oracle_fit_grads(m, betas, pose, transl, target, d_loss)
    rel_loss = np.abs(loss.detach().double().cpu().numpy() - want_loss) / want_loss
    assert rel_loss.max() <= 1e-5, (int(rel_loss.argmax()), float(rel_loss.max()))
    for got, want, name in ((tb.grad, gb, "betas"), (tp.grad, gp, "pose"), (tt.grad, gt, "transl")):
        g = got.double().cpu().numpy()
        assert np.isfinite(g).all(), name
        den = np.abs(want).max(axis=1)
        err = np.abs(g - want).max(axis=1)
        rel = np.where(den > 0, err / np.where(den > 0, den, 1.0), np.abs(g).max(axis=1))
        assert rel.max() <= GRAD_RTOL, (name, int(rel.argmax()), float(rel.max()))


def test_config5_full_100k_frame_amass_sequence(dev, golden_dir):
    """BASELINE config 5 at its full length: the real AMASS clip data/amsass/09_05_poses.npz (143 frames,
    156-D poses with both hands, lib/model2video.py:527-531) tiled to 100,000 frames with its
    trans - trans[0], the clip's single betas (16,) row broadcast.  8.3 GB of vertices stay on
    the device; 40 frames -- both sides of every 8192-frame chunk boundary plus the ends and the tile
    seams -- are checked against the float64 oracle."""
    c = clips.read_amsass(os.path.join(golden_dir, "amass_clip_09_05.npz"), full=True)
    n0 = c.poses.shape[0]
    assert c.poses.shape == (143, 156) and c.betas.shape == (16,) and np.all(c.trans[0] == 0)
    N = 100000
    reps = N // n0 + 1
    poses = np.tile(c.poses, (reps, 1))[:N].astype(np.float32)
    trans = np.tile(c.trans, (reps, 1))[:N].astype(np.float32)      # the looped clip restarts at its origin
    m = synthetic.make_model("smplh", seed=0)
    dm = smplk.DeviceModel(m, device=0)
    betas = _t(c.betas.reshape(1, 16), dev)
    dm.profile_enable(True)
    dm.profile_read()
    v, j, _, _ = body_model_apply(dm, betas, _t(poses, dev), transl=_t(trans, dev))
    torch.cuda.synchronize()
    prof = dm.profile_read()
    dm.profile_enable(False)
    chunks = (N + 8191) // 8192
    assert prof["blend_skin_fused"][1] == chunks and prof["pose_fwd"][1] == chunks
    assert v.shape == (N, 6890, 3)
    idx = sorted(set([0, 1, n0 - 1, n0, N - 2, N - 1] + [k * 8192 + d for k in range(1, chunks) for d in (-1, 0)] +
                     [50000, 77777, 12345, 99000 // n0 * n0, 99000 // n0 * n0 - 1]))
    assert len(idx) >= 32
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        torch.tensor(np.repeat(c.betas.reshape(1, 16), len(idx), 0), dtype=torch.float64),
        torch.tensor(poses[idx], dtype=torch.float64), torch.tensor(trans[idx], dtype=torch.float64))
    ii = torch.as_tensor(idx, device=dev)
    ev = float((v[ii].double().cpu() - ref.vertices).abs().max())
    ej = float((j[ii].double().cpu() - ref.joints).abs().max())
    assert ev <= TOL and ej <= TOL, (ev, ej)
    # periodicity over the whole sequence: frame i + 143 k repeats frame i, wherever it lands in a chunk
    fin = torch.isfinite(v.view(N, -1)).all(dim=1)
    assert bool(fin.all())
    last = (N // n0 - 1) * n0
    for off in (n0, 57 * n0, last):
        assert float((v[off:off + n0] - v[:n0]).abs().max()) <= 1e-6


def test_float64_interface_and_wrapper_attributes(dev, smplh_model):
    """lib/gen_smplh.py:66-67 may ask for float64: parameters, buffers, outputs and gradients carry it, the
    kernels compute in float32 (error bar unchanged: 1e-5 m / 1e-4 relative).  models/smplh.py:18-20 attributes:
    verts_numpy (template until the first forward), weigths, seg_index."""
    m = smplh_model
    B = 130
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=B, create_transl=True, dtype=torch.float64).to(dev)
    assert mod.betas.dtype == torch.float64 and mod.v_template.dtype == torch.float64
    assert np.array_equal(mod.verts_numpy, np.asarray(m["v_template"])) and mod.weigths.shape == (6890, 52)
    assert mod.seg_index == {}
    rng = np.random.default_rng(8)
    vals = dict(betas=rng.standard_normal((B, 16)), global_orient=rng.standard_normal((B, 3)) * 0.3,
                body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)),
                right_hand_pose=rng.standard_normal((B, 12)), transl=rng.standard_normal((B, 3)))
    mod.reset_params(**vals)
    out = mod(return_verts=True, return_full_pose=True)
    assert out.vertices.dtype == torch.float64 and out.joints.dtype == torch.float64 and out.full_pose.dtype == torch.float64
    ((out.vertices ** 2).sum() + (out.joints ** 2).sum()).backward()
    assert mod.body_pose.grad.dtype == torch.float64
    assert mod.verts_numpy.shape == (6890, 3) and np.allclose(mod.verts_numpy, out.vertices[0].detach().cpu().numpy())
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in vals.items()}
    ref = om.forward(t["betas"], t["global_orient"], t["body_pose"], t["left_hand_pose"], t["right_hand_pose"],
                     transl=t["transl"])
    ((ref.vertices ** 2).sum() + (ref.joints ** 2).sum()).backward()
    assert float((out.vertices.detach().cpu() - ref.vertices.detach()).abs().max()) <= TOL
    for name in ("betas", "body_pose", "left_hand_pose", "transl"):
        g, r = getattr(mod, name).grad.cpu(), t[name].grad
        assert float((g - r).abs().max() / r.abs().max()) <= GRAD_RTOL, name
    loss = mod.vertex_l2(torch.zeros(B, 6890, 3, dtype=torch.float64, device=dev), reduce="sum")
    assert loss.dtype == torch.float64 and abs(float(loss) / float((ref.vertices.detach() ** 2).sum()) - 1) <= 1e-5


def _err(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _lbs_forward(dm, pose, transl, dev):
    """LBS-only forward through the C ABI (smplk_forward) with device buffers."""
    import ctypes
    n = pose.shape[0]
    verts = torch.full((n, dm.V, 3), float("nan"), device=dev)
    ws = torch.empty(dm.workspace_bytes(n, 0), device=dev, dtype=torch.uint8)
    a = _lib.ForwardArgs()
    a.batch, a.flags = n, 0
    a.betas, a.betas_batch = None, 1
    a.pose = ctypes.c_void_p(pose.data_ptr())
    a.transl = ctypes.c_void_p(transl.data_ptr()) if transl is not None else None
    a.verts = ctypes.c_void_p(verts.data_ptr())
    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    dm.forward(a)
    torch.cuda.synchronize()
    return verts


@pytest.mark.parametrize("nv,frames", [(6890, 333), (50001, 97), (1000, 8192 + 77), (300, 64), (2 * 6890 + 1, 161)])
def test_rigged_mesh_replay_gemm_matches_oracle_and_streaming_kernel(dev, nv, frames):
    """config 5, rigged-mesh variant (lib/model2video.py:55-85): from 64 frames on, an LBS-only handle replays a clip as
    one tensor-core GEMM verts = P . T^T (lbs_replay_gemm.cuh; fp16 two-term operands, three passes).  Checked
    against the float64 oracle on frames at the 80-frame tile edges, the 8192-frame chunk edge and the clip's ends,
    and against the streaming skinning kernel (handle option replay_gemm = 0) on every frame; ragged vertex blocks
    (nv not a multiple of 256), odd vertex counts (4-byte aligned rows), with and without a translation."""
    rig = synthetic.make_rigged_mesh(nv, seed=nv)
    dm_g = smplk.DeviceModel(rig, device=0, lbs_only=True)
    dm_s = smplk.DeviceModel(rig, device=0, lbs_only=True, options={"replay_gemm": 0})
    rng = np.random.default_rng(frames)
    pose = rng.standard_normal((frames, 72)).astype(np.float32) * 0.4
    pose[0] = 0.0
    pose[-1] = pose[-1] / np.abs(pose[-1]).max() * 3.0
    trans = (rng.standard_normal((frames, 3)) * 3.0).astype(np.float32)
    dm_g.profile_enable(True)
    dm_s.profile_enable(True)
    vg = _lbs_forward(dm_g, _t(pose, dev), _t(trans, dev), dev)
    vs = _lbs_forward(dm_s, _t(pose, dev), _t(trans, dev), dev)
    chunks = (frames + 8191) // 8192
    full_chunks = sum(1 for c in range(chunks) if min(8192, frames - 8192 * c) >= 64)
    pg, ps = dm_g.profile_read(), dm_s.profile_read()
    assert pg["transpose"][1] == full_chunks and ps["transpose"][1] == 0, (pg, ps)     # the GEMM's operand pass ran
    assert torch.isfinite(vg).all() and torch.isfinite(vs).all()
    assert _err(vg, vs) <= 4e-6
    idx = sorted(i for i in {0, 1, 79, 80, frames // 2, frames - 2, frames - 1, 8191, 8192, 8193} if i < frames)
    for i in idx:
        ref = O.np_lbs_only(rig, pose[i].astype(np.float64), trans[i].astype(np.float64), ignore_joints=())["verts"]
        assert np.abs(vg[i].double().cpu().numpy() - ref).max() <= TOL, i
    v0 = _lbs_forward(dm_g, _t(pose, dev), None, dev)                # no translation
    assert _err(v0 + _t(trans, dev)[:, None, :], vg) <= 4e-6


def test_rigged_mesh_replay_gemm_holds_accuracy_on_large_coordinates(dev):
    """The GEMM's fp16 two-term operands carry a power-of-two scale chosen from the mesh: a rig in centimetres
    (coordinates ~ 100) keeps fp32-level RELATIVE accuracy."""
    rig = synthetic.make_rigged_mesh(4000, seed=5)
    rig = dict(rig)
    rig["v_template"] = np.asarray(rig["v_template"]) * 100.0
    rig["J"] = np.asarray(rig["J"]) * 100.0
    dm = smplk.DeviceModel(rig, device=0, lbs_only=True)
    rng = np.random.default_rng(0)
    pose = rng.standard_normal((128, 72)).astype(np.float32) * 0.5
    trans = (rng.standard_normal((128, 3)) * 100.0).astype(np.float32)
    v = _lbs_forward(dm, _t(pose, dev), _t(trans, dev), dev)
    for i in (0, 64, 127):
        ref = O.np_lbs_only(rig, pose[i].astype(np.float64), trans[i].astype(np.float64), ignore_joints=())["verts"]
        assert np.abs(v[i].double().cpu().numpy() - ref).max() <= 1e-5 * 100.0 * 3


@pytest.mark.parametrize("V,B", [(6890, 4097), (5002, 259), (84 * 5, 300), (1000, 300), (4001, 300)])
def test_fused_forward_tma_stores_are_bitwise_the_lane_stores(dev, V, B):
    """The fused kernel writes its result as TMA tensor stores when V is even and the buffer 16-byte aligned (even
    bodies one box, odd bodies a box shifted by two carried columns; tools/micro/tma_store_probe.cu shows why);
    handle option fused_tma_out = 0 keeps the per-lane stores.  Same arithmetic, so the outputs are bit-identical;
    odd batch sizes end on a half super-row, V = 420 ends exactly on a tile and, like V = 1000, has 16-byte aligned rows
    for every body (no shift); an odd V (4-byte aligned rows) takes the lane stores in both handles."""
    m = synthetic.make_model("smplh", seed=3, num_verts=V)
    dm_t = smplk.DeviceModel(m, device=0)
    dm_l = smplk.DeviceModel(m, device=0, options={"fused_tma_out": 0})
    betas, pose, transl = synthetic.make_inputs(m, B, seed=B)
    args = (_t(betas, dev), _t(pose, dev))
    dm_t.profile_enable(True)
    vt = body_model_apply(dm_t, *args, transl=_t(transl, dev))[0]
    vl = body_model_apply(dm_l, *args, transl=_t(transl, dev))[0]
    torch.cuda.synchronize()
    assert dm_t.profile_read()["blend_skin_fused"][1] == 1
    assert torch.isfinite(vt).all() and torch.equal(vt, vl)
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas[:64], pose[:64], transl[:64])])
    assert _err(vt[:64], ref.vertices) <= TOL


@pytest.mark.parametrize("kind,V,B", [("smplh", 6890, 333), ("smpl", 5003, 130)])
def test_two_kernel_forward_with_tensor_core_transform_blend(dev, kind, V, B):
    """Handle option skin_gemm = 1: the skinning pass of the two-kernel forward (SAVE_FOR_BACKWARD) blends the
    transforms as a GEMM (kSkin instance of lbs_replay_gemm_kernel) and applies them to v_posed in its epilogue.
    Same result as the streaming skinning kernel (default) and as the float64 oracle; gradients flow as before."""
    m = synthetic.make_model(kind, seed=4, num_verts=V)
    dm_g = smplk.DeviceModel(m, device=0, options={"skin_gemm": 1})
    dm_s = smplk.DeviceModel(m, device=0)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=V + B)
    outs = []
    for dm in (dm_g, dm_s):
        b, p, t = _t(betas, dev, True), _t(pose, dev, True), _t(transl, dev, True)
        dm.profile_enable(True)
        v = body_model_apply(dm, b, p, transl=t)[0]
        (v ** 2).sum().backward()
        torch.cuda.synchronize()
        outs.append((v.detach(), p.grad.clone(), dm.profile_read()))
    assert outs[0][2]["transpose"][1] == 1 and outs[1][2]["transpose"][1] == 0       # the GEMM's operand pass ran
    assert _err(outs[0][0], outs[1][0]) <= 3e-6
    assert float((outs[0][1] - outs[1][1]).abs().max() / outs[1][1].abs().max()) <= 1e-5
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert _err(outs[0][0], ref.vertices) <= TOL


@pytest.mark.parametrize("B", [37, 300, 1024, 4096 + 29])
def test_programmatic_dependent_launch_changes_no_bit(dev, smplh_model, B):
    """Handle option pdl (default 1): the kernels of a call are launched with programmatic stream serialization and
    run their set-up before `griddepcontrol.wait` -- the pose backward even recomputes its forward half there.  Only
    scheduling changes, so the forward, the gradients of back-to-back fitting steps and of a CUDA-graph replay must
    equal the plainly ordered launches (pdl = 0) bit for bit (the loss itself is summed with float atomics: 1e-6);
    with one shared betas row the beta gradient is an atomic sum too (memset moved to the head of the call): 1e-5.
    The batch sizes cover the warp-per-body pose kernel, 8-body blocks, one-wave 28-body blocks and several waves."""
    m = smplh_model
    dms = {k: smplk.DeviceModel(m, device=0, options={"pdl": k}) for k in (1, 0)}
    betas, pose, transl = synthetic.make_inputs(m, B, seed=B)
    res, shared = {}, {}
    for k, dm in dms.items():
        with torch.no_grad():
            v, j = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev))[:2]
        tgt = (v + 0.01).contiguous()

        def step(b, p, t):
            for x in (b, p, t):
                x.grad = None
            loss = fit_vertex_l2(dm, b, p, tgt, transl=t)
            loss.sum().backward()
            return loss.detach().clone()

        b, p, t = _t(betas, dev, True), _t(pose, dev, True), _t(transl, dev, True)
        losses = [step(b, p, t) for _ in range(3)]    # back to back: the next call's pose kernel follows the pose backward
        eager = [x.grad.clone() for x in (b, p, t)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step(b, p, t)
        torch.cuda.current_stream().wait_stream(s)
        for x in (b, p, t):
            x.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()
        g.replay()
        torch.cuda.synchronize()
        for e, x in zip(eager, (b, p, t)):
            assert torch.equal(e, x.grad)
        assert float((losses[0] - losses[-1]).abs().max() / losses[0].abs().max()) <= 1e-6
        res[k] = [v, j] + eager
        b1, p1 = _t(betas[:1], dev, True), _t(pose, dev, True)
        fit_vertex_l2(dm, b1, p1, tgt, transl=_t(transl, dev)).sum().backward()
        shared[k] = (losses[-1], b1.grad.clone(), p1.grad.clone())
    for x, y in zip(res[1], res[0]):
        assert torch.isfinite(x).all() and torch.equal(x, y)
    assert float((shared[1][0] - shared[0][0]).abs().max() / shared[0][0].abs().max()) <= 1e-6
    assert float((shared[1][1] - shared[0][1]).abs().max() / shared[0][1].abs().max()) <= 1e-5
    assert torch.equal(shared[1][2], shared[0][2])
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas[:40], pose[:40], transl[:40])])
    assert _err(res[1][0][:40], ref.vertices) <= TOL


def test_skip_pose_option_repeats_the_previous_call(dev, smplh_model):
    """Measurement aid of bench.py's roofline (handle option skip_pose): no pose kernel is launched, the blend /
    skinning kernel runs again on the workspace rows of the previous call -- whatever pose the call is given."""
    m = smplh_model
    B = 300
    dm = smplk.DeviceModel(m, device=0)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=5)
    b, p, t = _t(betas, dev), _t(pose, dev), _t(transl, dev)
    verts = torch.empty(B, dm.V, 3, device=dev)
    joints = torch.empty(B, dm.J + dm.E, 3, device=dev)
    ws = torch.empty(dm.workspace_bytes(B, 0), device=dev, dtype=torch.uint8)
    stream = torch.cuda.current_stream(dev)

    def call(pose_t):
        a = _lib.ForwardArgs()
        a.batch, a.flags = B, 0
        a.betas, a.betas_batch = ctypes.c_void_p(b.data_ptr()), B
        a.pose, a.transl = ctypes.c_void_p(pose_t.data_ptr()), ctypes.c_void_p(t.data_ptr())
        a.verts, a.joints = ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(joints.data_ptr())
        a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
        a.stream = ctypes.c_void_p(stream.cuda_stream)
        dm.forward(a)
        torch.cuda.synchronize()
        return verts.clone()

    n0 = _lib.launch_count()
    v_ref = call(p)
    per_call = _lib.launch_count() - n0
    dm.set_option("skip_pose", 1)
    verts.zero_()
    n0 = _lib.launch_count()
    v_again = call(torch.zeros_like(p))
    assert _lib.launch_count() - n0 == per_call - 1
    assert torch.equal(v_ref, v_again)
    dm.set_option("skip_pose", 0)
    v_zero = call(torch.zeros_like(p))
    assert not torch.equal(v_ref, v_zero)
