"""GPU parity tests proper: the CUDA path (through the C ABI / the drop-in modules) against the
oracle on the same seeded inputs, against the committed golden vectors made by the reference's own
code, and size-independent properties at the full BASELINE sizes.

Tolerances (BASELINE.json north_star): vertices / joints max abs error <= 1e-5 m against the fp32
torch oracle; gradients relative error <= 1e-4."""
import ctypes
import os

import numpy as np
import pytest
import torch

import smplk
from smplk import _lib, synthetic
from smplk.body_models import SMPL, SMPLH, body_model_apply
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


@pytest.fixture(scope="module")
def smpl_model():
    return synthetic.make_model("smpl", seed=22)


def _t(x, dev, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device=dev, requires_grad=grad)


def _maxerr(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def test_native_library_is_loaded_and_counts_launches(dev, smpl_model):
    lib = smplk.load()
    assert os.path.basename(lib._name) == "libsmplk.so"
    dm = smplk.DeviceModel(smpl_model, device=0)
    assert dm.info.has_tcgen05_path == 1 and dm.info.max_weights_per_vertex <= 4
    before = _lib.launch_count()
    b, p, t = synthetic.make_inputs(smpl_model, 3)
    body_model_apply(dm, _t(b, dev), _t(p, dev), transl=_t(t, dev))
    assert _lib.launch_count() >= before + 3


@pytest.mark.parametrize("B,flags", [(1, 0), (7, 0), (33, 0), (33, _lib.FLAG_BLEND_SIMT),
                                     (129, 0), (700, 0), (129, _lib.FLAG_BLEND_TF32),
                                     (700, _lib.FLAG_BLEND_TF32), (5, _lib.FLAG_BLEND_TCGEN05)])
def test_smplh_forward_matches_oracle(dev, smplh_model, B, flags):
    m = smplh_model
    dm = smplk.DeviceModel(m, device=0, extra_vertex_ids=m["extra_vertex_ids"],
                           regressor_posed=m["J_regressor_extra"])
    betas, pose, transl = synthetic.make_inputs(m, B, seed=B)
    pose[0] = 0.0                                  # zero pose: R(0) = I
    if B > 2:
        pose[1] *= 1e-6                            # tiny angles
        pose[2] = pose[2] / np.abs(pose[2]).max() * 4.8   # |theta| > pi as in the AMASS clips
    om32 = O.TorchOracleModel(m, dtype=torch.float32)
    om64 = O.TorchOracleModel(m, dtype=torch.float64)
    args64 = [torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)]
    ref32 = om32.forward_full_pose(*[a.float() for a in args64])
    ref64 = om64.forward_full_pose(*args64)
    try:
        v, j, jr, fp = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev),
                                        want_regressed=True, flags=flags)
    except RuntimeError as e:
        if "SMPLK_AB" in str(e):      # the exact-fp32 SIMT kernel ships only in -DSMPLK_AB builds
            pytest.skip("library built without -DSMPLK_AB")
        raise
    assert _maxerr(v, ref32.vertices) <= TOL and _maxerr(v, ref64.vertices) <= TOL
    assert _maxerr(j[:, :52], ref32.joints) <= TOL
    picks = ref64.vertices[:, torch.as_tensor(m["extra_vertex_ids"], dtype=torch.long)]
    assert _maxerr(j[:, 52:], picks) <= TOL
    extra = torch.einsum("bik,ji->bjk", ref64.vertices, torch.tensor(m["J_regressor_extra"]))
    assert _maxerr(jr, extra) <= TOL
    assert _maxerr(fp, torch.tensor(pose)) == 0.0


def test_blend_operand_formats_agree(dev, smplh_model):
    """exact-fp32 SIMT kernel vs tcgen05 fp16 two-term split (default) vs tcgen05 3xTF32."""
    dm = smplk.DeviceModel(smplh_model, device=0)
    betas, pose, transl = synthetic.make_inputs(smplh_model, 200, seed=9)
    run = lambda f: body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev), flags=f)[0]
    f16, tf32 = run(_lib.FLAG_BLEND_TCGEN05), run(_lib.FLAG_BLEND_TF32)
    assert _maxerr(f16, tf32) <= 2e-6
    try:
        simt = run(_lib.FLAG_BLEND_SIMT)
    except RuntimeError as e:         # A/B kernel: only in -DSMPLK_AB builds
        assert "SMPLK_AB" in str(e)
        return
    assert _maxerr(simt, f16) <= 5e-6 and _maxerr(simt, tf32) <= 5e-6


@pytest.mark.parametrize("B", [129, 257, 300, 1000, 2049])
def test_fused_blend_skinning_matches_two_kernel_path_and_oracle(dev, smplh_model, B, monkeypatch):
    """The forward without SAVE_FOR_BACKWARD runs the fused kernel (blend GEMM whose epilogue skins
    straight from TMEM); the handle option fused=0 selects blend GEMM -> v_posed -> skinning kernel.  Ragged
    batches exercise partial 256-body blocks and the last 85-vertex tile (5 vertices)."""
    m = smplh_model
    betas, pose, transl = synthetic.make_inputs(m, B, seed=100 + B)
    pose[0] = 0.0
    pose[B - 1] = pose[B - 1] / np.abs(pose[B - 1]).max() * 4.8
    dm_f = smplk.DeviceModel(m, device=0, extra_vertex_ids=m["extra_vertex_ids"])
    dm_u = smplk.DeviceModel(m, device=0, extra_vertex_ids=m["extra_vertex_ids"], options={"fused": 0})
    args = (_t(betas, dev), _t(pose, dev))
    dm_f.profile_enable(True)
    vf, jf, _, _ = body_model_apply(dm_f, *args, transl=_t(transl, dev))
    vf0, _, _, _ = body_model_apply(dm_f, *args)                     # no translation
    torch.cuda.synchronize()
    prof = dm_f.profile_read()
    assert prof["blend_skin_fused"][1] == 2 and prof["skin"][1] == 0 and prof["blend_tcgen05"][1] == 0
    dm_u.profile_enable(True)
    vu, ju, _, _ = body_model_apply(dm_u, *args, transl=_t(transl, dev))
    torch.cuda.synchronize()
    prof = dm_u.profile_read()
    assert prof["blend_skin_fused"][1] == 0 and prof["skin"][1] == 1
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert _maxerr(vf, ref.vertices) <= TOL and _maxerr(vu, ref.vertices) <= TOL
    assert _maxerr(vf, vu) <= 3e-6 and _maxerr(jf[:, :52], ju[:, :52]) == 0.0 and _maxerr(jf, ju) <= 3e-6
    assert _maxerr(vf0 + _t(transl, dev)[:, None, :], vf) <= 3e-6
    assert torch.isfinite(vf).all()


@pytest.mark.parametrize("V,B", [(1000, 300), (7001, 260)])
def test_fused_path_other_vertex_counts(dev, V, B):
    """Vertex counts other than 6890 take the runtime-row-pitch instance of the fused kernel; the
    last 84-vertex tile / 12-vertex chunk are ragged."""
    m = synthetic.make_model("smplh", seed=7, num_verts=V)
    dm = smplk.DeviceModel(m, device=0)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=V)
    dm.profile_enable(True)
    v, j, _, _ = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev))
    torch.cuda.synchronize()
    assert dm.profile_read()["blend_skin_fused"][1] == 1
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert _maxerr(v, ref.vertices) <= TOL and _maxerr(j, ref.joints) <= TOL


def test_fused_forward_is_deterministic_and_agrees_with_two_kernel_path_on_random_batches(dev, smplh_model, monkeypatch):
    """Race / hazard screen for the fused kernel (cp.async ring, rolling staging window, shared
    accumulators): repeated runs must be BIT-identical, and every batch size must agree with the
    two-kernel forward.  Sizes straddle the 32-body pose blocks, the 256-body GEMM blocks and the
    8192-body chunks; outputs are pre-filled with NaN to expose unwritten elements."""
    m = smplh_model
    dm_f = smplk.DeviceModel(m, device=0)
    dm_u = smplk.DeviceModel(m, device=0, options={"fused": 0})
    rng = np.random.default_rng(77)
    sizes = [129, 255, 256, 288, 1023, 4097, 8192 + 130] + [int(x) for x in rng.integers(130, 3000, size=4)]
    for B in sizes:
        betas, pose, transl = synthetic.make_inputs(m, B, seed=B)
        args = (_t(betas, dev), _t(pose, dev))
        tr = _t(transl, dev)
        runs = []
        for _ in range(3):
            v = torch.full((B, dm_f.V, 3), float("nan"), device=dev)
            j = torch.full((B, dm_f.J, 3), float("nan"), device=dev)
            ws = torch.empty(dm_f.workspace_bytes(B, 0), device=dev, dtype=torch.uint8)
            a = _lib.ForwardArgs()
            a.batch, a.flags = B, 0
            a.betas, a.betas_batch = ctypes.c_void_p(args[0].data_ptr()), B
            a.pose, a.transl = ctypes.c_void_p(args[1].data_ptr()), ctypes.c_void_p(tr.data_ptr())
            a.verts, a.joints = ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(j.data_ptr())
            a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
            a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            dm_f.forward(a)
            torch.cuda.synchronize()
            runs.append((v, j))
        assert all(torch.equal(runs[0][0], r[0]) and torch.equal(runs[0][1], r[1]) for r in runs[1:]), B
        assert torch.isfinite(runs[0][0]).all()
        vu, ju, _, _ = body_model_apply(dm_u, *args, transl=tr)
        assert _maxerr(runs[0][0], vu) <= 3e-6 and _maxerr(runs[0][1], ju) == 0.0, B


def test_fused_path_smpl_24_joints_and_broadcast_betas(dev, smpl_model):
    m = smpl_model
    B = 513
    betas, pose, transl = synthetic.make_inputs(m, B, seed=5)
    dm = smplk.DeviceModel(m, device=0)
    dm.profile_enable(True)
    v, j, _, _ = body_model_apply(dm, _t(betas[:1], dev), _t(pose, dev), transl=_t(transl, dev))
    torch.cuda.synchronize()
    assert dm.profile_read()["blend_skin_fused"][1] == 1
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        torch.tensor(betas[:1], dtype=torch.float64).expand(B, -1),
        torch.tensor(pose, dtype=torch.float64), torch.tensor(transl, dtype=torch.float64))
    assert _maxerr(v, ref.vertices) <= TOL and _maxerr(j, ref.joints) <= TOL


def test_split_operands_hold_fp32_accuracy_on_large_blendshapes(dev):
    """Stress the two-term splits: blendshape magnitudes 30x the synthetic default (cm-scale pose
    correctives, dm-scale shape directions), betas up to |5|, pose up to pi."""
    m = synthetic.make_model("smplh", seed=41)
    m["posedirs"] = m["posedirs"] * 30.0
    m["shapedirs"] = m["shapedirs"] * 10.0
    m["posedirs"][5, 1, 7] = 0.9            # one huge entry sets the fp16 scale; tiny ones must survive
    m["posedirs"][6, 2, 8] = 1e-7
    dm = smplk.DeviceModel(m, device=0)
    B = 256
    betas, pose, transl = synthetic.make_inputs(m, B, seed=3, pose_sigma=0.8)
    betas = np.clip(betas * 2.5, -5, 5).astype(np.float32)
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    for flags in (0, _lib.FLAG_BLEND_TF32):
        v = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev), flags=flags)[0]
        assert _maxerr(v, ref.vertices) <= TOL, flags


def test_smpl_module_and_broadcast_betas(dev, smpl_model):
    m = smpl_model
    mod = SMPL(model=m, batch_size=1).to(dev)
    B = 40
    betas, pose, transl = synthetic.make_inputs(m, B, seed=2, broadcast_betas=True)
    out = mod(betas=_t(betas, dev), global_orient=_t(pose[:, :3], dev), body_pose=_t(pose[:, 3:], dev),
              transl=_t(transl, dev), return_full_pose=True)
    om = O.TorchOracleModel(m, dtype=torch.float64)
    ref = om.forward(torch.tensor(betas, dtype=torch.float64).expand(B, -1),
                     torch.tensor(pose[:, :3], dtype=torch.float64), torch.tensor(pose[:, 3:], dtype=torch.float64),
                     transl=torch.tensor(transl, dtype=torch.float64))
    assert out.vertices.shape == (B, 6890, 3) and out.joints.shape == (B, 24 + 21, 3)
    assert _maxerr(out.vertices, ref.vertices) <= TOL and _maxerr(out.joints, ref.joints) <= TOL
    assert _maxerr(out.full_pose, ref.full_pose) == 0.0
    # module defaults: zero parameters -> rest template, verts_numpy lazily materialised
    rest = mod()
    assert _maxerr(rest.vertices[0], torch.tensor(m["v_template"])) <= 1e-6
    assert mod.verts_numpy.shape == (6890, 3)
    assert mod(return_verts=False).vertices is None


def test_smplh_module_pca_mean_mapper_and_wrapper_extras(dev, smplh_model):
    """Everything models/smplh.py:26-39 + smplx.create(...) of lib/gen_smplh.py:75-90 does."""
    m = smplh_model
    mapper_idx = np.array([52, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 53, 54, 55, 56, 57, 58,
                           59, 60, 61, 62, 20, 34, 35, 36, 63, 21, 49, 50, 51, 68, 72])
    class Mapper(torch.nn.Module):
        def forward(self, joints):
            return torch.index_select(joints, 1, torch.as_tensor(mapper_idx, device=joints.device))
    joint_map = np.array([3, 0, 36 + 8, 36 + 2, 7])
    B = 5
    mod = SMPLH(model=m, joint_mapper=Mapper(), use_pca=True, num_pca_comps=12, flat_hand_mean=False,
                create_transl=False, batch_size=B, J_regressor_extra=m["J_regressor_extra"],
                joint_map=joint_map).to(dev)
    rng = np.random.default_rng(4)
    kw = dict(betas=rng.standard_normal((B, 16)), global_orient=rng.standard_normal((B, 3)) * 0.4,
              body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)),
              right_hand_pose=rng.standard_normal((B, 12)))
    out = mod(**{k: _t(v, dev) for k, v in kw.items()}, return_full_pose=True, get_skin=True)
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12, joint_mapper=mapper_idx, joint_map=joint_map)
    ref = om.forward(*[torch.tensor(kw[k]) for k in ("betas", "global_orient", "body_pose", "left_hand_pose",
                                                    "right_hand_pose")], wrapper_extra=True)
    assert _maxerr(out.vertices, ref.vertices) <= TOL
    assert out.joints.shape == ref.joints.shape == (B, 5, 3)
    assert _maxerr(out.joints, ref.joints) <= TOL
    assert _maxerr(out.full_pose, ref.full_pose) <= 1e-6
    # flat_hand_mean + axis-angle hands
    mod2 = SMPLH(model=m, use_pca=False, flat_hand_mean=True, batch_size=2).to(dev)
    lh, rh = rng.standard_normal((2, 45)) * 0.2, rng.standard_normal((2, 45)) * 0.2
    out2 = mod2(left_hand_pose=_t(lh, dev), right_hand_pose=_t(rh, dev))
    om2 = O.TorchOracleModel(m, dtype=torch.float64, flat_hand_mean=True)
    z = lambda *s: torch.zeros(*s, dtype=torch.float64)
    ref2 = om2.forward(z(2, 16), z(2, 3), z(2, 63), torch.tensor(lh), torch.tensor(rh), transl=z(2, 3), use_pca=False)
    assert _maxerr(out2.vertices, ref2.vertices) <= TOL and _maxerr(out2.joints, ref2.joints) <= TOL


@pytest.mark.parametrize("B", [300, 33])
def test_smplh_module_inference_batches_with_pca_hands_and_pose_mean(dev, smplh_model, B):
    """The module under torch.no_grad() at a GEMM-sized batch: block pose kernel (PCA hands through
    warp broadcasts, pose mean from smem) -> fused blend+skinning; and at a batch below the tcgen05
    threshold of 128 rows.  One betas row broadcast, only the left hand given as PCA input."""
    m = smplh_model
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=6, flat_hand_mean=False, create_transl=True,
                batch_size=B).to(dev)
    rng = np.random.default_rng(B)
    kw = dict(betas=rng.standard_normal((1, 16)), global_orient=rng.standard_normal((B, 3)) * 0.4,
              body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 6)),
              right_hand_pose=rng.standard_normal((B, 6)) * 0.0, transl=rng.standard_normal((B, 3)))
    with torch.no_grad():
        out = mod(**{k: _t(v, dev) for k, v in kw.items()}, return_full_pose=True)
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=6)
    ref = om.forward(torch.tensor(kw["betas"]).expand(B, -1), torch.tensor(kw["global_orient"]),
                     torch.tensor(kw["body_pose"]), torch.tensor(kw["left_hand_pose"]),
                     torch.tensor(kw["right_hand_pose"]), transl=torch.tensor(kw["transl"]))
    assert _maxerr(out.vertices, ref.vertices) <= TOL and _maxerr(out.joints, ref.joints) <= TOL
    assert _maxerr(out.full_pose, ref.full_pose) <= 1e-6


def test_dense_weights_take_the_generic_path(dev):
    m = synthetic.make_model("smplh", seed=5, dense_weights=True, dense_regressor=True)
    dm = smplk.DeviceModel(m, device=0)
    assert dm.info.max_weights_per_vertex == 52
    betas, pose, transl = synthetic.make_inputs(m, 6, seed=1)
    v = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev))[0]
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert _maxerr(v, ref.vertices) <= TOL
    # dense weights at a GEMM-sized batch: every 16-vertex chunk would list all 52 joints, so the
    # packer keeps this model off the fused kernel (blend GEMM + generic skinning kernel instead)
    B = 200
    betas, pose, transl = synthetic.make_inputs(m, B, seed=2)
    dm.profile_enable(True)
    v = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev))[0]
    torch.cuda.synchronize()
    prof = dm.profile_read()
    assert prof["blend_skin_fused"][1] == 0 and prof["skin"][1] == 1
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert _maxerr(v, ref.vertices) <= TOL


def test_rigid_and_smooth_weights_take_the_8_vertex_group_kernel(dev):
    """Weights whose 8-vertex groups touch <= 8 joints (rigid binding; smooth two-joint blends)
    run skin_grouped8_kernel; the synthetic default (random secondary joints) does not."""
    import ctypes as C
    for max_nnz in (1, 2):
        m = synthetic.make_model("smplh", seed=17, max_nnz=max_nnz)
        if max_nnz == 2:   # smooth blend between a joint and its parent along each range
            W = np.zeros_like(m["weights"])
            prim = m["weights"].argmax(1)
            par = np.array(synthetic.SMPLH_PARENTS)
            t = (np.arange(6890) % 50) / 50.0
            for v in range(6890):
                j = prim[v]; p = par[j] if par[j] >= 0 else j
                W[v, j] += 1.0 - 0.5 * t[v]; W[v, p] += 0.5 * t[v]
            m["weights"] = W
        dm = smplk.DeviceModel(m, device=0)
        betas, pose, transl = synthetic.make_inputs(m, 70, seed=2)
        v = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(transl, dev))[0]
        ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
            *[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
        assert _maxerr(v, ref.vertices) <= TOL


def test_numpy_twins_reproduce_reference_golden_vectors(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "smplh_np_twin.npz"))
    m = synthetic.make_model("smplh", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    twin = smplk.SMPLHModel(m)
    assert np.abs(twin.verts[::53] - g["rest_verts_sub64"]).max() <= TOL      # constructor runs update()
    for i in range(g["pose"].shape[0]):
        v = twin.set_params(pose=g["pose"][i].reshape(52, 3), beta=g["beta"][i], trans=g["trans"][i])
        assert v.dtype == np.float64 and v.shape == (6890, 3)
        assert np.abs(v - g["verts"][i]).max() <= TOL
        assert np.abs(twin.gen_J_3d() - g["j3d"][i]).max() <= TOL
    allv = twin.forward_batch(g["pose"], g["beta"], g["trans"])
    assert np.abs(allv - g["verts"]).max() <= TOL
    g = np.load(os.path.join(golden_dir, "smpl_np_twin.npz"))
    m = synthetic.make_model("smpl", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    twin = smplk.SMPLModel(m)
    for i in range(g["pose"].shape[0]):
        v = twin.set_params(pose=g["pose"][i].reshape(24, 3), beta=g["beta"][i], trans=g["trans"][i])
        assert np.abs(v - g["verts"][i]).max() <= TOL
        assert np.abs(twin.gen_J_3d() - g["j3d"][i]).max() <= TOL


def test_recover_model_matches_reference_golden_vectors(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "recover_lbs.npz"))
    rig = synthetic.make_rigged_mesh(int(g["num_verts"]), seed=int(g["seed"]))
    rm = smplk.RecoverModel(rig)
    for i in range(g["pose"].shape[0]):
        v = rm.set_params(pose=g["pose"][i].reshape(24, 3).copy(), trans=g["trans"][i])
        assert np.abs(v - g["verts"][i]).max() <= TOL
    clip = rm.replay(g["pose"], g["trans"])
    assert clip.shape == g["verts"].shape and np.abs(clip - g["verts"]).max() <= TOL


def test_rodrigues_kernel_matches_geometry_golden(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "rodrigues_quat.npz"))
    aa = _t(g["theta"], dev).contiguous()
    out = torch.empty(aa.shape[0], 3, 3, device=dev)
    lib = smplk.load()
    _lib.check(lib.smplk_batch_rodrigues(aa.shape[0], ctypes.c_void_p(aa.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                         0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert _maxerr(out, torch.tensor(g["R"])) <= 2e-6


@pytest.mark.parametrize("kind,B,pca,joint_w", [("smplh", 4, True, 0.5), ("smpl", 130, False, 0.0),
                                                ("smplh", 64, False, 1.0)])
def test_backward_matches_autograd_oracle(dev, kind, B, pca, joint_w):
    m = synthetic.make_model(kind, seed=31)
    J = 52 if kind == "smplh" else 24
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    dm = smplk.DeviceModel(m, device=0, num_pca_comps=12 if pca else 0, extra_vertex_ids=m["extra_vertex_ids"],
                           regressor_posed=m["J_regressor_extra"])
    rng = np.random.default_rng(7)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=11)
    if not pca:
        pose[0, 3:9] = 0.0      # exercise the theta -> 0 branch of the Rodrigues backward
    lh, rh = rng.standard_normal((B, 12)), rng.standard_normal((B, 12))
    tgt_v = rng.standard_normal((B, 6890, 3))
    tgt_j = rng.standard_normal((B, J + 21, 3))
    tgt_r = rng.standard_normal((B, 9, 3))

    def loss_fn(v, j, r):
        l = ((v - tgt_v_t(v)) ** 2).sum()
        if joint_w:
            l = l + joint_w * ((j - tgt_j_t(j)) ** 2).sum() + joint_w * ((r - tgt_r_t(r)) ** 2).sum()
        return l
    tgt_v_t = lambda x: torch.as_tensor(tgt_v, dtype=x.dtype, device=x.device)
    tgt_j_t = lambda x: torch.as_tensor(tgt_j, dtype=x.dtype, device=x.device)
    tgt_r_t = lambda x: torch.as_tensor(tgt_r, dtype=x.dtype, device=x.device)

    # oracle (float64 autograd)
    ob, op, ot = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (betas, pose, transl))
    ol, orr = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (lh, rh))
    if pca:
        o = om.forward(ob, op[:, :3], op[:, 3:66], ol, orr, transl=ot, wrapper_extra=True)
    else:
        o = om.forward_full_pose(ob, op, ot)
        picks = o.vertices[:, torch.as_tensor(m["extra_vertex_ids"], dtype=torch.long)]
        extra = torch.einsum("bik,ji->bjk", o.vertices, om.J_regressor_extra)
        o = O.OracleOutput(o.vertices, torch.cat([o.joints, picks, extra], 1), None, None, None, None)
    loss_fn(o.vertices, o.joints[:, :J + 21], o.joints[:, J + 21:]).backward()

    gb, gp, gt = (_t(x, dev, True) for x in (betas, pose, transl))
    gl, gr = (_t(x, dev, True) for x in (lh, rh))
    v, j, r, _ = body_model_apply(dm, gb, gp, pca_l=gl if pca else None, pca_r=gr if pca else None, transl=gt,
                                  add_pose_mean=pca, want_regressed=True)
    loss_fn(v, j, r).backward()
    pairs = [("betas", gb.grad, ob.grad), ("transl", gt.grad, ot.grad)]
    if pca:
        pairs += [("pose", gp.grad[:, :66], op.grad[:, :66]), ("lh", gl.grad, ol.grad), ("rh", gr.grad, orr.grad)]
        assert float(gp.grad[:, 66:].abs().max()) == 0.0
    else:
        pairs += [("pose", gp.grad, op.grad)]
    for name, got, ref in pairs:
        rel = _maxerr(got, ref) / float(ref.abs().max())
        assert rel <= GRAD_RTOL, (name, rel)


@pytest.mark.parametrize("kind,B,pca", [("smplh", 6, True), ("smpl", 33, False)])
def test_keypoint_only_backward_takes_the_sparse_path_and_matches_oracle(dev, kind, B, pca, monkeypatch):
    """A loss on joints alone (FK joints + vertex picks: the data term of lib/Gen_SMPLH/fitting.py:369-381)
    runs the sparse pick kernel instead of the dense vertex backward: same gradients as the float64
    autograd oracle and as the dense path (handle option sparse_picks=0), fewer launches."""
    from smplk import _lib
    m = synthetic.make_model(kind, seed=17)
    J = 52 if kind == "smplh" else 24
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    rng = np.random.default_rng(5)
    betas, pose, transl = synthetic.make_inputs(m, B, seed=12)
    lh, rh = rng.standard_normal((B, 12)), rng.standard_normal((B, 12))
    tgt_j = rng.standard_normal((B, J + 21, 3))
    wj = rng.uniform(0.2, 2.0, size=(1, J + 21, 1))

    def loss_fn(j):
        t = torch.as_tensor(tgt_j, dtype=j.dtype, device=j.device)
        w = torch.as_tensor(wj, dtype=j.dtype, device=j.device)
        return (w * (j - t) ** 2).sum()

    ob, op, ot = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (betas, pose, transl))
    ol, orr = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (lh, rh))
    if pca:
        o = om.forward(ob, op[:, :3], op[:, 3:66], ol, orr, transl=ot)
        oj = o.joints[:, :J + 21]
    else:
        o = om.forward_full_pose(ob, op, ot)
        oj = torch.cat([o.joints, o.vertices[:, torch.as_tensor(m["extra_vertex_ids"], dtype=torch.long)]], 1)
    loss_fn(oj).backward()

    results = {}
    for mode in ("sparse", "dense"):
        dm = smplk.DeviceModel(m, device=0, num_pca_comps=12 if pca else 0, extra_vertex_ids=m["extra_vertex_ids"],
                               options={"sparse_picks": 0} if mode == "dense" else None)
        gb, gp, gt = (_t(x, dev, True) for x in (betas, pose, transl))
        gl, gr = (_t(x, dev, True) for x in (lh, rh))
        v, j, _, _ = body_model_apply(dm, gb, gp, pca_l=gl if pca else None, pca_r=gr if pca else None, transl=gt,
                                      add_pose_mean=pca)
        loss = loss_fn(j)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        loss.backward()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        grads = [gb.grad, gt.grad, gp.grad[:, :66] if pca else gp.grad] + ([gl.grad, gr.grad] if pca else [])
        results[mode] = (launches, grads)
    refs = [ob.grad, ot.grad, op.grad[:, :66] if pca else op.grad] + ([ol.grad, orr.grad] if pca else [])
    for got, ref in zip(results["sparse"][1], refs):
        assert _maxerr(got, ref) / float(ref.abs().max()) <= GRAD_RTOL
    for got, ref in zip(results["sparse"][1], results["dense"][1]):
        assert _maxerr(got, ref) <= 1e-5 * float(ref.abs().max())
    assert results["sparse"][0] == 2 and results["dense"][0] > results["sparse"][0], (results["sparse"][0], results["dense"][0])


def test_joints_only_forward_matches_dense_forward_and_backward(dev, smplh_model):
    """return_verts=False (fit_single_frame.py:313, fitting.py:82): only the 21 picked vertices are
    blended and skinned.  Joints agree with the full forward to 1e-6 m, gradients of a joints loss with
    the dense route, 3 launches forward+backward... and the oracle."""
    from smplk import _lib
    from smplk.body_models import SMPLH
    B = 11
    mod = SMPLH(model=smplh_model, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)
    rng = np.random.default_rng(8)
    vals = dict(betas=rng.standard_normal((B, 16)), global_orient=rng.standard_normal((B, 3)) * 0.3,
                body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)),
                right_hand_pose=rng.standard_normal((B, 12)), transl=rng.standard_normal((B, 3)))
    mod.reset_params(**vals)
    tgt = _t(rng.standard_normal((B, 73, 3)), dev)
    grads, joints, launches = [], [], []
    for rv in (False, True):
        mod.zero_grad()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        out = mod(return_verts=rv)
        ((out.joints - tgt) ** 2).sum().backward()
        torch.cuda.synchronize()
        launches.append(_lib.launch_count() - n0)
        assert (out.vertices is None) == (not rv)
        joints.append(out.joints.detach().clone())
        grads.append({n: p.grad.clone() for n, p in mod.named_parameters()})
    assert _maxerr(joints[0], joints[1]) <= 1e-6
    for n in grads[0]:
        assert _maxerr(grads[0][n], grads[1][n]) <= 1e-5 * float(grads[1][n].abs().max()), n
    assert launches[0] == 4 and launches[1] > launches[0], launches      # pose, picks | picks backward, pose backward
    # oracle joints
    om = O.TorchOracleModel(smplh_model, dtype=torch.float64, num_pca_comps=12)
    t = {k: torch.tensor(v) for k, v in vals.items()}
    ref = om.forward(t["betas"], t["global_orient"], t["body_pose"], t["left_hand_pose"], t["right_hand_pose"],
                     transl=t["transl"])
    assert _maxerr(joints[0].double().cpu(), ref.joints[:, :73]) <= TOL
    # no-grad inference and the rigged-mesh-free SMPL case
    with torch.no_grad():
        j2 = mod(return_verts=False).joints
    assert _maxerr(j2, joints[0]) <= 1e-6


def test_handles_of_different_skeletons_coexist(dev, smplh_model, smpl_model):
    """Dynamic shared-memory limits belong to the kernels, not to a handle: creating a handle for a
    smaller skeleton (24 joints, 10 betas) after a larger one (52 joints, 16 betas) must not break
    the larger model's launches (pose block kernel, grouped skinning, backward kernels)."""
    dm_h = smplk.DeviceModel(smplh_model, device=0)
    dm_s = smplk.DeviceModel(smpl_model, device=0)          # created second: smaller tables
    rig = smplk.DeviceModel(synthetic.make_rigged_mesh(3000, seed=1), device=0, lbs_only=True)
    for dm, m in ((dm_h, smplh_model), (dm_s, smpl_model), (dm_h, smplh_model)):
        B = 200
        betas, pose, transl = synthetic.make_inputs(m, B, seed=3)
        tb, tp, tt = (_t(x, dev, True) for x in (betas, pose, transl))
        v = body_model_apply(dm, tb, tp, transl=tt)[0]        # SAVE_FOR_BACKWARD: two-kernel forward
        (v ** 2).sum().backward()
        with torch.no_grad():
            v2 = body_model_apply(dm, tb, tp, transl=tt)[0]   # fused forward
        assert torch.isfinite(tp.grad).all() and _maxerr(v, v2) <= 3e-6
    del rig


def test_backward_is_safe_for_any_gradient_range(dev, smplh_model):
    """The backward GEMM runs on fp16 two-term-split operands; every body's d_v_posed row is scaled
    into the fp16 range by its own power of two (from max|d_verts| of that body).  Upstream gradients
    spanning 16 decades across the batch, a body with zero gradient and one with a single non-zero
    entry must all keep the 1e-4 relative bound PER BODY."""
    m = smplh_model
    dm = smplk.DeviceModel(m, device=0)
    B = 140
    betas, pose, transl = synthetic.make_inputs(m, B, seed=31)
    rng = np.random.default_rng(5)
    scale = 10.0 ** rng.uniform(-8, 8, size=B)
    scale[3] = 0.0
    cot = rng.standard_normal((B, 6890, 3)) * scale[:, None, None]
    cot[7] = 0.0
    cot[7, 1234, 1] = 3e-5
    tb, tp, tt = (_t(x, dev, True) for x in (betas, pose, transl))
    v = body_model_apply(dm, tb, tp, transl=tt)[0]
    (v * _t(cot, dev)).sum().backward()
    ob, op_, ot = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (betas, pose, transl))
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(ob, op_, ot)
    (ref.vertices * torch.tensor(cot)).sum().backward()
    for got, want, name in ((tb.grad, ob.grad, "betas"), (tp.grad, op_.grad, "pose"), (tt.grad, ot.grad, "transl")):
        g = got.double().cpu()
        assert torch.isfinite(g).all(), name
        den = want.abs().amax(dim=1).clamp_min(1e-300)
        rel = ((g - want).abs().amax(dim=1) / den)
        rel[want.abs().amax(dim=1) == 0] = g.abs().amax(dim=1)[want.abs().amax(dim=1) == 0]
        # fp32 cotangents: bodies whose scale underflows float32 are exactly zero on both sides
        assert float(rel.max()) <= GRAD_RTOL, (name, int(rel.argmax()), float(rel.max()))


def test_backward_broadcast_betas_and_module_autograd(dev, smplh_model):
    m = smplh_model
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=1, create_transl=False).to(dev)
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12)
    B = 9
    rng = np.random.default_rng(5)
    bp = rng.standard_normal((B, 63)) * 0.3
    with torch.no_grad():
        mod.betas.copy_(_t(rng.standard_normal((1, 16)), dev))
        mod.left_hand_pose.copy_(_t(rng.standard_normal((1, 12)), dev))
    tgt = rng.standard_normal((B, 73, 3))
    out = mod(body_pose=_t(bp, dev), return_verts=True)
    ((out.joints - _t(tgt, dev)) ** 2).sum().backward()      # SMPLifyCameraInitLoss-style joints L2
    ob = mod.betas.detach().double().cpu().requires_grad_(True)
    ol = mod.left_hand_pose.detach().double().cpu().requires_grad_(True)
    z = lambda *s: torch.zeros(*s, dtype=torch.float64)
    o = om.forward(ob.expand(B, -1), z(B, 3), torch.tensor(bp), ol.expand(B, -1), z(B, 12))
    ((o.joints - torch.tensor(tgt)) ** 2).sum().backward()
    for got, ref in ((mod.betas.grad, ob.grad), (mod.left_hand_pose.grad, ol.grad)):
        assert _maxerr(got, ref) / float(ref.abs().max()) <= GRAD_RTOL
    assert mod.global_orient.grad is not None and mod.right_hand_pose.grad is not None


def test_full_size_properties_batch_4096(dev, smplh_model):
    """BASELINE config 2 size; properties that need no oracle at this size."""
    m = smplh_model
    dm = smplk.DeviceModel(m, device=0)
    B = 4096
    betas, pose, transl = synthetic.make_inputs(m, B, seed=77)
    tb, tp, tt = _t(betas, dev), _t(pose, dev), _t(transl, dev)
    v = body_model_apply(dm, tb, tp, transl=tt)[0]
    assert torch.isfinite(v).all()
    # (1) shard equivalence, bitwise: concat of shard results == single-call result
    from smplk.sharding import shard_bounds
    for world in (2, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_bounds(B, world, r)
            parts.append(body_model_apply(dm, tb[lo:hi].contiguous(), tp[lo:hi].contiguous(),
                                          transl=tt[lo:hi].contiguous())[0])
        assert torch.equal(torch.cat(parts), v)
    # (2) translation equivariance: verts(t + d) - verts(t) == d up to fp32 rounding of the add
    d = torch.tensor([0.25, -1.5, 3.0], device=dev)
    v2 = body_model_apply(dm, tb, tp, transl=tt + d)[0]
    assert float((v2 - v - d).abs().max()) <= 2e-6
    # (3) determinism
    assert torch.equal(body_model_apply(dm, tb, tp, transl=tt)[0], v)
    # (4) spot-check 16 random bodies of the big batch against the float64 oracle
    idx = np.random.default_rng(0).choice(B, 16, replace=False)
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        *[torch.tensor(x[idx], dtype=torch.float64) for x in (betas, pose, transl)])
    assert _maxerr(v[torch.as_tensor(idx, device=dev)], ref.vertices) <= TOL
    # (5) global rotation by pi about z maps (x,y,z) -> (-x,-y,z) when only the root is posed
    p0 = torch.zeros(4, 156, device=dev)
    p1 = p0.clone(); p1[:, 2] = float(np.pi)
    a = body_model_apply(dm, tb[:4].contiguous(), p0)[0]
    b = body_model_apply(dm, tb[:4].contiguous(), p1)[0]
    j0 = body_model_apply(dm, tb[:4].contiguous(), p0)[1][:, :1]
    assert float(((a - j0) * torch.tensor([-1., -1., 1.], device=dev) - (b - j0)).abs().max()) <= 5e-6


def test_chunked_large_batch_and_sequence_replay(dev, smplh_model):
    """BASELINE configs 4/5 shapes: batches beyond the 8192-body chunk and a long motion sequence
    with ONE broadcast betas row (the 100k-frame replay of config 5, shortened to 20k frames here;
    frames are independent so the property is size independent)."""
    m = smplh_model
    dm = smplk.DeviceModel(m, device=0)
    N = 20000
    rng = np.random.default_rng(3)
    base = rng.standard_normal((97, 156)).astype(np.float32) * 0.3     # a 97-frame clip, tiled
    pose = np.tile(base, (N // 97 + 1, 1))[:N]
    trans = np.cumsum(rng.standard_normal((N, 3)).astype(np.float32) * 0.01, axis=0)
    betas = rng.standard_normal((1, 16)).astype(np.float32)
    v, j, _, _ = body_model_apply(dm, _t(betas, dev), _t(pose, dev), transl=_t(trans, dev))
    assert v.shape == (N, 6890, 3) and torch.isfinite(v).all()
    # periodicity: frame i and i+97 differ exactly by their translation difference
    d = (v[97:97 + 500] - v[:500]) - (_t(trans[97:597], dev) - _t(trans[:500], dev))[:, None]
    assert float(d.abs().max()) <= 4e-6
    # spot-check frames on both sides of the 8192 chunk boundaries against the oracle
    idx = np.array([0, 8191, 8192, 8193, 16383, 16384, 19999])
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        torch.tensor(np.repeat(betas, len(idx), 0), dtype=torch.float64),
        torch.tensor(pose[idx], dtype=torch.float64), torch.tensor(trans[idx], dtype=torch.float64))
    ii = torch.as_tensor(idx, device=dev)
    assert _maxerr(v[ii], ref.vertices) <= TOL and _maxerr(j[ii], ref.joints) <= TOL


def test_rigged_mesh_replay_large_vertex_count(dev):
    """config 5, LBS-only variant: recovered mesh with Nv = 50,001 (odd 3*Nv -> scalar store path)
    and Nv = 200,000, replayed over a clip in one call."""
    for nv in (50001, 200000):
        rig = synthetic.make_rigged_mesh(nv, seed=13)
        rm = smplk.RecoverModel(rig)
        rng = np.random.default_rng(nv)
        poses = rng.standard_normal((40, 72)) * 0.4
        trans = rng.standard_normal((40, 3))
        out = rm.replay(poses, trans)
        assert out.shape == (40, nv, 3)
        for i in (0, 39):
            ref = O.np_lbs_only(rig, poses[i], trans[i])["verts"]
            assert np.abs(out[i] - ref).max() <= TOL


def test_fused_vertex_l2_loss_matches_torch(dev, smpl_model):
    from smplk.body_models import vertex_l2_loss
    m = smpl_model
    dm = smplk.DeviceModel(m, device=0)
    B = 37
    betas, pose, transl = synthetic.make_inputs(m, B, seed=4)
    tgt = torch.randn(B, 6890, 3, device=dev)
    grads = []
    for fused in (True, False):
        tb, tp, tt = (_t(x, dev, True) for x in (betas, pose, transl))
        v = body_model_apply(dm, tb, tp, transl=tt)[0]
        per_body = vertex_l2_loss(v, tgt, 0.5) if fused else 0.5 * ((v - tgt) ** 2).sum(dim=(1, 2))
        (per_body * torch.arange(1, B + 1, device=dev)).sum().backward()
        grads.append((per_body.detach(), tb.grad, tp.grad, tt.grad))
    for a, b in zip(*grads):
        assert _maxerr(a, b) <= 1e-5 * float(b.abs().max())
    no_grad = vertex_l2_loss(v.detach(), tgt)
    assert no_grad.shape == (B,) and not no_grad.requires_grad


def test_one_node_fitting_step_matches_two_nodes(dev, smplh_model, smpl_model):
    """smplk.fit_vertex_l2 (body model + loss as one autograd node, BASELINE config 3) against the
    torch loss on the differentiable vertices: per-body weights on the loss, per-body and shared
    betas, hand PCA through the module method."""
    from smplk.body_models import fit_vertex_l2, SMPLH
    m = smpl_model
    dm = smplk.DeviceModel(m, device=0)
    for B, shared in ((37, False), (5, True), (1, False)):
        betas, pose, transl = synthetic.make_inputs(m, B, seed=4)
        if shared:
            betas = betas[:1]
        tgt = torch.randn(B, 6890, 3, device=dev)
        wts = torch.arange(1, B + 1, device=dev, dtype=torch.float32)
        res = []
        for node in (True, False):
            tb, tp, tt = (_t(x, dev, True) for x in (betas, pose, transl))
            if node:
                per_body = fit_vertex_l2(dm, tb, tp, tgt, transl=tt, scale=0.5)
            else:
                v = body_model_apply(dm, tb, tp, transl=tt)[0]
                per_body = 0.5 * ((v - tgt) ** 2).sum(dim=(1, 2))
            (per_body * wts).sum().backward()
            res.append((per_body.detach(), tb.grad, tp.grad, tt.grad))
        for a, b in zip(*res):
            assert a.shape == b.shape
            assert _maxerr(a, b) <= 2e-5 * float(b.abs().max()), (B, shared)
    # module method, SMPL-H with PCA hands and pose mean
    B = 9
    mod = SMPLH(model=smplh_model, use_pca=True, num_pca_comps=12, batch_size=B).to(dev)
    rng = np.random.default_rng(3)
    vals = dict(betas=rng.standard_normal((B, 16)), global_orient=rng.standard_normal((B, 3)) * 0.3,
                body_pose=rng.standard_normal((B, 63)) * 0.3, left_hand_pose=rng.standard_normal((B, 12)),
                right_hand_pose=rng.standard_normal((B, 12)), transl=rng.standard_normal((B, 3)))
    tgt = torch.randn(B, 6890, 3, device=dev)
    grads = []
    for node in (True, False):
        mod.reset_params(**vals)
        mod.zero_grad()
        loss = mod.vertex_l2(tgt).sum() if node else ((mod(return_verts=True).vertices - tgt) ** 2).sum()
        loss.backward()
        grads.append([float(loss)] + [p.grad.clone() for _, p in sorted(mod.named_parameters())])
    assert abs(grads[0][0] - grads[1][0]) <= 1e-5 * abs(grads[1][0])
    for a, b in zip(grads[0][1:], grads[1][1:]):
        assert _maxerr(a, b) <= 2e-5 * float(b.abs().max())
    # a target view at an odd float offset (4-byte aligned only) is copied, not read misaligned
    flat = torch.zeros(B * 6890 * 3 + 1, device=dev)
    flat[1:] = tgt.reshape(-1)
    odd = flat[1:].view(B, 6890, 3)
    assert odd.data_ptr() % 8 == 4
    mod.reset_params(**vals)
    assert abs(float(mod.vertex_l2(odd).sum()) - grads[1][0]) <= 1e-5 * abs(grads[1][0])
    # reduce="sum": the scalar accumulated in the kernel, scaled by an upstream factor in backward
    mod.reset_params(**vals)
    mod.zero_grad()
    total = mod.vertex_l2(tgt, reduce="sum")
    assert total.dim() == 0
    (0.5 * total).backward()
    assert abs(float(total) - grads[1][0]) <= 1e-5 * abs(grads[1][0])
    for (_, p), b in zip(sorted(mod.named_parameters()), grads[1][1:]):
        assert _maxerr(p.grad, 0.5 * b) <= 2e-5 * float(b.abs().max())


def test_errors_are_loud(dev, smpl_model):
    dm = smplk.DeviceModel(smpl_model, device=0)
    b, p, t = synthetic.make_inputs(smpl_model, 2)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        body_model_apply(dm, torch.tensor(b), torch.tensor(p))
    with pytest.raises(RuntimeError, match="float32"):
        body_model_apply(dm, _t(b, dev).double(), _t(p, dev).double())
    a = _lib.ForwardArgs()
    a.batch = 2
    a.pose = ctypes.c_void_p(_t(p, dev).data_ptr())
    with pytest.raises(RuntimeError, match="workspace"):
        dm.forward(a)
    # shapes the C side would stride over blindly are refused before the call (upstream raises on these too)
    with pytest.raises(ValueError, match="betas"):
        body_model_apply(dm, _t(np.zeros((3, 10)), dev), _t(p, dev))            # 3 rows for a batch of 2
    with pytest.raises(ValueError, match="betas"):
        body_model_apply(dm, _t(np.zeros((2, 16)), dev), _t(p, dev))            # 16 betas on a 10-beta model
    with pytest.raises(ValueError, match="pose"):
        body_model_apply(dm, _t(b, dev), _t(np.zeros((2, 156)), dev))           # SMPL-H pose on the 24-joint model
    with pytest.raises(ValueError, match="transl"):
        body_model_apply(dm, _t(b, dev), _t(p, dev), transl=_t(np.zeros((1, 3)), dev))
    from smplk.body_models import vertex_l2_loss, fit_vertex_l2
    v = body_model_apply(dm, _t(b, dev), _t(p, dev))[0]
    with pytest.raises(ValueError, match="shape"):
        vertex_l2_loss(v, v[0])                                                  # (V,3) target for (B,V,3) vertices
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        vertex_l2_loss(v, v.cpu())
    with pytest.raises(ValueError, match="target"):
        fit_vertex_l2(dm, _t(b, dev), _t(p, dev), v[:1])
    # the C ABI itself still refuses an inconsistent betas_batch
    a.betas, a.betas_batch = ctypes.c_void_p(_t(b, dev).data_ptr()), 3
    ws = torch.empty(dm.workspace_bytes(2, 0), device=dev, dtype=torch.uint8)
    a.workspace, a.workspace_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    with pytest.raises(RuntimeError, match="betas_batch"):
        dm.forward(a)
