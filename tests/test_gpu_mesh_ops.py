"""GPU parity of the mesh operations either side of the forward against the
oracle and the reference-made golden vectors.  Index work (face split, re-indexing) is bit-exact;
floating point within the tolerances written below."""
import os

import numpy as np
import pytest
import torch

import smplk
from smplk import synthetic
from smplk.body_models import body_model_apply
from smplk.mesh_ops import MeshTopology, inverse_joints, inverse_lbs, transforms
from oracle import smpl_oracle as O

pytestmark = pytest.mark.gpu


def _t(x):
    return torch.tensor(np.asarray(x, dtype=np.float32), device="cuda:0")


def test_transforms_only_flag_and_inverse_lbs_roundtrip():
    """to_T_pose (lib/mesh2smpl_model.py:183-207): pose a rigged mesh, then remove the pose again."""
    rig = synthetic.make_rigged_mesh(5003, seed=2)
    dm = smplk.DeviceModel(rig, device=0, lbs_only=True)
    rng = np.random.default_rng(0)
    B = 5
    pose = rng.standard_normal((B, 72)) * 0.5
    trans = rng.standard_normal((B, 3))
    A, joints = transforms(dm, None, _t(pose), _t(trans))
    posed = body_model_apply(dm, None, _t(pose), transl=_t(trans))[0]
    for b in (0, B - 1):
        ref = O.np_lbs_only(rig, pose[b], trans[b], ignore_joints=())
        assert np.abs(A[b].double().cpu().numpy() - ref["A"][:, :3, :]).max() <= 2e-6
        assert np.abs(posed[b].double().cpu().numpy() - ref["verts"]).max() <= 1e-5
        assert np.abs(joints[b].double().cpu().numpy() - (ref["G"][:, :3, 3] + trans[b])).max() <= 1e-5
    rest = inverse_lbs(dm, A, posed, _t(trans))
    # fp32 inverse of a blended transform: error scales with cond(T); these rigs stay below 3e-5 m
    assert float((rest - _t(rig["v_template"])[None]).abs().max()) <= 3e-5
    for b in (0, B - 1):
        ref = O.np_lbs_only(rig, pose[b], trans[b], ignore_joints=())
        want = O.np_inverse_lbs(rig["weights"], ref["A"], posed[b].double().cpu().numpy() - trans[b])
        assert np.abs(rest[b].double().cpu().numpy() - want).max() <= 3e-5
    jr = inverse_joints(A, joints, _t(trans))
    assert float((jr - _t(rig["J"])[None]).abs().max()) <= 1e-5


def test_inverse_lbs_with_smplh_weights_and_shape():
    """The un-posing of models/smpl_np.py:239-246: weights of the body model itself, betas != 0."""
    m = synthetic.make_model("smplh", seed=4)
    dm = smplk.DeviceModel(m, device=0)
    B = 3
    betas, pose, transl = synthetic.make_inputs(m, B, seed=8)
    A, _ = transforms(dm, _t(betas), _t(pose))
    v = body_model_apply(dm, _t(betas), _t(pose), transl=_t(transl))[0]
    rest = inverse_lbs(dm, A, v, _t(transl))
    om = O.TorchOracleModel(m, dtype=torch.float64)
    ref = om.forward_full_pose(*[torch.tensor(x, dtype=torch.float64) for x in (betas, pose, transl)])
    assert float((rest.double().cpu() - ref.v_posed).abs().max()) <= 3e-5


def test_vertex_normals_match_oracle():
    m = synthetic.make_model("smpl", seed=6)
    dm = smplk.DeviceModel(m, device=0)
    topo = MeshTopology(m["f"], dm.V)
    B = 4
    betas, pose, transl = synthetic.make_inputs(m, B, seed=2)
    v = body_model_apply(dm, _t(betas), _t(pose), transl=_t(transl))[0]
    n = topo.vertex_normals(v)
    vn = v.double().cpu().numpy()
    for b in range(B):
        ref = O.np_vertex_normals(vn[b], m["f"])
        # unit vectors from fp32 cross products of random (large) triangles
        assert np.abs(n[b].double().cpu().numpy() - ref).max() <= 2e-5
    used = np.zeros(dm.V, bool)
    used[np.asarray(m["f"]).ravel()] = True
    ln = n.norm(dim=2).cpu().numpy()
    assert np.abs(ln[:, used] - 1).max() <= 1e-5 and np.all(ln[:, ~used] == 0)


def test_divide_face_matches_reference_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "divide_face.npz"))
    m = synthetic.make_model("smplh", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    faces = np.asarray(m["f"])[:int(g["num_faces"])]
    topo = MeshTopology(faces, 6890)
    # (1) on the reference's own vertices: identical split and re-indexing as SMPLHModel.divide_face
    verts = _t(np.stack([g["verts0"], g["verts1"]]))
    res = topo.divide_face(verts)
    checked_against_reference = 0
    for i in range(2):
        ff, fv, fi, bf, bv, bi = res[i]
        v64 = g["verts%d" % i]
        # the kernel sees the float32 rounding of the reference's float64 vertices: bit-exact against
        # the oracle on those float32 values ...
        want = O.np_divide_face(verts[i].cpu().numpy().astype(np.float64), faces)
        for w, got in zip(want, (ff, fv, fi, bf, bv, bi)):
            assert np.array_equal(np.asarray(w), got.double().cpu().numpy() if got.is_floating_point() else got.cpu().numpy())
        # ... and against the reference's own output unless a triangle's z is within float32 rounding of 0
        m_, n_ = v64[faces[:, 1]] - v64[faces[:, 0]], v64[faces[:, 2]] - v64[faces[:, 1]]
        z = m_[:, 0] * n_[:, 1] - n_[:, 0] * m_[:, 1]
        if not np.any((np.abs(z) < 1e-6) & (z != 0)):      # exact zeros (degenerate faces) are front in both
            checked_against_reference += 1
            assert np.array_equal(ff.cpu().numpy(), g["front_face%d" % i])
            assert np.array_equal(fi.cpu().numpy(), g["front_index%d" % i])
            assert np.array_equal(bf.cpu().numpy(), g["back_face%d" % i])
            assert np.array_equal(bi.cpu().numpy(), g["back_index%d" % i])
        assert torch.equal(fv, verts[i][fi]) and torch.equal(bv, verts[i][bi])
    assert checked_against_reference >= 1
    # (2) full topology (13,776 faces), vertices from the CUDA forward, batch of 6: same as the oracle
    dm = smplk.DeviceModel(m, device=0)
    topo_full = MeshTopology(m["f"], 6890)
    betas, pose, transl = synthetic.make_inputs(m, 6, seed=12)
    v = body_model_apply(dm, _t(betas), _t(pose), transl=_t(transl))[0]
    res = topo_full.divide_face(v)
    vn = v.cpu().numpy().astype(np.float64)
    for b in (0, 5):
        want = O.np_divide_face(vn[b], m["f"])
        got = res[b]
        for w, gt in zip(want, got):
            assert np.array_equal(np.asarray(w), gt.cpu().numpy())


def test_twin_replays_reference_unposing_call_sequence(golden_dir):
    """The numpy twin offers what lib/mesh2smpl_model.py:146-207 reads from its `smpl` object (`.J`,
    `compute_R_G()`), and the un-posing kernels reproduce the reference's own outputs
    (tests/golden/unpose.npz: SMPLModel.inverse and RecoverModel.to_T_pose executed by
    oracle/make_golden_inverse.py).  fp32 tolerances: 1e-5 m forward, 3e-5 m through the inverse."""
    from smplk.mesh_ops import remove_rest
    g = np.load(os.path.join(golden_dir, "unpose.npz"))
    m = synthetic.make_model("smpl", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    smpl = smplk.SMPLModel(m, device=0)
    for i in range(g["inv_pose"].shape[0]):
        v = smpl.set_params(pose=g["inv_pose"][i].copy(), beta=g["inv_beta"][i].copy(), trans=g["inv_trans"][i].copy())
        assert np.abs(v[::53] - g["inv_posed_sub"][i]).max() <= 1e-5
        assert np.abs(smpl.J - g["inv_J"][i]).max() <= 1e-5           # models/smpl_np.py:169-170
        assert np.abs(smpl.R - g["inv_R"][i]).max() <= 2e-6
        assert np.abs(smpl.v_posed[::53] - g["inv_v_posed_sub"][i]).max() <= 1e-5
        G = smpl.compute_R_G()
        assert G.shape == (24, 4, 4) and np.abs(G - g["inv_G"][i]).max() <= 1e-5
        smpl.do_skinning(G)                                             # the second half of update()
        assert np.abs(smpl.verts[::53] - g["inv_posed_sub"][i]).max() <= 1e-5
        smpl.inverse()                                                  # models/smpl_np.py:239-246
        assert np.abs(smpl.verts - g["inv_unposed"][i]).max() <= 3e-5
    # ---- lib/mesh2smpl_model.py:183-207, line by line on the twin
    rig = synthetic.make_rigged_mesh(num_verts=int(g["tp_num_verts"]), seed=int(g["tp_rig_seed"]))
    W = np.asarray(rig["weights"], np.float64)
    rig["weights"] = W / W.sum(axis=1)[:, None]                         # :153
    rig_dm = smplk.DeviceModel(rig, device=0, lbs_only=True)
    smpl.set_params(g["tp_or_pose"].copy(), g["tp_or_shape"].copy())    # :188
    G = smpl.compute_R_G()                                              # :193
    assert np.abs(smpl.J - g["tp_smpl_J"]).max() <= 1e-5                # :197 reads smpl.J
    A = remove_rest(_t(G[None]), _t(smpl.J[None]))                      # :194-199
    v_template = inverse_lbs(rig_dm, A.reshape(1, 24, 12), _t(g["tp_or_verts"][None]))   # :200-203
    J = inverse_joints(A.reshape(1, 24, 12), _t(g["tp_or_J"][None]))    # :205-207
    assert np.abs(v_template[0].double().cpu().numpy() - g["tp_v_template"]).max() <= 3e-5
    assert np.abs(J[0].double().cpu().numpy() - g["tp_J"]).max() <= 1e-5


def test_twin_rejects_wrong_widths():
    """A 72-value clip row on a 52-joint model must raise, not read past the array."""
    m = synthetic.make_model("smplh", num_betas=10, seed=3)
    tw = smplk.SMPLHModel(m, device=0)
    with pytest.raises(ValueError):
        tw.forward_batch(np.zeros((4, 72)))
    with pytest.raises(ValueError):
        tw.forward_batch(np.zeros((4, 156)), betas=np.zeros((4, 16)).reshape(-1, 10)[:3])
    with pytest.raises(ValueError):
        tw.forward_batch(np.zeros((4, 156)), trans=np.zeros((3, 3)))
    with pytest.raises(ValueError):
        tw.set_params(pose=np.zeros((24, 3)))


def test_recover_model_twin_exposes_compute_R_G_and_do_skinning():
    """lib/model2video.py:55-81: the rigged-mesh twin's two halves of update() and the attributes they leave."""
    rig = synthetic.make_rigged_mesh(3001, seed=9)
    rm = smplk.RecoverModel(rig)
    rng = np.random.default_rng(3)
    pose = rng.standard_normal((24, 3)) * 0.4
    trans = rng.standard_normal(3)
    v = rm.set_params(pose=pose.copy(), trans=trans.copy()).copy()
    pose[[13, 14, 22, 23]] = 0.0                                   # set_params zeroes these joints (:44-45)
    ref = O.np_lbs_only(rig, pose, trans, ignore_joints=())
    assert np.abs(v - ref["verts"]).max() <= 1e-5
    G = rm.compute_R_G()
    assert G.shape == (24, 4, 4) and np.abs(G - ref["G"]).max() <= 1e-5
    assert np.abs(rm.R - O.np_rodrigues(pose)).max() <= 2e-6
    assert np.array_equal(rm.J, np.asarray(rig["J"]))              # fixed rest joints of the rig
    rm.do_skinning(G)
    assert np.abs(rm.verts - ref["verts"]).max() <= 1e-5
    G2 = G.copy()
    G2[:, :3, 3] += 0.25                                           # an edited G moves every vertex by the same offset
    rm.do_skinning(G2)
    assert np.abs(rm.verts - (ref["verts"] + 0.25)).max() <= 1e-5
