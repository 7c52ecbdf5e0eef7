"""CPU: the oracle restatements reproduce the outputs of the reference's own numpy classes
(fixtures made by oracle/make_golden.py from /root/reference) and agree with each other."""
import os

import numpy as np
import pytest
import torch

import smplk
from smplk import synthetic
from oracle import smpl_oracle as O
from oracle.make_golden import model_checksum


@pytest.fixture(scope="module")
def smplh_fix(golden_dir):
    g = np.load(os.path.join(golden_dir, "smplh_np_twin.npz"))
    m = synthetic.make_model("smplh", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    assert abs(model_checksum(m) - float(g["checksum"])) < 1e-9, "synthetic generator drifted"
    return g, m


@pytest.fixture(scope="module")
def smpl_fix(golden_dir):
    g = np.load(os.path.join(golden_dir, "smpl_np_twin.npz"))
    m = synthetic.make_model("smpl", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    assert abs(model_checksum(m) - float(g["checksum"])) < 1e-9
    return g, m


def test_numpy_oracle_matches_reference_smplh(smplh_fix):
    g, m = smplh_fix
    for i in range(g["pose"].shape[0]):
        r = O.np_forward(m, g["pose"][i], g["beta"][i], g["trans"][i])
        assert np.abs(r["verts"][::53] - g["verts_sub64"][i]).max() < 1e-12
        assert np.abs(r["verts"] - g["verts"][i]).max() < 1e-6          # f32-stored full field
        assert np.abs(r["G"] - g["G"][i]).max() < 1e-12
        assert np.abs(O.np_gen_J_3d(m, r["verts"]) - g["j3d"][i]).max() < 1e-12
    rest = O.np_forward(m)["verts"]
    assert np.abs(rest[::53] - g["rest_verts_sub64"]).max() < 1e-12


def test_numpy_oracle_matches_reference_smpl(smpl_fix):
    g, m = smpl_fix
    for i in range(g["pose"].shape[0]):
        r = O.np_forward(m, g["pose"][i], g["beta"][i], g["trans"][i])
        assert np.abs(r["verts"][::53] - g["verts_sub64"][i]).max() < 1e-12
        assert np.abs(O.np_gen_J_3d(m, r["verts"]) - g["j3d"][i]).max() < 1e-12


def test_lbs_only_oracle_matches_reference_recovermodel(golden_dir):
    g = np.load(os.path.join(golden_dir, "recover_lbs.npz"))
    rig = synthetic.make_rigged_mesh(int(g["num_verts"]), seed=int(g["seed"]))
    assert abs(model_checksum(rig) - float(g["checksum"])) < 1e-9
    for i in range(g["pose"].shape[0]):
        v = O.np_lbs_only(rig, g["pose"][i], g["trans"][i])["verts"]
        assert np.abs(v - g["verts"][i]).max() < 1e-12


def test_rodrigues_forms_agree(golden_dir):
    g = np.load(os.path.join(golden_dir, "rodrigues_quat.npz"))
    th = torch.tensor(g["theta"])
    assert np.abs(O.torch_rodrigues_quat(th).numpy() - g["R"]).max() < 1e-12   # utils/geometry.py
    assert np.abs(O.torch_rodrigues(th).numpy() - g["R"]).max() < 5e-8          # upstream K-route (eps)
    assert np.abs(O.np_rodrigues(g["theta"]) - g["R"]).max() < 5e-8             # numpy twin
    assert np.abs(O.np_rodrigues(np.zeros((1, 3)))[0] - np.eye(3)).max() == 0.0


def test_torch_restatement_matches_numpy_twin(smplh_fix):
    g, m = smplh_fix
    om = O.TorchOracleModel(m, dtype=torch.float64)
    out = om.forward_full_pose(torch.tensor(g["beta"]), torch.tensor(g["pose"]), torch.tensor(g["trans"]))
    assert np.abs(out.vertices.numpy()[:, ::53] - g["verts_sub64"]).max() < 1e-7   # 1e-8 eps in Rodrigues
    om32 = O.TorchOracleModel(m, dtype=torch.float32)
    out32 = om32.forward_full_pose(torch.tensor(g["beta"]).float(), torch.tensor(g["pose"]).float(),
                                   torch.tensor(g["trans"]).float())
    assert np.abs(out32.vertices.numpy() - g["verts"]).max() < 1e-5


def test_torch_oracle_pca_joints_and_broadcast():
    m = synthetic.make_model("smplh", seed=2)
    om = O.TorchOracleModel(m, dtype=torch.float64, num_pca_comps=12, joint_mapper=[52, 12, 0, 60, 72])
    B = 3
    rng = np.random.default_rng(0)
    t = lambda *s: torch.tensor(rng.standard_normal(s) * 0.3)
    out = om.forward(t(1, 16), t(B, 3), t(B, 63), t(B, 12), t(B, 12), transl=t(B, 3), wrapper_extra=True)
    assert out.vertices.shape == (B, 6890, 3)
    assert out.joints.shape == (B, 5 + 9, 3)
    assert out.full_pose.shape == (B, 156)
    # vertex pick 52 is extra vertex id 0 of the posed, translated mesh
    vid = int(m["extra_vertex_ids"][0])
    assert torch.allclose(out.joints[:, 0], out.vertices[:, vid])


def test_inverse_lbs_roundtrip():
    rig = synthetic.make_rigged_mesh(500, seed=5)
    rng = np.random.default_rng(1)
    pose = rng.standard_normal(72) * 0.3
    r = O.np_lbs_only(rig, pose, None, ignore_joints=())
    back = O.np_inverse_lbs(rig["weights"], r["A"], r["verts"])
    assert np.abs(back - rig["v_template"]).max() < 1e-9


def test_oracle_gradients_match_finite_differences():
    m = synthetic.make_model("smpl", seed=3, num_verts=300)
    m["extra_vertex_ids"] = None
    om = O.TorchOracleModel(m, dtype=torch.float64)
    rng = np.random.default_rng(2)
    betas = torch.tensor(rng.standard_normal((2, 10)))
    pose = torch.tensor(rng.standard_normal((2, 72)) * 0.3)
    transl = torch.tensor(rng.standard_normal((2, 3)))
    tgt = om.forward_full_pose(betas * 0.5, pose * 0.5, transl).vertices
    loss, gb, gp, gt = O.torch_vertex_l2_grads(om, betas, pose, transl, tgt)
    f = lambda p: float(((om.forward_full_pose(betas, p, transl).vertices - tgt) ** 2).sum())
    for idx in [(0, 0), (1, 5), (0, 40), (1, 71)]:
        e = torch.zeros_like(pose)
        e[idx] = 1e-6
        fd = (f(pose + e) - f(pose - e)) / 2e-6
        assert abs(fd - float(gp[idx])) < 1e-5 * max(1.0, abs(fd))


def test_divide_face_oracle_matches_reference(golden_dir):
    """oracle np_divide_face == SMPLHModel.divide_face of the reference (models/smplh_np.py:126-182),
    on the reference's own float64 vertices: same faces per side, same first-appearance order."""
    g = np.load(os.path.join(golden_dir, "divide_face.npz"))
    m = synthetic.make_model("smplh", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    assert abs(model_checksum(m) - float(g["checksum"])) < 1e-9
    faces = np.asarray(m["f"])[:int(g["num_faces"])]
    for i in range(2):
        verts = g["verts%d" % i]
        # the vertices themselves are the reference forward's
        r = O.np_forward(m, g["pose"][i].reshape(-1), g["beta"][i], g["trans"][i])
        assert np.abs(r["verts"] - verts).max() < 1e-12
        ff, fv, fi, bf, bv, bi = O.np_divide_face(verts, faces)
        assert np.array_equal(ff, g["front_face%d" % i]) and np.array_equal(fi, g["front_index%d" % i])
        assert np.array_equal(bf, g["back_face%d" % i]) and np.array_equal(bi, g["back_index%d" % i])
        assert np.array_equal(fv, verts[fi]) and np.array_equal(bv, verts[bi])
        assert len(ff) + len(bf) == len(faces)


def test_vertex_normals_and_inverse_joints_oracle_properties():
    rng = np.random.default_rng(4)
    m = synthetic.make_model("smpl", seed=3)
    r = O.np_forward(m, rng.standard_normal(72) * 0.3, rng.standard_normal(10), rng.standard_normal(3))
    faces = np.asarray(m["f"])
    n = O.np_vertex_normals(r["verts"], faces)
    used = np.zeros(len(n), bool)
    used[faces.ravel()] = True
    assert np.abs(np.linalg.norm(n[used], axis=1) - 1).max() < 1e-12 and np.all(n[~used] == 0)
    # brute force for a few vertices
    for v in np.flatnonzero(used)[:20]:
        acc = np.zeros(3)
        for f in faces[(faces == v).any(1)]:
            acc += np.cross(r["verts"][f[1]] - r["verts"][f[0]], r["verts"][f[2]] - r["verts"][f[0]]) * (f == v).sum()
        assert np.abs(acc / np.linalg.norm(acc) - n[v]).max() < 1e-12
    # a flipped face flips its contribution; a rigid rotation rotates the normals
    Rz = O.np_rodrigues(np.array([[0.3, -0.2, 0.5]]))[0]
    assert np.abs(O.np_vertex_normals(r["verts"] @ Rz.T, faces) - n @ Rz.T).max() < 1e-10
    # inverse joints: A_j^-1 maps the posed joint back to the rest joint
    posed_J = r["G"][:, :3, 3]
    back = O.np_inverse_joints(r["A"][:, :3, :], posed_J)
    assert np.abs(back - r["J"]).max() < 1e-10


def test_unposing_oracle_matches_reference_inverse_and_to_T_pose(golden_dir):
    """SURVEY 8f row 2, pinned by reference execution (oracle/make_golden_inverse.py):
    models/smpl_np.py SMPLModel.inverse (:239-246) with the compute_R_G by-products (J, R, v_posed, G),
    and lib/mesh2smpl_model.py RecoverModel.to_T_pose (:183-207)."""
    g = np.load(os.path.join(golden_dir, "unpose.npz"))
    m = synthetic.make_model("smpl", num_betas=int(g["num_betas"]), seed=int(g["seed"]))
    assert abs(model_checksum(m) - float(g["checksum"])) < 1e-9
    for i in range(g["inv_pose"].shape[0]):
        r = O.np_forward(m, g["inv_pose"][i], g["inv_beta"][i], g["inv_trans"][i])
        assert np.abs(r["J"] - g["inv_J"][i]).max() < 1e-12
        assert np.abs(r["R"] - g["inv_R"][i]).max() < 1e-12
        assert np.abs(r["G"] - g["inv_G"][i]).max() < 1e-12
        assert np.abs(r["v_posed"][::53] - g["inv_v_posed_sub"][i]).max() < 1e-12
        assert np.abs(r["verts"][::53] - g["inv_posed_sub"][i]).max() < 1e-12
        un = O.np_inverse_lbs(m["weights"], r["A"], r["verts"] - g["inv_trans"][i][None])
        assert np.abs(un[::53] - g["inv_unposed_sub"][i]).max() < 1e-11
        assert np.abs(un - g["inv_unposed"][i]).max() < 1e-6          # f32-stored full field
    # to_T_pose: SMPL transforms of (or_pose, or_shape) blended with the RECOVERED mesh's weights
    rig = synthetic.make_rigged_mesh(num_verts=int(g["tp_num_verts"]), seed=int(g["tp_rig_seed"]))
    W = np.asarray(rig["weights"], np.float64)
    W = W / W.sum(axis=1)[:, None]
    assert abs(float(np.abs(W).sum()) - float(g["tp_weights_checksum"])) < 1e-9
    r = O.np_forward(m, g["tp_or_pose"], g["tp_or_shape"])
    assert np.abs(r["J"] - g["tp_smpl_J"]).max() < 1e-12
    assert np.abs(O.np_inverse_lbs(W, r["A"], g["tp_or_verts"]) - g["tp_v_template"]).max() < 1e-10
    assert np.abs(O.np_inverse_joints(r["A"], g["tp_or_J"]) - g["tp_J"]).max() < 1e-10
