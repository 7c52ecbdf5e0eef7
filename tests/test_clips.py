"""Clip / fit-result ingestion: the readers return what the reference's readers
return for files in the reference's formats (fixtures cut from the reference's data by
oracle/make_golden_clips.py); the GPU test replays a clip through the CUDA path."""
import os

import numpy as np
import pytest

from smplk import clips, synthetic
from oracle import smpl_oracle as O


def test_readers_match_reference_readers(golden_dir):
    exp = np.load(os.path.join(golden_dir, "clip_expected.npz"))
    c = clips.read_amsass(os.path.join(golden_dir, "amass_clip_fixture.npz"))
    assert np.array_equal(c.poses, exp["amass_poses"]) and np.array_equal(c.trans, exp["amass_trans"])
    assert c.poses.shape == (5, 72) and np.all(c.trans[0] == 0)
    cf = clips.read_amsass(os.path.join(golden_dir, "amass_clip_fixture.npz"), full=True)
    assert cf.poses.shape == (5, 156) and cf.betas.shape == (16,)
    m = clips.read_mixamo(os.path.join(golden_dir, "mixamo_result_fixture.pkl"))
    assert np.array_equal(m.poses, exp["mixamo_pose"]) and m.extra["cam_array"].shape == (4, 3)
    f = clips.read_fit_result(os.path.join(golden_dir, "fit_result_fixture.pkl"))
    assert np.array_equal(f.pose, exp["fit_pose"]) and np.array_equal(f.betas, exp["fit_shape"])
    assert f.pose.dtype == np.float64 and f.camera_rotation.shape == (3, 3)


def test_pack_clip_pads_truncates_and_zeroes_ignored_joints(golden_dir):
    c = clips.read_amsass(os.path.join(golden_dir, "amass_clip_fixture.npz"), full=True)
    p52 = clips.pack_clip(c, 52)
    assert p52.shape == (5, 156) and p52.dtype == np.float32
    assert np.array_equal(p52, c.poses.astype(np.float32))
    p24 = clips.pack_clip(c, 24, ignore_joints=(13, 14, 22, 23))
    ref = c.poses[:, :72].astype(np.float32).reshape(5, 24, 3).copy()
    ref[:, [13, 14, 22, 23]] = 0
    assert np.array_equal(p24, ref.reshape(5, 72))
    m = clips.read_mixamo(os.path.join(golden_dir, "mixamo_result_fixture.pkl"))
    assert clips.pack_clip(m, 52)[:, 72:].max() == 0          # 24-joint clip on the 52-joint skeleton


def test_load_model_roundtrip(tmp_path):
    import pickle
    m = synthetic.make_model("smpl", seed=1)
    p = tmp_path / "m.pkl"
    with open(p, "wb") as f:
        pickle.dump({k: m[k] for k in ("J_regressor", "weights", "v_template", "shapedirs", "posedirs", "f", "kintree_table")}, f)
    q = tmp_path / "m.npz"
    np.savez(q, **{k: m[k] for k in ("J_regressor", "weights", "v_template", "shapedirs", "posedirs", "f", "kintree_table")})
    for path in (p, q):
        mm = clips.load_model(str(path))
        assert np.array_equal(mm["posedirs"], m["posedirs"]) and np.array_equal(mm["kintree_table"], m["kintree_table"])


@pytest.mark.gpu
def test_amass_clip_replays_through_the_cuda_path(golden_dir):
    import torch
    import smplk
    from smplk.body_models import body_model_apply
    c = clips.read_amsass(os.path.join(golden_dir, "amass_clip_fixture.npz"), full=True)
    m = synthetic.make_model("smplh", seed=0)
    dm = smplk.DeviceModel(m, device=0)
    pose, transl, betas = clips.clip_to_device(c, dm.J)
    v, j, _, _ = body_model_apply(dm, betas, pose, transl=transl)
    ref = O.TorchOracleModel(m, dtype=torch.float64).forward_full_pose(
        torch.tensor(np.repeat(c.betas[None], 5, 0)), torch.tensor(c.poses), torch.tensor(c.trans))
    assert float((v.double().cpu() - ref.vertices).abs().max()) <= 1e-5
    assert float((j.double().cpu() - ref.joints).abs().max()) <= 1e-5
    # LBS-only replay of the 24-joint part on a rigged mesh (lib/model2video.py:514-518)
    rig = synthetic.make_rigged_mesh(4001, seed=5)
    rm = smplk.RecoverModel(rig)
    out = rm.replay(c.poses[:, :72], c.trans)
    for i in (0, 4):
        want = O.np_lbs_only(rig, c.poses[i, :72], c.trans[i])["verts"]
        assert np.abs(out[i] - want).max() <= 1e-5
