"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE's own numpy implementations.

Runs only in the build container (needs /root/reference). The reference has no tests or golden
outputs of its own, so these vectors are outputs of the unmodified reference
classes on seeded synthetic models + the real pose inputs shipped under /root/reference/data:

  * models/smplh_np.py  SMPLHModel.set_params / gen_J_3d / compute_R_G      (J=52)
  * models/smpl_np.py   SMPLModel.set_params / gen_J_3d                     (J=24)
  * lib/model2video.py  RecoverModel.set_params (LBS-only rigged mesh)      (J=24, Nv arbitrary)

Inputs (small) are stored in the fixture next to the outputs so the tests never read
/root/reference. Models are regenerated from their seed at test time; a checksum of the
tensors is stored to detect generator drift.

Usage:  python oracle/make_golden.py
"""
import importlib.util
import os
import pickle
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
import smplk  # noqa: E402
from smplk import synthetic  # noqa: E402


def _load_ref_module(name, relpath, stubs=()):
    for s in stubs:
        if s not in sys.modules:
            try:
                __import__(s)
            except Exception:
                sys.modules[s] = types.ModuleType(s)
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def model_checksum(m):
    acc = 0.0
    for k in ("v_template", "shapedirs", "posedirs", "weights"):
        if k in m:
            a = np.asarray(m[k], dtype=np.float64)
            acc += float(np.abs(a).sum()) + float(a.ravel()[::97].sum())
    return acc


def real_smplh_poses():
    """(pose(156), betas(10)) from data/tests/*/smplh.pkl and a few AMASS frames with hands."""
    poses, betas = [], []
    for t in ("test01", "test02", "test03_1024"):
        with open(os.path.join(REF, "data/tests", t, "smplh.pkl"), "rb") as f:
            p = pickle.load(f, encoding="latin1")
        poses.append(np.asarray(p["spmlh_pose"], np.float64).reshape(156))
        betas.append(np.asarray(p["spmlh_shape"], np.float64).reshape(10))
    d = np.load(os.path.join(REF, "data/amsass/09_05_poses.npz"))
    for fr in (0, 70, 142):
        poses.append(d["poses"][fr].astype(np.float64))
        betas.append(d["betas"][:10].astype(np.float64))
    trans = np.concatenate([np.zeros((3, 3)), (d["trans"] - d["trans"][0])[[0, 70, 142]]])
    return np.stack(poses), np.stack(betas), trans


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp()

    # ---------------- SMPL-H numpy twin ----------------
    smplh_np = _load_ref_module("ref_smplh_np", "models/smplh_np.py")
    mh = synthetic.make_model("smplh", num_betas=10, seed=7)
    ph = os.path.join(tmp, "smplh.pkl")
    with open(ph, "wb") as f:
        pickle.dump({k: mh[k] for k in ("J_regressor", "weights", "v_template", "shapedirs",
                                        "posedirs", "f", "kintree_table")}, f)
    ref = smplh_np.SMPLHModel(ph)
    poses, betas, trans = real_smplh_poses()
    rng = np.random.default_rng(11)
    # stress rows: zero pose, tiny pose, |theta| > pi
    extra_pose = np.stack([np.zeros(156), rng.standard_normal(156) * 1e-7,
                           rng.standard_normal(156) * 2.5])
    poses = np.concatenate([poses, extra_pose])
    betas = np.concatenate([betas, rng.standard_normal((3, 10))])
    trans = np.concatenate([trans, rng.standard_normal((3, 3))])
    verts, j3d, Gs, rest_verts = [], [], [], ref.verts.copy()
    for i in range(poses.shape[0]):
        v = ref.set_params(pose=poses[i].reshape(52, 3).copy(), beta=betas[i].copy(),
                           trans=trans[i].copy())
        verts.append(v.copy())
        j3d.append(np.asarray(ref.gen_J_3d()).copy())
        Gs.append(ref.compute_R_G().copy())
    np.savez_compressed(os.path.join(out_dir, "smplh_np_twin.npz"),
                        seed=7, num_betas=10, checksum=model_checksum(mh),
                        pose=poses, beta=betas, trans=trans,
                        verts=np.stack(verts).astype(np.float32),
                        verts_sub64=np.stack(verts)[:, ::53],  # exact f64 subsample
                        j3d=np.stack(j3d), G=np.stack(Gs), rest_verts_sub64=rest_verts[::53])
    print("smplh twin:", np.stack(verts).shape)

    # ---------------- SMPL numpy twin ----------------
    smpl_np = _load_ref_module("ref_smpl_np", "models/smpl_np.py",
                               stubs=("cv2", "transforms3d", "trimesh"))
    ms = synthetic.make_model("smpl", num_betas=10, seed=8)
    ps = os.path.join(tmp, "smpl.pkl")
    with open(ps, "wb") as f:
        pickle.dump({k: ms[k] for k in ("J_regressor", "weights", "v_template", "shapedirs",
                                        "posedirs", "f", "kintree_table")}, f)
    refs = smpl_np.SMPLModel(ps)
    mix = pickle.load(open(os.path.join(REF, "data/mixamo/0007/result.pkl"), "rb"),
                      encoding="latin1")
    sp = np.asarray(mix["smpl_array"], np.float64)[[0, 30, 60, 110]].reshape(4, 72)
    sp = np.concatenate([sp, poses[:3, :72], np.zeros((1, 72))])
    sb = np.concatenate([rng.standard_normal((7, 10)), np.zeros((1, 10))])
    st = rng.standard_normal((8, 3))
    sv, sj = [], []
    for i in range(sp.shape[0]):
        v = refs.set_params(pose=sp[i].reshape(24, 3).copy(), beta=sb[i].copy(), trans=st[i].copy())
        sv.append(v.copy())
        sj.append(np.asarray(refs.gen_J_3d()).copy())
    np.savez_compressed(os.path.join(out_dir, "smpl_np_twin.npz"),
                        seed=8, num_betas=10, checksum=model_checksum(ms),
                        pose=sp, beta=sb, trans=st,
                        verts=np.stack(sv).astype(np.float32),
                        verts_sub64=np.stack(sv)[:, ::53], j3d=np.stack(sj))
    print("smpl twin:", np.stack(sv).shape)

    # ---------------- LBS-only rigged mesh (RecoverModel) ----------------
    m2v = _load_ref_module("ref_model2video", "lib/model2video.py",
                           stubs=("cv2", "trimesh", "open3d"))
    rig = synthetic.make_rigged_mesh(num_verts=3001, seed=9)
    pr = os.path.join(tmp, "recover.pkl")
    with open(pr, "wb") as f:
        pickle.dump(rig, f)
    rm = m2v.RecoverModel(pr)
    d = np.load(os.path.join(REF, "data/amsass/35_01_poses.npz"))
    frames = [0, 100, 200, 357]
    rp = d["poses"][frames, :72].astype(np.float64)
    rt = (d["trans"] - d["trans"][0])[frames]
    rv = []
    for i in range(len(frames)):
        rv.append(rm.set_params(pose=rp[i].reshape(24, 3).copy(), trans=rt[i].copy()).copy())
    np.savez_compressed(os.path.join(out_dir, "recover_lbs.npz"),
                        seed=9, num_verts=3001, checksum=model_checksum(rig),
                        pose=rp, trans=rt, verts=np.stack(rv))
    print("recover lbs:", np.stack(rv).shape)

    # ---------------- Rodrigues, quaternion route (utils/geometry.py) ----------------
    try:
        import torch
        geo = _load_ref_module("ref_geometry", "utils/geometry.py")
        th = np.concatenate([rng.standard_normal((60, 3)) * 0.5, np.zeros((1, 3)),
                             rng.standard_normal((3, 3)) * 3.0]).astype(np.float64)
        Rq = geo.batch_rodrigues(torch.from_numpy(th)).numpy()
        np.savez_compressed(os.path.join(out_dir, "rodrigues_quat.npz"), theta=th, R=Rq)
        print("rodrigues:", Rq.shape)
    except Exception as e:  # pragma: no cover
        print("rodrigues golden skipped:", e)


if __name__ == "__main__":
    main()
