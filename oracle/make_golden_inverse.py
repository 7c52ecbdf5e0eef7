"""Generate tests/golden/unpose.npz by EXECUTING the reference's un-posing code:

  * models/smpl_np.py  SMPLModel.set_params -> do_skinning (stores T_inverse, :199) -> inverse (:239-246),
    together with the by-products compute_R_G leaves on the object (J, R, v_posed, G; :168-189);
  * lib/mesh2smpl_model.py  RecoverModel.to_T_pose (:183-207): the recovered mesh and its joints taken
    back to the T pose through the inverse of the blended SMPL transforms.

`RecoverModel.__init__` cannot run unmodified (np.int, removed from numpy 2; Replace_Hands / trimesh
imports): the object is made with __new__ and given exactly the attributes `to_T_pose` reads
(smpl, weigths, or_pose, or_shape, or_verts, or_J); `to_T_pose` itself is the reference's code.
`np.int = int` is shimmed in THIS generator only.  Runs only in the build container
(needs /root/reference); the tests never read /root/reference.

Usage:  python oracle/make_golden_inverse.py
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import _load_ref_module, model_checksum  # noqa: E402
from smplk import synthetic  # noqa: E402


def main():
    if not hasattr(np, "int"):
        np.int = int          # lib/mesh2smpl_model.py:154 (generator-only shim)
    smpl_np = _load_ref_module("ref_smpl_np", "models/smpl_np.py", stubs=("cv2", "transforms3d", "trimesh"))
    lib_pkg = types.ModuleType("lib")
    lib_pkg.Replace_Hands = types.ModuleType("lib.Replace_Hands")
    sys.modules.setdefault("lib", lib_pkg)
    sys.modules.setdefault("lib.Replace_Hands", lib_pkg.Replace_Hands)
    m2s = _load_ref_module("ref_mesh2smpl", "lib/mesh2smpl_model.py", stubs=("trimesh",))

    ms = synthetic.make_model("smpl", num_betas=10, seed=8)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "smpl.pkl")
    with open(path, "wb") as f:
        pickle.dump({k: ms[k] for k in ("J_regressor", "weights", "v_template", "shapedirs",
                                        "posedirs", "f", "kintree_table")}, f)
    rng = np.random.default_rng(31)
    out = dict(seed=8, num_betas=10, checksum=model_checksum(ms))

    # ---- SMPLModel.inverse (models/smpl_np.py:239-246) and the compute_R_G by-products
    smpl = smpl_np.SMPLModel(path)
    n = 3
    poses = rng.standard_normal((n, 24, 3)) * 0.4
    poses[1] *= 2.0
    betas = rng.standard_normal((n, 10))
    trans = rng.standard_normal((n, 3))
    posed, unposed, Js, Rs, Gs, vps = [], [], [], [], [], []
    for i in range(n):
        posed.append(smpl.set_params(pose=poses[i].copy(), beta=betas[i].copy(), trans=trans[i].copy()).copy())
        Js.append(smpl.J.copy()); Rs.append(smpl.R.copy()); vps.append(smpl.v_posed.copy())
        Gs.append(smpl.compute_R_G().copy())
        smpl.inverse()
        unposed.append(smpl.verts.copy())
    out.update(inv_pose=poses, inv_beta=betas, inv_trans=trans, inv_posed_sub=np.stack(posed)[:, ::53],
               inv_unposed=np.stack(unposed).astype(np.float32), inv_unposed_sub=np.stack(unposed)[:, ::53], inv_J=np.stack(Js), inv_R=np.stack(Rs), inv_G=np.stack(Gs),
               inv_v_posed_sub=np.stack(vps)[:, ::53])
    print("inverse: max |unposed - v_posed| =", float(np.abs(np.stack(unposed) - np.stack(vps)).max()))

    # ---- RecoverModel.to_T_pose (lib/mesh2smpl_model.py:183-207)
    nv = 2003
    rig = synthetic.make_rigged_mesh(num_verts=nv, seed=12)
    W = np.asarray(rig["weights"], np.float64)
    rm = m2s.RecoverModel.__new__(m2s.RecoverModel)
    rm.smpl = smpl_np.SMPLModel(path)
    rm.weigths = W / W.sum(axis=1)[:, None]
    rm.or_pose = rng.standard_normal((24, 3)) * 0.35
    rm.or_shape = rng.standard_normal(10)
    rm.or_verts = rng.standard_normal((nv, 3)) * np.array([0.3, 0.5, 0.1])
    rm.or_J = rng.standard_normal((24, 3)) * np.array([0.3, 0.5, 0.1])
    rm.to_T_pose()
    out.update(tp_num_verts=nv, tp_rig_seed=12, tp_weights_checksum=float(np.abs(rm.weigths).sum()),
               tp_or_pose=rm.or_pose, tp_or_shape=rm.or_shape, tp_or_verts=rm.or_verts, tp_or_J=rm.or_J,
               tp_smpl_J=rm.smpl.J.copy(), tp_v_template=rm.v_template, tp_J=rm.J)
    print("to_T_pose:", rm.v_template.shape, rm.J.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "unpose.npz"), **out)


if __name__ == "__main__":
    main()
