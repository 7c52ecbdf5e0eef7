"""Generate tests/golden/divide_face.npz by EXECUTING the reference's SMPLHModel.divide_face
(models/smplh_np.py:126-182) on a seeded synthetic model (first 2,000 faces of the synthetic face
list, to keep the fixture small) for two posed bodies.  Runs only in the build container
(needs /root/reference); the tests never read /root/reference.

Usage:  python oracle/make_golden_mesh_ops.py
"""
import contextlib
import io
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import _load_ref_module, model_checksum  # noqa: E402
from smplk import synthetic  # noqa: E402


def main():
    smplh_np = _load_ref_module("ref_smplh_np", "models/smplh_np.py")
    m = synthetic.make_model("smplh", num_betas=10, seed=7)
    faces = np.asarray(m["f"])[:2000].copy()
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "smplh.pkl")
    with open(path, "wb") as f:
        d = {k: m[k] for k in ("J_regressor", "weights", "v_template", "shapedirs", "posedirs", "kintree_table")}
        d["f"] = faces
        pickle.dump(d, f)
    ref = smplh_np.SMPLHModel(path)
    rng = np.random.default_rng(21)
    poses = rng.standard_normal((2, 52, 3)) * 0.4
    betas = rng.standard_normal((2, 10))
    trans = rng.standard_normal((2, 3))
    out = dict(seed=7, num_betas=10, num_faces=2000, checksum=model_checksum(m), pose=poses, beta=betas, trans=trans)
    for i in range(2):
        ref.set_params(pose=poses[i].copy(), beta=betas[i].copy(), trans=trans[i].copy())
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints while it walks the faces
            ff, fv, fi, bf, bv, bi = ref.divide_face()
        out["verts%d" % i] = ref.verts.copy()
        out["front_face%d" % i] = np.asarray(ff, np.int32)
        out["front_index%d" % i] = np.asarray(fi, np.int32)
        out["back_face%d" % i] = np.asarray(bf, np.int32)
        out["back_index%d" % i] = np.asarray(bi, np.int32)
        assert np.array_equal(fv, ref.verts[fi]) and np.array_equal(bv, ref.verts[bi])
        print("body", i, "front faces", len(ff), "front verts", len(fi), "back faces", len(bf), "back verts", len(bi))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "divide_face.npz"), **out)


if __name__ == "__main__":
    main()
