"""BASELINE.json configs[0] and SURVEY 8(d) CPU-baseline item (i): time the REFERENCE's own numpy body
models (models/smpl_np.py SMPLModel, models/smplh_np.py SMPLHModel; float64, one body per call) on this
machine's CPU.  Needs /root/reference, so it only runs in the build container; the result is recorded
in profiles/r01_reference_cpu_timing.json.   Usage: python tools/ref_cpu_timing.py
"""
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import _load_ref_module  # noqa: E402
from smplk import synthetic  # noqa: E402


def bench(fn, n=30):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    tmp = tempfile.mkdtemp()
    keys = ("J_regressor", "weights", "v_template", "shapedirs", "posedirs", "f", "kintree_table")
    out = {"cpu_count": os.cpu_count(), "numpy": np.__version__}
    rng = np.random.default_rng(0)
    smpl_np = _load_ref_module("ref_smpl_np", "models/smpl_np.py", stubs=("cv2", "transforms3d", "trimesh"))
    m = synthetic.make_model("smpl", num_betas=10, seed=8)
    p = os.path.join(tmp, "smpl.pkl")
    pickle.dump({k: m[k] for k in keys}, open(p, "wb"))
    ref = smpl_np.SMPLModel(p)
    pose, beta, trans = rng.standard_normal((24, 3)) * 0.3, rng.standard_normal(10), rng.standard_normal(3)
    out["smpl_np_set_params_ms"] = bench(lambda: ref.set_params(pose=pose.copy(), beta=beta.copy(), trans=trans.copy()))
    smplh_np = _load_ref_module("ref_smplh_np", "models/smplh_np.py")
    mh = synthetic.make_model("smplh", num_betas=10, seed=7)
    ph = os.path.join(tmp, "smplh.pkl")
    pickle.dump({k: mh[k] for k in keys}, open(ph, "wb"))
    refh = smplh_np.SMPLHModel(ph)
    poseh = rng.standard_normal((52, 3)) * 0.3
    out["smplh_np_set_params_ms"] = bench(lambda: refh.set_params(pose=poseh.copy(), beta=beta.copy(), trans=trans.copy()))
    out["smpl_np_meshes_per_s"] = 1e3 / out["smpl_np_set_params_ms"]
    out["smplh_np_meshes_per_s"] = 1e3 / out["smplh_np_set_params_ms"]
    out["note"] = ("reference numpy classes, float64, one body per call, median of 30 calls on the build container's CPU "
                   "(numpy BLAS threads as configured); synthetic canonical-shape model tensors")
    print(json.dumps(out, indent=1))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r01_reference_cpu_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
