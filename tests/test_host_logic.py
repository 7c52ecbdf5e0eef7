"""CPU: the C-ABI library loads and exports every declared symbol, host-side packing helpers and
error behaviour without a GPU, synthetic generator invariants, header/binding consistency."""
import ctypes
import os
import re

import numpy as np
import pytest

import smplk
from smplk import _lib, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_header_symbol():
    lib = smplk.load()
    hdr = open(os.path.join(ROOT, "include", "smplk.h")).read()
    declared = set(re.findall(r"\b(smplk_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smplk_version() == 2


def test_header_is_plain_c():
    import subprocess, tempfile
    src = '#include "smplk.h"\nint main(void){ smplk_forward_args a; (void)a; return smplk_version() < 0; }\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                            "-c", p, "-o", os.path.join(d, "t.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_struct_sizes_match_header_layout():
    # pointers and int32 interleaved exactly as in the header (natural alignment)
    assert ctypes.sizeof(_lib.ForwardArgs) == 8 + 8 + 8 + 8 * 8 + 8 + 8 + 8 or ctypes.sizeof(_lib.ForwardArgs) % 8 == 0
    assert _lib.ModelInfo._fields_[0][0] == "num_verts" and len(_lib.ModelInfo._fields_) == 11


def test_ctypes_structs_match_the_compiled_header_layout():
    """sizeof / offsetof of every argument struct as gcc lays out include/smplk.h == the ctypes mirror."""
    import subprocess, tempfile
    pairs = [("smplk_model_desc", _lib.ModelDesc), ("smplk_forward_args", _lib.ForwardArgs),
             ("smplk_backward_args", _lib.BackwardArgs), ("smplk_reprojection_args", _lib.ReprojArgs),
             ("smplk_prior_args", _lib.PriorArgs), ("smplk_model_info", _lib.ModelInfo)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smplk.h"', 'int main(void){']
    for cname, cls in pairs:
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('printf("\\n");')
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write("\n".join(lines))
        exe = os.path.join(d, "t")
        r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), p, "-o", exe],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = subprocess.run([exe], capture_output=True, text=True).stdout.strip().splitlines()
    for (cname, cls), line in zip(pairs, out):
        nums = [int(x) for x in line.split()[1:]]
        assert nums[0] == ctypes.sizeof(cls), cname
        assert nums[1:] == [getattr(cls, f).offset for f, _ in cls._fields_], cname


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = synthetic.make_model("smpl", seed=1, num_verts=200)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        smplk.DeviceModel(m)
    mod = smplk.SMPL(model=m)
    with pytest.raises(RuntimeError):
        mod()


def test_bad_descriptor_arguments_are_rejected_before_touching_the_gpu():
    lib = smplk.load()
    h = ctypes.c_void_p()
    d = _lib.ModelDesc()
    d.num_joints, d.num_verts = 100, 10
    rc = lib.smplk_model_create(ctypes.byref(d), 0, ctypes.byref(h))
    assert rc == -2 and b"num_joints" in lib.smplk_last_error_string()
    d.num_joints = 24
    rc = lib.smplk_model_create(ctypes.byref(d), 0, ctypes.byref(h))
    assert rc == -1
    assert lib.smplk_model_create(None, 0, ctypes.byref(h)) == -1
    assert lib.smplk_workspace_bytes(None, 4, 0) == 0


def test_parent_tables():
    for kind, J in (("smpl", 24), ("smplh", 52)):
        m = synthetic.make_model(kind, seed=0, num_verts=100)
        p = _lib.parents_from_model(m)
        assert p.shape == (J,) and p[0] == -1 and all(p[i] < i for i in range(1, J))
        assert list(p) == list(synthetic.parents_from_kintree(m["kintree_table"]))
    rig = synthetic.make_rigged_mesh(50)
    assert list(_lib.parents_from_model(rig)) == synthetic.SMPL_PARENTS
    # canonical depths quoted in SURVEY.md H4
    def depths(par):
        d = [0] * len(par)
        for i in range(1, len(par)):
            d[i] = d[par[i]] + 1
        return d
    assert max(depths(synthetic.SMPL_PARENTS)) == 8
    assert max(depths(synthetic.SMPLH_PARENTS)) == 10


def test_synthetic_model_invariants():
    m = synthetic.make_model("smplh", seed=1)
    assert m["v_template"].shape == (6890, 3) and m["shapedirs"].shape == (6890, 3, 16)
    assert m["posedirs"].shape == (6890, 3, 459) and m["J_regressor"].shape == (52, 6890)
    W = m["weights"]
    assert np.allclose(W.sum(1), 1) and (W >= 0).all() and ((W != 0).sum(1) <= 4).all()
    assert np.allclose(m["J_regressor"].sum(1), 1)
    q = m["hands_componentsl"]
    assert np.allclose(q @ q.T, np.eye(45), atol=1e-10)
    d = synthetic.make_model("smpl", seed=1, dense_weights=True, num_verts=64)
    assert ((d["weights"] != 0).sum(1) == 24).all()
    b, p, t = synthetic.make_inputs(m, 7, broadcast_betas=True)
    assert b.shape == (1, 16) and p.shape == (7, 156) and t.shape == (7, 3)


def test_module_signatures_mirror_reference():
    import inspect
    from smplk.body_models import SMPL, SMPLH, ModelOutput
    sig = inspect.signature(SMPLH.forward).parameters
    for name in ("betas", "global_orient", "body_pose", "left_hand_pose", "right_hand_pose", "transl",
                 "return_verts", "return_full_pose", "pose2rot"):
        assert name in sig
    assert ModelOutput._fields[:6] == ("vertices", "joints", "full_pose", "betas", "global_orient", "body_pose")
    m = synthetic.make_model("smplh", seed=1, num_verts=128)
    mod = SMPLH(model=m, use_pca=True, num_pca_comps=12, batch_size=2, create_transl=False)
    names = dict(mod.named_parameters())
    assert names["left_hand_pose"].shape == (2, 12) and names["body_pose"].shape == (2, 63)
    assert "transl" not in names and names["betas"].shape == (2, 16)
    mod.reset_params(betas=np.ones((2, 16)))
    assert float(mod.betas.sum()) == 32.0 and float(mod.body_pose.abs().sum()) == 0.0
    s = SMPL(model=synthetic.make_model("smpl", seed=1, num_verts=128), create_body_pose=False)
    assert "body_pose" not in dict(s.named_parameters())
    from smplk import np_twins
    for cls in (np_twins.SMPLHModel, np_twins.SMPLModel, np_twins.RecoverModel):
        assert list(inspect.signature(cls.set_params).parameters)[1:] == ["pose", "beta", "trans"]


def test_flag_constants_match_the_header():
    """Every SMPLK_FLAG_* of include/smplk.h has the same value in the ctypes mirror, and vice versa."""
    import re
    text = open(os.path.join(ROOT, "include", "smplk.h")).read()
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SMPLK_FLAG_(\w+)\s+(\d+)u", text)}
    mirror = {k[len("FLAG_"):]: v for k, v in vars(_lib).items() if k.startswith("FLAG_")}
    assert header == mirror, (header, mirror)
    vals = sorted(header.values())
    assert vals == [1 << i for i in range(len(vals))]          # distinct single bits
    slots = int(re.search(r"#define\s+SMPLK_PROF_SLOTS\s+(\d+)", text).group(1))
    assert slots == len(_lib.PROF_SLOTS)


def test_graphed_closure_refuses_cpu_parameters():
    """No CPU path anywhere: the graph wrapper fails loudly on CPU parameters instead of running eagerly."""
    import pytest
    import torch
    from smplk.fitting import GraphedClosure
    p = torch.nn.Parameter(torch.zeros(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedClosure(lambda: (p ** 2).sum(), [p])
    with pytest.raises(ValueError):
        GraphedClosure(lambda: None, [torch.zeros(3)])


def test_create_resolves_model_folders(tmp_path):
    """`smplx.create(model_path=<folder>, model_type=..., gender=...)` as lib/gen_smplh.py:75-90 calls it:
    <folder>/<type>/<TYPE>_<GENDER>.pkl (or the file directly inside <folder>, or an .npz)."""
    import pickle
    import numpy as np
    from smplk import synthetic
    from smplk.body_models import SMPL, SMPLH, create, resolve_model_path
    keys = ("J_regressor", "weights", "v_template", "shapedirs", "posedirs", "f", "kintree_table",
            "hands_componentsl", "hands_componentsr", "hands_meanl", "hands_meanr")
    mh = synthetic.make_model("smplh", seed=1)
    ms = synthetic.make_model("smpl", seed=2)
    (tmp_path / "smplh").mkdir()
    (tmp_path / "smpl").mkdir()
    with open(tmp_path / "smplh" / "SMPLH_MALE.pkl", "wb") as f:
        pickle.dump({k: mh[k] for k in keys}, f)
    np.savez(tmp_path / "smpl" / "SMPL_NEUTRAL.npz", **{k: ms[k] for k in keys[:7]})
    m = create(model_path=str(tmp_path), model_type="smplh", gender="male", use_pca=True, num_pca_comps=12,
               create_transl=False, batch_size=2)
    assert isinstance(m, SMPLH) and m.num_joints == 52 and m.betas.shape == (2, 16)
    assert np.array_equal(np.asarray(m._model_dict["v_template"]), mh["v_template"])
    s = create(model_path=str(tmp_path), model_type="smpl", gender="neutral")          # .npz next to a missing .pkl
    assert isinstance(s, SMPL) and s.num_joints == 24
    assert resolve_model_path(str(tmp_path / "smplh"), "smplh", "male").endswith("SMPLH_MALE.pkl")
    direct = SMPLH(model_path=str(tmp_path / "smplh" / "SMPLH_MALE.pkl"), num_pca_comps=6)   # a file path is taken as is
    assert direct.left_hand_pose.shape == (1, 6)
    with pytest.raises(FileNotFoundError):
        create(model_path=str(tmp_path), model_type="smplh", gender="female")
    with pytest.raises(ValueError):
        create(model_path=str(tmp_path), model_type="flame")


def test_prior_inputs_broadcast_or_raise():
    """shared betas (1,NB) with B > 1 poses must not leave rows 1..B-1 unwritten."""
    import torch
    from smplk.fitting import _prior_rows
    b, p = torch.zeros(1, 10), torch.zeros(4, 63)
    B, ts, bc = _prior_rows([b, None, p, None, None])
    assert B == 4 and ts[0].shape == (4, 10) and ts[0].is_contiguous() and bc == [True, False, False, False, False]
    B, ts, bc = _prior_rows([b, None, None, None, None], 3)
    assert ts[0].shape == (3, 10) and bc[0]
    with pytest.raises(ValueError):
        _prior_rows([torch.zeros(2, 10), None, p, None, None])


def _csrc_text():
    d = os.path.join(ROOT, "3d-human-body-reconstruction_b200", "csrc")
    return {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d))}


def _kernel_body(text, name):
    """Source of the __global__ function `name` (brace matching from its definition)."""
    for m in re.finditer(r"\b%s\s*\(" % re.escape(name), text):
        head = text[max(0, m.start() - 400):m.start()]
        if "__global__" not in head.split(";")[-1] and "__global__" not in head.split("}")[-1]:
            continue
        i = text.index("{", m.end())
        depth, j = 1, i + 1
        while depth:
            depth += {"{": 1, "}": -1}.get(text[j], 0)
            j += 1
        return text[i:j]
    return None


def test_every_programmatically_launched_kernel_waits_for_its_predecessor():
    """Static guard of the PDL rule (ptx_sm100.cuh): a kernel that can be launched with the programmatic-stream-
    serialization attribute (launch_k with a pdl argument) must execute griddepcontrol.wait; and a kernel that
    triggers its dependents early must also wait, since its successor's wait relies on the chain being transitive."""
    src = _csrc_text()
    everything = "\n".join(src.values())
    launched = set(re.findall(r"launch_k\(\s*[^,()]+(?:\(\))?,\s*([A-Za-z_][A-Za-z0-9_]*_kernel)\b", everything))
    assert {"blend_skin_fused_kernel", "pose_forward_block_kernel", "pose_backward_kernel", "skin_fit_l2_kernel",
            "dA_seg_kernel", "blend_tcgen05_2cta_kernel", "reduce_splits_kernel", "lbs_replay_gemm_kernel"} <= launched
    for name in sorted(launched):
        body = next((b for b in (_kernel_body(t, name) for t in src.values()) if b), None)
        assert body is not None, name
        assert "pdl_wait()" in body, "%s is launched through launch_k but never waits" % name
    for fname, text in src.items():
        for m in re.finditer(r"__global__[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*_kernel)\s*\(", text):
            body = _kernel_body(text, m.group(1))
            if body and "pdl_launch_dependents()" in body:
                assert "pdl_wait()" in body, "%s:%s triggers dependents but never waits" % (fname, m.group(1))


def test_header_documents_every_handle_option():
    """include/smplk.h lists the names smplk_model_set_option accepts: the list and the implementation agree."""
    api = _csrc_text()["smplk_api.cu"]
    fn = api[api.index('extern "C" int smplk_model_set_option'):]
    fn = fn[:fn.index("\n}\n")]
    names = set(re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', fn))
    header = open(os.path.join(ROOT, "include", "smplk.h")).read()
    doc = header[header.index("Kernel choices of one handle"):header.index("int smplk_model_set_option")]
    documented = set(re.findall(r'"([a-z0-9_]+)" \(\d\)', doc))
    assert names == documented, (sorted(names - documented), sorted(documented - names))
    assert {"pdl", "skip_pose", "fused", "replay_gemm"} <= names
