"""ctypes binding of the C ABI in include/smplk.h (the shared library built from csrc/).

PyTorch is used for device memory and streams only; every math step of the path runs in
libsmplk.so.  There is no fallback: if the library is missing or no sm_100 GPU is present the
calls raise.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMPLK_LIB") or os.path.join(_HERE, "libsmplk.so")  # SMPLK_LIB: A/B variants
CSRC = os.path.join(_HERE, "csrc")

FLAG_SAVE_FOR_BACKWARD = 1
FLAG_ADD_POSE_MEAN = 2
FLAG_BLEND_SIMT = 4
FLAG_BLEND_TCGEN05 = 8
FLAG_BLEND_TF32 = 16
FLAG_TRANSFORMS_ONLY = 32
FLAG_FIT_VERTEX_L2 = 64
FLAG_LOSS_SUM = 128

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("num_verts", ctypes.c_int32), ("num_joints", ctypes.c_int32),
        ("num_betas", ctypes.c_int32),
        ("v_template", ctypes.c_void_p), ("shapedirs", ctypes.c_void_p),
        ("posedirs", ctypes.c_void_p), ("J_regressor", ctypes.c_void_p),
        ("joints_fixed", ctypes.c_void_p), ("weights", ctypes.c_void_p),
        ("parents", ctypes.c_void_p),
        ("num_pca", ctypes.c_int32),
        ("hand_comp_l", ctypes.c_void_p), ("hand_comp_r", ctypes.c_void_p),
        ("pose_mean", ctypes.c_void_p),
        ("num_extra_verts", ctypes.c_int32), ("extra_vertex_ids", ctypes.c_void_p),
        ("num_regressors", ctypes.c_int32), ("regressor_posed", ctypes.c_void_p),
    ]


class ModelInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "num_verts", "num_joints", "num_betas", "num_pose_feats", "num_extra_verts",
        "num_regressors", "num_pca", "max_weights_per_vertex", "lbs_only", "device",
        "has_tcgen05_path")]


class ForwardArgs(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("betas", ctypes.c_void_p), ("betas_batch", ctypes.c_int32),
        ("pose", ctypes.c_void_p), ("hand_pca_l", ctypes.c_void_p),
        ("hand_pca_r", ctypes.c_void_p), ("transl", ctypes.c_void_p),
        ("verts", ctypes.c_void_p), ("joints", ctypes.c_void_p),
        ("joints_regressed", ctypes.c_void_p), ("full_pose", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("stream", ctypes.c_void_p),
    ]


class BackwardArgs(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("betas", ctypes.c_void_p), ("betas_batch", ctypes.c_int32),
        ("pose", ctypes.c_void_p), ("hand_pca_l", ctypes.c_void_p),
        ("hand_pca_r", ctypes.c_void_p),
        ("d_verts", ctypes.c_void_p), ("d_joints", ctypes.c_void_p),
        ("d_joints_regressed", ctypes.c_void_p),
        ("d_betas", ctypes.c_void_p), ("d_pose", ctypes.c_void_p),
        ("d_hand_pca_l", ctypes.c_void_p), ("d_hand_pca_r", ctypes.c_void_p),
        ("d_transl", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("scratch", ctypes.c_void_p), ("scratch_bytes", ctypes.c_size_t),
        ("stream", ctypes.c_void_p), ("d_loss", ctypes.c_void_p), ("d_full_pose", ctypes.c_void_p),
    ]


class ReprojArgs(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("num_joints", ctypes.c_int32),
        ("joints", ctypes.c_void_p), ("rotation", ctypes.c_void_p), ("translation", ctypes.c_void_p),
        ("focal", ctypes.c_void_p), ("center", ctypes.c_void_p), ("camera_batch", ctypes.c_int32),
        ("gt_joints", ctypes.c_void_p), ("weights", ctypes.c_void_p), ("weights_batch", ctypes.c_int32),
        ("rho", ctypes.c_float), ("data_weight", ctypes.c_float),
        ("loss", ctypes.c_void_p), ("d_joints", ctypes.c_void_p), ("d_translation", ctypes.c_void_p),
        ("device", ctypes.c_int), ("stream", ctypes.c_void_p),
    ]


class PriorArgs(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32),
        ("betas", ctypes.c_void_p), ("num_betas", ctypes.c_int32),
        ("pose_embedding", ctypes.c_void_p), ("num_embedding", ctypes.c_int32),
        ("body_pose", ctypes.c_void_p), ("num_body_pose", ctypes.c_int32),
        ("left_hand_pose", ctypes.c_void_p), ("right_hand_pose", ctypes.c_void_p), ("num_hand", ctypes.c_int32),
        ("shape_weight", ctypes.c_float), ("body_pose_weight", ctypes.c_float),
        ("bending_prior_weight", ctypes.c_float), ("hand_prior_weight", ctypes.c_float),
        ("loss", ctypes.c_void_p),
        ("d_betas", ctypes.c_void_p), ("d_pose_embedding", ctypes.c_void_p), ("d_body_pose", ctypes.c_void_p),
        ("d_left_hand_pose", ctypes.c_void_p), ("d_right_hand_pose", ctypes.c_void_p),
        ("device", ctypes.c_int), ("stream", ctypes.c_void_p),
    ]


# every symbol include/smplk.h declares
EXPORTED_SYMBOLS = [
    "smplk_model_create", "smplk_model_destroy", "smplk_model_get_info", "smplk_workspace_bytes",
    "smplk_forward", "smplk_backward_scratch_bytes", "smplk_backward", "smplk_regress_joints",
    "smplk_batch_rodrigues", "smplk_forward_host", "smplk_last_error_string", "smplk_version",
    "smplk_launch_count", "smplk_workspace_layout", "smplk_profile_enable", "smplk_profile_read", "smplk_vertex_l2",
    "smplk_inverse_lbs", "smplk_inverse_joints", "smplk_vertex_normals", "smplk_divide_faces",
    "smplk_reprojection_loss", "smplk_fit_priors", "smplk_fit_vertex_l2", "smplk_skin_transforms", "smplk_remove_rest", "smplk_model_set_option",
]
PROF_SLOTS = ["pose_fwd", "blend_tcgen05", "blend_simt", "skin", "dA", "skin_bwd", "blend_bwd",
              "pose_bwd", "blend_skin_fused", "transpose"]


def build(force=False, verbose=False, ab=False):
    """Compile csrc/ into libsmplk.so for sm_100a with nvcc (cross-compiles without a GPU).
    ab=True builds libsmplk_ab.so with -DSMPLK_AB: the product library plus the A/B kernels that are
    on no default path (exact-fp32 SIMT blend, bulk-copy skinning) and their tuning switches; select
    it with SMPLK_LIB=<path> (tools/, the SIMT cross-checks of the parity tests)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    out = os.path.join(_HERE, "libsmplk_ab.so") if ab else LIB_PATH
    if not force and os.path.exists(out):
        if os.path.getmtime(out) >= max(os.path.getmtime(s) for s in srcs):
            return out
    cmd = ["nvcc"] + NVCC_FLAGS + (["-DSMPLK_AB"] if ab else []) + ["-o", out, os.path.join(CSRC, "smplk_api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


_lib = None


def load():
    """Load libsmplk.so (building it if nvcc is available and the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            build()
        except Exception as e:
            raise RuntimeError(
                "libsmplk.so is missing and could not be built (%s). smplk has no fallback path."
                % e)
    lib = ctypes.CDLL(LIB_PATH)
    lib.smplk_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p)]
    lib.smplk_model_create.restype = ctypes.c_int
    lib.smplk_model_destroy.argtypes = [ctypes.c_void_p]
    lib.smplk_model_destroy.restype = ctypes.c_int
    lib.smplk_model_get_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ModelInfo)]
    lib.smplk_model_get_info.restype = ctypes.c_int
    lib.smplk_workspace_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32]
    lib.smplk_workspace_bytes.restype = ctypes.c_size_t
    lib.smplk_forward.argtypes = [ctypes.c_void_p, ctypes.POINTER(ForwardArgs)]
    lib.smplk_forward.restype = ctypes.c_int
    lib.smplk_backward_scratch_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.smplk_backward_scratch_bytes.restype = ctypes.c_size_t
    lib.smplk_backward.argtypes = [ctypes.c_void_p, ctypes.POINTER(BackwardArgs)]
    lib.smplk_backward.restype = ctypes.c_int
    lib.smplk_regress_joints.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p]
    lib.smplk_regress_joints.restype = ctypes.c_int
    lib.smplk_batch_rodrigues.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_void_p]
    lib.smplk_batch_rodrigues.restype = ctypes.c_int
    lib.smplk_forward_host.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32,
                                       ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
    lib.smplk_forward_host.restype = ctypes.c_int
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.smplk_inverse_lbs.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.smplk_inverse_lbs.restype = ctypes.c_int
    lib.smplk_inverse_joints.argtypes = [i32, i32, vp, vp, i32, vp, vp, ctypes.c_int, vp]
    lib.smplk_inverse_joints.restype = ctypes.c_int
    lib.smplk_vertex_normals.argtypes = [i32, i32, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    lib.smplk_vertex_normals.restype = ctypes.c_int
    lib.smplk_divide_faces.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    lib.smplk_divide_faces.restype = ctypes.c_int
    lib.smplk_reprojection_loss.argtypes = [ctypes.POINTER(ReprojArgs)]
    lib.smplk_reprojection_loss.restype = ctypes.c_int
    lib.smplk_fit_priors.argtypes = [ctypes.POINTER(PriorArgs)]
    lib.smplk_fit_priors.restype = ctypes.c_int
    lib.smplk_workspace_layout.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32,
                                           ctypes.POINTER(ctypes.c_size_t),
                                           ctypes.POINTER(ctypes.c_int32)]
    lib.smplk_workspace_layout.restype = ctypes.c_int
    lib.smplk_vertex_l2.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p]
    lib.smplk_vertex_l2.restype = ctypes.c_int
    lib.smplk_fit_vertex_l2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                        ctypes.c_void_p]
    lib.smplk_fit_vertex_l2.restype = ctypes.c_int
    lib.smplk_skin_transforms.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.smplk_skin_transforms.restype = ctypes.c_int
    lib.smplk_model_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    lib.smplk_model_set_option.restype = ctypes.c_int
    lib.smplk_remove_rest.argtypes = [i32, i32, vp, vp, vp, ctypes.c_int, vp]
    lib.smplk_remove_rest.restype = ctypes.c_int
    lib.smplk_profile_enable.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.smplk_profile_enable.restype = ctypes.c_int
    lib.smplk_profile_read.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
    lib.smplk_profile_read.restype = ctypes.c_int
    lib.smplk_last_error_string.restype = ctypes.c_char_p
    lib.smplk_version.restype = ctypes.c_int
    lib.smplk_launch_count.restype = ctypes.c_uint64
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("smplk error %d: %s" % (rc, load().smplk_last_error_string().decode()))


def launch_count():
    return int(load().smplk_launch_count())


def _f64(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def parents_from_model(model):
    """Parent index per joint (root -1): `parent` dict of a rigged mesh (lib/model2video.py:25) or
    the kintree table of the pickles (models/smplh_np.py:19-23)."""
    kt = np.asarray(model["kintree_table"]).astype(np.int64)
    J = kt.shape[1]
    if model.get("parent") is not None and not isinstance(model.get("parent"), np.ndarray):
        par = model["parent"]
        return np.asarray([-1] + [int(par[i]) for i in range(1, J)], dtype=np.int32)
    col = {int(kt[1, i]): i for i in range(J)}
    return np.asarray([-1] + [col[int(kt[0, i])] for i in range(1, J)], dtype=np.int32)


class DeviceModel:
    """Owns one `smplk_model*` (packed constants on one GPU)."""

    def __init__(self, model, device=0, num_betas=None, num_pca_comps=0, flat_hand_mean=False,
                 regressor_posed=None, extra_vertex_ids=None, lbs_only=False, options=None):
        lib = load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        vt = _f64(model["v_template"])
        V = vt.shape[0]
        W = model["weights"]
        W = _f64(W.toarray() if hasattr(W, "toarray") else W)
        parents = np.ascontiguousarray(parents_from_model(model), dtype=np.int32)
        J = parents.shape[0]
        keep = [vt, W, parents]
        d = ModelDesc()
        d.num_verts, d.num_joints = V, J
        d.v_template, d.weights, d.parents = _ptr(vt), _ptr(W), _ptr(parents)
        if lbs_only:
            jf = _f64(model["J"])
            keep.append(jf)
            d.num_betas = 0
            d.joints_fixed = _ptr(jf)
        else:
            sd = _f64(model["shapedirs"])
            nb = sd.shape[2] if num_betas is None else min(num_betas, sd.shape[2])
            sd = np.ascontiguousarray(sd[:, :, :nb])
            pd = _f64(np.asarray(model["posedirs"]).reshape(V, 3, -1))
            if pd.shape[2] != 9 * (J - 1):
                raise ValueError("posedirs has %d pose features, expected %d" % (pd.shape[2], 9 * (J - 1)))
            Jr = model["J_regressor"]
            Jr = _f64(Jr.toarray() if hasattr(Jr, "toarray") else Jr)
            keep += [sd, pd, Jr]
            d.num_betas = nb
            d.shapedirs, d.posedirs, d.J_regressor = _ptr(sd), _ptr(pd), _ptr(Jr)
        self.pose_mean = None
        if num_pca_comps and "hands_componentsl" in model:
            cl = np.ascontiguousarray(_f64(model["hands_componentsl"])[:num_pca_comps])
            cr = np.ascontiguousarray(_f64(model["hands_componentsr"])[:num_pca_comps])
            keep += [cl, cr]
            d.num_pca, d.hand_comp_l, d.hand_comp_r = num_pca_comps, _ptr(cl), _ptr(cr)
        if "hands_meanl" in model and not lbs_only:
            pm = np.zeros(3 * J)
            if not flat_hand_mean:
                pm[3 * (J - 30):3 * (J - 15)] = np.asarray(model["hands_meanl"], np.float64)
                pm[3 * (J - 15):] = np.asarray(model["hands_meanr"], np.float64)
            keep.append(pm)
            d.pose_mean = _ptr(pm)
            self.pose_mean = pm
        if extra_vertex_ids is not None and len(extra_vertex_ids) > 0:
            ev = np.ascontiguousarray(np.asarray(extra_vertex_ids), dtype=np.int32)
            keep.append(ev)
            d.num_extra_verts, d.extra_vertex_ids = ev.shape[0], _ptr(ev)
        if regressor_posed is not None:
            rp = regressor_posed
            rp = _f64(rp.toarray() if hasattr(rp, "toarray") else rp)
            keep.append(rp)
            d.num_regressors, d.regressor_posed = rp.shape[0], _ptr(rp)
        check(lib.smplk_model_create(ctypes.byref(d), int(device), ctypes.byref(self.handle)))
        info = ModelInfo()
        check(lib.smplk_model_get_info(self.handle, ctypes.byref(info)))
        self.info = info
        self.V, self.J, self.NB = info.num_verts, info.num_joints, info.num_betas
        self.E, self.R, self.C = info.num_extra_verts, info.num_regressors, info.num_pca
        self.device = int(device)
        self.parents = parents
        for name, value in (options or {}).items():
            self.set_option(name, value)

    def set_option(self, name, value):
        """Kernel choice of this handle (smplk_model_set_option): 'fused', 'pose_block', 'blend_tf32',
        'backward_tf32', 'gemm_2cta', 'fit_fused', 'sparse_picks'."""
        check(self._lib.smplk_model_set_option(self.handle, name.encode(), int(value)))

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self._lib.smplk_model_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def workspace_bytes(self, batch, flags=0):
        return int(self._lib.smplk_workspace_bytes(self.handle, int(batch), int(flags)))

    def workspace_layout(self, batch, flags=0):
        offs = (ctypes.c_size_t * 4)()
        chunk = ctypes.c_int32()
        check(self._lib.smplk_workspace_layout(self.handle, int(batch), int(flags), offs,
                                               ctypes.byref(chunk)))
        return dict(F_hi=offs[0], F_lo=offs[1], A=offs[2], v_posed=offs[3], chunk=chunk.value)

    def profile_enable(self, on=True):
        check(self._lib.smplk_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self, reset=True):
        """{kernel: (total_ms, launches)} measured with CUDA events on the launching stream."""
        ms = (ctypes.c_double * len(PROF_SLOTS))()
        n = (ctypes.c_int64 * len(PROF_SLOTS))()
        check(self._lib.smplk_profile_read(self.handle, ms, n, 1 if reset else 0))
        return {k: (ms[i], n[i]) for i, k in enumerate(PROF_SLOTS)}

    def backward_scratch_bytes(self, batch):
        return int(self._lib.smplk_backward_scratch_bytes(self.handle, int(batch)))

    def forward(self, args):
        check(self._lib.smplk_forward(self.handle, ctypes.byref(args)))

    def backward(self, args):
        check(self._lib.smplk_backward(self.handle, ctypes.byref(args)))

    def fit_vertex_l2(self, args, target_ptr, scale, loss_ptr):
        check(self._lib.smplk_fit_vertex_l2(self.handle, ctypes.byref(args), target_ptr, float(scale), loss_ptr))
