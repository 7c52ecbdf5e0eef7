"""Drop-in numpy-facing twins of the reference's CPU body models, executed on the GPU.

    SMPLHModel   <-> models/smplh_np.py:5-117
    SMPLModel    <-> models/smpl_np.py:121-231 (forward part)
    RecoverModel <-> lib/model2video.py:12-130 == lib/mesh2smpl_model.py:131-313 (LBS-only rig)

Same constructor (`model_path` of a pickle, or a model dict), same `set_params(pose, beta, trans)
-> verts` contract, same attributes (`verts`, `J`, `R`, `faces`, `weights`, `parent`,
`kintree_table`, `v_template`) and `gen_J_3d()`.  Each call goes through the host-buffer C-ABI
entry `smplk_forward_host` (H2D copy -> sm_100a kernels -> D2H copy); `forward_batch` evaluates a
whole motion clip (frames are independent bodies) in one call, which is how the per-frame loops of
lib/model2video.py:514-518 should be driven on a GPU.
"""
import ctypes

import numpy as np

from . import _lib
from .body_models import load_model_file


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _TwinBase:
    NUM_JOINTS = 24
    LBS_ONLY = False

    def __init__(self, model_path, device=0, num_betas=10):
        params = load_model_file(model_path) if isinstance(model_path, str) else model_path
        self._params = params
        W = params["weights"]
        self.weights = W.toarray() if hasattr(W, "toarray") else np.asarray(W)
        self.v_template = np.asarray(params["v_template"])
        self.faces = np.asarray(params["f"])
        self.kintree_table = np.asarray(params["kintree_table"])
        pars = _lib.parents_from_model(params)
        self.parent = {i: int(pars[i]) for i in range(1, pars.shape[0])}
        J = pars.shape[0]
        self.pose_shape = [J, 3]
        self.beta_shape = [num_betas]
        self.trans_shape = [3]
        self.pose = np.zeros(self.pose_shape)
        self.beta = np.zeros(self.beta_shape)
        self.trans = np.zeros(self.trans_shape)
        self.verts = None
        self.J = None
        self.R = None
        self._device = device
        self._lib = _lib.load()

    # ---- C-ABI calls with host buffers
    def _forward_host(self, pose, beta, trans, want_joints=False):
        B = pose.shape[0]
        dm = self._dm
        verts = np.empty((B, dm.V, 3), dtype=np.float32)
        joints = np.empty((B, dm.J + dm.E, 3), dtype=np.float32) if want_joints else None
        beta_c = None if beta is None else _c(beta)
        pose_c, trans_c = _c(pose), (None if trans is None else _c(trans))
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self._lib.smplk_forward_host(
            dm.handle, B, 0, p(beta_c), 1 if beta_c is None else beta_c.shape[0], p(pose_c),
            p(trans_c), p(verts), p(joints), None))
        return verts, joints

    def forward_batch(self, poses, betas=None, trans=None, return_joints=False):
        """poses (N, J*3) [, betas (NB,) or (N,NB)] [, trans (N,3)] -> verts (N,V,3) float32."""
        poses = np.asarray(poses).reshape(len(poses), -1)
        if betas is not None:
            betas = np.asarray(betas, dtype=np.float32).reshape(-1, self._dm.NB) if self._dm.NB else None
        verts, joints = self._forward_host(poses, betas, trans, want_joints=return_joints)
        return (verts, joints) if return_joints else verts

    def set_params(self, pose=None, beta=None, trans=None):
        if pose is not None:
            self.pose = pose
        if beta is not None:
            self.beta = beta
        if trans is not None:
            self.trans = trans
        self.update()
        return self.verts

    def update(self):
        pose = np.asarray(self.pose, dtype=np.float64).reshape(1, -1)
        beta = None if self.LBS_ONLY else np.asarray(self.beta, dtype=np.float64).reshape(1, -1)
        trans = np.asarray(self.trans, dtype=np.float64).reshape(1, 3)
        verts, joints = self._forward_host(pose, beta, trans, want_joints=True)
        self.verts = verts[0].astype(np.float64)
        self._joints_fk = joints[0, :self._dm.J].astype(np.float64)

    def gen_J_3d(self):
        """J_regressor . posed verts (models/smplh_np.py:116-117), on the GPU."""
        import torch
        dm = self._dm
        v = torch.as_tensor(self.verts, dtype=torch.float32, device="cuda:%d" % self._device).reshape(1, dm.V, 3).contiguous()
        out = torch.empty(1, dm.R, 3, dtype=torch.float32, device=v.device)
        _lib.check(self._lib.smplk_regress_joints(
            dm.handle, 1, ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)))
        return out[0].cpu().numpy().astype(np.float64)

    def output_mesh(self, path):
        with open(path, "w") as fp:
            for v in self.verts:
                fp.write("v %f %f %f\n" % (v[0], v[1], v[2]))
            for f in self.faces + 1:
                fp.write("f %d %d %d\n" % (f[0], f[1], f[2]))


class SMPLHModel(_TwinBase):
    """models/smplh_np.py:5 -- 52 joints, full axis-angle pose (52,3), no PCA, no pose mean."""

    def __init__(self, model_path, device=0, num_betas=10):
        super().__init__(model_path, device, num_betas)
        p = self._params
        self.J_regressor = p["J_regressor"]
        self.shapedirs = np.asarray(p["shapedirs"])
        self.posedirs = np.asarray(p["posedirs"])
        self._dm = _lib.DeviceModel(p, device=device, num_betas=num_betas,
                                    regressor_posed=p["J_regressor"])
        self.update()


class SMPLModel(SMPLHModel):
    """models/smpl_np.py:121 -- 24 joints."""


class RecoverModel(_TwinBase):
    """lib/model2video.py:12 -- recovered mesh bound to the SMPL skeleton; LBS only, fixed joints;
    joints 13,14,22,23 are zeroed in `set_params` (:44-45)."""
    LBS_ONLY = True

    def __init__(self, model_path, device=0):
        super().__init__(model_path, device, 10)
        p = self._params
        self.weigths = self.weights
        self.color = p.get("color")
        self.J = np.asarray(p["J"])
        self.or_pose = p.get("or_pose")
        self.ignor_J = [13, 14, 22, 23]
        self._dm = _lib.DeviceModel(p, device=device, lbs_only=True)
        self.verts = self.v_template
        self.update()

    def set_params(self, pose=None, beta=None, trans=None):
        if pose is not None:
            for i in self.ignor_J:
                pose[i] = [0, 0, 0]
        return super().set_params(pose=pose, beta=beta, trans=trans)

    def replay(self, poses, trans=None):
        """Whole clip at once: poses (N,72) as read by read_amsass (lib/model2video.py:527-531)."""
        poses = np.array(poses, dtype=np.float32).reshape(len(poses), -1, 3)
        poses[:, self.ignor_J] = 0.0
        return self.forward_batch(poses.reshape(len(poses), -1), None, trans)
