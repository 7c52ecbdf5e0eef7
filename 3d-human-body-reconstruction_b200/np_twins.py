"""Drop-in numpy-facing twins of the reference's CPU body models, executed on the GPU.

    SMPLHModel   <-> models/smplh_np.py:5-117
    SMPLModel    <-> models/smpl_np.py:121-246 (forward part + `inverse`)
    RecoverModel <-> lib/model2video.py:12-130 == lib/mesh2smpl_model.py:131-313 (LBS-only rig)

Same constructor (`model_path` of a pickle, or a model dict), same `set_params(pose, beta, trans)
-> verts` contract, same attributes (`verts`, `J`, `R`, `v_posed`, `faces`, `weights`, `parent`,
`kintree_table`, `v_template`) and methods: `compute_R_G() -> G (J,4,4)`, `do_skinning(G)`,
`update()`, `gen_J_3d()`, `inverse()` (models/smpl_np.py:239-246), `output_mesh`.

`set_params` / `update` go through the host-buffer C-ABI entry `smplk_forward_host` (H2D copy ->
sm_100a kernels -> D2H copy) in ONE call; the by-products the reference's `compute_R_G` leaves on the
object (`J`, `R`, `v_posed`) are fetched from the GPU the first time they are read after a
`set_params` (lib/mesh2smpl_model.py:146,159,193-197 read `smpl.J` and call `smpl.compute_R_G()`).
`forward_batch` evaluates a whole motion clip (frames are independent bodies) in one call, which is
how the per-frame loops of lib/model2video.py:514-518 should be driven on a GPU.  There is no CPU
path: all arithmetic runs in libsmplk.so.
"""
import ctypes

import numpy as np

from . import _lib
from .body_models import load_model_file


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _vp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


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
        self._J = self._R = self._v_posed = None
        self._aux_stale = True
        self._A_dev = None
        self._device = device
        self._lib = _lib.load()

    # ---- reference attributes that compute_R_G leaves behind; fetched from the GPU on first read
    def _aux(self, name):
        if self._aux_stale:
            self.compute_R_G()
        return getattr(self, name)

    J = property(lambda self: self._aux("_J"), lambda self, v: setattr(self, "_J", v))
    R = property(lambda self: self._aux("_R"), lambda self, v: setattr(self, "_R", v))
    v_posed = property(lambda self: self._aux("_v_posed"), lambda self, v: setattr(self, "_v_posed", v))

    # ---- argument checks (the C side strides over fixed widths)
    def _check(self, pose, beta, trans):
        dm = self._dm
        B = pose.shape[0]
        if pose.ndim != 2 or pose.shape[1] != 3 * dm.J:
            raise ValueError("pose must have %d values per body (%d joints x 3), got %s"
                             % (3 * dm.J, dm.J, tuple(pose.shape)))
        if beta is not None and (beta.ndim != 2 or beta.shape[1] != dm.NB or beta.shape[0] not in (1, B)):
            raise ValueError("beta must have %d values (1 or %d rows), got %s" % (dm.NB, B, tuple(beta.shape)))
        if trans is not None and tuple(trans.shape) != (B, 3):
            raise ValueError("trans must be (%d, 3), got %s" % (B, tuple(trans.shape)))

    # ---- C-ABI calls with host buffers
    def _forward_host(self, pose, beta, trans, want_joints=False):
        B = pose.shape[0]
        dm = self._dm
        beta_c = None if beta is None else _c(beta)
        pose_c, trans_c = _c(pose), (None if trans is None else _c(trans))
        self._check(pose_c, beta_c, trans_c)
        verts = np.empty((B, dm.V, 3), dtype=np.float32)
        joints = np.empty((B, dm.J + dm.E, 3), dtype=np.float32) if want_joints else None
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
        if trans is not None:
            trans = np.asarray(trans).reshape(-1, 3)
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

    def _host_params(self):
        pose = np.asarray(self.pose, dtype=np.float64).reshape(1, -1)
        beta = None if self.LBS_ONLY else np.asarray(self.beta, dtype=np.float64).reshape(1, -1)
        trans = np.asarray(self.trans, dtype=np.float64).reshape(1, 3)
        return pose, beta, trans

    def update(self):
        """compute_R_G + do_skinning of the reference in one host-buffer call."""
        pose, beta, trans = self._host_params()
        verts, joints = self._forward_host(pose, beta, trans, want_joints=True)
        self.verts = verts[0].astype(np.float64)
        self._joints_fk = joints[0, :self._dm.J].astype(np.float64)
        self._aux_stale = True
        self._A_dev = None

    # ---- the two halves of update(), as the reference exposes them
    def _run_device(self, pose_t, beta_t, flags):
        """One smplk_forward on device tensors (no transl: compute_R_G never sees it).
        Returns (workspace, layout, fk_joints)."""
        import torch
        dm = self._dm
        dev = pose_t.device
        ws = torch.empty(dm.workspace_bytes(1, flags), device=dev, dtype=torch.uint8)
        joints = torch.empty(1, dm.J + dm.E, 3, device=dev, dtype=torch.float32)
        a = _lib.ForwardArgs()
        a.batch, a.flags = 1, flags
        a.betas, a.betas_batch = _vp(beta_t), 1
        a.pose, a.joints = _vp(pose_t), _vp(joints)
        a.workspace, a.workspace_bytes = _vp(ws), ws.numel()
        a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        dm.forward(a)
        return ws, dm.workspace_layout(1, flags), joints

    def compute_R_G(self):
        """models/smplh_np.py:49-70 (lib/model2video.py:55-66 for a rigged mesh): sets `J` (rest joints),
        `R` (J,3,3), `v_posed` (blendshape models) and returns the global joint transforms G (J,4,4)."""
        import torch
        dm = self._dm
        dev = torch.device("cuda", self._device)
        pose, beta, _ = self._host_params()
        self._check(pose, beta, None)
        with torch.cuda.device(dev):
            pose_t = torch.as_tensor(_c(pose), device=dev)
            beta_t = None if beta is None else torch.as_tensor(_c(beta), device=dev)
            flags = _lib.FLAG_SAVE_FOR_BACKWARD | (_lib.FLAG_TRANSFORMS_ONLY if self.LBS_ONLY else 0)
            ws, lay, fk = self._run_device(pose_t, beta_t, flags)
            A = ws[lay["A"]:lay["A"] + dm.J * 48].view(torch.float32).view(dm.J, 3, 4)
            R = torch.empty(dm.J, 3, 3, device=dev, dtype=torch.float32)
            _lib.check(self._lib.smplk_batch_rodrigues(dm.J, _vp(pose_t), _vp(R), self._device,
                                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            G = torch.zeros(dm.J, 4, 4, device=dev, dtype=torch.float32)
            G[:, :3, :3] = A[:, :, :3]
            G[:, :3, 3] = fk[0, :dm.J]
            G[:, 3, 3] = 1.0
            if not self.LBS_ONLY:
                npad = (3 * dm.V + 255) // 256 * 256
                vp = ws[lay["v_posed"]:lay["v_posed"] + npad * 4].view(torch.float32)[:3 * dm.V]
                self._v_posed = vp.view(dm.V, 3).double().cpu().numpy()
                # rest joints: the FK joints of the zero pose are J_template + J_shapedirs . beta
                _, _, j0 = self._run_device(torch.zeros_like(pose_t), beta_t,
                                            _lib.FLAG_SAVE_FOR_BACKWARD | _lib.FLAG_TRANSFORMS_ONLY)
                self._J = j0[0, :dm.J].double().cpu().numpy()
            self._R = R.double().cpu().numpy()
            G = G.double().cpu().numpy()
        self._aux_stale = False
        return G

    def do_skinning(self, G):
        """models/smplh_np.py:72-82: rest-pose removal, weight blend, apply to `v_posed`, add `trans`."""
        import torch
        dm = self._dm
        dev = torch.device("cuda", self._device)
        G = np.asarray(G, dtype=np.float64)
        if G.shape != (dm.J, 4, 4):
            raise ValueError("G must be (%d, 4, 4), got %s" % (dm.J, G.shape))
        Jr = np.asarray(self.J, dtype=np.float64).reshape(dm.J, 3)
        with torch.cuda.device(dev):
            G_t = torch.as_tensor(_c(G.reshape(1, dm.J, 16)), device=dev)
            J_t = torch.as_tensor(_c(Jr.reshape(1, dm.J, 3)), device=dev)
            tr = torch.as_tensor(_c(np.asarray(self.trans, np.float64).reshape(1, 3)), device=dev)
            vp_t, ld = None, 0
            if not self.LBS_ONLY:
                ld = (3 * dm.V + 3) // 4 * 4
                vp_t = torch.zeros(1, ld, device=dev, dtype=torch.float32)
                vp_t[0, :3 * dm.V] = torch.as_tensor(_c(np.asarray(self.v_posed).reshape(-1)), device=dev)
            A = torch.empty(1, dm.J, 12, device=dev, dtype=torch.float32)
            out = torch.empty(1, dm.V, 3, device=dev, dtype=torch.float32)
            _lib.check(self._lib.smplk_skin_transforms(
                dm.handle, 1, _vp(G_t), _vp(J_t), _vp(vp_t), ld, _vp(tr), _vp(A), _vp(out),
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            self._A_dev = A
            self.verts = out[0].double().cpu().numpy()

    def inverse(self):
        """models/smpl_np.py:239-246: verts <- T^-1 [verts - trans; 1] with the per-vertex blended
        transforms of the last skinning (un-posing)."""
        import torch
        dm = self._dm
        dev = torch.device("cuda", self._device)
        if self._A_dev is None:
            self.do_skinning(self.compute_R_G())
        with torch.cuda.device(dev):
            v = torch.as_tensor(_c(np.asarray(self.verts).reshape(1, dm.V, 3)), device=dev)
            tr = torch.as_tensor(_c(np.asarray(self.trans, np.float64).reshape(1, 3)), device=dev)
            out = torch.empty_like(v)
            _lib.check(self._lib.smplk_inverse_lbs(dm.handle, 1, _vp(self._A_dev), _vp(v), _vp(tr), _vp(out),
                                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            self.verts = out[0].double().cpu().numpy()

    def gen_J_3d(self):
        """J_regressor . posed verts (models/smplh_np.py:116-117), on the GPU."""
        import torch
        dm = self._dm
        dev = torch.device("cuda", self._device)
        with torch.cuda.device(dev):
            v = torch.as_tensor(self.verts, dtype=torch.float32, device=dev).reshape(1, dm.V, 3).contiguous()
            out = torch.empty(1, dm.R, 3, dtype=torch.float32, device=dev)
            _lib.check(self._lib.smplk_regress_joints(
                dm.handle, 1, _vp(v), _vp(out), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
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
        self._J = np.asarray(p["J"])
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
