"""Drop-in torch modules for the reference's body models.

    SMPLH  <-> models/smplh.py:13-39  (subclass of upstream smplx.SMPLH; created through
               smplx.create(...) at lib/gen_smplh.py:75-90 and called at
               lib/Gen_SMPLH/fitting.py:243-245)
    SMPL   <-> models/smpl.py:11-37

Same constructor keywords, same `forward` signature, same `ModelOutput` fields.  The math runs in
libsmplk.so (hand-written sm_100a kernels behind the C ABI of include/smplk.h) through a
`torch.autograd.Function`, so the modules are differentiable w.r.t. betas, global_orient,
body_pose, hand poses (axis-angle or PCA) and transl.  torch only provides device memory, the
current stream and the autograd glue.
"""
import ctypes
import os
import pickle
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

ModelOutput = namedtuple("ModelOutput", ["vertices", "joints", "full_pose", "betas",
                                         "global_orient", "body_pose", "expression",
                                         "left_hand_pose", "right_hand_pose", "jaw_pose"])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)


def resolve_model_path(model_path, kind, gender="neutral", ext="pkl"):
    """Upstream smplx path rule (lib/gen_smplh.py:75-90 passes a FOLDER): a directory holds
    `<KIND>_<GENDER>.<ext>` (SMPLH_MALE.pkl, SMPL_NEUTRAL.pkl ...); a file path is taken as is."""
    if os.path.isdir(model_path):
        fn = "%s_%s.%s" % (kind.upper(), str(gender).upper(), ext)
        cand = os.path.join(model_path, fn)
        if not os.path.exists(cand) and ext == "pkl" and os.path.exists(cand[:-3] + "npz"):
            cand = cand[:-3] + "npz"
        if not os.path.exists(cand):
            raise FileNotFoundError("Path %s does not exist!" % cand)
        return cand
    return model_path


def load_model_file(path):
    """Model dict from a `.pkl` (official layout, latin1 pickles) or `.npz` file."""
    if str(path).endswith(".npz"):
        d = np.load(path, allow_pickle=True)
        return {k: d[k] for k in d.files}
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _prep(t, device):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("smplk needs CUDA tensors (no CPU path): got a tensor on %s" % t.device)
    if t.dtype != torch.float32:
        raise RuntimeError("smplk computes in float32: got %s" % t.dtype)
    return t.contiguous()


def _check_inputs(dm, B, betas, pose, pca_l, pca_r, transl):
    """Shapes the C side strides over with fixed widths; upstream raises on these from its cat / matmul."""
    if pose.dim() != 2 or pose.shape[1] != 3 * dm.J:
        raise ValueError("pose must be (B, %d) axis-angle for this %d-joint model, got %s"
                         % (3 * dm.J, dm.J, tuple(pose.shape)))
    if betas is not None and (betas.dim() != 2 or betas.shape[1] != dm.NB or betas.shape[0] not in (1, B)):
        raise ValueError("betas must be (1|%d, %d), got %s" % (B, dm.NB, tuple(betas.shape)))
    for name, t in (("left_hand_pose", pca_l), ("right_hand_pose", pca_r)):
        if t is not None and (t.dim() != 2 or tuple(t.shape) != (B, dm.C)):
            raise ValueError("%s PCA coefficients must be (%d, %d), got %s" % (name, B, dm.C, tuple(t.shape)))
    if transl is not None and tuple(transl.shape) != (B, 3):
        raise ValueError("transl must be (%d, 3), got %s" % (B, tuple(transl.shape)))


def _grad_flows(*ts):
    """True when autograd will record the node: grad mode on and an input that requires grad.  Under
    torch.no_grad() module Parameters still report requires_grad, and the inference forward must not
    keep the whole batch's v_posed (it takes the fused blend+skinning kernel and 8192-body chunks)."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class _BodyModelFn(torch.autograd.Function):
    """verts, joints, joints_regressed, full_pose = f(betas, pose, pca_l, pca_r, transl)."""

    @staticmethod
    def forward(ctx, dm, flags, want_regressed, want_verts, betas, pose, pca_l, pca_r, transl):
        B = pose.shape[0]
        dev = pose.device
        # body_model_apply sets SAVE_FOR_BACKWARD when a gradient can flow (grad mode is always off in here)
        needs_grad = bool(flags & _lib.FLAG_SAVE_FOR_BACKWARD)
        betas_c, pose_c = _prep(betas, dev), _prep(pose, dev)
        pl, pr, tr = _prep(pca_l, dev), _prep(pca_r, dev), _prep(transl, dev)
        _check_inputs(dm, B, betas_c, pose_c, pl, pr, tr)
        jreg = (torch.empty(B, dm.R, 3, device=dev, dtype=torch.float32)
                if (want_regressed and dm.R > 0) else None)
        # joints only (return_verts=False): the library blends and skins the picked vertices alone
        want_verts = want_verts or jreg is not None
        verts = torch.empty(B, dm.V, 3, device=dev, dtype=torch.float32) if want_verts else None
        joints = torch.empty(B, dm.J + dm.E, 3, device=dev, dtype=torch.float32)
        full_pose = torch.empty(B, 3 * dm.J, device=dev, dtype=torch.float32)
        ws_bytes = dm.workspace_bytes(B, flags)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        a = _lib.ForwardArgs()
        a.batch, a.flags = B, flags
        a.betas, a.betas_batch = _ptr(betas_c), (betas_c.shape[0] if betas_c is not None else 1)
        a.pose, a.hand_pca_l, a.hand_pca_r, a.transl = _ptr(pose_c), _ptr(pl), _ptr(pr), _ptr(tr)
        a.verts, a.joints, a.joints_regressed = _ptr(verts), _ptr(joints), _ptr(jreg)
        a.full_pose = _ptr(full_pose)
        a.workspace, a.workspace_bytes = _ptr(ws), ws_bytes
        a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            dm.forward(a)
        ctx.dm, ctx.flags = dm, flags
        # outputs the loss never touched arrive as None in backward instead of zero tensors: a zero
        # d_joints would send the backward through the vertex-pick scatter (a copy of d_verts)
        ctx.set_materialize_grads(False)
        if needs_grad:
            ctx.ws = ws
            ctx.inputs = (betas_c, pose_c, pl, pr, tr)
        if jreg is None:
            jreg = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(jreg)
        if verts is None:
            verts = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(verts)
        return verts, joints, jreg, full_pose

    @staticmethod
    def backward(ctx, d_verts, d_joints, d_jreg, d_full_pose):
        dm = ctx.dm
        betas, pose, pl, pr, tr = ctx.inputs
        B = pose.shape[0]
        dev = pose.device
        need = ctx.needs_input_grad[1:]  # (flags, want_reg, want_verts, betas, pose, pca_l, pca_r, transl)

        def grad_in(g):
            if g is None:
                return None
            return g.contiguous().float()

        d_verts = grad_in(d_verts) if (d_verts is not None and d_verts.numel() > 0) else None
        d_joints = grad_in(d_joints)
        d_jreg = grad_in(d_jreg) if (d_jreg is not None and d_jreg.numel() > 0) else None
        d_full_pose = grad_in(d_full_pose)
        d_betas = torch.empty_like(betas) if (betas is not None and need[3]) else None
        d_pose = torch.empty_like(pose) if need[4] else None
        d_pl = torch.empty_like(pl) if (pl is not None and need[5]) else None
        d_pr = torch.empty_like(pr) if (pr is not None and need[6]) else None
        d_tr = torch.empty_like(tr) if (tr is not None and need[7]) else None
        sc_bytes = dm.backward_scratch_bytes(B)
        sc = torch.empty(sc_bytes, device=dev, dtype=torch.uint8)
        a = _lib.BackwardArgs()
        a.batch, a.flags = B, ctx.flags
        a.betas, a.betas_batch = _ptr(betas), (betas.shape[0] if betas is not None else 1)
        a.pose, a.hand_pca_l, a.hand_pca_r = _ptr(pose), _ptr(pl), _ptr(pr)
        a.d_verts, a.d_joints, a.d_joints_regressed = _ptr(d_verts), _ptr(d_joints), _ptr(d_jreg)
        a.d_full_pose = _ptr(d_full_pose)
        a.d_betas, a.d_pose = _ptr(d_betas), _ptr(d_pose)
        a.d_hand_pca_l, a.d_hand_pca_r, a.d_transl = _ptr(d_pl), _ptr(d_pr), _ptr(d_tr)
        a.workspace, a.workspace_bytes = _ptr(ctx.ws), ctx.ws.numel()
        a.scratch, a.scratch_bytes = _ptr(sc), sc_bytes
        a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            dm.backward(a)
        return None, None, None, None, d_betas, d_pose, d_pl, d_pr, d_tr


class _VertexL2Fn(torch.autograd.Function):
    """loss (B,) = scale * sum ||V - V*||^2 per body, gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, verts, target, scale):
        if tuple(target.shape) != tuple(verts.shape):
            raise ValueError("target %s must have the shape of verts %s" % (tuple(target.shape), tuple(verts.shape)))
        v, t = _prep(verts, verts.device), _prep(target, verts.device)
        B = v.shape[0]
        n = v[0].numel()
        loss = torch.empty(B, device=v.device, dtype=torch.float32)
        grad = torch.empty_like(v) if verts.requires_grad else None
        lib = _lib.load()
        with torch.cuda.device(v.device):
            _lib.check(lib.smplk_vertex_l2(B, n, _ptr(v), _ptr(t), float(scale), _ptr(grad), _ptr(loss),
                                           v.device.index or 0,
                                           ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)))
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        g = ctx.grad
        if g is None:
            return None, None, None
        return g * d_loss.view(-1, *([1] * (g.dim() - 1))), None, None


def vertex_l2_loss(verts, target, scale=1.0):
    """Per-body squared-L2 vertex (or joint) data term, fused loss + gradient kernel.
    `vertex_l2_loss(v, t).sum().backward()` is the fitting step of BASELINE config 3."""
    return _VertexL2Fn.apply(verts, target, scale)


class _FitVertexL2Fn(torch.autograd.Function):
    """loss (B,) = scale * sum ||V(betas, pose, ...) - V*||^2 in one autograd node (BASELINE config 3).

    Forward: smplk_fit_vertex_l2 (pose kernel, blend GEMM, then skinning + loss + gradient +
    skinning backward as one kernel); the vertices never reach HBM, their gradient stays internal.
    Backward: the body-model backward on the stored gradient; the chain is linear in d_loss, so the
    pose backward kernel scales body b's parameter gradients by d_loss[b] (smplk_backward_args.d_loss)
    instead of a pass over the (B,V,3) vertex gradient, which the separate vertex_l2_loss node pays."""

    @staticmethod
    def forward(ctx, dm, flags, scale, target, betas, pose, pca_l, pca_r, transl):
        B = pose.shape[0]
        dev = pose.device
        flags |= _lib.FLAG_SAVE_FOR_BACKWARD | _lib.FLAG_FIT_VERTEX_L2
        betas_c, pose_c = _prep(betas, dev), _prep(pose, dev)
        pl, pr, tr = _prep(pca_l, dev), _prep(pca_r, dev), _prep(transl, dev)
        tgt = _prep(target, dev)
        _check_inputs(dm, B, betas_c, pose_c, pl, pr, tr)
        if tuple(tgt.shape) != (B, dm.V, 3):
            raise ValueError("target must be (%d, %d, 3), got %s" % (B, dm.V, tuple(tgt.shape)))
        if tgt.data_ptr() % 8:       # a view at an odd float offset: the kernel reads 8-byte pairs
            tgt = tgt.clone()
        verts = torch.empty(B, dm.V, 3, device=dev, dtype=torch.float32)
        ws_bytes = dm.workspace_bytes(B, flags)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        a = _lib.ForwardArgs()
        a.batch, a.flags = B, flags
        a.betas, a.betas_batch = _ptr(betas_c), (betas_c.shape[0] if betas_c is not None else 1)
        a.pose, a.hand_pca_l, a.hand_pca_r, a.transl = _ptr(pose_c), _ptr(pl), _ptr(pr), _ptr(tr)
        a.verts = _ptr(verts)
        a.workspace, a.workspace_bytes = _ptr(ws), ws_bytes
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        a.stream = stream
        loss = torch.empty((() if (flags & _lib.FLAG_LOSS_SUM) else (B,)), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            # `verts` comes back holding the vertex gradient 2 scale (V - V*), not the vertices
            dm.fit_vertex_l2(a, _ptr(tgt), scale, _ptr(loss))
        ctx.dm, ctx.flags = dm, flags
        ctx.ws, ctx.g = ws, verts
        ctx.inputs = (betas_c, pose_c, pl, pr, tr)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        if d_loss is None:
            return (None,) * 9
        dm = ctx.dm
        betas, pose, pl, pr, tr = ctx.inputs
        B = pose.shape[0]
        dev = pose.device
        need = ctx.needs_input_grad   # (dm, flags, scale, target, betas, pose, pca_l, pca_r, transl)
        d_loss = d_loss.contiguous().float()
        g = ctx.g
        d_betas = torch.empty_like(betas) if (betas is not None and need[4]) else None
        d_pose = torch.empty_like(pose) if need[5] else None
        d_pl = torch.empty_like(pl) if (pl is not None and need[6]) else None
        d_pr = torch.empty_like(pr) if (pr is not None and need[7]) else None
        d_tr = torch.empty_like(tr) if (tr is not None and need[8]) else None
        sc_bytes = dm.backward_scratch_bytes(B)
        sc = torch.empty(sc_bytes, device=dev, dtype=torch.uint8)
        a = _lib.BackwardArgs()
        a.batch, a.flags = B, ctx.flags
        a.betas, a.betas_batch = _ptr(betas), (betas.shape[0] if betas is not None else 1)
        a.pose, a.hand_pca_l, a.hand_pca_r = _ptr(pose), _ptr(pl), _ptr(pr)
        a.d_verts = _ptr(g)
        a.d_betas, a.d_pose = _ptr(d_betas), _ptr(d_pose)
        a.d_hand_pca_l, a.d_hand_pca_r, a.d_transl = _ptr(d_pl), _ptr(d_pr), _ptr(d_tr)
        a.workspace, a.workspace_bytes = _ptr(ctx.ws), ctx.ws.numel()
        a.scratch, a.scratch_bytes = _ptr(sc), sc_bytes
        a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        a.d_loss = _ptr(d_loss)      # the pose backward kernel scales body b's gradients by d_loss[b]
        with torch.cuda.device(dev):
            dm.backward(a)
        return None, None, None, None, d_betas, d_pose, d_pl, d_pr, d_tr


def fit_vertex_l2(dm, betas, pose, target, pca_l=None, pca_r=None, transl=None, add_pose_mean=False,
                  scale=1.0, flags=0, reduce="none"):
    """Per-body loss (B,) = scale * sum ||V - target||^2 of the posed vertices against target meshes,
    as ONE autograd node: body-model forward + fused loss/gradient kernel forward, body-model backward
    on the stored vertex gradient.  `fit_vertex_l2(...).sum().backward()` is the fitting step of
    BASELINE config 3; it equals `vertex_l2_loss(body_model_apply(...)[0], target)`.
    reduce="sum" returns the scalar sum over the bodies (accumulated inside the kernel)."""
    if add_pose_mean:
        flags |= _lib.FLAG_ADD_POSE_MEAN
    if reduce == "sum":       # the kernel accumulates one scalar: no separate sum / expand kernels around the node
        flags |= _lib.FLAG_LOSS_SUM
    elif reduce != "none":
        raise ValueError("reduce must be 'none' or 'sum'")
    return _FitVertexL2Fn.apply(dm, flags, scale, target, betas, pose, pca_l, pca_r, transl)


def body_model_apply(dm, betas, pose, pca_l=None, pca_r=None, transl=None, add_pose_mean=False,
                     want_regressed=False, flags=0, want_verts=True):
    """Functional entry point: (verts|None, joints_fk_plus_picks, joints_regressed|None, full_pose).
    want_verts=False (and no regressed joints): only the vertex-pick joints' vertices are blended
    and skinned (pick_forward_kernel), verts is None."""
    if add_pose_mean:
        flags |= _lib.FLAG_ADD_POSE_MEAN
    if _grad_flows(betas, pose, pca_l, pca_r, transl):
        flags |= _lib.FLAG_SAVE_FOR_BACKWARD
    v, j, r, fp = _BodyModelFn.apply(dm, flags, want_regressed, want_verts, betas, pose, pca_l, pca_r, transl)
    return (v if v.numel() > 0 else None), j, (r if r.numel() > 0 else None), fp


class _BodyModelBase(nn.Module):
    NUM_BODY_JOINTS = 23
    NUM_HAND_JOINTS = 0
    KIND = "smpl"

    def __init__(self, model_path=None, model=None, joint_mapper=None, create_betas=True,
                 betas=None, num_betas=None, create_global_orient=True, global_orient=None,
                 create_body_pose=True, body_pose=None, create_transl=True, transl=None,
                 dtype=torch.float32, batch_size=1, gender="neutral", J_regressor_extra=None,
                 joint_map=None, vertex_ids=None, device=None, ext="pkl", **kwargs):
        super().__init__()
        # float64 (lib/gen_smplh.py:66-67 float_dtype) is accepted at the interface: parameters, buffers and
        # outputs carry it, the kernels compute in float32 (inputs are cast down, outputs cast up)
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("Unknown float type %s, exiting!" % dtype)
        if model is None:
            if model_path is None:
                raise ValueError("give `model` (dict) or `model_path`")
            model = load_model_file(resolve_model_path(model_path, self.KIND, gender, ext))
        self.batch_size = batch_size
        self.gender = gender
        self.dtype = dtype
        self.joint_mapper = joint_mapper
        self._model_dict = model
        self._dm = {}
        self._dm_kwargs = {}
        self.faces = np.asarray(model["f"]).astype(np.int64)
        self.register_buffer("faces_tensor", torch.from_numpy(self.faces.copy()))
        self.register_buffer("v_template", torch.tensor(np.asarray(model["v_template"]), dtype=dtype))
        W = model["weights"]
        W = W.toarray() if hasattr(W, "toarray") else np.asarray(W)
        self.register_buffer("lbs_weights", torch.tensor(W, dtype=dtype))
        self.weigths = np.asarray(W)          # models/smplh.py:20 (sic)
        self.seg_index = {}                   # models/smplh.py:18
        sd = np.asarray(model["shapedirs"])
        self.num_betas = sd.shape[2] if num_betas is None else min(num_betas, sd.shape[2])
        self.parents = _lib.parents_from_model(model)
        self.num_joints = int(self.parents.shape[0])
        if vertex_ids is None:
            vertex_ids = model.get("extra_vertex_ids")
        self.extra_vertex_ids = None if vertex_ids is None else np.asarray(vertex_ids, np.int32)
        if J_regressor_extra is None:
            J_regressor_extra = kwargs.pop("regressor_extra", None)
        self._regressor_extra = None
        if J_regressor_extra is not None:
            self._regressor_extra = np.asarray(J_regressor_extra, np.float64)
            self.register_buffer("J_regressor_extra", torch.tensor(self._regressor_extra, dtype=dtype))
        # a buffer: follows .to(device) (a per-call H2D copy is illegal during CUDA-graph capture)
        self.register_buffer("joint_map", None if joint_map is None else
                             torch.as_tensor(np.asarray(joint_map), dtype=torch.long))

        def param(flag, value, shape, name):
            if not flag:
                return
            if value is None:
                t = torch.zeros(shape, dtype=dtype)
            else:
                t = torch.as_tensor(value, dtype=dtype).clone().reshape(shape)
            self.register_parameter(name, nn.Parameter(t, requires_grad=True))

        param(create_betas, betas, [batch_size, self.num_betas], "betas")
        param(create_global_orient, global_orient, [batch_size, 3], "global_orient")
        param(create_body_pose, body_pose, [batch_size, self.NUM_BODY_JOINTS * 3], "body_pose")
        param(create_transl, transl, [batch_size, 3], "transl")
        self._verts_cache = None
        if device is not None:
            self.to(device)

    # ---- device-side packed model, one per GPU (lazy: built on first forward on that device)
    def device_model(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        dm = self._dm.get(idx)
        if dm is None:
            dm = _lib.DeviceModel(self._model_dict, device=idx, num_betas=self.num_betas,
                                  extra_vertex_ids=self.extra_vertex_ids,
                                  regressor_posed=self._regressor_extra, **self._dm_kwargs)
            self._dm[idx] = dm
        return dm

    @property
    def verts_numpy(self):
        """models/smplh.py:38 caches vertices[0] on the host every forward (a device sync);
        here the copy happens only when the attribute is read."""
        if self._verts_cache is None:          # models/smplh.py:19: the template until the first forward
            return self.v_template.detach().cpu().numpy()
        return self._verts_cache[0].detach().cpu().numpy()

    @torch.no_grad()
    def reset_params(self, **params_dict):
        """upstream smplx reset_params (lib/Gen_SMPLH/fit_single_frame.py:283,361)."""
        for name, p in self.named_parameters():
            if name in params_dict:
                p[:] = torch.as_tensor(params_dict[name], dtype=p.dtype, device=p.device).reshape(p.shape)
            else:
                p.fill_(0)

    def _default(self, value, name, batch, cols, device):
        if value is None:
            value = getattr(self, name, None)
        if value is None:
            value = torch.zeros(batch, cols, dtype=self.dtype, device=device)
        return value

    @staticmethod
    def _f32(t):
        """Kernel-side view of an interface tensor (float64 interface: cast down; autograd casts the gradient up)."""
        return t if (t is None or t.dtype == torch.float32) else t.float()

    def _out(self, t):
        return t if (t is None or t.dtype == self.dtype) else t.to(self.dtype)

    @staticmethod
    def _expand(t, B):
        if t.shape[0] == B:
            return t
        if t.shape[0] == 1:
            return t.expand(B, -1)
        raise ValueError("batch mismatch: %d vs %d" % (t.shape[0], B))

    def _finish(self, verts, joints, jreg, transl_used):
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints)
        if jreg is not None:
            joints = torch.cat([joints, jreg], dim=1)
        if self.joint_map is not None:
            joints = joints[:, self.joint_map]
        return joints


class SMPL(_BodyModelBase):
    """24-joint SMPL; drop-in for models/smpl.py:11-37 (upstream smplx.SMPL + extra joints)."""
    NUM_BODY_JOINTS = 23
    KIND = "smpl"

    def _assemble(self, betas, body_pose, global_orient, transl):
        dev = self.v_template.device
        go = self._default(global_orient, "global_orient", self.batch_size, 3, dev)
        bp = self._default(body_pose, "body_pose", self.batch_size, 69, dev)
        be = self._default(betas, "betas", self.batch_size, self.num_betas, dev)
        tr = transl if transl is not None else getattr(self, "transl", None)
        B = max(go.shape[0], bp.shape[0], be.shape[0])
        pose = self._f32(torch.cat([self._expand(go, B), self._expand(bp, B)], dim=1))
        if tr is not None:
            tr = self._f32(self._expand(tr, B))
        return self.device_model(dev), go, bp, be, tr, pose

    def vertex_l2(self, target, scale=1.0, betas=None, body_pose=None, global_orient=None, transl=None, reduce="none"):
        """Per-body scale * sum ||vertices - target||^2 as one autograd node (see fit_vertex_l2)."""
        dm, go, bp, be, tr, pose = self._assemble(betas, body_pose, global_orient, transl)
        return self._out(fit_vertex_l2(dm, self._f32(be), pose, self._f32(target), transl=tr, scale=scale, reduce=reduce))

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                return_verts=True, return_full_pose=False, pose2rot=True, **kwargs):
        if not pose2rot:
            raise NotImplementedError("pose2rot=False (rotation-matrix input) is not supported")
        dm, go, bp, be, tr, pose = self._assemble(betas, body_pose, global_orient, transl)
        verts, joints, jreg, full_pose = body_model_apply(
            dm, self._f32(be), pose, transl=tr, want_regressed=self._regressor_extra is not None, want_verts=return_verts)
        self._verts_cache = None if verts is None else verts.detach()   # never keeps the autograd graph alive
        joints = self._finish(verts, joints, jreg, tr)
        return ModelOutput(vertices=self._out(verts) if return_verts else None, joints=self._out(joints),
                           full_pose=self._out(full_pose) if return_full_pose else None, betas=be,
                           global_orient=go, body_pose=bp)


class SMPLH(_BodyModelBase):
    """52-joint SMPL-H; drop-in for models/smplh.py:13-39 and for the object
    `smplx.create(model_type='smplh', ...)` returns at lib/gen_smplh.py:90."""
    NUM_BODY_JOINTS = 21
    NUM_HAND_JOINTS = 15
    KIND = "smplh"

    def __init__(self, model_path=None, model=None, create_left_hand_pose=True,
                 left_hand_pose=None, create_right_hand_pose=True, right_hand_pose=None,
                 use_pca=True, num_pca_comps=6, flat_hand_mean=False, **kwargs):
        dtype = kwargs.get("dtype", torch.float32)
        batch_size = kwargs.get("batch_size", 1)
        super().__init__(model_path=model_path, model=model, **kwargs)
        self.use_pca = use_pca
        self._pad_cache = {}
        self.num_pca_comps = num_pca_comps
        self.flat_hand_mean = flat_hand_mean
        m = self._model_dict
        if self.num_joints != 52:
            raise ValueError("SMPLH needs a 52-joint model (got %d joints)" % self.num_joints)
        self._dm_kwargs = dict(num_pca_comps=num_pca_comps if use_pca else 0,
                               flat_hand_mean=flat_hand_mean)
        if use_pca:
            self.register_buffer("left_hand_components", torch.tensor(
                np.asarray(m["hands_componentsl"])[:num_pca_comps], dtype=dtype))
            self.register_buffer("right_hand_components", torch.tensor(
                np.asarray(m["hands_componentsr"])[:num_pca_comps], dtype=dtype))
        ml = np.zeros(45) if flat_hand_mean else np.asarray(m["hands_meanl"], np.float64)
        mr = np.zeros(45) if flat_hand_mean else np.asarray(m["hands_meanr"], np.float64)
        self.register_buffer("left_hand_mean", torch.tensor(ml, dtype=dtype))
        self.register_buffer("right_hand_mean", torch.tensor(mr, dtype=dtype))
        self.register_buffer("pose_mean", torch.tensor(np.concatenate([np.zeros(66), ml, mr]), dtype=dtype))
        hand_dim = num_pca_comps if use_pca else 45
        for flag, val, name in ((create_left_hand_pose, left_hand_pose, "left_hand_pose"),
                                (create_right_hand_pose, right_hand_pose, "right_hand_pose")):
            if flag:
                t = (torch.zeros(batch_size, hand_dim, dtype=dtype) if val is None
                     else torch.as_tensor(val, dtype=dtype).clone().reshape(batch_size, hand_dim))
                self.register_parameter(name, nn.Parameter(t, requires_grad=True))
        dev = kwargs.get("device")
        if dev is not None:
            self.to(dev)

    def _assemble(self, betas, global_orient, body_pose, left_hand_pose, right_hand_pose, transl):
        dev = self.v_template.device
        hand_dim = self.num_pca_comps if self.use_pca else 45
        go = self._default(global_orient, "global_orient", self.batch_size, 3, dev)
        bp = self._default(body_pose, "body_pose", self.batch_size, 63, dev)
        be = self._default(betas, "betas", self.batch_size, self.num_betas, dev)
        lh = self._default(left_hand_pose, "left_hand_pose", self.batch_size, hand_dim, dev)
        rh = self._default(right_hand_pose, "right_hand_pose", self.batch_size, hand_dim, dev)
        tr = transl if transl is not None else getattr(self, "transl", None)
        B = max(go.shape[0], bp.shape[0], be.shape[0], lh.shape[0], rh.shape[0])
        go_e, bp_e = self._expand(go, B), self._expand(bp, B)
        lh_e, rh_e = self._expand(lh, B), self._expand(rh, B)
        if tr is not None:
            tr = self._f32(self._expand(tr, B))
        dm = self.device_model(dev)
        if self.use_pca:
            pad = self._pad_cache.get((B, str(dev)))
            if pad is None:       # hand slots of the axis-angle row (filled from the PCA coefficients in the kernel)
                pad = self._pad_cache[(B, str(dev))] = torch.zeros(B, 90, dtype=torch.float32, device=dev)
            pose = torch.cat([self._f32(go_e), self._f32(bp_e), pad], dim=1)
            pca_l, pca_r = self._f32(lh_e), self._f32(rh_e)
        else:
            pose = self._f32(torch.cat([go_e, bp_e, lh_e, rh_e], dim=1))
            pca_l = pca_r = None
        return dm, go, bp, be, lh, rh, tr, pose, pca_l, pca_r, (go_e, bp_e, lh_e, rh_e)

    def vertex_l2(self, target, scale=1.0, betas=None, global_orient=None, body_pose=None,
                  left_hand_pose=None, right_hand_pose=None, transl=None, reduce="none"):
        """Per-body scale * sum ||vertices - target||^2 as one autograd node (see fit_vertex_l2)."""
        dm, _, _, be, _, _, tr, pose, pca_l, pca_r, _ = self._assemble(
            betas, global_orient, body_pose, left_hand_pose, right_hand_pose, transl)
        return self._out(fit_vertex_l2(dm, self._f32(be), pose, self._f32(target), pca_l=pca_l, pca_r=pca_r, transl=tr,
                                       add_pose_mean=True, scale=scale, reduce=reduce))

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                right_hand_pose=None, transl=None, return_verts=True, return_full_pose=False,
                pose2rot=True, get_skin=True, **kwargs):
        if not pose2rot:
            raise NotImplementedError("pose2rot=False (rotation-matrix input) is not supported")
        dm, go, bp, be, lh, rh, tr, pose, pca_l, pca_r, (go_e, bp_e, lh_e, rh_e) = self._assemble(
            betas, global_orient, body_pose, left_hand_pose, right_hand_pose, transl)
        verts, joints, jreg, full_pose = body_model_apply(
            dm, self._f32(be), pose, pca_l=pca_l, pca_r=pca_r, transl=tr, add_pose_mean=True,
            want_regressed=self._regressor_extra is not None, want_verts=return_verts)
        self._verts_cache = None if verts is None else verts.detach()   # never keeps the autograd graph alive
        joints = self._finish(verts, joints, jreg, tr)
        # full_pose (PCA hands + pose mean applied) is an output of the pose kernel, differentiable through
        # the same autograd node (upstream assembles it with cat + einsum + add)
        if self.use_pca:
            # upstream returns the PROJECTED 45-D hand poses (einsum with the components, before the mean
            # is added); here they are the hand slots of full_pose minus the mean, one kernel for both hands
            hands = self._out(full_pose[:, 66:156]) - self.pose_mean[66:156]
            lh, rh = hands[:, :45], hands[:, 45:]
        return ModelOutput(vertices=self._out(verts) if return_verts else None, joints=self._out(joints),
                           full_pose=self._out(full_pose) if return_full_pose else None, betas=be,
                           global_orient=go, body_pose=bp, left_hand_pose=lh, right_hand_pose=rh)


def create(model_path=None, model_type="smplh", **kwargs):
    """Mirror of `smplx.create` as used at lib/gen_smplh.py:90."""
    kwargs.pop("create_expression", None)
    kwargs.pop("create_jaw_pose", None)
    kwargs.pop("create_leye_pose", None)
    kwargs.pop("create_reye_pose", None)
    # upstream: a folder holds one sub-folder per model type (<dir>/smplh/SMPLH_MALE.pkl)
    if model_path is not None and os.path.isdir(model_path) and os.path.isdir(os.path.join(model_path, model_type)):
        model_path = os.path.join(model_path, model_type)
    if model_type.lower() == "smpl":
        return SMPL(model_path=model_path, **kwargs)
    if model_type.lower() == "smplh":
        return SMPLH(model_path=model_path, **kwargs)
    raise ValueError("Unknown model type %s, exiting!" % model_type)
