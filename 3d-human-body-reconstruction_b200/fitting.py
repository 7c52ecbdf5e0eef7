"""Batched fitting loss around the body model as fused loss-and-gradient CUDA
kernels behind torch autograd Functions.

    reprojection_loss  <-> PerspectiveCamera.forward (lib/Gen_SMPLH/camera.py:93-117) + GMoF
                           (lib/Gen_SMPLH/util.py:60-71) + the joint term of SMPLifyLoss.forward
                           (lib/Gen_SMPLH/fitting.py:369-381); rho=0 -> SMPLifyCameraInitLoss
                           (fitting.py:486-495)
    fit_priors         <-> shape / pose / bending / hand priors (fitting.py:383-413, prior.py:53-97)
    SMPLifyLoss        <-> fitting.py:297-449 (same forward signature; interpenetration, face and
                           jaw terms are outside this path), without the batch_size == 1 restriction
                           of lib/Gen_SMPLH/fit_single_frame.py:97.

Losses are returned per body (B,); the reference's scalar is their sum.
"""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f(t):
    return None if t is None else t.contiguous().float()


def _launch_reproj(joints, translation, rotation, focal, center, gt, weights, rho, data_weight, loss):
    """Runs smplk_reprojection_loss into `loss` (B,); returns (d_joints, d_translation) for d_loss = 1."""
    B, Jn = joints.shape[0], joints.shape[1]
    dev = joints.device
    dj = torch.empty_like(joints)
    dt = torch.empty(B, 3, device=dev)
    a = _lib.ReprojArgs()
    a.batch, a.num_joints = B, Jn
    a.joints, a.rotation, a.translation = _ptr(joints), _ptr(rotation), _ptr(translation)
    a.focal, a.center, a.camera_batch = _ptr(focal), _ptr(center), translation.shape[0]
    a.gt_joints, a.weights = _ptr(gt), _ptr(weights)
    a.weights_batch = weights.shape[0] if weights is not None else 1
    a.rho, a.data_weight = float(rho), float(data_weight)
    a.loss, a.d_joints, a.d_translation = _ptr(loss), _ptr(dj), _ptr(dt)
    a.device = dev.index or 0
    a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.load().smplk_reprojection_loss(ctypes.byref(a)))
    return dj, dt


def _prior_rows(ts, B=None):
    """The kernel strides every prior input over B rows.  Inputs with one row (shared betas,
    upstream lbs batch_size = max(...)) are expanded; anything else must have B rows.
    Returns (B, expanded inputs, per-input `was broadcast` flags)."""
    live = [t for t in ts if t is not None]
    if B is None:
        B = max(t.shape[0] for t in live)
    out, bc = [], []
    for t in ts:
        if t is None:
            out.append(None)
            bc.append(False)
            continue
        if t.dim() != 2 or t.shape[0] not in (1, B):
            raise ValueError("prior inputs must be (1|%d, n), got %s" % (B, tuple(t.shape)))
        bc.append(t.shape[0] == 1 and B > 1)
        out.append(t.expand(B, -1).contiguous() if bc[-1] else t)
    return B, out, bc


def _launch_priors(ts, shape_w, pose_w, bend_w, hand_w, loss):
    """Runs smplk_fit_priors into `loss` (B,) on inputs that all have B rows; returns the gradient
    list (None where the input is None)."""
    ref = next(t for t in ts if t is not None)
    B, dev = ref.shape[0], ref.device
    grads = [None if t is None else torch.empty_like(t) for t in ts]
    a = _lib.PriorArgs()
    a.batch = B
    a.betas, a.num_betas = _ptr(ts[0]), (ts[0].shape[1] if ts[0] is not None else 0)
    a.pose_embedding, a.num_embedding = _ptr(ts[1]), (ts[1].shape[1] if ts[1] is not None else 0)
    a.body_pose, a.num_body_pose = _ptr(ts[2]), (ts[2].shape[1] if ts[2] is not None else 0)
    a.left_hand_pose, a.right_hand_pose = _ptr(ts[3]), _ptr(ts[4])
    if ts[3] is not None and ts[4] is not None and ts[3].shape[1] != ts[4].shape[1]:
        raise ValueError("left / right hand pose widths differ: %d vs %d" % (ts[3].shape[1], ts[4].shape[1]))
    a.num_hand = ts[3].shape[1] if ts[3] is not None else (ts[4].shape[1] if ts[4] is not None else 0)
    a.shape_weight, a.body_pose_weight = float(shape_w), float(pose_w)
    a.bending_prior_weight, a.hand_prior_weight = float(bend_w), float(hand_w)
    a.loss = _ptr(loss)
    a.d_betas, a.d_pose_embedding, a.d_body_pose = _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2])
    a.d_left_hand_pose, a.d_right_hand_pose = _ptr(grads[3]), _ptr(grads[4])
    a.device = dev.index or 0
    a.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.load().smplk_fit_priors(ctypes.byref(a)))
    return grads


class _Reproj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, joints, translation, rotation, focal, center, gt, weights, rho, data_weight):
        if not joints.is_cuda:
            raise RuntimeError("smplk fitting losses need CUDA tensors (no CPU fallback)")
        joints, translation, rotation = _f(joints), _f(translation), _f(rotation)
        focal, center, gt, weights = _f(focal), _f(center), _f(gt), _f(weights)
        loss = torch.empty(joints.shape[0], device=joints.device)
        with torch.cuda.device(joints.device):
            dj, dt = _launch_reproj(joints, translation, rotation, focal, center, gt, weights, rho, data_weight, loss)
        ctx.save_for_backward(dj, dt)
        ctx.cam_batch = translation.shape[0]
        return loss

    @staticmethod
    def backward(ctx, g):
        dj, dt = ctx.saved_tensors
        gj = dj * g[:, None, None]
        gt = dt * g[:, None]
        if ctx.cam_batch == 1:
            gt = gt.sum(0, keepdim=True)
        return gj, gt, None, None, None, None, None, None, None


def reprojection_loss(joints, rotation, translation, focal, center, gt_joints, weights=None, rho=100.0,
                      data_weight=1.0):
    """(B,) robust reprojection loss; differentiable w.r.t. `joints` (B,Jn,3) and the camera
    `translation` ((1,3) shared or (B,3)).  focal / center: (1|B, 2)."""
    return _Reproj.apply(joints, translation, rotation, focal, center, gt_joints, weights, rho, data_weight)


class _Priors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, betas, emb, body_pose, lh, rh, shape_w, pose_w, bend_w, hand_w):
        ts = [betas, emb, body_pose, lh, rh]
        ref = next(t for t in ts if t is not None)
        if not ref.is_cuda:
            raise RuntimeError("smplk fitting losses need CUDA tensors (no CPU fallback)")
        B, ts, ctx.bc = _prior_rows([_f(t) for t in ts])
        loss = torch.empty(B, device=ref.device)
        with torch.cuda.device(ref.device):
            ctx.grads = _launch_priors(ts, shape_w, pose_w, bend_w, hand_w, loss)
        return loss

    @staticmethod
    def backward(ctx, g):
        out = [None if d is None else d * g[:, None] for d in ctx.grads]
        out = [o.sum(0, keepdim=True) if (o is not None and bc) else o for o, bc in zip(out, ctx.bc)]
        return tuple(out) + (None, None, None, None)


def fit_priors(betas=None, pose_embedding=None, body_pose=None, left_hand_pose=None, right_hand_pose=None,
               shape_weight=0.0, body_pose_weight=0.0, bending_prior_weight=0.0, hand_prior_weight=0.0):
    """(B,) prior loss; `body_pose` is full_pose[:, 3:66].  One-row inputs (shared betas) broadcast."""
    return _Priors.apply(betas, pose_embedding, body_pose, left_hand_pose, right_hand_pose,
                         shape_weight, body_pose_weight, bending_prior_weight, hand_prior_weight)


class _SMPLifyTotal(torch.autograd.Function):
    """Scalar total of SMPLifyLoss (data term + priors, lib/Gen_SMPLH/fitting.py:365-449) as ONE autograd
    node: the two loss kernels write into one (2,B) buffer summed by a single reduction, and backward
    scales every stored gradient by the upstream scalar with one multi-tensor kernel -- instead of two
    nodes, three reductions / adds and six element-wise kernels (the closure at batch 1 is made of such
    2-3 us kernels)."""

    @staticmethod
    def forward(ctx, joints, translation, betas, emb, body_pose, lh, rh, rotation, focal, center, gt, weights,
                rho, data_weight, shape_w, pose_w, bend_w, hand_w):
        if not joints.is_cuda:
            raise RuntimeError("smplk fitting losses need CUDA tensors (no CPU fallback)")
        joints, translation, rotation = _f(joints), _f(translation), _f(rotation)
        focal, center, gt, weights = _f(focal), _f(center), _f(gt), _f(weights)
        B, dev = joints.shape[0], joints.device
        # the batch is the joints' (shared betas of shape (1,NB) broadcast over it)
        _, ts, ctx.bc = _prior_rows([_f(t) for t in (betas, emb, body_pose, lh, rh)], B)
        both = torch.empty(2, B, device=dev)
        with torch.cuda.device(dev):
            dj, dt = _launch_reproj(joints, translation, rotation, focal, center, gt, weights, rho, data_weight, both[0])
            if any(t is not None for t in ts):
                pg = _launch_priors(ts, shape_w, pose_w, bend_w, hand_w, both[1])
            else:
                pg = [None] * 5
                both[1].zero_()
        ctx.cam_batch = translation.shape[0]
        ctx.grads = [dj, dt] + pg
        return both.sum()

    @staticmethod
    def backward(ctx, g):
        live = [t for t in ctx.grads if t is not None]
        scaled = iter(torch._foreach_mul(live, g))
        out = [None if t is None else next(scaled) for t in ctx.grads]
        if ctx.cam_batch == 1 and out[1].shape[0] != 1:
            out[1] = out[1].sum(0, keepdim=True)
        for i, bc in enumerate(ctx.bc):
            if bc and out[2 + i] is not None:
                out[2 + i] = out[2 + i].sum(0, keepdim=True)
        return tuple(out) + (None,) * 11


class PerspectiveCamera(torch.nn.Module):
    """Parameter holder with the attributes of lib/Gen_SMPLH/camera.py:52-117."""

    def __init__(self, rotation=None, translation=None, focal_length_x=5000.0, focal_length_y=5000.0,
                 center=None, batch_size=1, device="cuda"):
        super().__init__()
        self.batch_size = batch_size
        rot = torch.eye(3).repeat(batch_size, 1, 1) if rotation is None else torch.as_tensor(rotation, dtype=torch.float32)
        tr = torch.zeros(batch_size, 3) if translation is None else torch.as_tensor(translation, dtype=torch.float32)
        self.rotation = torch.nn.Parameter(rot.to(device), requires_grad=False)
        self.translation = torch.nn.Parameter(tr.to(device), requires_grad=True)
        self.register_buffer("focal", torch.tensor([[focal_length_x, focal_length_y]], dtype=torch.float32).repeat(batch_size, 1).to(device))
        c = torch.zeros(batch_size, 2) if center is None else torch.as_tensor(center, dtype=torch.float32).reshape(-1, 2)
        self.register_buffer("center", c.to(device))


class SMPLifyLoss(torch.nn.Module):
    """lib/Gen_SMPLH/fitting.py:297-449 with the same forward signature; returns the summed loss."""

    def __init__(self, rho=100.0, data_weight=1.0, body_pose_weight=0.0, shape_weight=0.0,
                 bending_prior_weight=0.0, hand_prior_weight=0.0, use_joints_conf=True, use_hands=True, **kwargs):
        super().__init__()
        self.rho, self.data_weight = rho, data_weight
        self.body_pose_weight, self.shape_weight = body_pose_weight, shape_weight
        self.bending_prior_weight, self.hand_prior_weight = bending_prior_weight, hand_prior_weight
        self.use_joints_conf, self.use_hands = use_joints_conf, use_hands

    def reset_loss_weights(self, loss_weight_dict):
        for k, v in loss_weight_dict.items():
            if hasattr(self, k):
                setattr(self, k, float(v))

    def forward(self, body_model_output, camera, gt_joints, joints_conf, body_model_faces=None,
                joint_weights=None, use_vposer=False, pose_embedding=None, **kwargs):
        w = joint_weights * joints_conf if self.use_joints_conf else joint_weights
        body_pose = body_model_output.full_pose[:, 3:66]
        return _SMPLifyTotal.apply(
            body_model_output.joints, camera.translation, body_model_output.betas,
            pose_embedding if use_vposer else None, body_pose,
            body_model_output.left_hand_pose if self.use_hands else None,
            body_model_output.right_hand_pose if self.use_hands else None,
            camera.rotation, camera.focal, camera.center, gt_joints, w, self.rho, self.data_weight,
            self.shape_weight, self.body_pose_weight, self.bending_prior_weight,
            self.hand_prior_weight if self.use_hands else 0.0)


class GraphedClosure:
    """A fitting closure captured once into a CUDA graph and replayed.

    The reference's closure (lib/Gen_SMPLH/fitting.py:230-262, `create_fitting_closure`: zero_grad ->
    body_model(...) -> loss -> backward, called by the optimiser at batch size 1,
    fit_single_frame.py:97) is launch-bound on this path: ~15 short kernels, 0.24 ms from Python
    against 0.13 ms of device time.  `GraphedClosure(fn, params)` runs `fn()` (which must build the
    scalar loss from `params` and whatever static tensors it closes over) a few times eagerly,
    captures `loss = fn(); loss.backward()` and afterwards replays that graph on every call:

        closure = GraphedClosure(lambda: loss_fn(model(return_verts=True, return_full_pose=True), cam, gt, conf,
                                                 joint_weights=jw), model.parameters())
        for _ in range(n):
            optimiser.step(closure)

    Parameters are updated in place by the optimiser, so the replay sees their new values; the
    gradients land in the same `.grad` tensors every time (re-attached on each call, so
    `zero_grad(set_to_none=True)` between calls is harmless).  Inputs other than the parameters must be
    changed in place (`tensor.copy_`), never rebound.  Shapes are frozen at capture."""

    def __init__(self, fn, params, warmup=3):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GraphedClosure needs at least one parameter that requires grad")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedClosure needs CUDA parameters (smplk has no CPU path)")
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):
                for p in self.params:
                    p.grad = None
                fn().backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = fn()
            self.loss.backward()
        self.grads = [p.grad for p in self.params]

    def __call__(self):
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        return self.loss
