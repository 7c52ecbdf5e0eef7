"""Mesh operations either side of the body-model forward, on torch CUDA
tensors through the C ABI:

    inverse_lbs / inverse_joints  <-> RecoverModel.to_T_pose, to_rest_pose
                                      (lib/mesh2smpl_model.py:183-207, :340-372), models/smpl_np.py:239-246
    MeshTopology.vertex_normals   <-> VertNormals(verts, faces, True)  (utils/render_model.py:36,63-81)
    MeshTopology.divide_face      <-> SMPLHModel.divide_face           (models/smplh_np.py:126-182)

There is no CPU fallback: non-CUDA tensors raise.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("smplk mesh ops need CUDA tensors (no CPU fallback)")


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def transforms(dm, betas, pose, transl=None):
    """Skinning transforms A (B,J,3,4) and FK joints (B,J,3) of a pose: the pose/FK kernel only
    (models/smplh_np.py:49-78 compute_R_G + rest-pose removal)."""
    _check_cuda(betas, pose, transl)
    B = pose.shape[0]
    dev = pose.device
    flags = _lib.FLAG_SAVE_FOR_BACKWARD | _lib.FLAG_TRANSFORMS_ONLY
    ws = torch.empty(dm.workspace_bytes(B, flags), device=dev, dtype=torch.uint8)
    joints = torch.empty(B, dm.J + dm.E, 3, device=dev)
    a = _lib.ForwardArgs()
    a.batch, a.flags = B, flags
    if betas is not None:
        betas = betas.contiguous().float()
        a.betas, a.betas_batch = _ptr(betas), betas.shape[0]
    else:
        a.betas_batch = 1
    pose = pose.contiguous().float()
    a.pose = _ptr(pose)
    if transl is not None:
        transl = transl.contiguous().float()
        a.transl = _ptr(transl)
    a.joints = _ptr(joints)
    a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
    a.stream = _stream(dev)
    dm.forward(a)
    off = dm.workspace_layout(B, flags)["A"]
    A = ws[off:off + B * dm.J * 48].view(torch.float32).view(B, dm.J, 3, 4).clone()
    return A, joints[:, :dm.J]


def remove_rest(G, joints_rest):
    """A (B,J,3,4) = [G_R | G_t - G_R J] from global transforms G (B,J,4,4) and rest joints (B,J,3): the
    `G - pack(G . [J;0])` step of lib/mesh2smpl_model.py:194-199 / models/smplh_np.py:73-78."""
    _check_cuda(G, joints_rest)
    B, J = G.shape[0], G.shape[1]
    if tuple(G.shape) != (B, J, 4, 4) or tuple(joints_rest.shape) != (B, J, 3):
        raise ValueError("G must be (B,J,4,4) and joints_rest (B,J,3)")
    G = G.contiguous().float()
    joints_rest = joints_rest.contiguous().float()
    A = torch.empty(B, J, 3, 4, device=G.device)
    _lib.check(_lib.load().smplk_remove_rest(B, J, _ptr(G), _ptr(joints_rest), _ptr(A), G.device.index or 0,
                                             _stream(G.device)))
    return A


def inverse_lbs(dm, A, verts, transl=None):
    """v_rest = (W.A)^-1 [verts - transl; 1] per vertex, with the skin weights of `dm`."""
    _check_cuda(A, verts, transl)
    B = verts.shape[0]
    A = A.contiguous().float()
    verts = verts.contiguous().float()
    transl = None if transl is None else transl.contiguous().float()
    out = torch.empty_like(verts)
    lib = _lib.load()
    _lib.check(lib.smplk_inverse_lbs(dm.handle, B, _ptr(A), _ptr(verts), _ptr(transl), _ptr(out), _stream(verts.device)))
    return out


def inverse_joints(A, joints, transl=None):
    """J_rest = A_j^-1 [J_posed - transl; 1]  (lib/mesh2smpl_model.py:205-207)."""
    _check_cuda(A, joints, transl)
    B, J = joints.shape[0], joints.shape[1]
    A = A.contiguous().float()
    joints = joints.contiguous().float()
    transl = None if transl is None else transl.contiguous().float()
    out = torch.empty(B, J, 3, device=joints.device)
    lib = _lib.load()
    _lib.check(lib.smplk_inverse_joints(B, J, _ptr(A), _ptr(joints), 3 * J, _ptr(transl), _ptr(out),
                                        joints.device.index or 0, _stream(joints.device)))
    return out


class MeshTopology:
    """Faces of a mesh on one GPU, with the vertex -> incident-face lists the normal kernel gathers."""

    def __init__(self, faces, num_verts, device=0):
        f = np.ascontiguousarray(np.asarray(faces), dtype=np.int32).reshape(-1, 3)
        if f.min() < 0 or f.max() >= num_verts:
            raise ValueError("face index out of range")
        self.F, self.V = f.shape[0], int(num_verts)
        order = np.argsort(f.ravel(), kind="stable")
        counts = np.bincount(f.ravel(), minlength=self.V)
        ptr = np.zeros(self.V + 1, dtype=np.int32)
        ptr[1:] = np.cumsum(counts)
        self.device = torch.device("cuda", device)
        self.faces = torch.tensor(f, device=self.device)
        self.vf_ptr = torch.tensor(ptr, device=self.device)
        self.vf_face = torch.tensor((order // 3).astype(np.int32), device=self.device)

    def vertex_normals(self, verts):
        _check_cuda(verts)
        verts = verts.contiguous().float()
        B = verts.shape[0]
        out = torch.empty_like(verts)
        lib = _lib.load()
        _lib.check(lib.smplk_vertex_normals(B, self.V, _ptr(self.faces), _ptr(self.vf_ptr), _ptr(self.vf_face),
                                            _ptr(verts), _ptr(out), self.device.index, _stream(self.device)))
        return out

    def divide_face(self, verts):
        """Per body: (front_face, front_verts, front_verts_index, back_face, back_verts,
        back_verts_index) exactly as models/smplh_np.py:126-182 returns them (torch tensors)."""
        _check_cuda(verts)
        verts = verts.contiguous().float()
        B = verts.shape[0]
        fo = torch.empty(B, 2, self.F, 3, device=self.device, dtype=torch.int32)
        vo = torch.empty(B, 2, self.V, device=self.device, dtype=torch.int32)
        cnt = torch.empty(B, 2, 2, device=self.device, dtype=torch.int32)
        lib = _lib.load()
        _lib.check(lib.smplk_divide_faces(B, self.V, self.F, _ptr(self.faces), _ptr(verts), _ptr(fo), _ptr(vo),
                                          _ptr(cnt), self.device.index, _stream(self.device)))
        c = cnt.cpu().numpy()
        res = []
        for b in range(B):
            item = []
            for side in (0, 1):
                nf, nv = int(c[b, side, 0]), int(c[b, side, 1])
                idx = vo[b, side, :nv].long()
                item += [fo[b, side, :nf], verts[b, idx], idx]
            res.append(tuple(item))
        return res
