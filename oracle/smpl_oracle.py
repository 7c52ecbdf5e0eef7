"""CPU oracle for the SMPL / SMPL-H body-model forward and backward.

TEST INFRASTRUCTURE ONLY. Nothing under `oracle/` is imported by the product package;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it, and only as the checker or the timed CPU baseline.

Parity pinning: the reference ships NO tests and NO golden outputs for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the reference's own
importable numpy twins (`/root/reference/models/smplh_np.py`, `models/smpl_np.py`) and the
LBS-only `RecoverModel` math (`lib/model2video.py:42-85`) executed in the build container on
seeded synthetic models; the vectors are committed under `tests/golden/` together with the
generating script `oracle/make_golden.py`. The torch path of the reference
(`models/smplh.py:26-39` -> third-party `smplx`, un-vendored, version unpinned, API era
<= 0.1.13) cannot be imported offline; `torch_forward` below restates its published
algorithm (smplx.lbs.lbs / batch_rodrigues / batch_rigid_transform, VertexJointSelector,
joint_mapper, transl) and is cross-checked against the numpy twin restatement.

Two restatements:
  * numpy float64, one body per call  -- follows models/smplh_np.py:49-117 line by line
  * torch, batched, differentiable    -- follows upstream smplx `lbs`
"""
from collections import namedtuple

import numpy as np

try:  # torch is only needed for the batched / autograd oracle
    import torch
except Exception:  # pragma: no cover
    torch = None


# ----------------------------------------------------------------------------------------
# numpy float64 twin restatement (one body)
# ----------------------------------------------------------------------------------------

def np_parents(kintree_table):
    """models/smplh_np.py:19-23 -- parent index per joint from the (2,J) kintree table."""
    kt = np.asarray(kintree_table).astype(np.int64)
    col = {int(kt[1, i]): i for i in range(kt.shape[1])}
    return [-1] + [col[int(kt[0, i])] for i in range(1, kt.shape[1])]


def np_rodrigues(r):
    """models/smplh_np.py:88-109 -- axis-angle (J,3) -> (J,3,3); theta clamped to `tiny`."""
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    theta = np.sqrt((r * r).sum(1))
    theta = np.maximum(theta, np.finfo(np.float64).tiny)
    n = r / theta[:, None]
    c = np.cos(theta)[:, None, None]
    s = np.sin(theta)[:, None, None]
    K = np.zeros((r.shape[0], 3, 3))
    K[:, 0, 1] = -n[:, 2]
    K[:, 0, 2] = n[:, 1]
    K[:, 1, 0] = n[:, 2]
    K[:, 1, 2] = -n[:, 0]
    K[:, 2, 0] = -n[:, 1]
    K[:, 2, 1] = n[:, 0]
    outer = n[:, :, None] * n[:, None, :]
    return c * np.eye(3)[None] + (1.0 - c) * outer + s * K


def np_chain(R, J, parents):
    """models/smplh_np.py:60-70 -- world transforms G (J,4,4) of the kinematic chain."""
    nj = R.shape[0]
    G = np.zeros((nj, 4, 4))
    G[:, 3, 3] = 1.0
    G[0, :3, :3] = R[0]
    G[0, :3, 3] = J[0]
    for i in range(1, nj):
        p = parents[i]
        L = np.eye(4)
        L[:3, :3] = R[i]
        L[:3, 3] = J[i] - J[p]
        G[i] = G[p] @ L
    return G


def np_remove_rest(G, J):
    """models/smplh_np.py:73-78 -- A_i = G_i - [0 | G_i [J_i;0]]."""
    A = G.copy()
    A[:, :3, 3] = G[:, :3, 3] - np.einsum("jab,jb->ja", G[:, :3, :3], J)
    return A


def np_forward(model, pose=None, beta=None, trans=None):
    """Full forward of models/smplh_np.py:39-86 (== models/smpl_np.py:158-206) in float64.

    Returns dict(verts, J, R, G, A, v_shaped, v_posed, joints_fk).
    `joints_fk` (= G[:, :3, 3] + trans) is what the torch path returns as joints; the numpy
    twin itself only offers gen_J_3d (see np_gen_J_3d)."""
    parents = np_parents(model["kintree_table"])
    nj = len(parents)
    shapedirs = np.asarray(model["shapedirs"], dtype=np.float64)
    nb = shapedirs.shape[2]
    pose = np.zeros((nj, 3)) if pose is None else np.asarray(pose, np.float64).reshape(nj, 3)
    beta = np.zeros(nb) if beta is None else np.asarray(beta, np.float64).reshape(-1)
    trans = np.zeros(3) if trans is None else np.asarray(trans, np.float64).reshape(3)
    v_shaped = shapedirs[:, :, :beta.shape[0]].dot(beta) + np.asarray(model["v_template"], np.float64)  # :50
    Jreg = model["J_regressor"]
    J = Jreg.dot(v_shaped)                                                        # :51
    J = np.asarray(J)
    R = np_rodrigues(pose)                                                        # :52-53
    feat = (R[1:] - np.eye(3)[None]).ravel()                                      # :54-58
    v_posed = v_shaped + np.asarray(model["posedirs"], np.float64).dot(feat)      # :59
    G = np_chain(R, J, parents)                                                   # :60-70
    A = np_remove_rest(G, J)                                                      # :73-78
    T = np.tensordot(np.asarray(model["weights"], np.float64), A, axes=[[1], [0]])  # :79
    vh = np.concatenate([v_posed, np.ones((v_posed.shape[0], 1))], 1)             # :80
    verts = np.einsum("vab,vb->va", T, vh)[:, :3] + trans[None]                   # :81-82
    return dict(verts=verts, J=J, R=R, G=G, A=A, v_shaped=v_shaped, v_posed=v_posed,
                joints_fk=G[:, :3, 3] + trans[None])


def np_gen_J_3d(model, verts):
    """models/smplh_np.py:116-117 -- joints regressed from the POSED vertices."""
    return np.asarray(model["J_regressor"].dot(verts))


def np_lbs_only(rig, pose, trans=None, ignore_joints=(13, 14, 22, 23)):
    """lib/model2video.py:42-81 (== lib/mesh2smpl_model.py:268-309): skinning of a rigged mesh
    with FIXED rest joints, no blendshapes; joints in `ignore_joints` are zeroed (:44-45)."""
    kt = np.asarray(rig["kintree_table"])
    nj = kt.shape[1]
    parent = rig.get("parent")
    if parent is None:
        parents = np_parents(kt)
    else:
        parents = [-1] + [int(parent[i]) for i in range(1, nj)]
    pose = np.array(pose, dtype=np.float64).reshape(nj, 3)
    for j in ignore_joints:
        pose[j] = 0.0
    trans = np.zeros(3) if trans is None else np.asarray(trans, np.float64).reshape(3)
    J = np.asarray(rig["J"], np.float64)
    R = np_rodrigues(pose)
    G = np_chain(R, J, parents)
    A = np_remove_rest(G, J)
    T = np.tensordot(np.asarray(rig["weights"], np.float64), A, axes=[[1], [0]])
    vt = np.asarray(rig["v_template"], np.float64)
    vh = np.concatenate([vt, np.ones((vt.shape[0], 1))], 1)
    verts = np.einsum("vab,vb->va", T, vh)[:, :3] + trans[None]
    return dict(verts=verts, G=G, A=A, R=R)


def np_inverse_lbs(rig_weights, A, verts):
    """lib/mesh2smpl_model.py:183-207 / models/smpl_np.py:239-246: un-pose vertices with the
    per-vertex inverse of T = W.A (4x4)."""
    T = np.tensordot(np.asarray(rig_weights, np.float64), A, axes=[[1], [0]])
    Tinv = np.linalg.inv(T)
    vh = np.concatenate([verts, np.ones((verts.shape[0], 1))], 1)
    return np.einsum("vab,vb->va", Tinv, vh)[:, :3]


def np_inverse_joints(A, joints):
    """lib/mesh2smpl_model.py:205-207: J = inv(G) [or_J; 1] with G the rest-removed transforms."""
    A = np.asarray(A, np.float64)
    if A.shape[-2:] == (4, 4):
        A = A[:, :3, :]
    A4 = np.zeros((A.shape[0], 4, 4))
    A4[:, :3, :] = A.reshape(-1, 3, 4)
    A4[:, 3, 3] = 1.0
    jh = np.concatenate([np.asarray(joints, np.float64), np.ones((joints.shape[0], 1))], 1)
    return np.einsum("jab,jb->ja", np.linalg.inv(A4), jh)[:, :3]


def np_vertex_normals(verts, faces):
    """utils/render_model.py:36 VertNormals(verts, faces, True) [upstream opendr, unpinned]: per-vertex
    sum of the incident triangles' (un-normalised, i.e. area-weighted) cross products
    (v1-v0)x(v2-v0), normalised to unit length."""
    v = np.asarray(verts, np.float64)
    f = np.asarray(faces, np.int64)
    tn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v)
    for c in range(3):
        np.add.at(n, f[:, c], tn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    return n / np.where(ln > 0, ln, 1.0)


def np_divide_face(verts, faces):
    """models/smplh_np.py:126-182 divide_face, restated with the same sequential semantics:
    z = m0*n1 - n0*m1 with m = v1-v0, n = v2-v1 (:149-152); z <= 0 -> front (:155), else back;
    vertices re-indexed in order of first appearance (:157-163)."""
    v = np.asarray(verts)
    f = np.asarray(faces)
    out = []
    v0, v1, v2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    m, n = v1 - v0, v2 - v1
    z = m[:, 0] * n[:, 1] - n[:, 0] * m[:, 1]
    for side_mask in (z <= 0, z > 0):
        index_of = {}
        order = []
        new_faces = []
        for tri in f[side_mask]:
            row = []
            for idx in tri:
                idx = int(idx)
                if idx not in index_of:
                    index_of[idx] = len(order)
                    order.append(idx)
                row.append(index_of[idx])
            new_faces.append(row)
        order = np.asarray(order, dtype=np.int64)
        out += [np.asarray(new_faces, dtype=np.int64).reshape(-1, 3), v[order], order]
    return tuple(out)


# ----------------------------------------------------------------------------------------
# torch batched restatement of upstream smplx (differentiable)
# ----------------------------------------------------------------------------------------

OracleOutput = namedtuple("OracleOutput", ["vertices", "joints", "full_pose", "v_posed", "A",
                                           "joints_fk", "left_hand_pose", "right_hand_pose"])
OracleOutput.__new__.__defaults__ = (None, None)


def torch_rodrigues_quat(theta):
    """utils/geometry.py:9-45 -- quaternion route; eps 1e-8 added to the VECTOR before the norm."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    n = theta / angle
    half = angle * 0.5
    q = torch.cat([torch.cos(half), torch.sin(half) * n], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
            2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
            2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def torch_rodrigues(rot_vecs):
    """[upstream-smplx] lbs.batch_rodrigues -- K-matrix route: R = I + sin K + (1-cos) K^2,
    angle = || r + 1e-8 ||."""
    n_ = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    c = torch.cos(angle)[:, None]
    s = torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(n_, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def torch_lbs(betas, full_pose, v_template, shapedirs, posedirs, J_regressor, parents,
              lbs_weights):
    """[upstream-smplx] lbs.lbs with pose2rot=True.

    shapedirs (V,3,NB); posedirs (P,3V) [= reshape(pkl.posedirs,(3V,P)).T]; J_regressor (J,V);
    lbs_weights (V,J); parents list, parents[0] = -1.
    Returns verts (B,V,3), J_transformed (B,J,3), v_posed (B,V,3), A (B,J,4,4)."""
    B = max(betas.shape[0], full_pose.shape[0])
    nj = J_regressor.shape[0]
    V = v_template.shape[0]
    dt = betas.dtype
    v_shaped = v_template[None] + torch.einsum("bl,mkl->bmk", betas, shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)
    R = torch_rodrigues(full_pose.reshape(-1, 3)).view(B, nj, 3, 3)
    eye = torch.eye(3, dtype=dt, device=betas.device)
    feat = (R[:, 1:] - eye).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(feat, posedirs).view(B, V, 3)
    if J.shape[0] != B:
        J = J.expand(B, -1, -1)
    # batch_rigid_transform
    rel = J.clone()
    par = torch.as_tensor(parents[1:], dtype=torch.long, device=betas.device)
    rel = torch.cat([J[:, :1], J[:, 1:] - J[:, par]], dim=1)
    L = torch.zeros(B, nj, 4, 4, dtype=dt, device=betas.device)
    L[:, :, :3, :3] = R
    L[:, :, :3, 3] = rel
    L[:, :, 3, 3] = 1.0
    chain = [L[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[parents[i]], L[:, i]))
    G = torch.stack(chain, dim=1)
    J_tr = G[:, :, :3, 3]
    Jh = torch.cat([J, torch.zeros(B, nj, 1, dtype=dt, device=betas.device)], dim=2)
    corr = torch.matmul(G, Jh[..., None])[..., 0]                   # (B,J,4)
    A = G.clone()
    A = torch.cat([G[..., :3], (G[..., 3] - corr)[..., None]], dim=-1)
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.view(B, nj, 16)).view(B, V, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, V, 1, dtype=dt, device=betas.device)], dim=2)
    verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
    return verts, J_tr, v_posed, A


class TorchOracleModel:
    """Holds the model tensors in the layout upstream smplx registers them, plus the extras of
    models/smplh.py:16-24 (J_regressor_extra, joint_map)."""

    def __init__(self, model, dtype=None, num_pca_comps=12, flat_hand_mean=False,
                 joint_map=None, joint_mapper=None, device="cpu"):
        dtype = dtype or torch.float32
        self.dtype = dtype
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dtype, device=device)
        self.v_template = t(model["v_template"])
        self.shapedirs = t(model["shapedirs"])
        V = self.v_template.shape[0]
        pd = np.asarray(model["posedirs"], dtype=np.float64)
        self.posedirs = t(pd.reshape(V * 3, -1).T)
        Jr = model["J_regressor"]
        Jr = Jr.toarray() if hasattr(Jr, "toarray") else np.asarray(Jr)
        self.J_regressor = t(Jr)
        self.lbs_weights = t(model["weights"])
        self.parents = np_parents(model["kintree_table"])
        self.nj = len(self.parents)
        self.extra_vertex_ids = None
        if model.get("extra_vertex_ids") is not None:
            self.extra_vertex_ids = torch.as_tensor(np.asarray(model["extra_vertex_ids"]),
                                                    dtype=torch.long, device=device)
        self.J_regressor_extra = None
        if model.get("J_regressor_extra") is not None:
            self.J_regressor_extra = t(model["J_regressor_extra"])
        self.joint_map = None if joint_map is None else torch.as_tensor(
            np.asarray(joint_map), dtype=torch.long, device=device)
        self.joint_mapper = None if joint_mapper is None else torch.as_tensor(
            np.asarray(joint_mapper), dtype=torch.long, device=device)
        self.has_hands = "hands_componentsl" in model
        self.num_pca_comps = num_pca_comps
        if self.has_hands:
            self.comp_l = t(model["hands_componentsl"])[:num_pca_comps]
            self.comp_r = t(model["hands_componentsr"])[:num_pca_comps]
            if flat_hand_mean:
                ml = torch.zeros(45, dtype=dtype, device=device)
                mr = torch.zeros(45, dtype=dtype, device=device)
            else:
                ml, mr = t(model["hands_meanl"]), t(model["hands_meanr"])
            self.pose_mean = torch.cat([torch.zeros(66, dtype=dtype, device=device), ml, mr])
        else:
            self.pose_mean = torch.zeros(self.nj * 3, dtype=dtype, device=device)

    def forward(self, betas, global_orient, body_pose, left_hand_pose=None,
                right_hand_pose=None, transl=None, use_pca=True, wrapper_extra=False):
        """[upstream-smplx] SMPLH.forward / SMPL.forward and, when
        wrapper_extra, the tail of models/smplh.py:29-31."""
        if self.has_hands:
            if use_pca:
                left_hand_pose = torch.einsum("bi,ij->bj", left_hand_pose, self.comp_l)
                right_hand_pose = torch.einsum("bi,ij->bj", right_hand_pose, self.comp_r)
            full_pose = torch.cat([global_orient, body_pose, left_hand_pose, right_hand_pose], 1)
            full_pose = full_pose + self.pose_mean
        else:
            full_pose = torch.cat([global_orient, body_pose], 1)
        verts, J_tr, v_posed, A = torch_lbs(betas, full_pose, self.v_template, self.shapedirs,
                                            self.posedirs, self.J_regressor, self.parents,
                                            self.lbs_weights)
        joints = J_tr
        if self.extra_vertex_ids is not None:
            joints = torch.cat([joints, verts[:, self.extra_vertex_ids]], dim=1)
        if self.joint_mapper is not None:
            joints = torch.index_select(joints, 1, self.joint_mapper)
        joints_fk = J_tr
        if transl is not None:
            joints = joints + transl[:, None]
            verts = verts + transl[:, None]
            joints_fk = joints_fk + transl[:, None]
        if wrapper_extra and self.J_regressor_extra is not None:
            extra = torch.einsum("bik,ji->bjk", verts, self.J_regressor_extra)
            joints = torch.cat([joints, extra], dim=1)
            if self.joint_map is not None:
                joints = joints[:, self.joint_map]
        # upstream returns the hand poses AFTER the PCA projection (45-D), before the mean is added
        return OracleOutput(verts, joints, full_pose, v_posed, A, joints_fk, left_hand_pose, right_hand_pose)

    def forward_full_pose(self, betas, full_pose, transl=None):
        """lbs on an already assembled (B,3J) axis-angle pose (numpy-twin style input)."""
        verts, J_tr, v_posed, A = torch_lbs(betas, full_pose, self.v_template, self.shapedirs,
                                            self.posedirs, self.J_regressor, self.parents,
                                            self.lbs_weights)
        joints = J_tr
        if transl is not None:
            joints = joints + transl[:, None]
            verts = verts + transl[:, None]
        return OracleOutput(verts, joints, full_pose, v_posed, A, joints)


def torch_vertex_l2_grads(om, betas, full_pose, transl, target_verts, joints_weight=0.0,
                          target_joints=None):
    """Config-3 loss: L = sum ||V - V*||^2 (+ w * sum ||J - J*||^2); returns
    (loss, d_betas, d_pose, d_transl) by autograd."""
    betas = betas.clone().requires_grad_(True)
    full_pose = full_pose.clone().requires_grad_(True)
    transl = transl.clone().requires_grad_(True)
    out = om.forward_full_pose(betas, full_pose, transl)
    loss = ((out.vertices - target_verts) ** 2).sum()
    if joints_weight and target_joints is not None:
        loss = loss + joints_weight * ((out.joints - target_joints) ** 2).sum()
    loss.backward()
    return loss.detach(), betas.grad, full_pose.grad, transl.grad
