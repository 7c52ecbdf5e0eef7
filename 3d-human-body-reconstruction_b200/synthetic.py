"""Seeded synthetic SMPL / SMPL-H model tensors of the canonical shapes.

The official model pickles (`models/model/smpl/*.pkl`, `models/model/smplh/*.pkl`,
listed in the reference's `.MISSING_LARGE_BLOBS`) are not available offline, so every
test and benchmark uses a model dict with the same keys and shapes the reference
loads at `models/smplh_np.py:11-17` / `models/smpl_np.py:127-133`:

    v_template (V,3)  shapedirs (V,3,NB)  posedirs (V,3,P)  J_regressor (J,V)
    weights (V,J)     kintree_table (2,J) f (F,3)

plus, for SMPL-H, the MANO hand-PCA keys upstream smplx reads
(`hands_componentsl/r` (45,45), `hands_meanl/r` (45,)).

Value distributions follow SURVEY.md section 8(d).
"""
import numpy as np

V_SMPL = 6890
F_SMPL = 13776

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17,
                18, 19, 20, 21]
SMPLH_PARENTS = (SMPL_PARENTS[:22]
                 + [20, 22, 23, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35]
                 + [21, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50])

# upstream smplx VertexJointSelector ids for 'smplh' (data, not logic; see SURVEY 8a footnote)
SMPLH_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583,            # nose, reye, leye, rear, lear
                          3216, 3226, 3387, 6617, 6624, 6787,    # feet
                          2746, 2319, 2445, 2556, 2673,          # left finger tips
                          6191, 5782, 5905, 6016, 6133]          # right finger tips


def kintree_from_parents(parents):
    """(2,J) table as stored in the pickles: row 0 = parent id (root = 2**32-1), row 1 = own id."""
    J = len(parents)
    kt = np.zeros((2, J), dtype=np.int64)
    kt[0] = [p if p >= 0 else 2 ** 32 - 1 for p in parents]
    kt[1] = np.arange(J)
    return kt


def parents_from_kintree(kintree_table):
    """Same mapping the reference builds at models/smplh_np.py:19-23; root gets -1."""
    kt = np.asarray(kintree_table).astype(np.int64)
    id_to_col = {int(kt[1, i]): i for i in range(kt.shape[1])}
    parents = [-1]
    for i in range(1, kt.shape[1]):
        parents.append(id_to_col[int(kt[0, i])])
    return np.asarray(parents, dtype=np.int32)


def _sparse_rows(rng, rows, cols, nnz, centers=None, spread=None):
    """Non-negative rows summing to 1 with `nnz` non-zeros each."""
    M = np.zeros((rows, cols), dtype=np.float64)
    for r in range(rows):
        if centers is None:
            idx = rng.choice(cols, size=nnz, replace=False)
        else:
            lo = max(0, int(centers[r] - spread))
            hi = min(cols, int(centers[r] + spread))
            idx = rng.choice(np.arange(lo, hi), size=min(nnz, hi - lo), replace=False)
        w = rng.random(len(idx)) + 0.05
        M[r, idx] = w / w.sum()
    return M


def _skin_weights(rng, V, parents, max_nnz):
    """Canonically sparse LBS weights: each vertex is bound to a primary joint (contiguous
    vertex ranges share a joint, as in the real models), plus up to max_nnz-1 joints from
    the primary's tree neighbourhood (parent, children, grand-parent)."""
    J = len(parents)
    children = [[] for _ in range(J)]
    for j, p in enumerate(parents):
        if p >= 0:
            children[p].append(j)
    W = np.zeros((V, J), dtype=np.float64)
    # uneven ranges so joints own different numbers of vertices
    cuts = np.sort(rng.choice(np.arange(1, V), size=J - 1, replace=False))
    bounds = np.concatenate([[0], cuts, [V]])
    owner = rng.permutation(J)
    for r in range(J):
        j = int(owner[r])
        nb = [j]
        if parents[j] >= 0:
            nb.append(parents[j])
            if parents[parents[j]] >= 0:
                nb.append(parents[parents[j]])
        nb += children[j]
        nb = list(dict.fromkeys(nb))
        for v in range(bounds[r], bounds[r + 1]):
            k = int(rng.integers(1, max_nnz + 1))
            sel = [j] + list(rng.permutation(nb[1:])[:k - 1])
            w = rng.random(len(sel)) + 0.02
            w[0] += 1.0
            W[v, sel] = w / w.sum()
    return W


def make_model(kind="smplh", num_betas=None, seed=0, dense_weights=False,
               dense_regressor=False, num_verts=V_SMPL, max_nnz=4, dtype=np.float64):
    """Return a model dict with the pickle keys of the reference.

    kind: 'smpl' (J=24, P=207, NB=10 default) or 'smplh' (J=52, P=459, NB=16 default).
    dense_weights=True gives Dirichlet-like dense skin weights (correctness case).
    """
    rng = np.random.default_rng(seed)
    if kind == "smpl":
        parents = SMPL_PARENTS
        nb = 10 if num_betas is None else num_betas
    elif kind == "smplh":
        parents = SMPLH_PARENTS
        nb = 16 if num_betas is None else num_betas
    else:
        raise ValueError("kind must be 'smpl' or 'smplh'")
    J = len(parents)
    P = 9 * (J - 1)
    V = num_verts
    m = {}
    m["v_template"] = rng.standard_normal((V, 3)) * np.array([0.3, 0.5, 0.1])
    m["shapedirs"] = rng.standard_normal((V, 3, nb)) * 0.01
    m["posedirs"] = rng.standard_normal((V, 3, P)) * 0.001
    if dense_regressor:
        R = rng.random((J, V))
        m["J_regressor"] = R / R.sum(1, keepdims=True)
    else:
        centers = rng.integers(0, V, size=J)
        m["J_regressor"] = _sparse_rows(rng, J, V, 32, centers=centers, spread=200)
    if dense_weights:
        W = rng.random((V, J)) ** 4
        m["weights"] = W / W.sum(1, keepdims=True)
    else:
        m["weights"] = _skin_weights(rng, V, parents, max_nnz)
    m["kintree_table"] = kintree_from_parents(parents)
    m["f"] = rng.integers(0, V, size=(F_SMPL, 3)).astype(np.uint32)
    if kind == "smplh":
        for side in ("l", "r"):
            q, _ = np.linalg.qr(rng.standard_normal((45, 45)))
            m["hands_components" + side] = q
            m["hands_mean" + side] = rng.standard_normal(45) * 0.1
        m["extra_vertex_ids"] = np.asarray(
            [i % V for i in SMPLH_EXTRA_VERTEX_IDS], dtype=np.int32)
    else:
        m["extra_vertex_ids"] = np.asarray(
            [i % V for i in SMPLH_EXTRA_VERTEX_IDS], dtype=np.int32)
    centers = rng.integers(0, V, size=9)
    m["J_regressor_extra"] = _sparse_rows(rng, 9, V, 32, centers=centers, spread=200)
    for k, a in m.items():
        if isinstance(a, np.ndarray) and a.dtype == np.float64:
            m[k] = a.astype(dtype)
    m["kind"] = kind
    return m


def make_rigged_mesh(num_verts=20000, seed=0, max_nnz=4):
    """Synthetic stand-in for the un-shipped `recover.pkl` the reference animates
    (keys per lib/model2video.py:16-26): an arbitrary-vertex-count mesh bound to the 24-joint
    SMPL skeleton with fixed rest joints `J` and a `parent` dict."""
    rng = np.random.default_rng(seed)
    parents = SMPL_PARENTS
    J = len(parents)
    m = {}
    m["v_template"] = rng.standard_normal((num_verts, 3)) * np.array([0.3, 0.5, 0.1])
    m["weights"] = _skin_weights(rng, num_verts, parents, max_nnz)
    m["kintree_table"] = kintree_from_parents(parents)
    m["J"] = rng.standard_normal((J, 3)) * np.array([0.3, 0.5, 0.1])
    m["parent"] = {i: parents[i] for i in range(1, J)}
    m["f"] = rng.integers(0, num_verts, size=(2 * num_verts, 3)).astype(np.uint32)
    m["color"] = rng.random((num_verts, 3))
    m["or_pose"] = np.zeros((J, 3))
    return m


def make_inputs(model, batch, seed=0, pose_sigma=0.3, dtype=np.float32, broadcast_betas=False):
    """Random (betas, full axis-angle pose, transl) per SURVEY 8(d)."""
    rng = np.random.default_rng(1000 + seed)
    J = model["kintree_table"].shape[1]
    nb = model["shapedirs"].shape[2]
    betas = rng.standard_normal((1 if broadcast_betas else batch, nb)).astype(dtype)
    pose = (rng.standard_normal((batch, J * 3)) * pose_sigma).astype(dtype)
    transl = rng.standard_normal((batch, 3)).astype(dtype)
    return betas, pose, transl
