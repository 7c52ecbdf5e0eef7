"""Motion-clip / fit-result / model ingestion: the on-disk side of the path.

    read_amsass(path)      <-> model2video.read_amsass   (lib/model2video.py:527-531):
                               poses[:, :72] and trans - trans[0]; `full=True` keeps all 156 pose
                               columns (hands) and the clip's betas for the SMPL-H replay.
    read_mixamo(path)      <-> model2video_miaxmo.read_mixamo (lib/model2video_miaxmo.py:544-551):
                               smpl_array.reshape(anim_len, -1), cam_array
    read_fit_result(path)  <-> the smplh.pkl a fit writes (main.py:50-59, lib/model2video.py:533-541):
                               spmlh_pose (156) / spmlh_shape (10) / camera_* as float64
    load_model(path)       <-> models/smplh_np.py:8-17, models/smpl_np.py:124-133 (pickle) or .npz
    clip_to_device(...)    packed float32 device buffers (pinned staging) ready for
                           DeviceModel.forward / SMPLH / RecoverModel.replay; frames are independent
                           bodies, so a clip is one batched call.
"""
import pickle
from collections import namedtuple

import numpy as np

Clip = namedtuple("Clip", ["poses", "trans", "betas", "extra"])
FitResult = namedtuple("FitResult", ["pose", "betas", "camera_rotation", "camera_translation",
                                     "camera_center", "raw"])


def read_amsass(path, full=False):
    """AMASS .npz: (poses, root_trans) exactly as the reference reads them; with `full` the 156-D
    poses (body + hands) and betas are kept (BASELINE config 5)."""
    d = np.load(path, allow_pickle=True)
    poses = np.asarray(d["poses"])
    poses = poses if full else poses[:, :72]
    trans = np.asarray(d["trans"]) - np.asarray(d["trans"])[0]
    betas = np.asarray(d["betas"]) if "betas" in d.files else None
    extra = {k: d[k] for k in d.files if k not in ("poses", "trans", "betas")}
    return Clip(poses, trans, betas, extra)


def read_mixamo(path):
    """Mixamo result.pkl: (pose_array (anim_len, 72), cam_array)."""
    with open(path, "rb") as f:
        m = pickle.load(f, encoding="iso-8859-1")
    n = int(m["anim_len"])
    poses = np.asarray(m["smpl_array"]).reshape(n, -1)
    return Clip(poses, None, None, {"cam_array": np.asarray(m["cam_array"])})


def read_fit_result(path):
    """smplh.pkl written by a fit: pose (52,3) float64, betas (10,) float64, camera parameters."""
    with open(path, "rb") as f:
        r = pickle.load(f, encoding="iso-8859-1")

    def get(k):
        return np.asarray(r[k]).astype("float64") if k in r else None
    return FitResult(get("spmlh_pose").reshape(-1, 3), get("spmlh_shape"), get("camera_rotation"),
                     get("camera_translation"), get("camera_center"), r)


def load_model(path):
    """Body-model tensors from a pickle (models/smplh_np.py:8-17) or an .npz with the same keys."""
    if str(path).endswith(".npz"):
        d = np.load(path, allow_pickle=True)
        return {k: d[k] for k in d.files}
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def clip_to_device(clip, num_joints, device=0, ignore_joints=(), betas=None):
    """(pose (N, 3*num_joints), transl (N,3) or None, betas (1,NB) or None) as float32 CUDA tensors,
    staged through pinned host memory.  `ignore_joints` are zeroed (RecoverModel.set_params,
    lib/model2video.py:44-45)."""
    import torch
    dev = torch.device("cuda", device)
    pose_np = pack_clip(clip, num_joints, ignore_joints)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().to(dev, non_blocking=True)
    pose = up(pose_np)
    transl = up(clip.trans) if clip.trans is not None else None
    b = betas if betas is not None else clip.betas
    bt = up(np.asarray(b).reshape(1, -1)) if b is not None else None
    return pose, transl, bt


def pack_clip(clip, num_joints, ignore_joints=()):
    """float32 (N, 3*num_joints) pose rows: the clip's columns, zero padded / truncated to the
    skeleton, with `ignore_joints` zeroed."""
    n = clip.poses.shape[0]
    p = np.zeros((n, 3 * num_joints), dtype=np.float32)
    k = min(p.shape[1], clip.poses.shape[1])
    p[:, :k] = clip.poses[:, :k]
    if len(ignore_joints):
        p.reshape(n, num_joints, 3)[:, list(ignore_joints)] = 0.0
    return p
