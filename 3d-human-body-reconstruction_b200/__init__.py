"""smplk -- B200-native SMPL / SMPL-H body-model forward + backward (sm_100a CUDA behind a C ABI).

Drop-in for the body-model path of bokchoy-mian/3D-human-body-reconstruction:
  torch modules   SMPL, SMPLH            (models/smpl.py, models/smplh.py)
  numpy twins     SMPLModel, SMPLHModel  (models/smpl_np.py, models/smplh_np.py)
  rigged mesh     RecoverModel           (lib/model2video.py:12-130, lib/mesh2smpl_model.py:131-313)
"""
from . import synthetic  # noqa: F401
from . import _lib  # noqa: F401
from ._lib import DeviceModel, build, load  # noqa: F401


def __getattr__(name):  # torch-dependent modules load lazily
    if name in ("SMPL", "SMPLH", "ModelOutput", "create", "body_model_apply", "load_model_file", "vertex_l2_loss",
                "fit_vertex_l2"):
        from . import body_models
        return getattr(body_models, name)
    if name in ("MeshTopology", "inverse_lbs", "inverse_joints", "transforms"):
        from . import mesh_ops
        return getattr(mesh_ops, name)
    if name in ("SMPLifyLoss", "PerspectiveCamera", "GraphedClosure", "reprojection_loss", "fit_priors"):
        from . import fitting
        return getattr(fitting, name)
    if name in ("SMPLModel", "SMPLHModel", "RecoverModel"):
        from . import np_twins
        return getattr(np_twins, name)
    raise AttributeError(name)
