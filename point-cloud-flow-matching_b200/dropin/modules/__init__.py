import _bootstrap  # noqa: F401
from pcfm.nn import (BallQuery, KLLoss, PointNetAModule, PointNetFPModule, PointNetSAModule, PVConv, SE3d,
                     SharedMLP, Voxelization)


def __getattr__(name):
    if name == "FrustumPointNetLoss":
        raise ImportError("FrustumPointNetLoss (KITTI detection loss) is outside the hot-path scope (SURVEY.md 8)")
    raise AttributeError(name)
