import _bootstrap  # noqa: F401
from pcfm.nn import PointNetAModule, PointNetFPModule, PointNetSAModule

__all__ = ["PointNetAModule", "PointNetSAModule", "PointNetFPModule"]
