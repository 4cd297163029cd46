import _bootstrap  # noqa: F401
from pcfm.nn import Voxelization

__all__ = ["Voxelization"]
