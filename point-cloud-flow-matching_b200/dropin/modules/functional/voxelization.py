import _bootstrap  # noqa: F401
from pcfm.functional import AvgVoxelization, avg_voxelize

__all__ = ["avg_voxelize"]
